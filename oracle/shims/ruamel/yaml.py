"""ORACLE shim for `ruamel.yaml` (reference crafter/constants.py:3,6-7): PyYAML's safe loader
keeps mapping order, which is semantics here (inventory / achievement order)."""
import yaml as _pyyaml


class YAML:

  def __init__(self, typ=None, pure=False):
    pass

  def load(self, text):
    return _pyyaml.safe_load(text)
