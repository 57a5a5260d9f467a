"""ORACLE -- test infrastructure only.

CPU restatement of the reference hot path (danijar/crafter `Env.reset/step/render`) used as the
checker by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs.  Nothing under `crafter_b200/` may import this package; the product path is CUDA only.
"""
