"""ORACLE shim: lets the read-only reference import without ruamel.yaml (absent here)."""
