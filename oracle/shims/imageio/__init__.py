"""ORACLE shim: lets the read-only reference import without imageio (absent here)."""
