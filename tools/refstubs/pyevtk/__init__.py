"""``pyevtk`` stand-in (tools/export.py:18)."""
