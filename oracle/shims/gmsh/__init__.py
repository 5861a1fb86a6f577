"""Shim for gmsh (room_check.py:11 via deism/__init__.py:5); geometry import is out of scope."""
