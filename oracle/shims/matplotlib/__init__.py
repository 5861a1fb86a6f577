"""Shim for matplotlib (plotting is out of scope; utilities.py:3 imports it)."""
rcParams = {}


def use(*args, **kwargs):
    return None
