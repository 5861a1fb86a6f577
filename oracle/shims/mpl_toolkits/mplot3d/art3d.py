"""Shim for mpl_toolkits.mplot3d.art3d (core_deism_arg.py:12-14)."""


class Poly3DCollection:  # pragma: no cover
    def __init__(self, *args, **kwargs):
        raise RuntimeError("matplotlib is a shim in this container")
