"""Shim for matplotlib.patches."""


class Patch:  # pragma: no cover
    pass


class Polygon(Patch):  # pragma: no cover
    pass
