"""Shim for matplotlib.pyplot."""
rcParams = {}


def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise RuntimeError("matplotlib is a shim in this container: pyplot.%s" % name)

    return _unavailable
