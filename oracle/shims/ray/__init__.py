"""Shim for `ray` (not installable offline).  core_deism_arg.py:16 imports it
unconditionally; parallel_backends.py:942-947 guards it.  The decorator returns the
function untouched so module import succeeds; the Ray backend itself is never run."""


def remote(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(fn):
        return fn

    return deco


def is_initialized():
    return False


def init(*args, **kwargs):
    raise RuntimeError("ray is a shim in this container; the Ray backend cannot run")


def shutdown():
    return None


def put(obj):
    return obj


def get(obj):
    return obj
