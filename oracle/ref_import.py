"""oracle/ref_import.py — TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference package from /root/reference (read-only) inside this
container so that golden vectors can be minted from the reference's own ``run_DEISM`` and
so the C restatement (oracle/deism_oracle.c) can be validated against it.

Nothing in the product (deism_b200/) may import this module.  /root/reference does not
exist on the GPU box; callers must check ``reference_available()`` first.

What is done:
  * oracle/shims/ is put on sys.path: sound_field_analysis.sph.sphankel2, ray,
    matplotlib, mpl_toolkits, gmsh — modules the reference imports at module level but
    which are not installed here.
  * scipy.special.sph_harm (removed from this SciPy) is re-created from sph_harm_y with the
    reference's (m, n, azimuth, polar) argument order (core_deism.py:1257,1328).
  * the pybind11 extension ``deism.libroom_deism`` is loaded from oracle/_ref/ (built by
    oracle/Makefile from the reference's own C++ sources) and registered under that name.
  * sys.argv is trimmed (data_loader.py:610 parses it) and CWD is set to the reference root
    while a DEISM object is built (YAML lookup is CWD-relative, data_loader.py:398-409).
"""
from __future__ import annotations

import contextlib
import glob
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("DEISM_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
_REF_BUILD = os.path.join(_HERE, "_ref")

_state = {"deism": None}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "deism"))


def libroom_path():
    hits = glob.glob(os.path.join(_REF_BUILD, "libroom_deism*.so"))
    return hits[0] if hits else None


def load_libroom(as_name: str = "libroom_deism"):
    """Load the reference's compiled pybind11 module from oracle/_ref (works without
    /root/reference: the .so is self-contained)."""
    path = libroom_path()
    if path is None:
        raise ImportError("oracle/_ref/libroom_deism*.so not built (make -C oracle ref)")
    # the pybind11 types can be registered once per process: one module object, two names
    for name in (as_name, "libroom_deism", "deism.libroom_deism"):
        if name in sys.modules:
            sys.modules[as_name] = sys.modules[name]
            return sys.modules[name]
    spec = importlib.util.spec_from_file_location(as_name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[as_name] = mod
    return mod


def import_reference():
    """Return the reference's ``deism`` package (imported once)."""
    if _state["deism"] is not None:
        return _state["deism"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REF_ROOT}")
    import scipy.special as sps

    if not hasattr(sps, "sph_harm"):
        sps.sph_harm = lambda m, n, az, pol: sps.sph_harm_y(n, m, pol, az)
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REF_ROOT not in sys.path:
        sys.path.insert(1, REF_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(_REF_BUILD, "numba_cache"))
    load_libroom("deism.libroom_deism")
    argv = sys.argv
    sys.argv = [argv[0] if argv else "oracle"]
    try:
        import deism  # noqa: F401  (the reference)
    finally:
        sys.argv = argv
    _state["deism"] = sys.modules["deism"]
    return _state["deism"]


@contextlib.contextmanager
def reference_cwd():
    """CWD = reference root and a clean argv while reference constructors run."""
    old = os.getcwd()
    argv = sys.argv
    sys.argv = [argv[0] if argv else "oracle"]
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)
        sys.argv = argv


def make_deism(mode: str, roomtype: str, silent: bool = True, **overrides):
    """Build a reference ``DEISM`` object; ``overrides`` are written into ``params``
    before the update_* calls are made by the caller."""
    import_reference()
    from deism.core_deism import DEISM  # type: ignore

    with reference_cwd():
        d = DEISM(mode, roomtype, silent=silent)
    for key, val in overrides.items():
        d.params[key] = val
    return d
