"""ctypes binding of libdeism_b200.so (the C ABI declared in include/deism_b200.h).

The shared library is built in-tree by ``deism_b200/csrc/Makefile`` (nvcc, sm_100a only).
There is no CPU fallback: if the library is missing or no sm_100 device is present, calls
raise ``RuntimeError`` with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeism_b200.so")
CSRC = os.path.join(_HERE, "csrc")

DEISM_OK = 0
ATTEN_C64, ATTEN_C128, ATTEN_FUSED_ROOM, ATTEN_FUSED_ANGLES = 0, 1, 2, 3


class ShoeboxAtten(C.Structure):
    """struct deism_shoebox_atten (include/deism_b200.h)."""

    _fields_ = [
        ("mode", C.c_int32),
        ("ang_dep_flag", C.c_int32),
        ("d_atten", C.c_void_p),
        ("d_A", C.c_void_p),
        ("d_R_sI_r", C.c_void_p),
        ("d_Z_S", C.c_void_p),
        ("room_size", C.c_double * 3),
        ("pos_source", C.c_double * 3),
        ("pos_receiver", C.c_double * 3),
    ]


class ConvexWalls(C.Structure):
    """struct deism_convex_walls (include/deism_b200.h)."""

    _fields_ = [
        ("n_walls", C.c_int32),
        ("max_corners", C.c_int32),
        ("h_normal", C.c_void_p),
        ("h_origin", C.c_void_p),
        ("h_basis", C.c_void_p),
        ("h_flat_corners", C.c_void_p),
        ("h_n_corners", C.c_void_p),
        ("h_reflection", C.c_void_p),
        ("h_impedance", C.c_void_p),
    ]


_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
_SIGNATURES = {
    "deism_b200_version": (C.c_int, []),
    "deism_b200_source_hash": (C.c_char_p, []),
    "deism_last_error": (C.c_char_p, []),
    "deism_device_info": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "deism_release_workspace": (C.c_int, []),
    "deism_kernel_launch_count": (C.c_uint64, []),
    "deism_chunk_plan": (C.c_int, [C.c_int64] * 6 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "deism_state_generation": (C.c_uint64, []),
    "deism_graph_replayed": (C.c_int, [_i32, _vp]),
    "deism_profile_enable": (C.c_int, [_i32]),
    "deism_profile_reset": (C.c_int, []),
    "deism_profile_get": (C.c_int, [C.c_char_p, _vp, _vp]),
    "deism_shoebox_count_images": (C.c_int, [_vp, _vp, _vp, _dbl, _i32, _i32, _i32, _vp, _vp, _vp]),
    "deism_shoebox_fill_images": (C.c_int, [_vp, _vp, _vp, _dbl, _i32, _i32, _i32, _vp, _vp]
                                  + [_vp] * 8 + [_vp]),
    "deism_shoebox_attenuation": (C.c_int, [_i64, _i64, _vp, _vp, _i32, _vp]),
    "deism_lc_shoebox": (C.c_int, [_i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _i32, _vp]),
    "deism_org_plan_create": (C.c_int, [_i32, _i32, _vp, _vp, _vp]),
    "deism_org_plan_destroy": (C.c_int, [_vp]),
    "deism_org_shoebox": (C.c_int, [_i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _i32, _i32, _vp, _vp]),
    "deism_org_shoebox_c128": (C.c_int, [_i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _i32, _vp, _vp]),
    "deism_lc_arg": (C.c_int, [_i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _i64, _vp, _i32, _vp]),
    "deism_org_arg": (C.c_int, [_i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _i64, _vp, _i32, _vp, _vp]),
    "deism_wigner_tables": (C.c_int, [_i32, _i32, _vp, _vp]),
    "deism_pchip_interp": (C.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "deism_arg_refit": (C.c_int, [_i64, _i64, _i32, _i64, _vp, _vp, _vp, _vp, C.c_double, _vp, _vp]),
    "deism_convex_images_count": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "deism_convex_images_fill": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "deism_fp64_peak": (C.c_int, [_i32, _i32, _vp, _vp]),
}
EXPORTED = tuple(_SIGNATURES)

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libdeism_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=out)
    return LIB_PATH


_HASHED = ("api.cu", "images.cu", "lc.cu", "org.cu", "arg.cu", "convex.cu", "wigner.cu", "refit.cu",
           "materials.cu", "common.cuh", os.path.join("..", "..", "include", "deism_b200.h"))


def source_hash() -> str:
    """Fingerprint of the sources next to this file — the Makefile's SRCHASH (same files, same
    order); None when the sources are not there (a binary-only deployment)."""
    import hashlib

    h = hashlib.sha256()
    try:
        for name in _HASHED:
            with open(os.path.join(CSRC, name), "rb") as fh:
                h.update(fh.read())
    except OSError:
        return None
    return h.hexdigest()[:16]


def _stale(path) -> bool:
    """True when the .so at `path` was built from other sources than the ones in the tree.  The
    fingerprint is read from the file (marker string), not through dlopen: a probe handle would
    pin the old image in the process and the rebuilt file could not replace it."""
    want = source_hash()
    if want is None or os.environ.get("DEISM_B200_SKIP_HASH") == "1":
        return False
    marker = b"DEISM_SRC_HASH="
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return True
    i = blob.find(marker)
    return i < 0 or blob[i + len(marker):i + len(marker) + 16].decode("ascii", "replace") != want


def lib():
    """Load the shared library.  It is (re)built first when the in-tree .so is absent or was built
    from other sources than the ones next to it (the .so is git-ignored, so a checkout or an edit
    can leave an old binary behind); a binary that is still stale after the build is refused."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        elif _stale(LIB_PATH):
            build()
            if _stale(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} does not match the sources in {CSRC} "
                                   "(deism_b200_source_hash) and rebuilding did not help")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the header and the .so diverge
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != DEISM_OK:
        msg = lib().deism_last_error().decode("utf-8", "replace")
        exc = {1: ValueError, 3: MemoryError}.get(rc, RuntimeError)
        raise exc(f"libdeism_b200 {what}: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().deism_kernel_launch_count())
