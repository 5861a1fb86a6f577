"""CPU: pins oracle/convex_oracle.c against the compiled reference engine (oracle/_ref, when
built) and against the reference-minted fixtures.  Bit-exact everywhere, attenuation included
(glibc acosf/cosf transcribed)."""
import os

import numpy as np
import pytest

from conftest import Golden, golden_names
from oracle import oracle

CASES = golden_names("x") + ["c4_full"]


@pytest.mark.parametrize("name", CASES)
def test_convex_engine_outputs_bit_equal(name):
    g = Golden(name)
    wt = oracle.walls_from_golden(g)
    e, p = g.group("engine"), g.group("params")
    out = oracle.convex_images(wt, p["posSource"], p["posReceiver"], int(p["maxReflOrder"]))
    assert out["sources"].shape == e["sources"].shape
    for key in ("sources", "attenuations", "reflection_matrix"):
        assert np.array_equal(out[key].view(np.uint32), e[key].view(np.uint32)), key
    assert np.array_equal(out["orders"], e["orders"])
    assert np.array_equal(out["gen_walls"], e["gen_walls"])


def _libroom():
    try:
        from oracle import ref_import
        return ref_import.load_libroom()
    except Exception:
        return None


@pytest.mark.skipif(_libroom() is None, reason="oracle/_ref/libroom_deism not built")
def test_primitives_against_compiled_reference():
    """acosf/cosf chain via Wall_deism.get_attenuation, reflect, side on random inputs."""
    import ctypes as C

    lr = _libroom()
    rng = np.random.default_rng(5)
    corners = np.array([[0, 0, 0], [4, 0, 0.5], [4, 3, 0.5], [0, 3, 0]], dtype=np.float32).T
    centroid = np.array([2.0, 1.5, -3.0], dtype=np.float32)
    z = np.array([18.0, 7.5, 33.25], dtype=np.float32)
    w = lr.Wall_deism(corners, centroid, z, np.zeros(3, np.float32), np.zeros(3, np.float32), "t")
    L = oracle.lib()
    L.oracle_wall_attenuation.restype = C.c_float
    L.oracle_acosf.restype = C.c_float
    for _ in range(20000):
        theta = np.float32(rng.uniform(0, np.pi / 2))
        got = np.array([L.oracle_wall_attenuation(C.c_float(float(zz)), C.c_float(float(theta))) for zz in z],
                       dtype=np.float32)
        assert np.array_equal(got, np.asarray(w.get_attenuation(float(theta)), dtype=np.float32))
    xs = rng.uniform(-1, 1, 200000).astype(np.float32)
    libm = C.CDLL("libm.so.6")       # the acosf the reference binary resolves to
    libm.acosf.restype = C.c_float
    libm.acosf.argtypes = [C.c_float]
    for x in xs[:50000]:
        assert L.oracle_acosf(C.c_float(float(x))) == libm.acosf(float(x))
    nrm, org = np.asarray(w.normal, np.float32), np.asarray(w.origin, np.float32)
    for _ in range(5000):
        pt = rng.uniform(-6, 6, 3).astype(np.float32)
        ref_out = np.zeros(3, np.float32)
        d_ref = w.reflect(pt, ref_out)
        out = np.zeros(3, np.float32)
        d = L.oracle_wall_reflect(nrm.ctypes.data_as(C.c_void_p), org.ctypes.data_as(C.c_void_p),
                                  pt.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert d == d_ref and np.array_equal(out, ref_out)
        assert L.oracle_wall_side(nrm.ctypes.data_as(C.c_void_p), org.ctypes.data_as(C.c_void_p),
                                  pt.ctypes.data_as(C.c_void_p)) == w.side(pt)
