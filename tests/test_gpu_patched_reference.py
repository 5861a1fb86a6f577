"""The UNMODIFIED reference class (deism.core_deism.DEISM, core_deism.py:44-769) driving this engine
through deism_b200.patch.install() on a GPU — the drop-in contract of SURVEY 8(b) end to end.

The reference's Python package must be importable on the GPU box.  /root/reference does not exist
there; the one sanctioned carrier is an offline `pip install --no-deps --target baseline/_ref
/root/reference` done in the build container (git-ignored, travels with the snapshot).  Without it
these tests skip.  Nothing else in the suite, smoke() or bench.py touches baseline/_ref.

What is compared: params["RTF"] of the reference class with the plug points swapped (image
generation, Wigner tables, run_DEISM / run_DEISM_ARG, the convex room engine) against the RTF the
same class produced with its own Numba / libroom code when the fixtures were minted
(oracle/make_golden.py, same call sequences).
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, Golden, rel_l2

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
CONFIGS = os.path.join(ROOT, "deism_b200", "configs")
RTF_TOL = 1e-9     # north star: 1e-8


@pytest.fixture
def reference(monkeypatch):
    if not os.path.isdir(os.path.join(REF, "deism")):
        pytest.skip("baseline/_ref/deism not installed (pip install --target baseline/_ref /root/reference)")
    monkeypatch.setenv("DEISM_REFERENCE_ROOT", REF)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_import

    monkeypatch.setattr(ref_import, "REF_ROOT", REF)
    ref_import.import_reference()
    import deism_b200.patch as patch

    # the reference looks its YAML up relative to the CWD (data_loader.py:398-409); the packaged
    # configs are the reference's, key for key and value for value (tests/test_class_cpu.py)
    monkeypatch.chdir(CONFIGS)
    monkeypatch.setattr(sys, "argv", [sys.argv[0]])
    names = patch.install()
    assert "run_DEISM" in names and "_init_room" in names
    yield ref_import
    patch.uninstall()


def _inject(d, N, V, seed=0):
    """oracle/make_golden.py inject_directivities: synthetic coefficients + the (patched) Wigner tables."""
    from deism.core_deism import pre_calc_Wigner, vectorize_C_nm_s, vectorize_C_vu_r

    from deism_b200 import workloads

    p = d.params
    K = len(p["waveNumbers"])
    p["C_nm_s"], p["C_vu_r"] = workloads.synth_coeffs(K, N, V, seed)
    p["sourceOrder"], p["receiverOrder"] = N, V
    if p["DEISM_method"] in ("LC", "MIX"):
        d.params = vectorize_C_nm_s(d.params)
        d.params = vectorize_C_vu_r(d.params)
    if p["DEISM_method"] in ("ORG", "MIX"):
        d.params = pre_calc_Wigner(d.params)


def _shoebox(name, mode, method, order, N=None, V=None, overrides=None, materials=None):
    from deism.core_deism import DEISM

    from deism_b200 import _lib

    d = DEISM(mode, "shoebox", silent=True)
    p = d.params
    p["DEISM_method"], p["maxReflOrder"] = method, order
    for key, val in (overrides or {}).items():
        p[key] = val
    d.update_wall_materials(*(materials or ()))
    d.update_freqs()
    if N is None:
        d.update_directivities()
    else:
        _inject(d, N, V)
    n0 = _lib.launch_count()
    d.update_source_receiver()
    assert d.params["images"].get("storage") == "fused"          # this engine's generator ran
    d.run_DEISM(if_clean_up=False)
    assert _lib.launch_count() > n0
    g = Golden(name)
    assert d.params["RTF"].dtype == np.complex128 and d.params["RTF"].shape == g.rtf.shape
    assert rel_l2(d.params["RTF"], g.rtf) < RTF_TOL, name
    return d


def test_reference_class_shoebox_rtf_mix_monopole(reference):
    d = _shoebox("s1_shoebox_rtf_mix_o5_mono", "RTF", "MIX", 5, overrides={"endFreq": 200})
    # SURVEY 8c smoke values of the reference
    assert abs(d.params["RTF"][0] - (1.12751211 - 0.32203787j)) < 1e-7


def test_reference_class_shoebox_org_and_lc_with_directivities(reference):
    _shoebox("s3_shoebox_rtf_org_o4_N3V2", "RTF", "ORG", 4, N=3, V=2,
             overrides={"startFreq": 50, "endFreq": 2000, "freqStep": 50})
    Zb = np.array([[18, 20, 8, 5, 30, 19], [12 + 3j, 25 - 2j, 9 + 1j, 6, 28 + 4j, 15 - 1j],
                   [10, 30, 7, 4, 35, 12]]).T
    _shoebox("s2_shoebox_rtf_lc_o6_N3", "RTF", "LC", 6, N=3, V=3,
             overrides={"startFreq": 20, "endFreq": 1280, "freqStep": 20},
             materials=(Zb, np.array([100.0, 500.0, 1000.0]), "impedance"))


def test_reference_class_shoebox_rir_mode(reference):
    d = _shoebox("s4_shoebox_rir_mix_o9_N2", "RIR", "MIX", 9, N=2, V=2,
                 overrides={"sampleRate": 1500, "ifRemoveDirectPath": 1},
                 materials=(0.08, None, "reverberationTime"))
    g = Golden("s4_shoebox_rir_mix_o9_N2")
    rir = d.get_results()
    want = g.z["RIR"]
    assert rir.shape == want.shape and np.linalg.norm(rir - want) / np.linalg.norm(want) < 1e-8


def test_reference_class_convex_rir_mix_monopole(reference):
    from deism.core_deism import DEISM

    from deism_b200 import convex

    d = DEISM("RIR", "convex", silent=True)
    p = d.params
    p["DEISM_method"], p["maxReflOrder"], p["sampleRate"] = "MIX", 4, 4000
    d.update_wall_materials()
    d.update_freqs()
    d.update_source_receiver()
    assert isinstance(d.room_convex.room_engine, convex.Room_deism)    # the GPU reflection tree
    d.update_directivities()
    d.run_DEISM(if_clean_up=False)
    g = Golden("x1_convex_rir_mix_o4_mono")
    eng = d.room_convex.room_engine
    assert np.array_equal(np.asarray(eng.sources).view(np.uint32), g.z["engine/sources"].view(np.uint32))
    assert np.array_equal(np.asarray(eng.orders), g.z["engine/orders"])
    assert rel_l2(d.params["RTF"], g.rtf) < RTF_TOL
