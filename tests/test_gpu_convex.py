"""GPU parity tests for the convex (DEISM-ARG) accumulation kernels against reference-minted
fixtures (reference image lists as input) and the CPU oracle."""
import numpy as np
import pytest

from conftest import Golden, golden_names, rel_l2

pytestmark = pytest.mark.gpu
RTF_TOL = 1e-9


@pytest.mark.parametrize("name", golden_names("x"))
def test_arg_accumulation_reference_layout(name):
    from deism_b200 import backend

    g = Golden(name)
    got = backend.run_DEISM_ARG(g.params())
    assert got.dtype == np.complex128 and got.shape == g.rtf.shape
    assert rel_l2(got, g.rtf) < RTF_TOL


def test_arg_unknown_method():
    from deism_b200 import backend

    g = Golden("x1_convex_rir_mix_o4_mono")
    p = g.params()
    p["DEISM_method"] = "XYZ"
    with pytest.raises(ValueError):
        backend.run_DEISM_ARG(p)


# --------------------------------------------------------------------------------------
# reflection-tree enumeration on the GPU
# --------------------------------------------------------------------------------------
def _room_from_golden(g):
    from deism_b200 import convex
    from oracle import oracle

    wt = oracle.walls_from_golden(g)
    p = g.group("params")
    room = convex.Room_deism(wt, [], [], 343.0, int(p["maxReflOrder"]))
    room.add_mic(np.asarray(p["posReceiver"], np.float32))
    return room, p


@pytest.mark.parametrize("name", golden_names("x") + ["c4_full"])
def test_convex_enumeration_bit_equal(name):
    """sources / orders / gen_walls / float32 attenuations / reflection matrices of the GPU
    tree == libroom_deism.Room_deism (reference C++), in the reference's DFS order."""
    g = Golden(name)
    room, p = _room_from_golden(g)
    n = room.image_source_model(np.asarray(p["posSource"], np.float32))
    e = g.group("engine")
    assert n == e["sources"].shape[1]
    assert np.array_equal(room.sources.view(np.uint32), e["sources"].view(np.uint32))
    assert np.array_equal(room.orders, e["orders"])
    assert np.array_equal(room.gen_walls, e["gen_walls"])
    assert np.array_equal(np.asarray(room.reflection_matrix).view(np.uint32),
                          e["reflection_matrix"].view(np.uint32))
    assert np.array_equal(room.attenuations.view(np.uint32), e["attenuations"].view(np.uint32))


@pytest.mark.parametrize("name", ["x1_convex_rir_mix_o4_mono", "c4_full"])
def test_convex_end_to_end_monopole(name):
    """BASELINE config C4 path: GPU tree -> get_ref_paths_ARG -> run_DEISM_ARG (MIX) == reference."""
    from deism_b200 import backend, convex, wigner, workloads

    g = Golden(name)
    room, gp = _room_from_golden(g)
    room.image_source_model(np.asarray(gp["posSource"], np.float32))
    k = np.asarray(gp["waveNumbers"])
    p = {"DEISM_method": "MIX", "silentMode": 1, "waveNumbers": k, "posReceiver": gp["posReceiver"],
         "ifRemoveDirectPath": int(gp["ifRemoveDirectPath"]), "mixEarlyOrder": int(gp["mixEarlyOrder"]),
         "sourceOrder": 0, "receiverOrder": 0}
    p = convex.get_ref_paths_ARG(p, room)
    ref_im = g.group("images")
    for key in ("R_sI_r_all", "atten_all", "early_indices", "late_indices"):
        assert np.array_equal(np.asarray(p["images"][key]), ref_im[key]), key
    I = p["images"]["R_sI_r_all"].shape[1]
    mono = workloads.monopole_coeffs(k)
    p["C_nm_s_ARG"] = (mono[..., None] * np.ones((1, 1, 1, I), np.complex64)).astype(np.complex64)
    p["C_vu_r"] = mono
    p["n_all"], p["m_all"], _ = workloads.vectorize(mono, 0)
    p["C_nm_s_ARG_vec"] = np.ascontiguousarray(p["C_nm_s_ARG"][:, 0, :, :])
    p["v_all"], p["u_all"], p["C_vu_r_vec"] = workloads.vectorize(mono, 0)
    p = wigner.pre_calc_Wigner(p)
    got = backend.run_DEISM_ARG(p)
    err = rel_l2(got, g.rtf)
    print(f"{name}: convex end-to-end rel L2 = {err:.3e}")
    assert err < RTF_TOL
