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


def _c64_close(got, want):
    """complex64 arrays computed from values that agree to ~1e-13: equal up to rare 1-ulp flips."""
    assert got.shape == want.shape and got.dtype == want.dtype == np.complex64
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-8
    assert np.mean(got != want) < 2e-3
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-30)) < 3e-7


def test_arg_refit_reference_fixture():
    """deism_arg_refit against the reference's cal_C_nm_s_new + vectorize_C_nm_s_ARG."""
    import os

    from deism_b200 import directivities as D

    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "r1_arg_refit_N3.npz"))
    params = {"waveNumbers": f["waveNumbers"], "sourceOrder": int(f["sourceOrder"]),
              "radiusSource": float(f["radiusSource"]), "freqs": f["freqs"]}
    got = D.cal_C_nm_s_ARG_vec_device(f["reflection_matrix"], f["Psh"], f["coords"], params).cpu().numpy()
    _c64_close(got, f["C_nm_s_ARG_vec"])
    # engine layout [I,3,3] gives the same, and the 4-D mirror matches the reference's array
    got2 = D.cal_C_nm_s_ARG_vec_device(np.ascontiguousarray(np.transpose(f["reflection_matrix"], (2, 0, 1))),
                                       f["Psh"], f["coords"], params).cpu().numpy()
    assert np.array_equal(got, got2)
    _c64_close(D.cal_C_nm_s_new(f["reflection_matrix"], f["Psh"], f["coords"], params), f["C_nm_s_ARG"])


@pytest.mark.parametrize("I,K,N,n_dir", [(1, 1, 0, 1), (5, 3, 1, 4), (130, 37, 2, 31), (77, 20, 5, 200),
                                         (9, 33, 7, 150), (6, 17, 10, 300), (3, 9, 11, 333)])
def test_arg_refit_vs_oracle(I, K, N, n_dir):
    """Ragged sizes, orders 0..11, proper and improper orthogonal matrices (float32), against the
    NumPy restatement (pinv per image)."""
    from deism_b200 import directivities as D
    from oracle import oracle

    rng = np.random.default_rng(100 + I)
    Q = np.linalg.qr(rng.standard_normal((I, 3, 3)))[0]
    Q[::2] *= -1.0                                                   # mix reflections and rotations
    R = np.ascontiguousarray(np.transpose(Q, (1, 2, 0)).astype(np.float32))
    X = rng.standard_normal((3, n_dir))
    X /= np.linalg.norm(X, axis=0)
    if n_dir == 1:
        X[:] = [[0.3], [0.4], [0.5]]
    Psh = rng.standard_normal((K, n_dir)) + 1j * rng.standard_normal((K, n_dir))
    k = np.sort(rng.uniform(0.05, 300.0, K))
    params = {"waveNumbers": k, "sourceOrder": N, "radiusSource": 0.35}
    got = D.cal_C_nm_s_ARG_vec_device(R, Psh, X, params).cpu().numpy()
    want = oracle.cal_C_nm_s_ARG_vec(R, Psh, X, k, N, 0.35)
    _c64_close(got, want)


def test_arg_refit_errors():
    from deism_b200 import directivities as D

    X = np.random.default_rng(0).standard_normal((3, 10))
    P = np.ones((2, 10), complex)
    base = {"waveNumbers": np.array([1.0, 2.0]), "radiusSource": 0.4}
    with pytest.raises(ValueError):          # 10 directions cannot determine 16 coefficients
        D.cal_C_nm_s_ARG_vec_device(np.eye(3, dtype=np.float32)[None], P, X, dict(base, sourceOrder=3))
    with pytest.raises(ValueError):          # order above the kernel's budget
        D.cal_C_nm_s_ARG_vec_device(np.eye(3, dtype=np.float32)[None], np.ones((2, 200), complex),
                                    np.random.default_rng(0).standard_normal((3, 200)), dict(base, sourceOrder=12))
    with pytest.raises(ValueError):          # all directions identical: rank-deficient fit
        D.cal_C_nm_s_ARG_vec_device(np.eye(3, dtype=np.float32)[None], P, np.tile([[0.], [0.], [1.]], (1, 10)),
                                    dict(base, sourceOrder=1))


@pytest.mark.parametrize("name", ["x1_convex_rir_mix_o4_mono", "c4_full"])
def test_device_resident_convex_pipeline(name):
    """Tree enumeration -> geometry -> monopole coefficients -> accumulation with every array left
    on the GPU (get_ref_paths_ARG_device) gives the reference's RTF; the lazily copied host
    attributes are still the reference's arrays."""
    import os

    from deism_b200 import backend, convex, directivities as D, wigner, workloads

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))

    class W:
        pass

    walls = []
    for i in range(int(z["walls/n"])):
        w = W()
        w.normal, w.origin = z[f"walls/{i}/normal"], z[f"walls/{i}/origin"]
        w.basis, w.flat_corners = z[f"walls/{i}/basis"], z[f"walls/{i}/flat_corners"]
        w.reflection_matrix, w.impedence_bands = z[f"walls/{i}/reflection_matrix"], z[f"walls/{i}/impedance"]
        walls.append(w)
    room = convex.Room_deism(walls, [], [], 343.0, int(z["params/maxReflOrder"]))
    room.add_mic(z["params/posReceiver"].astype(np.float32))
    n = room.image_source_model(z["params/posSource"].astype(np.float32))
    k = z["params/waveNumbers"]
    mono = workloads.monopole_coeffs(k)
    q = {"DEISM_method": "MIX", "silentMode": 1, "waveNumbers": k, "posReceiver": z["params/posReceiver"],
         "ifRemoveDirectPath": int(z["params/ifRemoveDirectPath"]), "mixEarlyOrder": int(z["params/mixEarlyOrder"]),
         "sourceOrder": 0, "receiverOrder": 0, "sourceType": "monopole", "C_vu_r": mono}
    q["v_all"], q["u_all"], q["C_vu_r_vec"] = workloads.vectorize(mono, 0)
    q = wigner.pre_calc_Wigner(q)
    q = convex.get_ref_paths_ARG_device(q, room)
    assert q["images"]["atten_all"].is_cuda and q["images"]["R_sI_r_all"].is_cuda
    q = D.vectorize_C_nm_s_ARG(D.init_source_directivities_ARG(q))
    assert q["C_nm_s_ARG_vec"].is_cuda and tuple(q["C_nm_s_ARG_vec"].shape) == (len(k), 1, q["images"]["R_sI_r_all"].shape[1])
    got = backend.run_DEISM_ARG(q)
    assert rel_l2(got, z["RTF"]) < RTF_TOL
    # geometry computed on the GPU equals the host formula bit for bit; lazy attributes match
    host = convex.get_ref_paths_ARG({"DEISM_method": "MIX", "posReceiver": z["params/posReceiver"],
                                     "ifRemoveDirectPath": int(z["params/ifRemoveDirectPath"]),
                                     "mixEarlyOrder": int(z["params/mixEarlyOrder"])}, room)
    assert np.array_equal(host["images"]["R_sI_r_all"].view(np.uint32),
                          q["images"]["R_sI_r_all"].cpu().numpy().view(np.uint32))
    assert np.array_equal(host["images"]["late_indices"], q["images"]["late_indices"])
    assert np.array_equal(room.orders, z["engine/orders"]) and n == len(room.orders)
