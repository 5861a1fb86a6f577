"""GPU tests of the boundary's robustness contract (include/deism_b200.h "Threading", dtype handling,
size limits): calls on different streams and threads, complex128 coefficient tables, frequency counts
beyond one grid dimension, the NumPy-facing handles of device-resident arrays."""
import threading

import numpy as np
import pytest

from conftest import Golden, rel_l2

pytestmark = pytest.mark.gpu
RTF_TOL = 1e-9     # north star: 1e-8


def _lc_params(rng, I, K, N, V):
    from oracle import oracle

    cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
    n_all, m_all, csv = oracle.vectorize(cs, N)
    v_all, u_all, crv = oracle.vectorize(cr, V)
    R = lambda: np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I),
                          rng.uniform(0.5, 300.0, I)], axis=1).astype(np.float32)
    atten = (rng.standard_normal((I, K)) + 1j * rng.standard_normal((I, K))).astype(np.complex64)
    return {"DEISM_method": "LC", "waveNumbers": np.sort(rng.uniform(0.05, 400.0, K)),
            "n_all": n_all, "m_all": m_all, "v_all": v_all, "u_all": u_all,
            "C_nm_s_vec": csv, "C_vu_r_vec": crv, "silentMode": 1, "angDepFlag": 1, "numParaImages": 100,
            "images": {"A": np.zeros((I, 6), np.int32), "R_sI_r_all": R(), "R_s_rI_all": R(),
                       "R_r_sI_all": R(), "atten_all": atten}}


def _org_params(rng, I, K, N, V, dtype=np.complex64):
    from deism_b200 import wigner

    cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(dtype)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(dtype)
    A = rng.integers(-3, 4, size=(I, 6)).astype(np.int32)
    A[:, 3:] = rng.integers(0, 2, size=(I, 3))
    R = np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I),
                  rng.uniform(0.5, 60.0, I)], axis=1).astype(np.float32)
    atten = (rng.standard_normal((I, K)) + 1j * rng.standard_normal((I, K))).astype(np.complex64)
    return {"DEISM_method": "ORG", "waveNumbers": np.sort(rng.uniform(0.5, 100.0, K)),
            "sourceOrder": N, "receiverOrder": V, "C_nm_s": cs, "C_vu_r": cr,
            "Wigner": wigner.wigner_tables(N, V), "silentMode": 1, "angDepFlag": 1, "numParaImages": 50,
            "images": {"A": A, "R_sI_r_all": R, "R_s_rI_all": R, "R_r_sI_all": R, "atten_all": atten}}


def test_calls_on_different_streams_with_different_bases():
    """Two streams, different (n, m) lists, different sizes, no host synchronisation in between:
    the library orders the calls itself (one call in flight per device), so neither overwrites the
    constant tables or the scratch the other's kernels are still reading."""
    import torch

    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(3)
    jobs = [_lc_params(rng, 3000, 700, 5, 5), _lc_params(rng, 500, 90, 2, 7), _org_params(rng, 900, 120, 4, 3),
            _lc_params(rng, 2500, 300, 0, 0), _org_params(rng, 300, 64, 6, 2), _lc_params(rng, 1100, 260, 7, 1)]
    want = [oracle.run_DEISM(p) for p in jobs]
    dev = torch.device("cuda", torch.cuda.current_device())
    # inputs on the device first, so that the calls are nothing but back-to-back launches
    dj = []
    for p in jobs:
        q = dict(p)
        q["waveNumbers"] = backend.to_device(p["waveNumbers"], torch.float64, dev)
        for key in ("C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
            if key in p:
                q[key] = backend.to_device(p[key], torch.complex64, dev)
        q["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v)
                       for k, v in p["images"].items()}
        dj.append(q)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev) for _ in range(3)]
    for rep in range(3):
        outs = []
        for i, q in enumerate(dj):
            with torch.cuda.stream(streams[(i + rep) % 3]):
                outs.append(backend.run_DEISM_device(q, dev))
        torch.cuda.synchronize()
        for i, (o, w) in enumerate(zip(outs, want)):
            assert rel_l2(o.cpu().numpy(), w) < RTF_TOL, (rep, i)


def test_calls_from_several_host_threads():
    """The same from four host threads at once, each with its own stream and its own job."""
    import torch

    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(5)
    jobs = [_lc_params(rng, 800, 150, 3, 3), _org_params(rng, 400, 80, 3, 4), _lc_params(rng, 1500, 200, 5, 2),
            _lc_params(rng, 600, 333, 0, 0)]
    want = [oracle.run_DEISM(p) for p in jobs]
    got, errs = [None] * len(jobs), []

    def work(i):
        try:
            torch.cuda.set_device(0)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3):
                    got[i] = backend.run_DEISM(jobs[i])
        except Exception as exc:      # pragma: no cover
            errs.append((i, exc))

    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i, w in enumerate(want):
        assert rel_l2(got[i], w) < RTF_TOL, i


def test_org_complex128_coefficient_tables_keep_their_precision():
    """The reference's ORG path up-casts whatever coefficient dtype it is given (pb:556-557): a
    caller's own complex128 tables are used at full precision, not rounded to complex64 first."""
    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(17)
    p = _org_params(rng, 150, 40, 5, 4, dtype=np.complex128)
    want = oracle.run_DEISM(p)
    got = backend.run_DEISM(p)
    assert rel_l2(got, want) < RTF_TOL
    # rounding the tables to complex64 first would be visible at the 1e-8 level
    q = dict(p)
    q["C_nm_s"], q["C_vu_r"] = p["C_nm_s"].astype(np.complex64), p["C_vu_r"].astype(np.complex64)
    assert rel_l2(oracle.run_DEISM(q), want) > 1e-9
    # complex64-representable complex128 input takes the complex64 entry point: same numbers
    r = dict(q)
    r["C_nm_s"], r["C_vu_r"] = q["C_nm_s"].astype(np.complex128), q["C_vu_r"].astype(np.complex128)
    assert rel_l2(backend.run_DEISM(r), backend.run_DEISM(q)) == 0.0


def test_lc_accepts_complex128_vectors():
    """LC operands travel as complex64 (what the reference's own pipeline stores, cd:1184); other
    dtypes are converted, not refused."""
    from deism_b200 import backend
    from oracle import oracle

    rng = np.random.default_rng(19)
    p = _lc_params(rng, 200, 50, 3, 2)
    want = oracle.run_DEISM(p)
    q = dict(p)
    q["C_nm_s_vec"], q["C_vu_r_vec"] = p["C_nm_s_vec"].astype(np.complex128), p["C_vu_r_vec"].astype(np.complex128)
    assert rel_l2(backend.run_DEISM(q), want) < RTF_TOL


def test_convex_paths_beyond_65535_frequencies():
    """ARG LC / ORG and the convex attenuation export at K > 65535 (one grid dimension)."""
    from deism_b200 import backend, wigner
    from oracle import oracle

    rng = np.random.default_rng(23)
    I, K, N, V = 5, 70001, 1, 1
    Js, Jr = (N + 1) ** 2, (V + 1) ** 2
    cs = (rng.standard_normal((K, N + 1, 2 * N + 1, I)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1, I))).astype(np.complex64)
    cr = (rng.standard_normal((K, V + 1, 2 * V + 1)) + 1j * rng.standard_normal((K, V + 1, 2 * V + 1))).astype(np.complex64)
    from deism_b200 import directivities as D

    n_all, m_all, csv = D._flatten(cs, N)          # [K, J, I], image index innermost (cd:1193-1217)
    v_all, u_all, crv = oracle.vectorize(cr, V)
    R = np.stack([rng.uniform(-np.pi, np.pi, I), rng.uniform(0.05, 3.0, I), rng.uniform(0.5, 30.0, I)]).astype(np.float32)
    atten = rng.uniform(0.2, 1.0, (K, I)).astype(np.complex64)
    p = {"DEISM_method": "MIX", "waveNumbers": np.linspace(0.5, 300.0, K), "sourceOrder": N, "receiverOrder": V,
         "C_nm_s_ARG": cs, "C_vu_r": cr, "C_nm_s_ARG_vec": csv, "C_vu_r_vec": crv, "n_all": n_all, "m_all": m_all,
         "v_all": v_all, "u_all": u_all, "Wigner": wigner.wigner_tables(N, V), "silentMode": 1,
         "images": {"R_sI_r_all": R, "atten_all": atten, "early_indices": np.array([0, 1]),
                    "late_indices": np.array([2, 3, 4])}}
    want = oracle.run_DEISM_ARG(p)
    got = backend.run_DEISM_ARG(p)
    assert rel_l2(got, want) < RTF_TOL


def test_convex_attenuation_stays_on_the_device_until_read():
    """get_ref_paths_ARG on this package's engine hands run_DEISM_ARG a NumPy-facing handle of the
    device-resident [K, I] attenuation; reading it gives the reference's complex64 array."""
    from deism_b200 import backend, convex

    z = Golden("x1_convex_rir_mix_o4_mono")
    walls = []

    class W:
        pass

    g = z.z
    for i in range(int(g["walls/n"])):
        w = W()
        w.normal, w.origin = g[f"walls/{i}/normal"], g[f"walls/{i}/origin"]
        w.basis, w.flat_corners = g[f"walls/{i}/basis"], g[f"walls/{i}/flat_corners"]
        w.reflection_matrix, w.impedence_bands = g[f"walls/{i}/reflection_matrix"], g[f"walls/{i}/impedance"]
        walls.append(w)
    p = z.params()
    room = convex.Room_deism(walls, [], [], 343.0, int(p["maxReflOrder"]))
    room.add_mic(np.asarray(p["posReceiver"], np.float32))
    room.image_source_model(np.asarray(p["posSource"], np.float32))
    q = convex.get_ref_paths_ARG({k: v for k, v in p.items() if k != "images"}, room)
    att = q["images"]["atten_all"]
    assert isinstance(att, backend.DeviceArray) and att._host is None
    assert att.shape == p["images"]["atten_all"].shape and att.dtype == np.complex64
    got = backend.run_DEISM_ARG(q)
    assert att._host is None                                   # the run did not need the host copy
    assert rel_l2(got, z.rtf) < RTF_TOL
    assert np.array_equal(np.ascontiguousarray(np.asarray(att)).view(np.uint32),
                          np.ascontiguousarray(p["images"]["atten_all"]).view(np.uint32))


def test_monopole_arg_coefficients_are_a_broadcast_view():
    """init_source_directivities_ARG (monopole): [K, 1, 1, I] with the reference's values and no
    K x I copies of one column; the flattened [K, 1, I] keeps the broadcast and the backend uploads
    the column once."""
    from deism_b200 import backend, directivities as D

    K, I = 300, 1000
    k = np.linspace(0.1, 50.0, K)
    p = {"sourceType": "monopole", "waveNumbers": k, "sourceOrder": 3,
         "reflection_matrix": np.zeros((3, 3, I), np.float32)}
    p = D.vectorize_C_nm_s_ARG(D.init_source_directivities_ARG(p))
    want = (-1j * k / np.sqrt(4 * np.pi)).astype(np.complex64)
    assert p["C_nm_s_ARG"].shape == (K, 1, 1, I) and p["C_nm_s_ARG"].strides[3] == 0
    assert p["C_nm_s_ARG_vec"].shape == (K, 1, I) and p["C_nm_s_ARG_vec"].strides[2] == 0
    assert np.array_equal(p["C_nm_s_ARG"][:, 0, 0, 7], want) and np.array_equal(p["C_nm_s_ARG_vec"][:, 0, -1], want)
    import torch

    t = backend._c64_on_device(p["C_nm_s_ARG_vec"], torch.device("cuda", torch.cuda.current_device()))
    assert tuple(t.shape) == (K, 1, I) and t.is_contiguous()
    assert np.array_equal(t.cpu().numpy(), np.broadcast_to(want[:, None, None], (K, 1, I)))


def _device_params(p):
    """The same call with every bulk input resident on the device (what the class keeps with
    device_resident=True)."""
    import torch

    from deism_b200 import backend

    dev = torch.device("cuda", torch.cuda.current_device())
    q = dict(p)
    q["waveNumbers"] = backend.to_device(p["waveNumbers"], torch.float64, dev)
    for key in ("C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
        if key in p:
            q[key] = backend.to_device(p[key], torch.complex64, dev)
    q["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v)
                   for k, v in p["images"].items()}
    return q, dev


def test_cuda_graph_replay_of_repeated_calls():
    """Calls that come again with the same device buffers are captured on the second call and
    replayed afterwards: same numbers as the eager call, new buffer CONTENTS are picked up, the
    kernels of a replay are counted, and a graph is dropped when another workload changes the
    library's constant tables / scratch (deism_state_generation)."""
    import torch

    from deism_b200 import _lib, backend
    from oracle import oracle

    assert backend.USE_GRAPHS
    backend.clear_graphs()
    rng = np.random.default_rng(29)
    p1, p2 = _lc_params(rng, 700, 130, 4, 3), _org_params(rng, 260, 70, 3, 3)
    want1, want2 = oracle.run_DEISM(p1), oracle.run_DEISM(p2)
    q1, dev = _device_params(p1)
    q2, _ = _device_params(p2)
    outs = [backend.run_DEISM_device(q1, dev).cpu().numpy() for _ in range(4)]       # eager, capture, replay x2
    assert len(backend._graphs) == 1
    for o in outs:
        assert rel_l2(o, want1) < RTF_TOL
    assert np.array_equal(outs[2], outs[3])
    n0 = _lib.launch_count()
    backend.run_DEISM_device(q1, dev)
    assert _lib.launch_count() - n0 >= 4                    # a replay reports the kernels it launched
    # new contents in the same buffers
    q1["C_nm_s_vec"].mul_(2.0)
    assert rel_l2(backend.run_DEISM_device(q1, dev).cpu().numpy(), 2.0 * want1) < RTF_TOL
    q1["C_nm_s_vec"].mul_(0.5)
    # another workload moves the library state: the first graph must not be replayed against it
    for _ in range(3):
        assert rel_l2(backend.run_DEISM_device(q2, dev).cpu().numpy(), want2) < RTF_TOL
    for _ in range(3):
        assert rel_l2(backend.run_DEISM_device(q1, dev).cpu().numpy(), want1) < RTF_TOL
    # pinned host buffers are captured too (the copies are part of the graph); pageable ones never
    ph = dict(p1)
    keep = []
    for key in ("waveNumbers", "C_nm_s_vec", "C_vu_r_vec"):
        t = torch.from_numpy(np.ascontiguousarray(p1[key])).pin_memory()
        keep.append(t)
        ph[key] = t.numpy()
    ph["images"] = {}
    for k, v in p1["images"].items():
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        keep.append(t)
        ph["images"][k] = t.numpy()
    n_graphs = len(backend._graphs)
    for _ in range(4):
        assert rel_l2(backend.run_DEISM(ph), want1) < RTF_TOL
    assert len(backend._graphs) == n_graphs + 1
    ph["waveNumbers"][:] = ph["waveNumbers"] * 1.0          # same values, written in place: still valid
    assert rel_l2(backend.run_DEISM(ph), want1) < RTF_TOL
    for _ in range(3):
        assert rel_l2(backend.run_DEISM(p1), want1) < RTF_TOL                             # pageable: eager
    assert len(backend._graphs) == n_graphs + 1
    backend.clear_graphs()
