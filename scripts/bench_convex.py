#!/usr/bin/env python
"""BASELINE config C4 (convex ARG room, RIR yaml, order 10, K=4322, monopole, MIX) on one GPU:
times the reflection-tree enumeration and the ARG accumulation separately and checks both against
the reference's frozen outputs.  Prints one JSON line (not the driver's bench contract)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deism_b200 import backend, convex, wigner, workloads  # noqa: E402


class _W:
    pass


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "c4_full.npz"))
    walls = []
    for i in range(int(z["walls/n"])):
        w = _W()
        w.normal, w.origin = z[f"walls/{i}/normal"], z[f"walls/{i}/origin"]
        w.basis, w.flat_corners = z[f"walls/{i}/basis"], z[f"walls/{i}/flat_corners"]
        w.reflection_matrix, w.impedence_bands = z[f"walls/{i}/reflection_matrix"], z[f"walls/{i}/impedance"]
        walls.append(w)
    order = int(z["params/maxReflOrder"])
    src, mic = z["params/posSource"].astype(np.float32), z["params/posReceiver"].astype(np.float32)
    room = convex.Room_deism(walls, [], [], 343.0, order)
    room.add_mic(mic)
    for _ in range(2):
        room.image_source_model(src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        n = room.image_source_model(src)
    torch.cuda.synchronize()
    t_tree = (time.perf_counter() - t0) / reps * 1e3
    ok_tree = bool(np.array_equal(room.sources.view(np.uint32), z["engine/sources"].view(np.uint32))
                   and np.array_equal(room.orders, z["engine/orders"])
                   and np.array_equal(room.gen_walls, z["engine/gen_walls"])
                   and np.array_equal(room.attenuations.view(np.uint32), z["engine/attenuations"].view(np.uint32)))
    k = z["params/waveNumbers"]
    p = {"DEISM_method": "MIX", "silentMode": 1, "waveNumbers": k, "posReceiver": z["params/posReceiver"],
         "ifRemoveDirectPath": int(z["params/ifRemoveDirectPath"]), "mixEarlyOrder": int(z["params/mixEarlyOrder"]),
         "sourceOrder": 0, "receiverOrder": 0}
    p = convex.get_ref_paths_ARG(p, room)
    mono = workloads.monopole_coeffs(k)
    p["C_nm_s_ARG"] = (mono[..., None] * np.ones((1, 1, 1, n), np.complex64)).astype(np.complex64)
    p["C_vu_r"] = mono
    p["n_all"], p["m_all"], _ = workloads.vectorize(mono, 0)
    p["C_nm_s_ARG_vec"] = np.ascontiguousarray(p["C_nm_s_ARG"][:, 0, :, :])
    p["v_all"], p["u_all"], p["C_vu_r_vec"] = workloads.vectorize(mono, 0)
    p = wigner.pre_calc_Wigner(p)
    for _ in range(3):
        out = backend.run_DEISM_ARG(p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = backend.run_DEISM_ARG(p)
    torch.cuda.synchronize()
    t_run = (time.perf_counter() - t0) / 10 * 1e3
    err = float(np.linalg.norm(out - z["RTF"]) / np.linalg.norm(z["RTF"]))

    # ---- the same job without leaving the GPU: tree -> geometry -> coefficients -> accumulation ----
    from deism_b200 import directivities as D

    def device_pipeline():
        room.image_source_model(src)
        q = {"DEISM_method": "MIX", "silentMode": 1, "waveNumbers": k, "posReceiver": z["params/posReceiver"],
             "ifRemoveDirectPath": int(z["params/ifRemoveDirectPath"]),
             "mixEarlyOrder": int(z["params/mixEarlyOrder"]), "sourceOrder": 0, "receiverOrder": 0,
             "sourceType": "monopole", "Wigner": p["Wigner"], "C_vu_r": mono, "v_all": p["v_all"],
             "u_all": p["u_all"], "C_vu_r_vec": p["C_vu_r_vec"]}
        q = convex.get_ref_paths_ARG_device(q, room)
        q = D.vectorize_C_nm_s_ARG(D.init_source_directivities_ARG(q))
        return backend.run_DEISM_ARG(q)

    for _ in range(3):
        out_dev = device_pipeline()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out_dev = device_pipeline()
    torch.cuda.synchronize()
    t_pipe = (time.perf_counter() - t0) / 10 * 1e3
    err_dev = float(np.linalg.norm(out_dev - z["RTF"]) / np.linalg.norm(z["RTF"]))
    print(json.dumps({"workload": "C4 convex ARG RIR yaml, MIX order 10, monopole", "images": int(n),
                      "frequencies": int(len(k)), "tree_nodes": 720247,
                      "tree_enumeration_ms_e2e": round(t_tree, 3), "tree_bit_exact": ok_tree,
                      "run_DEISM_ARG_ms_e2e": round(t_run, 3), "terms_per_s_e2e": n * len(k) / (t_run * 1e-3),
                      "rel_l2_vs_reference": err,
                      "device_pipeline_ms": round(t_pipe, 3), "device_pipeline_rel_l2_vs_reference": err_dev,
                      "device_pipeline_note": "tree enumeration + geometry + monopole coefficients + "
                                              "accumulation, arrays never leave the GPU; RTF copied back",
                      "reference_cpu": {"tree_s": float(z["meta/t_images_s"]), "run_cold_s": float(z["meta/t_run_cold_s"])}}))


if __name__ == "__main__":
    main()
