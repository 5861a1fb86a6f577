"""Wall time of the whole user-facing workflow, stage by stage, through deism_b200.core_deism.DEISM
(the reference quotes its runs this way: examples/deism_singleparam_example.py).

    python scripts/bench_class.py [c1|c2|c2n5|c4|c5|c5mono] [--repeat 3]

Every stage is bracketed by torch.cuda.synchronize(); the first pass is cold (CUDA context, module
load, scratch allocation), the later ones are what a parameter sweep over one object pays.
Reference (Numba, 8 threads, build container): C1 3.7 s images + 0.4 s run; C5 16.8 s
images + 86 s run; C4 0.84 s tree + 0.84 s run.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deism_b200 import directivities, wigner, workloads  # noqa: E402
from deism_b200.core_deism import DEISM  # noqa: E402

CASES = {
    # name: (mode, roomtype, method, order, materials, overrides, injected (N, V) or None)
    "c1": ("RTF", "shoebox", "MIX", 25, (), {}, None),
    "c2": ("RIR", "shoebox", "LC", 20, (1.0 / 3.0, None, "reverberationTime"), {}, None),
    "c2n5": ("RIR", "shoebox", "LC", 20, (1.0 / 3.0, None, "reverberationTime"), {}, (5, 5)),
    "c4": ("RIR", "convex", "MIX", 10, (), {}, None),
    "c5": ("RIR", "shoebox", "MIX", 40, (2.0 / 3.0, None, "reverberationTime"), {}, (5, 5)),
    "c5mono": ("RIR", "shoebox", "MIX", 40, (2.0 / 3.0, None, "reverberationTime"), {}, None),
}


_SYNTH = {}


class Clock:
    def __init__(self):
        self.t, self.rows = None, {}

    def start(self):
        torch.cuda.synchronize()
        self.t = time.perf_counter()

    def lap(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.rows[name] = (now - self.t) * 1e3
        self.t = now


def run(case, d=None, device_resident=False):
    mode, roomtype, method, order, mats, overrides, inject = CASES[case]
    ck = Clock()
    ck.start()
    if d is None:
        d = DEISM(mode, roomtype, silent=True, argv=[], device_resident=device_resident)
        d.params["DEISM_method"], d.params["maxReflOrder"] = method, order
        d.params.update(overrides)
        ck.lap("init (yaml)")
    d.update_wall_materials(*mats)
    ck.lap("update_wall_materials")
    d.update_freqs()
    ck.lap("update_freqs")
    if roomtype == "convex":
        d.update_source_receiver()
        ck.lap("update_source_receiver")

    def dirs():
        if inject is None:
            d.update_directivities()
        else:
            p = d.params
            key = (len(p["waveNumbers"]),) + tuple(inject)
            if key not in _SYNTH:                    # stand-in for measured directivities: drawn once
                _SYNTH[key] = workloads.synth_coeffs(len(p["waveNumbers"]), *inject)
            cs, cr = _SYNTH[key]
            p["sourceOrder"], p["receiverOrder"], p["C_nm_s"], p["C_vu_r"] = inject[0], inject[1], cs, cr
            if method in ("LC", "MIX"):
                d.params = directivities.vectorize_C_vu_r(directivities.vectorize_C_nm_s(p))
            if method in ("ORG", "MIX"):
                d.params = wigner.pre_calc_Wigner(d.params)
    dirs()
    ck.lap("update_directivities" + (" (synthetic coefficients, cached after the first pass)" if inject else ""))
    if roomtype == "shoebox":
        d.update_source_receiver()
        ck.lap("update_source_receiver")
    im = d.params["images"]
    n_img = int(im["R_sI_r_all"].shape[1] if roomtype == "convex" else
                len(im["A"]) if "A" in im else len(im["A_early"]) + len(im["A_late"]))
    d.run_DEISM()
    ck.lap("run_DEISM")
    out = d.get_results()
    ck.lap("get_results")
    K = len(d.params["waveNumbers"])
    return d, ck.rows, n_img, K, np.asarray(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case", nargs="?", default="c5", choices=sorted(CASES))
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--device-resident", action="store_true", help="DEISM(..., device_resident=True)")
    a = ap.parse_args()
    d = None
    for i in range(a.repeat):
        mode = CASES[a.case][0]
        if d is not None and mode == "RIR":
            # get_results() leaves the windowed spectrum in params["RTF"], nothing else to undo; the
            # material input is given again so that the fit runs as on a fresh object
            d.params["impedance"] = [18] * 6
            d.params["givenMaterials"] = ["impedance"]
        elif d is not None:
            d.params["impedance"] = [18] * 6
        d, rows, n_img, K, out = run(a.case, d, a.device_resident)
        total = sum(rows.values())
        print(json.dumps({"case": a.case, "device_resident": a.device_resident, "pass": i, "images": n_img, "freqs": K, "total_ms": round(total, 2),
                          "stages_ms": {k: round(v, 2) for k, v in rows.items()},
                          "terms_per_s_whole_workflow": n_img * K / (total * 1e-3)}))
