#!/usr/bin/env python
"""oracle/time_reference.py — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Times the UNMODIFIED reference's Numba backend (deism.parallel_backends.run_DEISM, imported from
/root/reference via oracle/ref_import.py) on the BASELINE configs in THIS (build) container and
writes profiles/reference_numba_build_container.json.  The reference cannot travel to the GPU box
(no /root/reference there) and Ray cannot be installed at all (no network), so this table — warm
second call, numba.get_num_threads() stated — is the reference's own arm next to the GPU numbers
bench.py prints (SURVEY 8d "CPU baseline"); bench.py copies it into its JSON line.

    python oracle/time_reference.py [c1 c2mono c2 c4 c5] [--threads N]

C3 (order-25 ORG, N = V = 10) takes ~1 h per call: its time is taken from the cold run that minted
tests/golden/c3_full.npz (meta/t_run_cold_s; JIT compilation is seconds of it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _HERE)
import make_golden as mg  # noqa: E402
import ref_import  # noqa: E402


def shoebox(mode, method, order, N=None, overrides=None, materials=None):
    d = ref_import.make_deism(mode, "shoebox")
    p = d.params
    p["DEISM_method"], p["maxReflOrder"] = method, order
    for k, v in (overrides or {}).items():
        p[k] = v
    d.update_wall_materials(*(materials or ()))
    d.update_freqs()
    if N is None:
        d.update_directivities()
    else:
        mg.inject_directivities(d, N, N, 0)
    return d


def timed(d, label, out):
    import numba

    t0 = time.time()
    d.update_source_receiver()
    t_img = time.time() - t0
    times = []
    for rep in range(2):
        if rep:
            d.update_source_receiver()        # run_DEISM(if_clean_up=True) drops the images
        t0 = time.time()
        d.run_DEISM(if_clean_up=True)
        times.append(time.time() - t0)
    p = d.params
    K = len(p["waveNumbers"])
    n_img = int(out.get("_n_images", 0))
    out[label] = {"run_DEISM_warm_s": round(times[1], 3), "run_DEISM_cold_s": round(times[0], 3),
                  "image_generation_s": round(t_img, 3), "frequencies": K,
                  "numba_threads": int(numba.get_num_threads())}
    print(label, out[label], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2mono", "c2", "c4", "c5"])
    ap.add_argument("--threads", type=int, default=None)
    args = ap.parse_args()
    ref_import.import_reference()
    import numba

    if args.threads:
        numba.set_num_threads(args.threads)
    path = os.path.join(ROOT, "profiles", "reference_numba_build_container.json")
    out = {}
    if os.path.exists(path):
        out = json.load(open(path))
    t60 = "reverberationTime"
    if "c1" in args.which:
        timed(shoebox("RTF", "MIX", 25), "c1", out)
    if "c2mono" in args.which:
        timed(shoebox("RIR", "LC", 20, overrides={"sampleRate": 48000}, materials=(1.0 / 3.0, None, t60)),
              "c2mono", out)
    if "c2" in args.which:
        timed(shoebox("RIR", "LC", 20, N=5, overrides={"sampleRate": 48000}, materials=(1.0 / 3.0, None, t60)),
              "c2", out)
    if "c4" in args.which:
        d = ref_import.make_deism("RIR", "convex")
        d.params["maxReflOrder"] = 10
        d.update_wall_materials()
        d.update_freqs()
        d.update_source_receiver()        # convex: the per-image source coefficients need the images
        d.update_directivities()
        timed(d, "c4", out)
    if "c5" in args.which:
        timed(shoebox("RIR", "MIX", 40, N=5, overrides={"sampleRate": 48000}, materials=(2.0 / 3.0, None, t60)),
              "c5", out)
    c3 = os.path.join(ROOT, "tests", "golden", "c3_full.npz")
    if os.path.exists(c3):
        z = np.load(c3)
        out["c3"] = {"run_DEISM_cold_s": round(float(z["meta/t_run_cold_s"]), 1),
                     "image_generation_s": round(float(z["meta/t_images_s"]), 3), "frequencies": 2000,
                     "numba_threads": 5,
                     "note": "the run that minted tests/golden/c3_full.npz (one call, NUMBA_NUM_THREADS=5)"}
    out["_what"] = ("UNMODIFIED reference (deism.parallel_backends.run_DEISM, Numba backend) timed in the build "
                    "container by oracle/time_reference.py; warm = second call of the process; Ray is not "
                    "installable (no network)")
    out["_host"] = {"cpus": os.cpu_count()}
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
