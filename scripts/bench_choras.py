"""Wall time of a CHORAS job (deism_b200.choras.deism_method) at the YAML's full settings: the job files
frozen in tests/golden/j_choras_{shoebox,convex}.npz (band-wise absorption -> frequency-dependent
impedance), 48 kHz / 44.1 kHz, reflection order 25 (shoebox) / 10 (convex).

    python scripts/bench_choras.py [shoebox|convex] [--repeat 3]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deism_b200 import choras  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", nargs="?", default="shoebox", choices=["shoebox", "convex"])
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    z = np.load(os.path.join(root, "tests", "golden", f"j_choras_{a.kind}.npz"))
    overrides = {"maxReflOrder": 10} if a.kind == "convex" else {}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "job.json")
        for i in range(a.repeat):
            open(path, "w").write(str(z["job_json"]))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rir = choras.deism_method(path, overrides=overrides)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            print(json.dumps({"job": a.kind, "pass": i, "wall_ms_incl_json_io": round(ms, 2), "rir_samples": int(len(rir)),
                              "rir_peak": float(np.max(np.abs(rir)))}))
