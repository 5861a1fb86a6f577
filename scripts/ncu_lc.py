"""One device-resident run_DEISM LC step on a BASELINE-sized workload (for ncu captures).

    python scripts/ncu_lc.py [c2_full_N5|c5_full_N5] [N] [reps]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden
from deism_b200 import backend, images as im, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "c2_full_N5"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = Golden(name); p = g.params(); p["ifRemoveDirectPath"] = 0
k = p["waveNumbers"]; K = len(k)
rng = np.random.default_rng(0)
cs = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
cr = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
p["n_all"], p["m_all"], p["C_nm_s_vec"] = workloads.vectorize(cs, N)
p["v_all"], p["u_all"], p["C_vu_r_vec"] = workloads.vectorize(cr, N)
p["DEISM_method"] = "LC"
imgs = im.merge_images(im.shoebox_images(p))
dev = torch.device("cuda")
pd = dict(p)
pd["images"] = {kk: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for kk, v in imgs.items()}
for key, dt in (("waveNumbers", torch.float64), ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64), ("impedance", torch.complex128)):
    pd[key] = torch.from_numpy(np.ascontiguousarray(p[key])).to(dt).to(dev)
for _ in range(reps):
    out = backend.run_DEISM_device(pd)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
