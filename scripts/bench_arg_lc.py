"""Convex DEISM-ARG LC accumulation with per-image source coefficients at C4 size (1 577 images x
4 322 frequencies) and order-N directivities: the kernel streams C_s[K, J, I] complex64 once, so
its roofline is HBM bandwidth.

    python scripts/bench_arg_lc.py [N]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from deism_b200 import backend, workloads

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
raw = np.load(os.path.join(ROOT, "tests", "golden", "c4_full.npz"), allow_pickle=True)
k = raw["params/waveNumbers"]; K = k.size
R = raw["images/R_sI_r_all"]; att = raw["images/atten_all"]; I = R.shape[1]
J = (N + 1) ** 2
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(0)
Cs = torch.view_as_complex(torch.randn((K, J, I, 2), device=dev, generator=g, dtype=torch.float32))
rng = np.random.default_rng(0)
cr = (rng.standard_normal((K, N + 1, 2 * N + 1)) + 1j * rng.standard_normal((K, N + 1, 2 * N + 1))).astype(np.complex64)
v_all, u_all, crv = workloads.vectorize(cr, N)
p = {"DEISM_method": "LC", "waveNumbers": torch.from_numpy(k).to(dev), "n_all": v_all, "m_all": u_all,
     "v_all": v_all, "u_all": u_all, "C_nm_s_ARG_vec": Cs, "C_vu_r_vec": torch.from_numpy(crv).to(dev),
     "sourceOrder": N, "receiverOrder": N,
     "images": {"R_sI_r_all": torch.from_numpy(R).to(dev), "atten_all": torch.from_numpy(att).to(dev)}}
for _ in range(3):
    out = backend.run_DEISM_ARG_device(p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
reps = 10
e0.record()
for _ in range(reps):
    out = backend.run_DEISM_ARG_device(p)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
gb = Cs.numel() * 8 / 1e9
print(f"ARG LC I={I} K={K} N={N} (J={J}): {ms:.3f} ms/step, C_s stream {gb:.2f} GB -> {gb / ms * 1e3:.0f} GB/s, "
      f"{I * K / ms * 1e3:.3e} terms/s")
