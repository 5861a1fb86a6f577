"""ARG source-directivity refit at the size of the C4 convex room with an order-N
sampled source: GPU (deism_arg_refit, CUDA events).  With cpu_images > 0 the NumPy restatement of
the reference's per-image pinv + product (oracle/, the same CPU-baseline role it has in bench.py)
is timed on that many images beside it and the two outputs are compared.

    python scripts/bench_refit.py [N] [n_dir] [cpu_images=6]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import Golden
from deism_b200 import directivities as D

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n_dir = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cpu_images = int(sys.argv[3]) if len(sys.argv) > 3 else 6
g = Golden("c4_full"); p = g.params()
raw = np.load(os.path.join(ROOT, "tests", "golden", "c4_full.npz"), allow_pickle=True)
R = np.ascontiguousarray(np.transpose(raw["engine/reflection_matrix"], (1, 2, 0)).astype(np.float32))  # [3,3,I]
k = np.asarray(p["waveNumbers"]); K, I = k.size, R.shape[2]
rng = np.random.default_rng(0)
i = np.arange(n_dir) + 0.5
pol, az = np.arccos(1 - 2 * i / n_dir), np.pi * (1 + 5 ** 0.5) * i
X = np.vstack((np.sin(pol) * np.cos(az), np.sin(pol) * np.sin(az), np.cos(pol)))
Psh = rng.standard_normal((K, n_dir)) + 1j * rng.standard_normal((K, n_dir))
params = {"waveNumbers": k, "sourceOrder": N, "radiusSource": 0.4}
dev = torch.device("cuda")
Rd = torch.from_numpy(np.ascontiguousarray(np.transpose(R, (2, 0, 1)))).to(dev)
Pd = torch.from_numpy(Psh).to(dev)
for _ in range(2):
    out = D.cal_C_nm_s_ARG_vec_device(Rd, Pd, X, params)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
reps = 5
e0.record()
for _ in range(reps):
    out = D.cal_C_nm_s_ARG_vec_device(Rd, Pd, X, params)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
J = (N + 1) ** 2
flops = 4.0 * I * J * n_dir * K
line = (f"refit I={I} K={K} N={N} (J={J}) n_dir={n_dir}: GPU {ms:.2f} ms ({flops / ms / 1e9:.2f} TFLOP/s of the "
        f"real x complex contraction, out {out.numel() * 8 / 1e9:.2f} GB c64)")
if cpu_images > 0:
    from oracle import oracle          # CPU baseline leg only

    t0 = time.perf_counter()
    want = oracle.cal_C_nm_s_ARG_vec(R[:, :, :cpu_images], Psh, X, k, N, 0.4)
    t_cpu = (time.perf_counter() - t0) / cpu_images
    got = out[:, :, :cpu_images].cpu().numpy()
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    line += (f"; CPU restatement {t_cpu * 1e3:.1f} ms/image -> {t_cpu * I:.1f} s for all images "
             f"({t_cpu * I / ms * 1e3:.0f}x); rel-L2 vs CPU on {cpu_images} images {err:.2e}")
print(line)
