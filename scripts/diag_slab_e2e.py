import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deism_b200 import backend, workloads, distributed, images as gimages
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p, fx, desc = workloads.build("c5")
dev = torch.device("cuda", 0)
im = gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
ps, lo, hi = distributed.shard_frequencies(p, 0, world)
ps["images"] = im
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t.numpy(), t
keep = []
ph = dict(ps); ph["images"] = {}
for k, v in im.items():
    if isinstance(v, np.ndarray):
        a, t = pin(v); keep.append(t); ph["images"][k] = a
    else: ph["images"][k] = v
for k in ("waveNumbers", "impedance", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
    a, t = pin(np.asarray(ps[k])); keep.append(t); ph[k] = a
pd = dict(ps)
pd["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v) for k, v in im.items()}
for k, dt in (("waveNumbers", torch.float64), ("impedance", torch.complex128), ("C_nm_s", torch.complex64), ("C_vu_r", torch.complex64), ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64)):
    pd[k] = torch.from_numpy(np.ascontiguousarray(ps[k])).to(dt).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / n
for g in (True, False):
    backend.USE_GRAPHS = g; backend.clear_graphs()
    d = timeit(lambda: backend.run_DEISM_device(pd, dev))
    e = timeit(lambda: backend.run_DEISM_device(ph, dev).cpu())
    print(f"slab 1/{world} graphs={g}: device {d:.3f} ms, e2e pinned {e:.3f} ms, gap {e - d:.3f}")
# raw copies of the same arrays
arrs = [ph["images"][k] for k in ("A_late", "R_s_rI_all_late", "R_r_sI_all_late")] + [ph[k] for k in ("waveNumbers", "impedance", "C_nm_s_vec", "C_vu_r_vec")]
def copies():
    return [torch.from_numpy(a).to(dev, non_blocking=True) for a in arrs]
print("raw H2D of the LC inputs (%.2f MB): %.3f ms" % (sum(a.nbytes for a in arrs) / 1e6, timeit(copies)))
arrs2 = [ph[k] for k in ("C_nm_s", "C_vu_r")]
print("raw H2D of the ORG tables (%.2f MB): %.3f ms" % (sum(a.nbytes for a in arrs2) / 1e6, timeit(lambda: [torch.from_numpy(a).to(dev, non_blocking=True) for a in arrs2])))
out = backend.run_DEISM_device(pd, dev)
print("D2H of the result: %.3f ms" % timeit(lambda: out.cpu()))
