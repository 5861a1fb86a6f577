"""What one rank of an N-GPU C5 run executes (its frequency slab, all images), on one GPU: step time
and the profiled kernels, to see which part does not shrink with the slab.

    python scripts/slab_overhead.py [world=8]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deism_b200 import _lib, backend, distributed, images as gimages, workloads

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p, fixture, desc = workloads.build("c5")
dev = torch.device("cuda")
full_images = gimages.merge_images(gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)) \
    if p["DEISM_method"] in ("ORG", "LC") else gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
p["images"] = full_images
q, lo, hi = distributed.shard_frequencies(p, 0, world)
qd = dict(q)
qd["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev) if isinstance(v, np.ndarray) else v)
                for k, v in q["images"].items()}
for key, dt in (("waveNumbers", torch.float64), ("impedance", torch.complex128), ("C_nm_s", torch.complex64),
                ("C_vu_r", torch.complex64), ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64)):
    if key in q:
        qd[key] = torch.from_numpy(np.ascontiguousarray(q[key])).to(dt).to(dev)
lib = _lib.lib()
for _ in range(3):
    backend.run_DEISM_device(qd, dev)
torch.cuda.synchronize()
lib.deism_profile_enable(1); lib.deism_profile_reset()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n = 10
e0.record()
for _ in range(n):
    backend.run_DEISM_device(qd, dev)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
prof = {}
for name in ("lc_main", "org_consume", "org_btable"):
    t, c = C.c_double(0), C.c_int(0)
    lib.deism_profile_get(name.encode(), C.byref(t), C.byref(c))
    if c.value:
        prof[name] = round(t.value / c.value, 4)
lib.deism_profile_enable(0)
print(f"slab {lo}:{hi} of world {world}: {ms:.3f} ms/step back to back, kernels {prof}, rest {ms - sum(prof.values()):.3f} ms")
