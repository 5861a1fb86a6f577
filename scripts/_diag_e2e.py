import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deism_b200 import backend, workloads, images as gimages
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
p, fx, desc = workloads.build(name)
dev = torch.device("cuda", 0)
im = gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
if p["DEISM_method"] in ("ORG", "LC"): im = gimages.merge_images(im)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t.numpy(), t
keep = []
ph = dict(p); ph["images"] = {}
for k, v in im.items():
    if isinstance(v, np.ndarray):
        a, t = pin(v); keep.append(t); ph["images"][k] = a
    else: ph["images"][k] = v
for k in ("waveNumbers", "impedance", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
    if k in p:
        a, t = pin(np.asarray(p[k])); keep.append(t); ph[k] = a
pn = dict(p); pn["images"] = dict(im)
for label, q in (("pinned", ph), ("pageable", pn), ("pinned", ph), ("pageable", pn)):
    for _ in range(3): backend.run_DEISM_device(q, dev).cpu()
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        t0 = time.perf_counter(); r = backend.run_DEISM_device(q, dev).cpu(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, label, "wall ms per call:", [round(x, 2) for x in ts])
