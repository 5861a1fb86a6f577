"""Scratch timing of the LC kernel on BASELINE-like sizes (device-resident inputs)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import Golden
from deism_b200 import backend, images as im
from deism_b200 import workloads

def run(name, N, V, flatZ=True):
    g = Golden(name); p = g.params(); p["ifRemoveDirectPath"] = 0
    k = p["waveNumbers"]; K = len(k)
    rng = np.random.default_rng(0)
    if N == 0:
        cs = workloads.monopole_coeffs(k); cr = workloads.monopole_coeffs(k)
    else:
        cs = (rng.standard_normal((K, N+1, 2*N+1)) + 1j*rng.standard_normal((K, N+1, 2*N+1))).astype(np.complex64)
        cr = (rng.standard_normal((K, V+1, 2*V+1)) + 1j*rng.standard_normal((K, V+1, 2*V+1))).astype(np.complex64)
    p["n_all"], p["m_all"], p["C_nm_s_vec"] = workloads.vectorize(cs, N)
    p["v_all"], p["u_all"], p["C_vu_r_vec"] = workloads.vectorize(cr, V)
    p["DEISM_method"] = "LC"
    if not flatZ:
        Z = np.array(p["impedance"]); Z = Z * (1 + 0.1*np.linspace(0, 1, K))[None, :]; p["impedance"] = Z
    t0 = time.time(); imgs = im.merge_images(im.shoebox_images(p)); t_img = time.time() - t0
    p["images"] = imgs
    I = len(imgs["A"])
    dev = torch.device("cuda")
    # device-resident copies
    pd = dict(p)
    pd["images"] = {kk: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for kk, v in imgs.items()}
    for key, dt in (("waveNumbers", torch.float64), ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64), ("impedance", torch.complex128)):
        pd[key] = torch.from_numpy(np.ascontiguousarray(p[key])).to(dt).to(dev)
    for _ in range(2): backend.run_DEISM_device(pd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 5; e0.record()
    for _ in range(n): out = backend.run_DEISM_device(pd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t0 = time.time(); out_h = backend.run_DEISM(p); t_e2e = time.time() - t0
    J = (N+1)**2 + (V+1)**2
    print(f"{name} N={N} flatZ={flatZ}: I={I} K={K} terms={I*K:.3e} kernel {ms:.3f} ms -> {I*K/ms*1e3:.3e} terms/s, "
          f"contraction {4*J*I*K/ms*1e3/1e12:.2f} TFLOP/s (real-basis count); e2e {t_e2e*1e3:.1f} ms; images {t_img*1e3:.1f} ms", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:          # python scripts/quick_lc_perf.py c2_full_N5 3,3 6,6 10,10
        for nv in sys.argv[2:]:
            n, v = (int(t) for t in nv.split(","))
            run(sys.argv[1], n, v)
    else:
        run("c2_full_mono", 0, 0)
        run("c2_full_N5", 5, 5)
        run("c2_full_N5", 5, 5, flatZ=False)
        run("c5_full_N5", 5, 5)
        run("c5_full_N5", 5, 5, flatZ=False)
        run("c5_full_N5", 0, 0)
