#!/usr/bin/env python
"""bench.py — image x frequency terms/s of run_DEISM on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c2mono|c3|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: oracle port of the Numba backend

A "step" is one pass of the hot path (`run_DEISM`) over the workload: all images x all
frequencies of the named BASELINE config (default C5: shoebox RIR DEISM-MIX, order 40,
88 641 images x 16 000 frequencies, N=V=5).  `value` = I*K / t with inputs resident in HBM;
`e2e` = the same through deism_b200.backend.run_DEISM(params) with HOST (pinned NumPy)
buffers, host->device copies and the result read-back inside the timed region.

Multi-GPU: every rank owns a contiguous frequency slab and all images; the slabs
are combined with ONE all_gather per step (inside the timed region).  Total work is fixed as N
grows -> "scaling": "strong".

One JSON line is printed by rank 0.  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------------------
# algorithmic work per term (SURVEY.md §8d conventions:
# complex MAC = 8 flops, complex mul = 6, complex x real = 2).  The directivity contractions are
# counted in the real-harmonic basis the kernels use (Y_n^-m = (-1)^m conj Y_n^m makes them
# real x complex: 4 flops per harmonic and term instead of 8), i.e. the cheapest formulation known
# to us, not the reference's complex x complex loop.
# --------------------------------------------------------------------------------------
def wall_factor_flops(mean_order, real_impedance, same_opposite_walls=True):
    """Flops the per-term attenuation chain EXECUTES (common.cuh shoebox_atten_n), per term: per wall
    pair one wall factor and one power by squaring of the summed exponent (two of each when the two
    walls of the pair differ), then the product of the three pairs and the complex64 rounding.
      real chain   : factor = t, t+1, reciprocal seed + 2 Newton steps (8), 1 - 2r (2)      = 12
                     power  = ~1.5 log2(e) real multiplies (squarings + set bits)
      complex chain: factor = zr, zi, +-1 (4), |den|^2 (3), reciprocal (9), numerator (8)    = 24
                     power  = ~1.5 log2(e) complex multiplies of 6 flops
    e = mean exponent per pair = mean_order / 3."""
    import math

    e = max(2.0, mean_order / 3.0)
    mults = 1.5 * math.log2(e)
    chains = 3 * (1 if same_opposite_walls else 2)
    if real_impedance:
        return chains * (12 + mults) + 4 + 2
    return chains * (24 + 6 * mults) + 12 + 2


def lc_flops_per_term(Js, Jr, fused_per_term, mean_order, real_impedance=False, same_opposite_walls=True):
    if Js == 1 and Jr == 1:
        # monopole kernel: no contraction; per term one step of the rotation recurrence (complex
        # mul 6 + first-order grid correction 5), the accumulate (2) and 1/16 of a sincos
        f = 6 + 5 + 2 + 3
    else:
        f = 4 * (Js + Jr) + 52
    if fused_per_term:
        f += wall_factor_flops(mean_order, real_impedance, same_opposite_walls)
    return f


def org_consume_flops_per_term(N, V):
    L = N + V + 1
    return 4 * L * L + 14 * L + 40      # contraction + h_l recurrence/accumulate + sincos


def clocks_sampler(stop, out, gpu_index):
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True,
                               timeout=5)
            parts = [x.strip() for x in r.stdout.strip().split(",")]
            if len(parts) >= 7:
                out.append(parts)
        except Exception:
            pass
        stop.wait(0.1)


def summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = set()
    for s in samples:
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                              "sw_power_cap"), s[3:7]):
            if val.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]),
            "power_w_max": max(float(s[2]) for s in samples), "samples": len(samples),
            "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Numba backend on a bounded sample
# --------------------------------------------------------------------------------------
def cpu_sample_params(p, target_terms):
    """First images of the workload (all early + as many late as fit the term budget), all
    frequencies, attenuation materialised as the reference's default storage does."""
    from oracle import oracle

    K = len(p["waveNumbers"])
    q = dict(p)
    # a shallower lattice gives the same per-image work; the cut keeps CPU time bounded
    order = int(p["maxReflOrder"])
    while order > 2:
        n = 1 + 2 * order * (2 * order * order + 3 * order + 4) // 3
        if n * K <= target_terms:
            break
        order -= 1
    q["maxReflOrder"] = order
    t0 = time.time()
    images = oracle.shoebox_images(q)
    t_img = time.time() - t0
    if q["DEISM_method"] in ("ORG", "LC"):
        images = oracle.merge_images(images)
    q["images"] = images
    n_img = len(images["A"]) if "A" in images else len(images["A_early"]) + len(images["A_late"])
    return q, n_img, order, t_img


def run_cpu(p, steps, warmup, target_terms):
    from oracle import oracle

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread it is allowed
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        oracle.set_num_threads(os.cpu_count() or 1)
    q, n_img, order, t_img = cpu_sample_params(p, target_terms)
    K = len(q["waveNumbers"])
    for _ in range(max(0, min(warmup, 1))):
        oracle.run_DEISM(q)
    t0 = time.time()
    for _ in range(steps):
        oracle.run_DEISM(q)
    dt = (time.time() - t0) / steps
    return {"value": n_img * K / dt, "unit": "terms/s", "cores": oracle.num_threads(),
            "kind": "port",
            "sample": f"same workload cut to reflection order {order}: {n_img} images x {K} "
                      f"freqs = {n_img * K:.3e} terms, {dt:.2f} s/step; CPU image generation "
                      f"incl. attenuation materialisation {t_img:.2f} s (not in value)",
            "ms_per_step": dt * 1e3, "terms": n_img * K}


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c2mono", "c3", "c5"])
    ap.add_argument("--general-impedance", action="store_true",
                    help="perturb Z_S across frequency so that the per-term beta path runs")
    ap.add_argument("--cpu-terms", type=float, default=None,
                    help="term budget of the CPU sample (default: ~15 s of CPU work)")
    ap.add_argument("--sharding", default="frequency", choices=["frequency", "image"],
                    help="N>1: contiguous frequency slabs + all_gather (default) or, for the merged "
                         "ORG/LC image lists of few-frequency runs, image slabs + all_reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 0)

    from deism_b200 import workloads

    p, fixture, desc = workloads.build(args.workload)
    if args.general_impedance:
        K0 = len(p["waveNumbers"])
        p["impedance"] = np.ascontiguousarray(
            p["impedance"] * (1.0 + 0.1 * np.linspace(0.0, 1.0, K0))[None, :])
    method, N, V = p["DEISM_method"], p["sourceOrder"], p["receiverOrder"]
    K = len(p["waveNumbers"])
    cpu_rate = {"c5": 1.65e7, "c2": 1.7e7, "c2mono": 5e7, "c1": 5e7, "c3": 1.2e4}[args.workload]
    cpu_terms = args.cpu_terms or cpu_rate * 15.0

    config = {"workload": desc, "method": method, "source_order": N, "receiver_order": V,
              "frequencies": K, "sharding": (f"freq-slab x{world}" if args.sharding == "frequency" else f"image-slab x{world}")
              if world > 1 else "single GPU",
              "l2": "L2 flushed (256 MiB write) between timed steps",
              "impedance": "general (per-term beta)" if args.general_impedance
              else "as shipped (frequency-flat: beta per image)"}

    # ---------------- reference arm: CPU, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        res = run_cpu(p, max(1, steps), warmup, cpu_terms)
        line = {"impl": "reference", "metric": "image x frequency terms/s (run_DEISM)",
                "value": res["value"], "unit": "terms/s", "n_gpus": 0, "steps": max(1, steps),
                "warmup": min(warmup, 1), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "terms/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "note": "oracle port (C/OpenMP restatement of the Numba backend, pinned to the "
                        "reference's outputs); the Python/Numba reference and Ray cannot travel "
                        "to the GPU box"}
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist

    from deism_b200 import _lib, backend, distributed, images as gimages

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # image enumeration on the GPU (untimed here; timed separately below)
    full_images = gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    if method in ("ORG", "LC"):
        full_images = gimages.merge_images(full_images)
    I = workloads.n_images(full_images)
    terms = I * K

    # frequency slab (or image slab) of this rank
    by_image = args.sharding == "image" and world > 1
    if by_image:
        if method == "MIX":
            raise SystemExit("--sharding image needs a merged image list (ORG or LC workload)")
        p_all = dict(p)
        p_all["images"] = full_images
        ps, ilo, ihi = distributed.shard_images(p_all, rank, world)
        full_images = ps["images"]
        lo, hi = 0, K
    else:
        ps, lo, hi = distributed.shard_frequencies(p, rank, world)
        ps["images"] = full_images
    Ks = hi - lo
    I_rank = workloads.n_images(full_images)      # this rank's images (all of them unless --sharding image)

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    # host (pinned) copy for the e2e leg, device-resident copy for the kernel leg
    keep = []
    p_host = dict(ps)
    p_host["images"] = {}
    for key, val in full_images.items():
        if isinstance(val, np.ndarray):
            a, t = pin(val)
            keep.append(t)
            p_host["images"][key] = a
        else:
            p_host["images"][key] = val
    h2d_bytes = 0
    for key in ("waveNumbers", "impedance", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
        if key in ps:
            a, t = pin(np.asarray(ps[key]))
            keep.append(t)
            p_host[key] = a
    used_img = ["A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all"] if method != "MIX" else \
        ["A_early", "R_sI_r_all_early", "A_late", "R_sI_r_all_late", "R_s_rI_all_late",
         "R_r_sI_all_late"]
    if method == "ORG":
        used_img = ["A", "R_sI_r_all"]
    for key in used_img:
        h2d_bytes += p_host["images"][key].nbytes
    for key in ("waveNumbers", "impedance"):
        h2d_bytes += p_host[key].nbytes
    for key in (("C_nm_s_vec", "C_vu_r_vec") if method in ("LC", "MIX") else ()):
        h2d_bytes += p_host[key].nbytes
    for key in (("C_nm_s", "C_vu_r") if method in ("ORG", "MIX") else ()):
        h2d_bytes += p_host[key].nbytes
    d2h_bytes = Ks * 16

    p_dev = dict(ps)
    p_dev["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev)
                           if isinstance(v, np.ndarray) else v) for k, v in full_images.items()}
    for key, dt in (("waveNumbers", torch.float64), ("impedance", torch.complex128),
                    ("C_nm_s", torch.complex64), ("C_vu_r", torch.complex64),
                    ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64)):
        if key in ps:
            p_dev[key] = torch.from_numpy(np.ascontiguousarray(ps[key])).to(dt).to(dev)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    def combine(slab):
        """The one collective of the path: frequency slabs -> full RTF on every rank (all_gather),
        or the sum of the per-rank partial RTFs over image slabs (all_reduce)."""
        if by_image:
            return distributed.reduce_image_slabs(slab)
        return distributed.gather_frequency_slabs(slab, K)

    def step_device():
        return combine(backend.run_DEISM_device(p_dev, dev))

    def step_e2e():
        out = backend.run_DEISM_device(p_host, dev)   # H2D of every input inside
        res = combine(out)
        return res.cpu() if rank == 0 or world == 1 else res[:1].cpu()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n steps, each bracketed by CUDA events on the launching stream; L2 flushed between
        steps outside the brackets.  Returns (sum of device ms, wall ms incl. flushes)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(n)]
        barrier()
        w0 = time.perf_counter()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        barrier()
        wall = (time.perf_counter() - w0) * 1e3
        return sum(a.elapsed_time(b) for a, b in evs), wall

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also allocates the library's scratch) ----
    result = None
    for _ in range(max(warmup, 1)):
        result = step_device()
    torch.cuda.synchronize()

    # ---- parity of this very run against the reference's frozen RTF ----
    parity, parity_worst = None, None
    if "RTF" in fixture.files and not args.general_impedance and fixture["RTF"].shape[0] == K:
        ref = fixture["RTF"]
        got = result.cpu().numpy()
        parity = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        parity_worst = float(np.max(np.abs(got - ref) / np.abs(ref)))

    # ---- timed region: device-resident ----
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    lib = _lib.lib()
    lib.deism_profile_enable(1)
    lib.deism_profile_reset()
    launches0 = _lib.launch_count()
    dev_ms, wall_ms = timed(step_device, steps)
    launches = _lib.launch_count() - launches0
    prof = {}
    for name in ("lc_main", "org_consume", "org_btable", "org_direct"):
        ms, n = C.c_double(0), C.c_int(0)
        lib.deism_profile_get(name.encode(), C.byref(ms), C.byref(n))
        if n.value:
            prof[name] = (ms.value, n.value)
    lib.deism_profile_enable(0)
    kernel_ms = {name: round(ms / n, 4) for name, (ms, n) in prof.items()}
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / steps

    # ---- timed region: end to end from host buffers ----
    for _ in range(2):
        step_e2e()
    e2e_ms, _ = timed(step_e2e, steps)
    e2e_ms = max_over_ranks(e2e_ms) / steps
    stop.set()

    # ---- extras (rank 0, N=1): image enumeration time, FP64 peak, other workloads ----
    extras = {}
    peak_tf, peak_kind = None, None
    if rank == 0:
        best = 0.0
        for kind, nm in ((0, "dfma"), (1, "dmma_m8n8k4")):
            tf, ms = C.c_double(0), C.c_double(0)
            if lib.deism_fp64_peak(kind, 20000, C.byref(tf), C.byref(ms)) == 0:
                extras[f"fp64_peak_{nm}_tflops"] = round(tf.value, 2)
                if tf.value > best:
                    best, peak_kind = tf.value, nm
        peak_tf = best or None
    if rank == 0 and world == 1 and not args.no_extras:
        # third witness of the FP64 denominator: cuBLAS DGEMM through torch, 6144^3
        try:
            n = 6144
            A = torch.randn((n, n), dtype=torch.float64, device=dev)
            B = torch.randn((n, n), dtype=torch.float64, device=dev)
            torch.matmul(A, B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                torch.matmul(A, B)
            e1.record()
            torch.cuda.synchronize()
            extras["fp64_cublas_dgemm_tflops"] = round(3 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12, 2)
            del A, B
        except Exception as exc:       # pragma: no cover
            extras["fp64_cublas_dgemm_tflops"] = f"unavailable: {exc}"
        t0 = time.perf_counter()
        for _ in range(3):
            gimages.shoebox_images_device(p, dev)
        torch.cuda.synchronize()
        extras["image_enumeration_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)

    # ---- roofline of the dominant kernel ----
    roofline = None
    if rank == 0 and prof:
        dom = max(prof, key=lambda n: prof[n][0])
        dom_kernel = dom + "_kernel"
        ms_k, n_k = prof[dom]
        per_launch_ms = ms_k / n_k
        if dom == "lc_main":
            n_late = len(full_images["A_late"]) if method == "MIX" else I_rank
            Js, Jr = (N + 1) ** 2, (V + 1) ** 2
            Zh = np.asarray(p["impedance"])
            z_real = bool(np.all(np.abs(Zh.imag) <= 4e-16 * np.abs(Zh.real)))       # common.cuh z_is_real
            z_same = bool(np.array_equal(Zh[0::2], Zh[1::2]))
            fpt = round(lc_flops_per_term(Js, Jr, args.general_impedance, 0.75 * p["maxReflOrder"], z_real,
                                          z_same), 1)
            units = n_late * Ks
            if Js == 1 and Jr == 1:
                dom_kernel = "lc_mono_kernel"      # what deism_lc_shoebox launches for monopoles
                if extras.get("fp64_peak_dfma_tflops"):
                    # no DMMA in it: its roofline is the FP64 FMA pipe
                    peak_tf, peak_kind = extras["fp64_peak_dfma_tflops"], "dfma"
        elif dom == "org_consume":
            fpt = org_consume_flops_per_term(N, V)
            units = I_rank * Ks
        else:
            fpt, units = None, None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
                traffic = json.load(fh).get(dom_kernel, {}).get(args.workload)
            if world > 1 or args.general_impedance:
                traffic = None      # captured at N=1 on the shipped configuration only
        except Exception:
            traffic = None
        if fpt:
            achieved = fpt * units / (per_launch_ms * 1e-3) / 1e12
            roofline = {"bound": "fp64", "kernel": dom_kernel, "achieved": round(achieved, 3),
                        "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": round(achieved / peak_tf, 4) if peak_tf else None,
                        "traffic": traffic,
                        "algorithmic_flops_per_term": fpt, "terms_per_launch": units,
                        "kernel_ms_per_launch": round(per_launch_ms, 4),
                        "kernel_share_of_step": round(ms_k / steps / ms_per_step, 4),
                        "peak_source": f"measured in this run: deism_fp64_peak {peak_kind} "
                                       "micro-benchmark (MEASURED_PEAKS.json has no FP64 entry)"}

    # ---- CPU baseline (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = run_cpu(p, 1, 0, cpu_terms)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": "terms/s", "cores": 0, "kind": "port",
                   "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": "image x frequency terms/s (run_DEISM)",
            "value": terms / (ms_per_step * 1e-3), "unit": "terms/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 1), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(config, images=I, terms=terms),
            "e2e": {"value": terms / (e2e_ms * 1e-3), "unit": "terms/s",
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": int(launches),
            "clocks": summarise_clocks(samples),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_rel_l2_vs_reference": parity,
            "parity_worst_single_frequency": parity_worst,
            "wall_ms_incl_l2_flush": wall_ms,
            "kernel_ms_per_launch": kernel_ms,
            "extras": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
