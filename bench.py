#!/usr/bin/env python
"""bench.py — image x frequency terms/s of run_DEISM on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c2mono|c3|c1|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: oracle port of the Numba backend

A "step" is one pass of the hot path (`run_DEISM`) over the workload: all images x all
frequencies of the named BASELINE config.  The headline (top-level keys of the JSON line) is C5:
shoebox RIR DEISM-MIX, order 40, 88 641 images x 16 000 frequencies, N=V=5.  `value` = I*K / t
with inputs resident in HBM; `e2e` = the same through deism_b200.backend.run_DEISM_device(params)
with HOST (pinned NumPy) buffers, host->device copies and the result read-back inside the timed
region; `e2e_pageable` = the same from plain (pageable) NumPy arrays through backend.run_DEISM,
which is what a reference caller holds.

`configs` (N=1, default run): every other BASELINE configuration (the
monopole run of C2) and the per-term wall-factor path (frequency-dependent impedance, what a
CHORAS job with band-wise absorption runs), each measured the same way with fewer steps and each
carrying its own ms_per_step, terms/s, e2e, dominant-kernel roofline and parity against the
reference's frozen RTF.  At N>1 `configs` holds C3 (the ORG workload) sharded like the headline.

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

METRIC = "image x frequency terms/s (run_DEISM)"
L2_NOTE = "L2 flushed (256 MiB write) between timed steps"


# --------------------------------------------------------------------------------------
# algorithmic work per term (SURVEY.md §8d conventions:
# complex MAC = 8 flops, complex mul = 6, complex x real = 2).  The directivity contractions are
# counted in the real-harmonic basis the kernels use (Y_n^-m = (-1)^m conj Y_n^m makes them
# real x complex: 4 flops per harmonic and term instead of 8), i.e. the cheapest formulation known
# to us, not the reference's complex x complex loop.
# --------------------------------------------------------------------------------------
def wall_factor_flops(mean_order, real_impedance, same_opposite_walls=True):
    """Flops the per-term attenuation chain EXECUTES (common.cuh), per term.
      real tile (shoebox_atten_real_n): per wall a factor 1 - 2/(zeta c + 1) = t, t+1, reciprocal
                     seed + one Newton step (4), 1 - 2r (2) = 8; then ONE simultaneous
                     square-and-multiply over the exponents: per bit of the largest exponent a
                     squaring and ~0.9 multiplies per wall (a warp holds 8 images: a multiply runs
                     when any of them has the bit); float rounding and 1/r = 2
      complex chain (shoebox_atten_n): factor = zr, zi, +-1 (4), |den|^2 (3), reciprocal (9),
                     numerator (8) = 24; power = ~1.5 log2(e) complex multiplies of 6 flops per
                     wall pair; product of the pairs 12; rounding 2
    e = mean exponent per pair = mean_order / 3."""
    import math

    e = max(2.0, mean_order / 3.0)
    chains = 3 * (1 if same_opposite_walls else 2)
    if real_impedance:
        bits = math.log2(e) + 1.5          # the largest of the exponents a warp sees
        return chains * 8 + bits * (1 + 0.9 * chains) + 2
    mults = 1.5 * math.log2(e)
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


def org_consume_flops_per_term(N, V, upward=False):
    """Flops per (image, frequency) term of the ORG factorised sum, L = N + V + 1 degrees: the
    degree contractions (real-harmonic rows x Re/Im planes: 4 (2l+1) per degree = 4 L^2), the sum
    over degrees, and the pair's start values (sincos, attenuation product).  The kernel takes the
    sum over degrees by Clenshaw's recurrence with the tensor core adding e_{l+2}: one multiply and
    two FMAs per degree (5 L).  `upward=True` is the count of the upward Hankel recurrence +
    complex accumulate the kernel ran before (14 L) — the same mathematical sum, kept so that the
    two rounds' fractions can be compared."""
    L = N + V + 1
    return 4 * L * L + (14 if upward else 5) * L + 40


def clocks_sampler(stop, out, gpu_index):
    """The clocks line of B200_PROFILING.md (SM clock, max SM clock, power, throttle reasons), sampled
    during the timed region.  NVML in-process (nvidia_ml_py): forking nvidia-smi ten times a second
    out of a process with gigabytes of pinned and device mappings stalls the host threads that the
    end-to-end legs are timing; nvidia-smi is the fallback."""
    nv = None
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        nv = (pynvml, h)
    except Exception:
        nv = None
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            if nv is not None:
                m, h = nv
                r = m.nvmlDeviceGetCurrentClocksEventReasons(h)
                act = lambda bit: "Active" if r & bit else "Not Active"
                out.append([str(m.nvmlDeviceGetClockInfo(h, m.NVML_CLOCK_SM)),
                            str(m.nvmlDeviceGetMaxClockInfo(h, m.NVML_CLOCK_SM)),
                            str(m.nvmlDeviceGetPowerUsage(h) / 1000.0),
                            act(m.nvmlClocksEventReasonHwSlowdown), act(m.nvmlClocksEventReasonHwThermalSlowdown),
                            act(m.nvmlClocksEventReasonSwThermalSlowdown), act(m.nvmlClocksEventReasonSwPowerCap)])
            else:
                r = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                   timeout=5)
                parts = [x.strip() for x in r.stdout.strip().split(",")]
                if len(parts) >= 7:
                    out.append(parts)
        except Exception:
            pass
        stop.wait(0.05 if nv is not None else 0.1)


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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------
def build_case(name, general=False):
    """(params, fixture, description) of a shoebox BASELINE workload; `general` perturbs Z_S along
    the frequency axis so that the per-term wall-factor path runs."""
    from deism_b200 import workloads

    p, fixture, desc = workloads.build(name)
    if general:
        K0 = len(p["waveNumbers"])
        p["impedance"] = np.ascontiguousarray(
            p["impedance"] * (1.0 + 0.1 * np.linspace(0.0, 1.0, K0))[None, :])
    return p, fixture, desc


def case_config(p, desc, world, sharding, general):
    return {"workload": desc, "method": p["DEISM_method"], "source_order": int(p["sourceOrder"]),
            "receiver_order": int(p["receiverOrder"]), "frequencies": len(p["waveNumbers"]),
            "sharding": (f"freq-slab x{world}" if sharding == "frequency" else f"image-slab x{world}")
            if world > 1 else "single GPU",
            "l2": L2_NOTE,
            "impedance": "general (per-term beta)" if general
            else "as shipped (frequency-flat: beta per image)"}


def lattice_size(order):
    """Images of an un-cut shoebox lattice of this reflection order."""
    return 1 + 2 * order * (2 * order * order + 3 * order + 4) // 3


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Numba backend
# --------------------------------------------------------------------------------------
CPU_RATE = {"c5": 3.0e6, "c2": 3.2e6, "c2mono": 1.0e7, "c1": 1.0e7, "c3": 2.4e3}   # terms/s per host thread


def cpu_sample_params(p, target_terms):
    """The workload itself when it fits the term budget; otherwise the same workload cut to a
    shallower lattice (same per-image work, all frequencies).  Attenuation materialised as the
    reference's default storage does."""
    from oracle import oracle

    K = len(p["waveNumbers"])
    q = dict(p)
    full_order = int(p["maxReflOrder"])
    order = full_order
    while order > 2 and lattice_size(order) * K > target_terms:
        order -= 1
    q["maxReflOrder"] = order
    t0 = time.time()
    images = oracle.shoebox_images(q)
    t_img = time.time() - t0
    if q["DEISM_method"] in ("ORG", "LC"):
        images = oracle.merge_images(images)
    q["images"] = images
    n_img = len(images["A"]) if "A" in images else len(images["A_early"]) + len(images["A_late"])
    return q, n_img, order, order == full_order, t_img


def run_cpu(p, steps, warmup, target_terms):
    from oracle import oracle

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread it is allowed
    oracle.set_num_threads(host_threads())
    q, n_img, order, full, t_img = cpu_sample_params(p, target_terms)
    K = len(q["waveNumbers"])
    for _ in range(warmup):
        oracle.run_DEISM(q)
    t0 = time.time()
    for _ in range(steps):
        oracle.run_DEISM(q)
    dt = (time.time() - t0) / steps
    what = "the whole workload" if full else f"same workload cut to reflection order {order}"
    return {"value": n_img * K / dt, "unit": "terms/s", "cores": oracle.num_threads(),
            "kind": "port",
            "sample": f"{what}: {n_img} images x {K} freqs = {n_img * K:.3e} terms, {dt:.2f} s/step; "
                      f"CPU image generation incl. attenuation materialisation {t_img:.2f} s (not in "
                      "value); Wigner tables from the repo's native generator (hash-equal to sympy's)",
            "ms_per_step": dt * 1e3, "terms": n_img * K, "whole_workload": full}


def numba_reference_table():
    """Warm run_DEISM times of the UNMODIFIED reference (Numba backend) measured in the build
    container by oracle/time_reference.py; it cannot travel to the GPU box."""
    try:
        with open(os.path.join(ROOT, "profiles", "reference_numba_build_container.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx(args):
    import torch
    import torch.distributed as dist

    from deism_b200 import _lib

    c = Ctx()
    c.rank = int(os.environ.get("RANK", "0"))
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local_rank)
    c.dev = torch.device("cuda", c.local_rank)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    c.lib = _lib.lib()
    c.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=c.dev)
    c.peaks = {}
    for kind, nm in ((0, "dfma"), (1, "dmma_m8n8k4")):
        tf, ms = C.c_double(0), C.c_double(0)
        if c.lib.deism_fp64_peak(kind, 20000, C.byref(tf), C.byref(ms)) == 0:
            c.peaks[nm] = tf.value
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            c.traffic = json.load(fh)
    except Exception:
        c.traffic = {}
    return c


def barrier(c):
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if c.world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(c, fn, n):
    """n steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
    outside the brackets.  Returns (sum of device ms, wall ms incl. flushes)."""
    import torch

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(n)]
    barrier(c)
    w0 = time.perf_counter()
    for a, b in evs:
        c.flush.fill_(1)
        a.record()
        fn()
        b.record()
    barrier(c)
    wall = (time.perf_counter() - w0) * 1e3
    if os.environ.get("DEISM_BENCH_DEBUG") and c.rank == 0:
        print("per-step ms:", [round(a.elapsed_time(b), 3) for a, b in evs], file=sys.stderr)
    return sum(a.elapsed_time(b) for a, b in evs), wall


def max_over_ranks(c, x):
    import torch
    import torch.distributed as dist

    if c.world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=c.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


PROFILED = ("lc_main", "org_consume", "org_btable", "org_direct", "arg_lc")


def profile_read(c):
    prof = {}
    for name in PROFILED:
        ms, n = C.c_double(0), C.c_int(0)
        c.lib.deism_profile_get(name.encode(), C.byref(ms), C.byref(n))
        if n.value:
            prof[name] = (ms.value, n.value)
    return prof


def peak_of(c, kind):
    if kind == "dfma" and c.peaks.get("dfma"):
        return c.peaks["dfma"], "dfma"
    best = max(c.peaks, key=lambda k: c.peaks[k]) if c.peaks else None
    return (c.peaks[best], best) if best else (None, None)


def measure_shoebox(c, name, steps, warmup, general=False, sharding="frequency", pageable=True):
    """One shoebox workload: device-resident, e2e from pinned host buffers, e2e from pageable NumPy."""
    import torch

    from deism_b200 import _lib, backend, distributed, images as gimages, workloads

    p, fixture, desc = build_case(name, general)
    method, N, V = p["DEISM_method"], int(p["sourceOrder"]), int(p["receiverOrder"])
    K = len(p["waveNumbers"])
    dev, rank, world = c.dev, c.rank, c.world

    full_images = gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
    if method in ("ORG", "LC"):
        full_images = gimages.merge_images(full_images)
    I = workloads.n_images(full_images)
    terms = I * K

    by_image = sharding == "image" and world > 1
    if by_image:
        if method == "MIX":
            raise SystemExit("--sharding image needs a merged image list (ORG or LC workload)")
        p_all = dict(p)
        p_all["images"] = full_images
        ps, _, _ = distributed.shard_images(p_all, rank, world)
        full_images = ps["images"]
        lo, hi = 0, K
    else:
        ps, lo, hi = distributed.shard_frequencies(p, rank, world)
        ps["images"] = full_images
    Ks = hi - lo
    I_rank = workloads.n_images(full_images)

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

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
    h2d_bytes = sum(p_host["images"][key].nbytes for key in used_img)
    if world > 1 and not by_image and os.environ.get("DEISM_B200_REPLICATE_IMAGES", "0") == "1":   # row slabs
        h2d_bytes = sum(-(-p_host["images"][key].shape[0] // world) * p_host["images"][key][0].nbytes
                        if p_host["images"][key].shape[0] >= 4 * world else p_host["images"][key].nbytes
                        for key in used_img)
    h2d_bytes += sum(p_host[key].nbytes for key in ("waveNumbers", "impedance"))
    if method in ("LC", "MIX"):
        h2d_bytes += sum(p_host[key].nbytes for key in ("C_nm_s_vec", "C_vu_r_vec"))
    if method in ("ORG", "MIX"):
        h2d_bytes += sum(p_host[key].nbytes for key in ("C_nm_s", "C_vu_r"))
    d2h_bytes = Ks * 16

    p_dev = dict(ps)
    p_dev["images"] = {k: (torch.from_numpy(np.ascontiguousarray(v)).to(dev)
                           if isinstance(v, np.ndarray) else v) for k, v in full_images.items()}
    for key, dt in (("waveNumbers", torch.float64), ("impedance", torch.complex128),
                    ("C_nm_s", torch.complex64), ("C_vu_r", torch.complex64),
                    ("C_nm_s_vec", torch.complex64), ("C_vu_r_vec", torch.complex64)):
        if key in ps:
            p_dev[key] = torch.from_numpy(np.ascontiguousarray(ps[key])).to(dt).to(dev)

    # plain NumPy (pageable) inputs, as a caller of the reference holds them
    p_np = dict(ps)
    p_np["images"] = dict(full_images)

    def combine(slab):
        if by_image:
            return distributed.reduce_image_slabs(slab)
        return distributed.gather_frequency_slabs(slab, K)

    def step_device():
        return combine(backend.run_DEISM_device(p_dev, dev))

    def step_e2e():
        # H2D of every input inside (DEISM_B200_REPLICATE_IMAGES=1: the image list enters the node
        # once — every rank uploads its row slab, the slabs are all_gathered over NVLink)
        q = p_host if by_image else distributed.replicate_images(p_host, dev)
        res = combine(backend.run_DEISM_device(q, dev))
        return res.cpu() if rank == 0 or world == 1 else res[:1].cpu()

    def step_pageable():
        res = combine(backend.run_DEISM_device(p_np, dev))
        return res.cpu().numpy() if rank == 0 or world == 1 else res[:1].cpu()

    graphs_on = backend.USE_GRAPHS
    result = None
    for _ in range(max(warmup, 1)):
        result = step_device()
    torch.cuda.synchronize()

    # ---- parity of this very run against the reference's frozen RTF ----
    parity, parity_worst, parity_note = None, None, None
    if "RTF" in fixture.files and not general:
        ref = fixture["RTF"]
        got = result.cpu().numpy()
        if ref.shape[0] == K:
            parity = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
            parity_worst = float(np.max(np.abs(got - ref) / np.abs(ref)))
            parity_note = f"all {K} frequencies against the reference's own run_DEISM"

    # per-kernel times: a short pass with the library's event brackets on and graph replay off
    # (events cannot be timed inside a replayed graph); then the timed region proper
    backend.USE_GRAPHS = False
    c.lib.deism_profile_enable(1)
    c.lib.deism_profile_reset()
    n_prof = min(steps, 3)
    timed(c, step_device, n_prof)
    prof = profile_read(c)
    c.lib.deism_profile_enable(0)
    backend.USE_GRAPHS = graphs_on
    kernel_ms = {nm: round(ms / n, 4) for nm, (ms, n) in prof.items()}
    for _ in range(2):          # first call with these buffers runs eagerly, the second one captures
        step_device()
    launches0 = _lib.launch_count()
    dev_ms, wall_ms = timed(c, step_device, steps)
    launches = _lib.launch_count() - launches0
    ms_per_step = max_over_ranks(c, dev_ms) / steps

    for _ in range(2):
        step_e2e()
    if os.environ.get("DEISM_BENCH_PROFILE") == name and world == 1:     # debugging aid (single process only)
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        timed(c, step_e2e, steps)
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(25)
    e2e_ms, _ = timed(c, step_e2e, steps)
    e2e_ms = max_over_ranks(c, e2e_ms) / steps
    # the end-to-end path must return what the device-resident one returned.  step_e2e() holds a
    # collective when N > 1: EVERY rank calls it, rank 0 (which holds both results) compares.
    e2e_diff = None
    got_e2e = step_e2e()
    if rank == 0:
        got_dev = result.cpu().numpy()
        e2e_diff = float(np.linalg.norm(got_e2e.numpy() - got_dev) / np.linalg.norm(got_dev))
    pg_ms = None
    if pageable:
        for _ in range(2):
            step_pageable()
        pg_ms, _ = timed(c, step_pageable, steps)
        pg_ms = max_over_ranks(c, pg_ms) / steps

    roofline = None
    if rank == 0 and prof:
        dom = max(prof, key=lambda n: prof[n][0])
        dom_kernel = dom + "_kernel"
        ms_k, n_k = prof[dom]
        per_launch_ms = ms_k / n_k
        peak_tf, peak_kind = peak_of(c, "dmma")
        fpt, units = None, None
        if dom == "lc_main":
            n_late = len(full_images["A_late"]) if method == "MIX" else I_rank
            Js, Jr = (N + 1) ** 2, (V + 1) ** 2
            Zh = np.asarray(p["impedance"])
            z_real = bool(np.all(np.abs(Zh.imag) <= 4e-16 * np.abs(Zh.real)))       # common.cuh z_is_real
            z_same = bool(np.array_equal(Zh[0::2], Zh[1::2]))
            fpt = round(lc_flops_per_term(Js, Jr, general, 0.75 * p["maxReflOrder"], z_real, z_same), 1)
            units = n_late * Ks
            if Js == 1 and Jr == 1:
                dom_kernel = "lc_mono_kernel"      # what deism_lc_shoebox launches for monopoles
                peak_tf, peak_kind = peak_of(c, "dfma")   # no DMMA in it: the FP64 FMA pipe
        elif dom == "org_consume":
            fpt = org_consume_flops_per_term(N, V)
            units = I_rank * Ks
        traffic = None
        if world == 1 and not general:
            traffic = c.traffic.get(dom_kernel, {}).get(name)
        if fpt:
            achieved = fpt * units / (per_launch_ms * 1e-3) / 1e12
            roofline = {"bound": "fp64", "kernel": dom_kernel, "achieved": round(achieved, 3),
                        "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": round(achieved / peak_tf, 4) if peak_tf else None,
                        "traffic": traffic,
                        "algorithmic_flops_per_term": fpt, "terms_per_launch": units,
                        **({"frac_on_upward_recurrence_flop_count": round(
                            org_consume_flops_per_term(N, V, upward=True) * units / (per_launch_ms * 1e-3)
                            / 1e12 / peak_tf, 4)} if dom == "org_consume" and peak_tf else {}),
                        "kernel_ms_per_launch": round(per_launch_ms, 4),
                        "kernel_share_of_step": round(ms_k / n_prof / ms_per_step, 4),
                        "step_level_frac": round(fpt * units / (ms_per_step * 1e-3) / 1e12 / peak_tf, 4)
                        if peak_tf else None,
                        "peak_source": f"measured in this run: deism_fp64_peak {peak_kind} "
                                       "micro-benchmark (MEASURED_PEAKS.json has no FP64 entry)"}

    out = {
        "config": dict(case_config(p, desc, world, sharding, general), images=I, terms=terms),
        "ms_per_step": ms_per_step, "value": terms / (ms_per_step * 1e-3), "unit": "terms/s",
        "steps": steps, "warmup": max(warmup, 1),
        "e2e": {"value": terms / (e2e_ms * 1e-3), "unit": "terms/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "host_buffers": "pinned NumPy", "rel_diff_vs_device_resident_result": e2e_diff},
        "e2e_pageable": ({"value": terms / (pg_ms * 1e-3), "unit": "terms/s", "ms_per_step": pg_ms,
                          "host_buffers": "plain (pageable) NumPy through backend.run_DEISM, result as NumPy"}
                         if pg_ms else None),
        "gpu_launches": int(launches),
        "cuda_graph_replay": bool(graphs_on),
        "roofline": roofline,
        "parity_rel_l2_vs_reference": parity,
        "parity_worst_single_frequency": parity_worst,
        "parity_checked_on": parity_note,
        "wall_ms_incl_l2_flush": wall_ms,
        "kernel_ms_per_launch": kernel_ms,
    }
    del p_dev, p_host, keep
    return out, p


def measure_convex(c, steps, warmup):
    """BASELINE config C4 (convex ARG room, RIR yaml, order 10, K=4322, monopole, MIX): reflection
    tree on the GPU + ARG accumulation; device-resident pipeline and the reference's host-array
    interface."""
    import torch

    from deism_b200 import _lib, backend, convex, directivities as D, wigner, workloads

    z = np.load(os.path.join(ROOT, "tests", "golden", "c4_full.npz"))

    class _W:
        pass

    walls = []
    for i in range(int(z["walls/n"])):
        w = _W()
        w.normal, w.origin = z[f"walls/{i}/normal"], z[f"walls/{i}/origin"]
        w.basis, w.flat_corners = z[f"walls/{i}/basis"], z[f"walls/{i}/flat_corners"]
        w.reflection_matrix, w.impedence_bands = z[f"walls/{i}/reflection_matrix"], z[f"walls/{i}/impedance"]
        walls.append(w)
    order = int(z["params/maxReflOrder"])
    src, mic = z["params/posSource"].astype(np.float32), z["params/posReceiver"].astype(np.float32)
    k = z["params/waveNumbers"]
    K = len(k)
    room = convex.Room_deism(walls, [], [], 343.0, order)
    room.add_mic(mic)
    n = room.image_source_model(src)
    tree_ok = bool(np.array_equal(room.sources.view(np.uint32), z["engine/sources"].view(np.uint32))
                   and np.array_equal(room.orders, z["engine/orders"])
                   and np.array_equal(room.gen_walls, z["engine/gen_walls"])
                   and np.array_equal(room.attenuations.view(np.uint32), z["engine/attenuations"].view(np.uint32)))
    mono = workloads.monopole_coeffs(k)
    base = {"DEISM_method": "MIX", "silentMode": 1, "waveNumbers": k, "posReceiver": z["params/posReceiver"],
            "ifRemoveDirectPath": int(z["params/ifRemoveDirectPath"]),
            "mixEarlyOrder": int(z["params/mixEarlyOrder"]), "sourceOrder": 0, "receiverOrder": 0,
            "sourceType": "monopole", "C_vu_r": mono}
    base["v_all"], base["u_all"], base["C_vu_r_vec"] = workloads.vectorize(mono, 0)
    base = wigner.pre_calc_Wigner(base)

    def host_params():
        """What the reference's glue hands run_DEISM_ARG: NumPy arrays (core_deism_arg.py:1204-1262 +
        the monopole branch of init_source_directivities_ARG)."""
        q = convex.get_ref_paths_ARG(dict(base), room)
        q = D.vectorize_C_nm_s_ARG(D.init_source_directivities_ARG(q))
        return q

    def device_params():
        q = convex.get_ref_paths_ARG_device(dict(base), room)
        return D.vectorize_C_nm_s_ARG(D.init_source_directivities_ARG(q))

    def step_tree():
        return room.image_source_model(src)

    q_dev = device_params()

    def step_device():
        return backend.run_DEISM_ARG_device(q_dev, c.dev)

    def step_pipeline():       # tree -> geometry -> coefficients -> accumulation, nothing leaves the GPU
        room.image_source_model(src)
        return backend.run_DEISM_ARG_device(device_params(), c.dev)

    def step_host():           # the reference's contract: every intermediate a NumPy array
        room.image_source_model(src)
        return backend.run_DEISM_ARG(host_params())

    for _ in range(max(warmup, 1)):
        out = step_device()
    err = float(np.linalg.norm(out.cpu().numpy() - z["RTF"]) / np.linalg.norm(z["RTF"]))
    c.lib.deism_profile_enable(1)
    c.lib.deism_profile_reset()
    l0 = _lib.launch_count()
    dev_ms, _ = timed(c, step_device, steps)
    launches = _lib.launch_count() - l0
    prof = profile_read(c)
    c.lib.deism_profile_enable(0)
    ms = dev_ms / steps
    for _ in range(2):
        step_tree()
    tree_ms, _ = timed(c, step_tree, steps)
    for _ in range(2):
        out_p = step_pipeline()
    err_p = float(np.linalg.norm(out_p.cpu().numpy() - z["RTF"]) / np.linalg.norm(z["RTF"]))
    pipe_ms, _ = timed(c, lambda: step_pipeline().cpu(), steps)
    for _ in range(2):
        out_h = step_host()
    err_h = float(np.linalg.norm(out_h - z["RTF"]) / np.linalg.norm(z["RTF"]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) / steps * 1e3
    terms = int(n) * K
    roofline = None
    if "arg_lc" in prof:
        ms_k, n_k = prof["arg_lc"]
        n_late = len(q_dev["images"]["late_indices"])
        # arg_lc_kernel streams the per-image source coefficients C_s[K, J, I] (complex64) and the
        # attenuation [K, I] (complex64): 8 B per (k, j, image) + 8 B per term; monopole J = 1
        nbytes = 16.0 * n_late * K
        hbm = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm = float(json.load(fh)["hbm_gbs"])
        except Exception:
            hbm = 6553.0
        ach = nbytes / (ms_k / n_k * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "arg_lc_kernel", "achieved": round(ach, 1), "peak": hbm,
                    "unit": "GB/s", "frac": round(ach / hbm, 4), "traffic": None,
                    "algorithmic_bytes_per_term": 16, "terms_per_launch": n_late * K,
                    "kernel_ms_per_launch": round(ms_k / n_k, 4),
                    "note": "1 577 images x 4 322 frequencies is 0.1 GB of input: the launch is "
                            "latency-, not bandwidth-bound"}
    return {
        "config": {"workload": "C4 convex ARG room RIR yaml (configSingleParam_ARG_RIR.yml), MIX order 10, "
                               "monopole, K=4322", "method": "MIX", "images": int(n), "frequencies": K,
                   "terms": terms, "tree_nodes": 720247, "l2": L2_NOTE},
        "ms_per_step": ms, "value": terms / (ms * 1e-3), "unit": "terms/s", "steps": steps,
        "warmup": max(warmup, 1),
        "tree_enumeration_ms": tree_ms / steps, "tree_bit_exact_vs_reference": tree_ok,
        "device_pipeline": {"ms_per_step": pipe_ms / steps, "rel_l2_vs_reference": err_p,
                            "what": "reflection tree + geometry + monopole coefficients + accumulation, "
                                    "arrays stay on the GPU, RTF copied back"},
        "e2e": {"value": terms / (host_ms * 1e-3), "unit": "terms/s", "ms_per_step": host_ms,
                "rel_l2_vs_reference": err_h, "host_buffers": "plain NumPy",
                "what": "reflection tree + the reference's host-array interface (get_ref_paths_ARG, "
                        "init_source_directivities_ARG, run_DEISM_ARG with NumPy in / NumPy out)"},
        "gpu_launches": int(launches), "roofline": roofline,
        "parity_rel_l2_vs_reference": err,
        "reference_cpu_build_container": {"tree_s": float(z["meta/t_images_s"]),
                                          "run_cold_s": float(z["meta/t_run_cold_s"])},
    }


def slim(res):
    """Per-config entry of the `configs` block."""
    keys = ("config", "ms_per_step", "value", "unit", "steps", "warmup", "e2e", "e2e_pageable",
            "gpu_launches", "roofline", "parity_rel_l2_vs_reference", "parity_worst_single_frequency",
            "parity_checked_on", "kernel_ms_per_launch", "tree_enumeration_ms",
            "tree_bit_exact_vs_reference", "device_pipeline", "reference_cpu_build_container")
    return {k: res[k] for k in keys if k in res and res[k] is not None}


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c2mono", "c3", "c4", "c5"])
    ap.add_argument("--general-impedance", action="store_true",
                    help="perturb Z_S across frequency so that the per-term beta path runs")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="CPU time the reference arm may spend in total (default 150 s; 20 s for "
                         "the cpu_baseline leg of the B200 arm)")
    ap.add_argument("--sharding", default="frequency", choices=["frequency", "image"],
                    help="N>1: contiguous frequency slabs + all_gather (default) or, for the merged "
                         "ORG/LC image lists of few-frequency runs, image slabs + all_reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="headline workload only (skip the `configs` block)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(args.warmup, 0)
    cpu_name = args.workload if args.workload != "c4" else "c5"

    # ---------------- reference arm: CPU, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        p, fixture, desc = build_case(cpu_name, args.general_impedance)
        budget = args.cpu_seconds or 150.0
        rate = CPU_RATE[cpu_name] * host_threads()
        res = run_cpu(p, steps, warmup, rate * budget / (steps + warmup))
        K = len(p["waveNumbers"])
        I_full = int(np.asarray(fixture["meta/n_images"])) if "meta/n_images" in fixture.files else None
        cfg = case_config(p, desc, 1, "frequency", args.general_impedance)
        if I_full:
            cfg = dict(cfg, images=I_full, terms=I_full * K)
        line = {"impl": "reference", "metric": METRIC,
                "value": res["value"], "unit": "terms/s", "n_gpus": 0, "steps": steps,
                "warmup": warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample",
                                                     "whole_workload")},
                "e2e": {"value": res["value"], "unit": "terms/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "reference_numba_build_container": numba_reference_table(),
                "note": "oracle port (C/OpenMP restatement of the Numba backend, pinned to the "
                        "reference's outputs) on all host threads; terms/s does not depend on the cut "
                        "(per-image work is the same at every reflection order); the Python/Numba "
                        "reference and Ray cannot travel to the GPU box"}
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist

    c = make_ctx(args)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, c.local_rank), daemon=True)

    if args.workload == "c4":
        if world > 1:
            raise SystemExit("C4 is a single-GPU workload (1 577 images)")
        if rank == 0:
            th.start()
        head = measure_convex(c, steps, warmup)
        p_head = None
    else:
        if rank == 0:
            th.start()
        head, p_head = measure_shoebox(c, args.workload, steps, warmup, args.general_impedance,
                                       args.sharding)
    stop.set()

    # ---- the other BASELINE configurations ----
    configs = {}
    if not args.no_configs and args.workload == "c5" and not args.general_impedance:
        s2, w2 = min(steps, 5), max(min(warmup, 3), 3)
        if world == 1:
            plan = [("c1", "c1", False), ("c2", "c2", False), ("c2mono", "c2mono", False),
                    ("c3", "c3", False), ("c4", None, False), ("c5_general_impedance", "c5", True),
                    ("c2mono_general_impedance", "c2mono", True), ("c3_general_impedance", "c3", True)]
        else:
            plan = [("c3", "c3", False)]
        for key, wl, gen in plan:
            try:
                if wl is None:
                    configs[key] = slim(measure_convex(c, s2, w2))
                else:
                    configs[key] = slim(measure_shoebox(c, wl, s2, w2, gen, "frequency",
                                                        pageable=(world == 1))[0])
            except Exception as exc:       # a failing side config must not lose the headline
                configs[key] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- extras (rank 0, N=1): image enumeration time, cuBLAS DGEMM witness of the FP64 peak ----
    extras = {f"fp64_peak_{k}_tflops": round(v, 2) for k, v in c.peaks.items()}
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            n = 6144
            A = torch.randn((n, n), dtype=torch.float64, device=c.dev)
            B = torch.randn((n, n), dtype=torch.float64, device=c.dev)
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
        if p_head is not None:
            from deism_b200 import images as gimages

            gimages.shoebox_images_device(p_head, c.dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                gimages.shoebox_images_device(p_head, c.dev)
            torch.cuda.synchronize()
            extras["image_enumeration_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)

    # ---- CPU baseline (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            pc = p_head if p_head is not None else build_case(cpu_name)[0]
            rate = CPU_RATE[cpu_name] * host_threads()
            res = run_cpu(pc, 1, 0, rate * (args.cpu_seconds or 20.0))
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "whole_workload")}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": "terms/s", "cores": 0, "kind": "port",
                   "sample": f"failed: {exc}"}

    if rank == 0:
        total_launches = head["gpu_launches"]
        line = {
            "metric": METRIC,
            "value": head["value"], "unit": "terms/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 1), "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": head["config"],
            "e2e": head["e2e"],
            "e2e_pageable": head.get("e2e_pageable"),
            "gpu_launches": int(total_launches),
            "clocks": summarise_clocks(samples),
            "roofline": head["roofline"],
            "cpu_baseline": cpu,
            "parity_rel_l2_vs_reference": head.get("parity_rel_l2_vs_reference"),
            "parity_worst_single_frequency": head.get("parity_worst_single_frequency"),
            "wall_ms_incl_l2_flush": head.get("wall_ms_incl_l2_flush"),
            "kernel_ms_per_launch": head.get("kernel_ms_per_launch"),
            "configs": configs,
            "reference_numba_build_container": numba_reference_table(),
            "extras": extras,
        }
        for k in ("tree_enumeration_ms", "tree_bit_exact_vs_reference", "device_pipeline"):
            if k in head:
                line[k] = head[k]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
