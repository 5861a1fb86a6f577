"""Multi-GPU sharding of the run_DEISM hot path.

Every (image, frequency) term is independent and the only coupling is the sum over images per
frequency (parallel_backends.py:257-261), so the path shards without any data-path exchange:

  * primary: every rank owns a contiguous FREQUENCY slab [lo, hi) and all images; the slabs
    are concatenated with one all_gather of hi-lo complex128 values per rank;
  * secondary (few frequencies, many images): contiguous IMAGE slabs, all frequencies, one
    all_reduce(sum) over 2K doubles.

One process per GPU, ``torch.distributed`` for the plumbing (NCCL on GPUs; the same code runs
over gloo with CPU tensors, which is how the host logic is tested without GPUs).  The compute
callback is injected so that this module never decides *how* a slab is computed.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

_FREQ_KEYS = ("waveNumbers", "freqs", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec",
              "C_nm_s_ARG", "C_nm_s_ARG_vec", "pointSrcStrength")
_IMAGE_KEYS = ("A", "R_sI_r_all", "R_s_rI_all", "R_r_sI_all", "atten_all")


def slab_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced [lo, hi) of ``n`` items for ``rank`` (sizes differ by at most 1)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_counts(n: int, world: int):
    return [slab_bounds(n, world, r)[1] - slab_bounds(n, world, r)[0] for r in range(world)]


def _slice0(x, lo, hi):
    if isinstance(x, torch.Tensor):
        return x[lo:hi].contiguous()
    if hasattr(x, "tensor") and isinstance(getattr(x, "tensor"), torch.Tensor):   # backend.DeviceArray
        return type(x)(x.tensor[lo:hi].contiguous(), x.dtype)
    return np.ascontiguousarray(np.asarray(x)[lo:hi])


def check_same_job(params, group=None):
    """Sharding splits ONE simulation over the ranks; ranks that hold different simulations (a
    script that uses torchrun for data parallelism, one room per rank) must not be spliced together.
    Every rank hashes what defines the job (wave numbers, geometry, image count, method, orders);
    the hashes are compared with one all_gather and a mismatch raises on every rank."""
    import hashlib

    h = hashlib.sha256()
    wn = params["waveNumbers"]
    wn = wn.detach().cpu().numpy() if isinstance(wn, torch.Tensor) else np.asarray(wn)
    h.update(np.ascontiguousarray(wn, dtype=np.float64).tobytes())
    for key in ("posSource", "posReceiver", "roomSize", "vertices"):
        if params.get(key) is not None:
            h.update(np.ascontiguousarray(np.asarray(params[key], dtype=np.float64)).tobytes())
    images = params.get("images") or {}
    for key in ("A", "A_early", "A_late", "R_sI_r_all"):
        if key in images and images[key] is not None:
            h.update(f"{key}:{tuple(images[key].shape)}".encode())
    h.update(f"{params.get('DEISM_method')}:{params.get('sourceOrder')}:{params.get('receiverOrder')}".encode())
    mine = np.frombuffer(h.digest()[:8], dtype=np.int64).copy()
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.from_numpy(mine).to(dev)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t, group=group)
    vals = {int(b.item()) for b in bufs}
    if len(vals) != 1:
        raise RuntimeError(
            "deism_b200: params['b200Sharding'] splits ONE simulation over the ranks of the process "
            "group, but the ranks hold different jobs (wave numbers, geometry, image lists or orders "
            "differ).  Leave b200Sharding at 'off' when every rank simulates its own room.")


def shard_frequencies(params, rank: int, world: int):
    """params restricted to this rank's frequency slab (images are shared, not copied)."""
    K = int(np.asarray(params["waveNumbers"]).shape[0]) \
        if not isinstance(params["waveNumbers"], torch.Tensor) else int(params["waveNumbers"].shape[0])
    lo, hi = slab_bounds(K, world, rank)
    q = dict(params)
    for key in _FREQ_KEYS:
        if key in q and q[key] is not None:
            q[key] = _slice0(q[key], lo, hi)
    Z = q.get("impedance")
    if Z is not None:
        q["impedance"] = Z[:, lo:hi].contiguous() if isinstance(Z, torch.Tensor) \
            else np.ascontiguousarray(np.asarray(Z)[:, lo:hi])
    images = q.get("images")
    if images is not None:
        im = dict(images)
        shoebox = "A" in images or "A_early" in images
        for key, val in images.items():
            if key.startswith("atten_all") and val is not None:
                if shoebox:      # materialised shoebox attenuation is [I, K]
                    im[key] = val[:, lo:hi].contiguous() if isinstance(val, torch.Tensor) \
                        else np.ascontiguousarray(np.asarray(val)[:, lo:hi])
                else:            # convex attenuation is [K, I]
                    im[key] = _slice0(val, lo, hi)
        q["images"] = im
    return q, lo, hi


def shard_images(params, rank: int, world: int):
    """params restricted to this rank's slab of a merged (ORG / LC) shoebox image list."""
    images = params["images"]
    if "A" not in images:
        raise ValueError("image sharding needs the merged ORG/LC image dict")
    n = len(images["A"])
    lo, hi = slab_bounds(n, world, rank)
    q = dict(params)
    im = dict(images)
    for key in _IMAGE_KEYS:
        if key in im and im[key] is not None:
            im[key] = _slice0(im[key], lo, hi)
    q["images"] = im
    return q, lo, hi


class ImageReplicator:
    """Host image arrays -> full device copies on every rank, each host byte crossing PCIe once per
    NODE instead of once per GPU.

    With frequency slabs every rank needs the whole image list (8.8 MB at C5).  The ranks hold the
    same host arrays, so rank r uploads only rows [r, r+1) * ceil(I / G) of each array and the row
    slabs are all_gathered over NVLink / NVSwitch (one collective per array, ~1 MB per rank).  At 8
    GPUs that replaces 8 x 8.8 MB of simultaneous host -> device traffic — the ranks share PCIe
    switches — by 8 x 1.1 MB.  The gathered buffers are allocated once per shape, so repeated calls
    hand the kernels the same device pointers (graph replay keeps working).

    OPT-IN (``DEISM_B200_REPLICATE_IMAGES=1``): measured on 8 B200s the collectives cost what the
    saved PCIe time buys (C5 end to end 2.41 ms with, 2.42 ms without; C3 0.79 against 0.72 ms; on 2
    GPUs 7.92 against 7.71 ms) — at these sizes the end-to-end gap is copy and launch latency, not
    PCIe bandwidth.  It pays only when the image list is tens of MB per rank."""

    def __init__(self):
        self._bufs = {}

    def replicate(self, images, keys, dev, group=None):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        out = {}
        for key in keys:
            a = images.get(key)
            if a is None or isinstance(a, torch.Tensor) or not isinstance(a, np.ndarray) or a.shape[0] < 4 * world:
                continue
            a = np.ascontiguousarray(a)
            rows = a.shape[0]
            per = (rows + world - 1) // world
            tail = tuple(a.shape[1:])
            sig = (key, rows, tail, str(a.dtype), dev.index)
            ent = self._bufs.get(sig)
            if ent is None:
                tdt = torch.from_numpy(a[:0]).dtype
                ent = (torch.zeros((world * per,) + tail, dtype=tdt, device=dev),
                       torch.zeros((per,) + tail, dtype=tdt, device=dev))
                self._bufs[sig] = ent
            full, mine = ent
            lo, hi = min(rank * per, rows), min((rank + 1) * per, rows)
            if hi > lo:
                mine[:hi - lo].copy_(torch.from_numpy(a[lo:hi]), non_blocking=True)
            dist.all_gather_into_tensor(full, mine, group=group)
            out[key] = full[:rows]
        return out


_replicator = ImageReplicator()


def replicate_images(params, dev=None, group=None):
    """params whose (host) shoebox image arrays are replaced by device copies built with
    ImageReplicator; a no-op outside an NCCL process group or when the images are already tensors."""
    if os.environ.get("DEISM_B200_REPLICATE_IMAGES", "0") != "1":
        return params
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return params
    if dist.get_backend(group) != "nccl":
        return params
    images = params.get("images")
    if not images or not ("A" in images or "A_late" in images):
        return params
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    # only what the accumulation reads (backend._run_lc / _run_org): ORG takes A + R_sI_r, LC takes
    # R_s_rI + R_r_sI (+ A for the fused attenuation, + R_sI_r for the compact-storage one)
    method, storage = params.get("DEISM_method"), images.get("storage")

    def needed(k):
        if not k.startswith(("A", "R_")) or not isinstance(images[k], np.ndarray):
            return False
        org_part = method == "ORG" or (method == "MIX" and k.endswith("_early"))
        if org_part:
            return k.startswith(("A", "R_sI_r_all"))
        if k.startswith("R_sI_r_all"):
            return storage == "compact"
        return True

    keys = [k for k in images if needed(k)]
    rep = _replicator.replicate(images, keys, dev, group)
    if not rep:
        return params
    q = dict(params)
    q["images"] = dict(images)
    q["images"].update(rep)
    return q


def gather_frequency_slabs(slab: torch.Tensor, K: int, group=None) -> torch.Tensor:
    """The one collective of the frequency-sharded path: every rank ends up with the full
    complex128[K] result.  Equal slabs use all_gather_into_tensor; ragged ones are padded."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return slab
    counts = slab_counts(K, world)
    slab = slab.contiguous()
    if len(set(counts)) == 1:
        out = torch.empty(K, dtype=slab.dtype, device=slab.device)
        if slab.is_complex():
            dist.all_gather_into_tensor(torch.view_as_real(out), torch.view_as_real(slab), group=group)
        else:
            dist.all_gather_into_tensor(out, slab, group=group)
        return out
    width = max(counts)
    padded = torch.zeros(width, dtype=slab.dtype, device=slab.device)
    padded[:slab.shape[0]] = slab
    bufs = [torch.empty((width, 2), dtype=torch.float64, device=slab.device) for _ in range(world)]
    dist.all_gather(bufs, torch.view_as_real(padded).contiguous(), group=group)
    parts = [torch.view_as_complex(b.contiguous())[:c] for b, c in zip(bufs, counts)]
    return torch.cat(parts)


def reduce_image_slabs(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Image-sharded path: sum of the per-rank partial RTFs (one all_reduce over 2K doubles)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    buf = torch.view_as_real(partial.contiguous()).clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return torch.view_as_complex(buf)


def run_sharded(params, compute, mode: str = "frequency", group=None) -> torch.Tensor:
    """Run ``compute(params_slab) -> complex128 tensor`` on this rank's slab and combine.

    ``compute`` is e.g. ``deism_b200.backend.run_DEISM_device``; ``mode`` is "frequency"
    (all_gather) or "image" (all_reduce)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if mode == "frequency":
        K = int(np.asarray(params["waveNumbers"]).shape[0]) \
            if not isinstance(params["waveNumbers"], torch.Tensor) else int(params["waveNumbers"].shape[0])
        q, lo, hi = shard_frequencies(params, rank, world)
        q = replicate_images(q, group=group)      # opt-in: host image lists cross PCIe once per node
        return gather_frequency_slabs(compute(q), K, group)
    if mode == "image":
        q, lo, hi = shard_images(params, rank, world)
        return reduce_image_slabs(compute(q), group)
    raise ValueError(f"unknown sharding mode: {mode}")
