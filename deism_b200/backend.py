"""Host side of the accumulation hot path: the drop-in for the reference's Numba backend.

Mirrors ``deism/parallel_backends.py``:

    run_DEISM(params)      ~ pb:1572 -> run_DEISM_numba      pb:912-922   (shoebox)
    run_DEISM_ARG(params)  ~ pb:1577 -> run_DEISM_ARG_numba  pb:925-935   (convex)

Same ``params`` keys, same return value (``complex128[K]``), same error behaviour
(``ValueError`` for an unknown ``DEISM_method``; zero images -> zeros).  The work is done by
the sm_100a kernels of ``libdeism_b200.so`` through its C ABI; PyTorch is only the carrier of
device memory and streams.  NumPy inputs are staged host->device here and the result is read
back, so a call is end-to-end comparable with the reference's.

``params["images"]`` may be
  * the reference's own dict with a materialised ``atten_all*`` (complex64[I,K]) — streamed to
    the GPU in image batches (reference layout, 8 B/term of PCIe+HBM traffic);
  * the reference's "compact" dict (no ``atten_all*``, ``storage == "compact"``) — attenuation
    rebuilt in-kernel from the f32 angles (numerics of pb:140-186);
  * this package's dict (``storage == "fused"``, deism_b200.images) — attenuation fused
    in-kernel from A + room geometry with the complex64 rounding of the default reference
    storage reproduced (numerics of cd:2847-2928).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (ATTEN_C64, ATTEN_C128, ATTEN_FUSED_ANGLES, ATTEN_FUSED_ROOM, ShoeboxAtten,
                   check, lib)

# Device bytes a streamed batch of materialised attenuation may take (pb:34-35 uses 256 MB of
# host temporaries; HBM affords more).
ATTEN_BATCH_BYTES = 1 << 30


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("deism_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceArray:
    """NumPy-facing handle of an array that lives on the GPU.

    The reference's contract is "every intermediate in ``params`` is a NumPy array"; for the bulk
    arrays this engine produces on the device and consumes on the device again (the convex
    attenuation ``[K, I]``, 27 MB at C4) a round trip through host memory is the whole cost of the
    step.  A DeviceArray quacks like the NumPy array (``shape``, ``dtype``, ``len``, indexing,
    ``np.asarray``) and copies to the host the first time somebody actually looks at the values;
    ``to_device`` hands the kernels the tensor itself."""

    def __init__(self, tensor, dtype=None):
        self.tensor = tensor
        self._np_dtype = np.dtype(dtype) if dtype is not None else None
        self._host = None

    def _materialise(self):
        if self._host is None:
            a = self.tensor.cpu().numpy()
            self._host = a.astype(self._np_dtype) if self._np_dtype is not None and a.dtype != self._np_dtype else a
        return self._host

    shape = property(lambda self: tuple(self.tensor.shape))
    ndim = property(lambda self: self.tensor.dim())
    dtype = property(lambda self: self._np_dtype if self._np_dtype is not None else self._materialise().dtype)
    size = property(lambda self: self.tensor.numel())
    T = property(lambda self: self._materialise().T)
    real = property(lambda self: self._materialise().real)
    imag = property(lambda self: self._materialise().imag)

    def __len__(self):
        return int(self.tensor.shape[0])

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a.astype(dtype) if dtype is not None and np.dtype(dtype) != a.dtype else a

    def __getitem__(self, idx):
        return self._materialise()[idx]

    def astype(self, dtype, *args, **kwargs):
        return self._materialise().astype(dtype, *args, **kwargs)


def _broadcast_base(a):
    """(base, reps) when the NumPy array `a` is a broadcast view along its LAST axis (stride 0,
    e.g. the monopole coefficients of every image, which are one column repeated I times):
    base = a[..., :1] made contiguous; else None."""
    if isinstance(a, np.ndarray) and a.ndim >= 2 and a.shape[-1] > 1 and a.strides[-1] == 0:
        return np.ascontiguousarray(a[..., :1]), int(a.shape[-1])
    return None


class _NotCapturable(RuntimeError):
    """Raised inside a CUDA-graph capture when the call turns out not to be replayable."""


class _Staging:
    """Pinned staging buffers for pageable NumPy inputs.

    A reference caller holds plain NumPy arrays.  ``cudaMemcpyAsync`` from pageable memory goes
    through the driver's own bounce buffer on one thread and blocks the host (measured on the B200
    box: 1.76 ms for 33 MiB); copying into a pinned buffer first with torch's multi-threaded
    ``copy_`` and issuing a truly asynchronous H2D from there is twice as fast (0.81 ms) and leaves
    the host free to queue the kernels.  Buffers are recycled per power-of-two size class; a buffer
    is reused only after the event recorded behind its last H2D has completed."""

    MIN_BYTES = 1 << 18          # below this the pageable path is as fast
    MAX_BYTES = 1 << 28          # larger arrays (materialised atten slabs) go the pageable way
    PER_CLASS = 8
    CAP_BYTES = 1 << 30

    def __init__(self):
        self.pool = {}           # (device index, size class) -> list of [pinned uint8 tensor, event]
        self.total = 0

    def upload(self, a, dev):
        """a: C-contiguous NumPy array -> device tensor of the same dtype/shape, or None when the
        array should take the direct path."""
        nbytes = a.nbytes
        if nbytes < self.MIN_BYTES or nbytes > self.MAX_BYTES:
            return None
        src = torch.from_numpy(a)
        if src.is_pinned():
            return None
        size = 1 << (nbytes - 1).bit_length()
        slots = self.pool.setdefault((dev.index, size), [])
        slot = None
        for cand in slots:
            if cand[1] is None or cand[1].query():
                slot = cand
                break
        if slot is None:
            # never wait for a buffer: its H2D may sit behind a whole step of queued kernels
            if len(slots) >= self.PER_CLASS or self.total + size > self.CAP_BYTES:
                return None
            slot = [torch.empty(size, dtype=torch.uint8).pin_memory(), None]
            slots.append(slot)
            self.total += size
        # plain bytes: a vectorised, multi-threaded memcpy whatever the element type
        flat = torch.from_numpy(a.reshape(-1).view(np.uint8))
        slot[0][:nbytes].copy_(flat)
        t = slot[0][:nbytes].to(dev, non_blocking=True).view(src.dtype).reshape(src.shape)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        slot[1] = ev
        return t


_staging = _Staging()


def to_device(x, dtype, dev=None):
    """NumPy / torch (host or device) -> contiguous device tensor of ``dtype``."""
    dev = dev or _device()
    if isinstance(x, DeviceArray):
        x = x.tensor
    if isinstance(x, torch.Tensor):
        t = x
    else:
        bc = _broadcast_base(x)
        if bc is not None:
            # upload the one distinct column and repeat it on the device (HBM write, no PCIe)
            base = to_device(bc[0], dtype, dev)
            return base.expand(*base.shape[:-1], bc[1]).contiguous()
        a = np.asarray(x)
        if not a.flags["C_CONTIGUOUS"] or not a.flags["ALIGNED"]:
            a = np.ascontiguousarray(a)
        if not a.flags["WRITEABLE"]:
            a = a.copy()
        capturing = torch.cuda.is_current_stream_capturing()
        t = None if capturing else _staging.upload(a, dev)
        if t is None:
            t = torch.from_numpy(a)
            if capturing and not t.is_pinned():
                # a captured copy must re-read the CALLER's buffer at every replay: pageable
                # memory cannot be captured, and a staged copy would replay stale bytes
                raise _NotCapturable("pageable host array inside a graph capture")
    # host -> device in the source dtype, conversions on the device (no pageable temporaries)
    t = t.to(dev, non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _host_i64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64))


def _vec3(x):
    a = np.asarray(x, dtype=np.float64).reshape(3)
    return (C.c_double * 3)(float(a[0]), float(a[1]), float(a[2]))


def _c64_on_device(x, dev):
    """Directivity coefficients travel as complex64 — the reference's own pipeline stores them as
    complex64 (cd:1184,1236,1258,1329) and only up-casts inside the kernels (pb:556-557).  Any other
    input dtype (complex128, real) is converted, never refused; the shoebox ORG path keeps a
    caller's complex128 tables at full precision instead (``_coef_on_device``)."""
    if isinstance(x, (torch.Tensor, DeviceArray)):
        return to_device(x, torch.complex64, dev)
    a = x if isinstance(x, np.ndarray) else np.asarray(x)
    if a.dtype != np.complex64:
        a = a.astype(np.complex64)
    return to_device(a, torch.complex64, dev)


def _coef_on_device(x, dev):
    """ORG coefficient tables: (device tensor, is_complex128).  complex64 stays complex64; a
    caller's own complex128 tables go to the complex128 entry point when they are not
    complex64-representable (the reference up-casts whatever it is given, pb:556-557)."""
    if isinstance(x, DeviceArray):
        x = x.tensor
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.complex128:
            return to_device(x, torch.complex128, dev), True
        return to_device(x, torch.complex64, dev), False
    a = x if isinstance(x, np.ndarray) else np.asarray(x)
    if a.dtype == np.complex64:
        return to_device(a, torch.complex64, dev), False
    a128 = a.astype(np.complex128, copy=False)
    a64 = a128.astype(np.complex64)
    if np.array_equal(a64.astype(np.complex128), a128):
        return to_device(a64, torch.complex64, dev), False
    return to_device(a128, torch.complex128, dev), True


class _AttenSpec:
    """Builds struct deism_shoebox_atten and keeps the device tensors it points at alive."""

    def __init__(self, params, images, suffix, dev, lo=None, hi=None):
        self.keep = []
        s = ShoeboxAtten()
        s.ang_dep_flag = int(params["angDepFlag"])
        atten = images.get("atten_all" + suffix)
        storage = images.get("storage")
        sl = slice(lo, hi)
        if atten is not None:
            n_img = len(images["A" + suffix])
            a = atten
            if not isinstance(a, torch.Tensor):
                a = np.atleast_2d(np.asarray(a))           # pb:525-529
                if a.shape[0] != n_img and a.shape[1] == n_img:
                    a = a.T
            a = a[sl]
            is128 = (a.dtype == torch.complex128) if isinstance(a, torch.Tensor) \
                else (a.dtype == np.complex128)
            t = to_device(a, torch.complex128 if is128 else torch.complex64, dev)
            s.mode = ATTEN_C128 if is128 else ATTEN_C64
            s.d_atten = t.data_ptr()
            self.keep.append(t)
        else:
            if storage == "compact":
                s.mode = ATTEN_FUSED_ANGLES
            elif storage == "fused":
                s.mode = ATTEN_FUSED_ROOM
            else:
                raise KeyError(f"images has no 'atten_all{suffix}' and storage={storage!r}")
            A = to_device(images["A" + suffix][sl], torch.int32, dev)
            Z = params["impedance"]
            Z = to_device(Z if isinstance(Z, (torch.Tensor, DeviceArray)) else np.asarray(Z), torch.complex128, dev)
            if Z.dim() != 2 or Z.shape[0] != 6:
                raise ValueError("params['impedance'] must have shape [6, K]")
            s.d_A, s.d_Z_S = A.data_ptr(), Z.data_ptr()
            self.keep += [A, Z]
            if s.mode == ATTEN_FUSED_ANGLES:
                R = to_device(images["R_sI_r_all" + suffix][sl], torch.float32, dev)
                s.d_R_sI_r = R.data_ptr()
                self.keep.append(R)
            s.room_size = _vec3(params["roomSize"])
            s.pos_source = _vec3(params["posSource"])
            s.pos_receiver = _vec3(params["posReceiver"])
        self.struct = s

    def ref(self):
        return C.byref(self.struct)


def _image_batches(params, images, suffix, n_images, K):
    """Image ranges per kernel call: everything at once when the attenuation is fused, bounded
    slabs when a materialised atten[I,K] has to be streamed to the device (pb:471-522)."""
    atten = images.get("atten_all" + suffix)
    if atten is None or isinstance(atten, torch.Tensor) and atten.is_cuda:
        return [(0, n_images)]
    per_image = K * 16
    step = max(1, ATTEN_BATCH_BYTES // per_image)
    return [(s, min(s + step, n_images)) for s in range(0, n_images, step)]


def _run_lc(params, images, suffix, out, accumulate, dev):
    """pb:594-654 _run_numba_lc_matrix_in_batches."""
    Rs_all = images["R_s_rI_all" + suffix]
    n_images = int(Rs_all.shape[0])
    k = to_device(params["waveNumbers"], torch.float64, dev)
    K = int(k.shape[0])
    if n_images == 0:
        if not accumulate:
            out.zero_()
        return
    n_all, m_all = _host_i64(params["n_all"]), _host_i64(params["m_all"])
    v_all, u_all = _host_i64(params["v_all"]), _host_i64(params["u_all"])
    Cs = _c64_on_device(params["C_nm_s_vec"], dev)
    Cr = _c64_on_device(params["C_vu_r_vec"], dev)
    if Cs.shape != (K, n_all.shape[0]) or Cr.shape != (K, v_all.shape[0]):
        raise ValueError("C_nm_s_vec / C_vu_r_vec must be [K, (N+1)^2] / [K, (V+1)^2]")
    for lo, hi in _image_batches(params, images, suffix, n_images, K):
        Rs = to_device(Rs_all[lo:hi], torch.float32, dev)
        Rr = to_device(images["R_r_sI_all" + suffix][lo:hi], torch.float32, dev)
        spec = _AttenSpec(params, images, suffix, dev, lo, hi)
        rc = lib().deism_lc_shoebox(
            hi - lo, K, int(n_all.shape[0]), int(v_all.shape[0]),
            n_all.ctypes.data, m_all.ctypes.data, v_all.ctypes.data, u_all.ctypes.data,
            _ptr(Cs), _ptr(Cr), _ptr(Rs), _ptr(Rr), spec.ref(), _ptr(k), _ptr(out),
            1 if (accumulate or lo > 0) else 0, _stream_ptr())
        check(rc, "deism_lc_shoebox")
        del spec, Rs, Rr


def _wigner_host(Wigner):
    W1 = np.ascontiguousarray(np.asarray(Wigner["W_1_all"]).astype(np.complex64))
    W2 = np.ascontiguousarray(np.asarray(Wigner["W_2_all"]).astype(np.complex64))
    return W1, W2


class _OrgPlan:
    """deism_org_plan handle (coupling lists on the device), freed with the object."""

    def __init__(self, N, V, W1, W2):
        self.ptr = C.c_void_p(0)
        check(lib().deism_org_plan_create(N, V, W1.ctypes.data, W2.ctypes.data, C.byref(self.ptr)),
              "deism_org_plan_create")

    def __del__(self):
        try:
            if self.ptr:
                lib().deism_org_plan_destroy(self.ptr)
        except Exception:  # pragma: no cover  (interpreter shutdown)
            pass


_plan_cache = {}


def _org_plan(Wigner, N, V):
    """One plan per Wigner-table object (params['Wigner'] lives as long as the DEISM object); the
    tables are scanned on the host only the first time."""
    w2 = Wigner["W_2_all"]
    key = (id(w2), N, V, torch.cuda.current_device())
    hit = _plan_cache.get(key)
    if hit is not None and hit[1] is w2:
        return hit[0]
    W1, W2 = _wigner_host(Wigner)
    if W2.shape != (N + 1, V + 1, N + V + 1, 2 * N + 1, 2 * V + 1):
        raise ValueError("Wigner tables do not match sourceOrder/receiverOrder")
    plan = _OrgPlan(N, V, W1, W2)
    if len(_plan_cache) >= 8:
        _plan_cache.clear()
    _plan_cache[key] = (plan, w2)      # keeps w2 alive so that id() stays unique
    return plan


def _run_org(params, images, suffix, out, accumulate, dev):
    """pb:543-591 _run_numba_org_in_batches."""
    A_all = images["A" + suffix]
    n_images = len(A_all)
    k = to_device(params["waveNumbers"], torch.float64, dev)
    K = int(k.shape[0])
    if n_images == 0:
        if not accumulate:
            out.zero_()
        return
    N, V = int(params["sourceOrder"]), int(params["receiverOrder"])
    algo = int(params.get("b200OrgAlgo", 0))
    Cs, s128 = _coef_on_device(params["C_nm_s"], dev)
    Cr, r128 = _coef_on_device(params["C_vu_r"], dev)
    wide = (s128 or r128) and algo != 2
    if wide:
        Cs, Cr = Cs.to(torch.complex128), Cr.to(torch.complex128)
    else:
        Cs, Cr = Cs.to(torch.complex64), Cr.to(torch.complex64)
    if tuple(Cs.shape) != (K, N + 1, 2 * N + 1) or tuple(Cr.shape) != (K, V + 1, 2 * V + 1):
        raise ValueError("C_nm_s / C_vu_r must be [K, N+1, 2N+1] / [K, V+1, 2V+1]")
    plan = _org_plan(params["Wigner"], N, V)
    for lo, hi in _image_batches(params, images, suffix, n_images, K):
        A = to_device(A_all[lo:hi], torch.int32, dev)
        R = to_device(images["R_sI_r_all" + suffix][lo:hi], torch.float32, dev)
        spec = _AttenSpec(params, images, suffix, dev, lo, hi)
        if wide:
            rc = lib().deism_org_shoebox_c128(
                hi - lo, K, N, V, _ptr(Cs), _ptr(Cr), None, None, _ptr(A),
                _ptr(R), spec.ref(), _ptr(k), _ptr(out), 1 if (accumulate or lo > 0) else 0,
                plan.ptr, _stream_ptr())
        else:
            rc = lib().deism_org_shoebox(
                hi - lo, K, N, V, _ptr(Cs), _ptr(Cr), None, None, _ptr(A),
                _ptr(R), spec.ref(), _ptr(k), _ptr(out), 1 if (accumulate or lo > 0) else 0, algo,
                plan.ptr, _stream_ptr())
        check(rc, "deism_org_shoebox")
        del spec, A, R


def _run_DEISM_eager(params, dev):
    method = params["DEISM_method"]
    images = params["images"]
    K = int(np.asarray(params["waveNumbers"]).shape[0]) \
        if not isinstance(params["waveNumbers"], torch.Tensor) else int(params["waveNumbers"].shape[0])
    out = torch.zeros(K, dtype=torch.complex128, device=dev)
    if method == "ORG":
        _run_org(params, images, "", out, False, dev)
    elif method == "LC":
        _run_lc(params, images, "", out, False, dev)
    elif method == "MIX":
        # pb:725-777: early images -> ORG, late images -> LC, results added
        if isinstance(params.get("C_nm_s"), torch.Tensor) or len(images["A_late"]) == 0:
            _run_org(params, images, "_early", out, False, dev)
            _run_lc(params, images, "_late", out, True, dev)
        else:
            _run_mix_from_host(params, images, out, dev)
    else:
        raise ValueError(f"Unknown DEISM method: {method}")
    return out


# ----------------------------------------------------------------------------------------------
# CUDA-graph replay of repeated calls
# ----------------------------------------------------------------------------------------------
# A run_DEISM call is ~15-25 kernel launches and as many torch copies; for the sub-millisecond
# workloads (BASELINE C1, the monopole run of C2, an 8-GPU frequency slab) launch gaps and host
# dispatch are a third of the step.  When a call comes again with the SAME buffers — every array
# input a CUDA tensor or a PINNED host array at the same address and shape, the same scalars —
# the second call is captured into a CUDA graph (the pinned host->device copies become memcpy
# nodes: they re-read the host buffers at every replay) and later calls replay it.  Arrays in
# pageable memory are never captured (the driver stages them synchronously).  A graph bakes in the
# library's scratch pointers and constant tables: it is dropped when deism_state_generation()
# moves.  DEISM_B200_GRAPHS=0 (or backend.USE_GRAPHS = False) turns the mechanism off.
import os as _os

USE_GRAPHS = _os.environ.get("DEISM_B200_GRAPHS", "1") != "0"
_GRAPH_CACHE_MAX = 16
_graphs = {}        # key -> {"graph", "out", "gen", "n_kernels", "keep"}
_graph_seen = {}    # key -> generation at the last eager call with this key
_graph_never = set()    # keys whose capture failed: always eager


def _arr_key(x):
    """Identity of an array input for the graph cache, or None when it cannot be captured."""
    if isinstance(x, DeviceArray):
        x = x.tensor
    if isinstance(x, torch.Tensor):
        if x.is_cuda or x.is_pinned():
            return ("t", x.data_ptr(), tuple(x.shape), tuple(x.stride()), str(x.dtype), x.is_cuda)
        return None
    if isinstance(x, np.ndarray):
        if x.nbytes < 4096:          # small tables: by value
            return ("v", x.shape, str(x.dtype), x.tobytes())
        if not x.flags["C_CONTIGUOUS"] or _broadcast_base(x) is not None:
            return None
        if not torch.from_numpy(x if x.flags["WRITEABLE"] else x.copy()).is_pinned():
            return None
        return ("p", x.ctypes.data, x.shape, str(x.dtype))
    return None


def _graph_key(params, dev):
    images = params["images"]
    parts = [params["DEISM_method"], dev.index, int(params.get("angDepFlag", 0)),
             int(params.get("sourceOrder", 0)), int(params.get("receiverOrder", 0)),
             int(params.get("b200OrgAlgo", 0)), str(images.get("storage"))]
    for name in ("roomSize", "posSource", "posReceiver"):
        v = params.get(name)
        parts.append(None if v is None else np.asarray(v, dtype=np.float64).tobytes())
    for name in ("n_all", "m_all", "v_all", "u_all"):
        v = params.get(name)
        parts.append(None if v is None else np.asarray(v, dtype=np.int64).tobytes())
    w = params.get("Wigner")
    parts.append(None if w is None else id(w["W_2_all"]))
    for name in ("waveNumbers", "impedance", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
        v = params.get(name)
        if v is None:
            parts.append(None)
            continue
        k = _arr_key(v if isinstance(v, (torch.Tensor, DeviceArray)) else np.asarray(v))
        if k is None:
            return None
        parts.append(k)
    for name in sorted(images):
        v = images[name]
        if isinstance(v, (np.ndarray, torch.Tensor, DeviceArray)):
            k = _arr_key(v)
            if k is None:
                return None
            parts.append((name, k))
        else:
            parts.append((name, str(v)))
    return tuple(parts)


def _pin_small_tables(params):
    """Small NumPy tables are part of the graph key BY VALUE and may live in pageable memory, which a
    captured copy cannot read: the capture reads pinned copies (kept alive with the graph)."""
    keep = []

    def pinned(x):
        if isinstance(x, np.ndarray) and x.nbytes < 4096 and x.nbytes > 0:
            t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
            keep.append(t)
            return t.numpy()
        return x

    q = dict(params)
    for name in ("waveNumbers", "impedance", "C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec"):
        if name in q and q[name] is not None and not isinstance(q[name], (torch.Tensor, DeviceArray)):
            q[name] = pinned(np.asarray(q[name]))
    q["images"] = {k: pinned(v) for k, v in params["images"].items()}
    return q, keep


def _graphs_allowed():
    if not USE_GRAPHS:
        return False
    if torch.cuda.is_current_stream_capturing():
        return False
    return True


def run_DEISM_device(params, dev=None):
    """As run_DEISM but returns the complex128[K] result as a device tensor (no read-back)."""
    dev = dev or _device()
    if params["DEISM_method"] not in ("ORG", "LC", "MIX"):
        raise ValueError(f"Unknown DEISM method: {params['DEISM_method']}")
    key = _graph_key(params, dev) if _graphs_allowed() else None
    if key is None or key in _graph_never:
        return _run_DEISM_eager(params, dev)
    gen = int(lib().deism_state_generation())
    hit = _graphs.get(key)
    if hit is not None and hit["gen"] == gen:
        hit["graph"].replay()
        check(lib().deism_graph_replayed(hit["n_kernels"], _stream_ptr()), "deism_graph_replayed")
        return hit["out"].clone()
    if hit is not None:
        del _graphs[key]
    if _graph_seen.get(key) != gen:
        # first sight of these buffers (or the library state moved): run eagerly — this is also the
        # warm-up that sizes the scratch and uploads the constant tables
        out = _run_DEISM_eager(params, dev)
        _graph_seen[key] = int(lib().deism_state_generation())
        return out
    # second call with the same buffers and an unchanged library: capture it
    torch.cuda.synchronize(dev)
    captured, keep = _pin_small_tables(params)
    n0 = _lib.launch_count()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            out = _run_DEISM_eager(captured, dev)
    except Exception:
        # anything that cannot be captured: always eager for these buffers from now on
        if len(_graph_never) > 256:
            _graph_never.clear()
        _graph_never.add(key)
        torch.cuda.synchronize(dev)
        return _run_DEISM_eager(params, dev)
    n_kernels = _lib.launch_count() - n0
    if int(lib().deism_state_generation()) != gen:      # the capture itself moved the state: not replayable
        _graph_never.add(key)
        return _run_DEISM_eager(params, dev)
    if len(_graphs) >= _GRAPH_CACHE_MAX:
        _graphs.pop(next(iter(_graphs)))
    _graphs[key] = {"graph": graph, "out": out, "gen": gen, "n_kernels": int(n_kernels), "keep": keep}
    graph.replay()
    check(lib().deism_graph_replayed(int(n_kernels), _stream_ptr()), "deism_graph_replayed")
    return out.clone()


def clear_graphs():
    """Drop every captured graph (and the device memory their private pools hold)."""
    _graphs.clear()
    _graph_seen.clear()
    _graph_never.clear()


_copy_streams = {}


def _run_mix_from_host(params, images, out, dev):
    """MIX with host-resident inputs: the (long) LC part starts as soon as ITS inputs are on the
    device; the ORG part's inputs — the [K, N+1, 2N+1] coefficient tables are the largest arrays
    of the call — travel on a copy stream underneath it and the ORG kernels follow.  Same two
    partial sums, added in the other order."""
    p = dict(params)
    p["waveNumbers"] = to_device(params["waveNumbers"], torch.float64, dev)        # shared: once
    if params.get("impedance") is not None:
        Z = params["impedance"]
        p["impedance"] = to_device(Z if isinstance(Z, (torch.Tensor, DeviceArray)) else np.asarray(Z),
                                   torch.complex128, dev)
    main = torch.cuda.current_stream(dev)
    # fork point of the copy stream: BEFORE the LC work is queued, so that the ORG inputs travel
    # underneath the LC kernels (under graph capture this is also what joins the side stream to the
    # capture)
    fork = torch.cuda.Event()
    fork.record(main)
    _run_lc(p, images, "_late", out, False, dev)
    side = _copy_streams.get(dev.index)
    if side is None:
        side = _copy_streams[dev.index] = torch.cuda.Stream(dev)
    staged = []
    side.wait_event(fork)
    with torch.cuda.stream(side):
        for key in ("C_nm_s", "C_vu_r"):
            p[key] = _coef_on_device(params[key], dev)[0]
            staged.append(p[key])
        early = dict(images)
        for key, dt in (("A_early", torch.int32), ("R_sI_r_all_early", torch.float32)):
            if key in images:
                early[key] = to_device(images[key], dt, dev)
                staged.append(early[key])
        done = torch.cuda.Event()
        done.record(side)
    main.wait_event(done)
    for t in staged:
        t.record_stream(main)
    _run_org(p, early, "_early", out, True, dev)


def run_DEISM(params):
    """Drop-in for deism.parallel_backends.run_DEISM (pb:1572): complex128[K] NumPy array."""
    return run_DEISM_device(params).cpu().numpy()


# --------------------------------------------------------------------------------------
# convex (DEISM-ARG)
# --------------------------------------------------------------------------------------
def _arg_common(params, dev):
    images = params["images"]
    R = to_device(images["R_sI_r_all"], torch.float32, dev)
    if R.dim() != 2 or R.shape[0] != 3:
        raise ValueError("images['R_sI_r_all'] must be [3, I]")
    att = images["atten_all"]
    att = to_device(att, torch.complex64, dev)
    k = to_device(params["waveNumbers"], torch.float64, dev)
    return images, R, att, k


def _idx(sel):
    if sel is None:
        return None, 0, C.c_void_p(0)
    a = _host_i64(sel)
    return a, int(a.shape[0]), C.c_void_p(a.ctypes.data)


def _run_arg_lc(params, out, accumulate, dev, sel=None):
    images, R, att, k = _arg_common(params, dev)
    K, n_images = int(k.shape[0]), int(R.shape[1])
    n_all, m_all = _host_i64(params["n_all"]), _host_i64(params["m_all"])
    v_all, u_all = _host_i64(params["v_all"]), _host_i64(params["u_all"])
    Cs = _c64_on_device(params["C_nm_s_ARG_vec"], dev)
    Cr = _c64_on_device(params["C_vu_r_vec"], dev)
    if tuple(Cs.shape) != (K, n_all.shape[0], n_images):
        raise ValueError("C_nm_s_ARG_vec must be [K, (N+1)^2, I]")
    keep, n_sel, p_sel = _idx(sel)
    rc = lib().deism_lc_arg(n_images, K, int(n_all.shape[0]), int(v_all.shape[0]),
                            n_all.ctypes.data, m_all.ctypes.data, v_all.ctypes.data,
                            u_all.ctypes.data, _ptr(Cs), _ptr(Cr), _ptr(R), _ptr(att), _ptr(k),
                            p_sel, n_sel, _ptr(out), 1 if accumulate else 0, _stream_ptr())
    check(rc, "deism_lc_arg")
    del keep


def _run_arg_org(params, out, accumulate, dev, sel=None):
    images, R, att, k = _arg_common(params, dev)
    K, n_images = int(k.shape[0]), int(R.shape[1])
    N, V = int(params["sourceOrder"]), int(params["receiverOrder"])
    Cs = _c64_on_device(params["C_nm_s_ARG"], dev)
    Cr = _c64_on_device(params["C_vu_r"], dev)
    if tuple(Cs.shape) != (K, N + 1, 2 * N + 1, n_images):
        raise ValueError("C_nm_s_ARG must be [K, N+1, 2N+1, I]")
    plan = _org_plan(params["Wigner"], N, V)
    keep, n_sel, p_sel = _idx(sel)
    rc = lib().deism_org_arg(n_images, K, N, V, _ptr(Cs), _ptr(Cr), None, None,
                             _ptr(R), _ptr(att), _ptr(k), p_sel, n_sel, _ptr(out),
                             1 if accumulate else 0, plan.ptr, _stream_ptr())
    check(rc, "deism_org_arg")
    del keep


def run_DEISM_ARG_device(params, dev=None):
    dev = dev or _device()
    method = params["DEISM_method"]
    wn = params["waveNumbers"]
    K = int(wn.shape[0]) if isinstance(wn, torch.Tensor) else int(np.asarray(wn).shape[0])
    out = torch.zeros(K, dtype=torch.complex128, device=dev)
    if method == "LC":
        _run_arg_lc(params, out, False, dev)
    elif method == "ORG":
        _run_arg_org(params, out, False, dev)
    elif method == "MIX":
        im = params["images"]
        early, late = im["early_indices"], im["late_indices"]
        if len(early) > 0:
            _run_arg_org(params, out, False, dev, early)
        if len(late) > 0:
            _run_arg_lc(params, out, True, dev, late)
    else:
        raise ValueError(f"Unknown DEISM method: {method}")
    return out


def run_DEISM_ARG(params):
    """Drop-in for deism.parallel_backends.run_DEISM_ARG (pb:1577)."""
    return run_DEISM_ARG_device(params).cpu().numpy()
