"""CPU-only: the C-ABI library loads, exports every symbol include/deism_b200.h declares, the
host-side pieces (Wigner tables) are exact, and compute entry points fail loudly without a GPU."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, Golden


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "deism_b200.h")).read()
    return sorted(set(re.findall(r"DEISM_API\s+[\w\s\*]+?\b(deism_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from deism_b200 import _lib

    handle = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(handle, name), name
    assert set(declared) == set(_lib.EXPORTED)
    assert handle.deism_b200_version() >= 100


def test_no_cpu_fallback():
    import torch
    from deism_b200 import backend

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        backend.run_DEISM({"DEISM_method": "LC", "images": {}, "waveNumbers": [1.0]})
    from deism_b200._lib import lib

    rc = lib().deism_fp64_peak(0, 10, None, None)
    assert rc == 2 and b"no CPU fallback" in lib().deism_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deism_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # no import, dlopen or path use of the checker anywhere in the product
                bad = re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle[/.]_ref|oracle\.oracle|"
                                r"import_module\([\"']oracle", src, re.M)
                assert bad is None, (os.path.join(dirpath, f), bad.group(0))


@pytest.mark.parametrize("key", ["0_0", "3_2", "2_4", "5_5", "10_10"])
def test_native_wigner_tables_equal_sympy(key):
    """Hashes of the sympy tables (core_deism.py:1842-1913 run via the oracle) vs native."""
    from deism_b200 import wigner

    z = np.load(os.path.join(GOLDEN, "wigner_hashes.npz"))
    N, V = map(int, key.split("_"))
    W = wigner.wigner_tables(N, V)
    assert W["W_1_all"].dtype == np.complex64
    assert hashlib.sha256(W["W_1_all"].tobytes()).hexdigest() == str(z[key + "/sha_W1"])
    assert hashlib.sha256(W["W_2_all"].tobytes()).hexdigest() == str(z[key + "/sha_W2"])
    assert int(np.count_nonzero(W["W_2_all"])) == int(z[key + "/nnz_W2"])


def test_native_wigner_matches_reference_fixture_tables():
    from deism_b200 import wigner

    g = Golden("s3_shoebox_rtf_org_o4_N3V2")
    ref = g.group("wigner")
    p = {"sourceOrder": 3, "receiverOrder": 2, "track_updated_where": True, "updated_where": {}}
    p = wigner.pre_calc_Wigner(p)
    assert np.array_equal(p["Wigner"]["W_1_all"], ref["W_1_all"])
    assert np.array_equal(p["Wigner"]["W_2_all"], ref["W_2_all"])
    assert p["updated_where"]["Wigner"] == ["pre_calc_Wigner"]


def test_wigner_orthogonality():
    """sum_{m,u} (n v l; -m u m-u)^2 = 1 for every valid (n,v,l) (3j completeness)."""
    from deism_b200 import wigner

    N = V = 6
    W2 = wigner.wigner_tables(N, V)["W_2_all"].real.astype(np.float64)
    for n in range(N + 1):
        for v in range(V + 1):
            for l in range(abs(n - v), n + v + 1):
                s = np.sum(W2[n, v, l] ** 2)
                assert abs(s - 1.0) < 1e-6


def test_library_carries_the_fingerprint_of_its_sources():
    """The .so is git-ignored: a checkout or an edit can leave an old binary next to new sources.
    The build bakes the hash of csrc/*.cu, common.cuh and the header into the library; the loader
    compares it with the sources in the tree and rebuilds (or refuses) a stale one."""
    from deism_b200 import _lib

    want = _lib.source_hash()
    assert want is not None and len(want) == 16
    assert _lib.lib().deism_b200_source_hash().decode() == want
    assert not _lib._stale(_lib.LIB_PATH)
    # a binary built from other sources is recognised without being dlopen'ed
    import shutil
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        other = os.path.join(tmp, "libother.so")
        shutil.copy(_lib.LIB_PATH, other)
        blob = open(other, "rb").read()
        i = blob.find(b"DEISM_SRC_HASH=")
        assert i > 0
        with open(other, "wb") as fh:
            fh.write(blob[:i + 15] + b"0123456789abcdef" + blob[i + 31:])
        assert _lib._stale(other)


def test_generation_counter_and_graph_hooks_exist_without_a_gpu():
    from deism_b200 import _lib

    handle = _lib.lib()
    g = int(handle.deism_state_generation())
    assert g >= 1
    assert handle.deism_graph_replayed(0, None) == 2       # no device: DEISM_ERR_CUDA, loudly
    assert b"no CPU fallback" in handle.deism_last_error()


def _simulate_handout(sizes, n_ftiles, slots, fixed=0.5):
    """The block scheduler's hand-out as api.cu models it: CTAs in launch order (frequency tile
    fastest) to whichever slot frees up first; returns the makespan in tiles."""
    import heapq

    h = [0.0] * slots
    heapq.heapify(h)
    for c in sizes:
        for _ in range(n_ftiles):
            heapq.heappush(h, heapq.heappop(h) + c + fixed)
    return max(h)


def test_chunk_plan_invariants_and_benefit():
    """deism_chunk_plan (host logic, no GPU): a shape gets its plan when it is seen the second time (the
    simulations cost more host time than one call saves); whatever the shape, a plan covers every image
    tile exactly once, in at most 96 non-empty chunks of non-increasing size below the caller's bound,
    and finishes earlier than the uniform cut under the modelled hand-out; shapes it should not touch
    (tiny lists, too many chunks) come back as 'keep the uniform cut'."""
    from deism_b200 import _lib

    h = _lib.lib()
    rng = np.random.default_rng(5)
    shapes = [(1386, 250, 148), (1386, 32, 148), (1386, 63, 148), (1386, 125, 148), (181, 125, 148),
              (326, 63, 148), (322, 188, 148), (1386, 500, 296), (16, 1, 148), (40, 3, 148), (5000, 7, 148)]
    shapes += [(int(rng.integers(16, 4000)), int(rng.integers(1, 600)), int(rng.choice([148, 296, 132])))
               for _ in range(40)]
    used, seen = 0, set()
    for n_tiles, n_ftiles, slots in shapes:
        for waves in (4, 8, 16):
            n = max(1, min(n_tiles, slots * waves // n_ftiles))
            tpc = -(-n_tiles // n)
            n = -(-n_tiles // tpc)
            lo = (C.c_int32 * 97)()
            cnt = C.c_int32(0)
            bound = 1 << 20
            key = (n_tiles, n_ftiles, slots, n, tpc)
            if key in seen:
                continue
            seen.add(key)
            first = h.deism_chunk_plan(n_tiles, n_ftiles, slots, n, tpc, bound, lo, C.byref(cnt))
            assert first == 0                 # a shape is only recorded the first time it is seen
            got = h.deism_chunk_plan(n_tiles, n_ftiles, slots, n, tpc, bound, lo, C.byref(cnt))
            assert got == h.deism_chunk_plan(n_tiles, n_ftiles, slots, n, tpc, bound, lo, C.byref(cnt))
            if n > 96:
                assert got == 0
            if not got:
                continue
            used += 1
            b = np.array(lo[: cnt.value + 1])
            sizes = np.diff(b)
            assert 1 <= cnt.value <= 96 and b[0] == 0 and b[-1] == n_tiles, (n_tiles, n_ftiles, slots, b)
            assert np.all(sizes >= 1) and np.all(np.diff(sizes) <= 0) and sizes.max() <= bound
            uniform = [tpc] * (n - 1) + [n_tiles - tpc * (n - 1)]
            assert _simulate_handout(sizes.tolist(), n_ftiles, slots) < _simulate_handout(uniform, n_ftiles, slots)
    assert used >= 10        # the C5 / C3 / C2 shapes above do take a shrinking schedule
    lo = (C.c_int32 * 97)()
    cnt = C.c_int32(0)
    for _ in range(3):
        assert h.deism_chunk_plan(8, 4, 148, 2, 4, 1 << 20, lo, C.byref(cnt)) == 0      # tiny image list
    h.deism_chunk_plan(1386, 250, 148, 13, 107, 50, lo, C.byref(cnt))
    assert h.deism_chunk_plan(1386, 250, 148, 13, 107, 50, lo, C.byref(cnt)) == 0 or \
        np.diff(np.array(lo[: cnt.value + 1])).max() <= 50                                # chunk bound respected
