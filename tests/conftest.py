"""pytest configuration: `gpu` marker, repo root on sys.path, golden-fixture loader."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """A golden case frozen by oracle/make_golden.py from the reference's own run."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)

    def group(self, prefix):
        pre = prefix + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def params(self):
        """Rebuild the `params` dict the reference's run_DEISM(params) received."""
        p = {}
        for k, v in self.group("params").items():
            p[k] = v.item() if v.ndim == 0 else v
        if "impedance" not in p and "impedance_col0" in p:
            # frequency-flat impedance is stored as one column (oracle/make_golden.py)
            K = int(np.asarray(p["waveNumbers"]).shape[0])
            p["impedance"] = np.ascontiguousarray(
                np.repeat(np.asarray(p["impedance_col0"]).reshape(-1, 1), K, axis=1))
        p.pop("impedance_col0", None)
        p["DEISM_method"] = str(self.z["meta/method"])
        p["silentMode"] = 1
        img = {}
        for k, v in self.group("images").items():
            img[k] = str(v) if k == "storage" else v
        if img:
            p["images"] = img
        w = self.group("wigner")
        if w:
            p["Wigner"] = w
        return p

    @property
    def rtf(self):
        return self.z["RTF"]


def golden_names(prefix=""):
    if not os.path.isdir(GOLDEN):
        return []
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(prefix))


@pytest.fixture
def golden():
    return Golden


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def rel_max(a, b):
    """Worst per-frequency relative error max_k |a_k - b_k| / |b_k| — stricter than rel_l2 when a
    few frequencies dominate the norm (ORG terms reach 1e20 at small kr)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b) / np.abs(b)))
