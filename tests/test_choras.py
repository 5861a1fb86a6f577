"""The CHORAS job runner (deism_b200.choras.deism_method ~ examples/DeismInterface.py:113-260) against
the reference's class driven through the same call sequence (fixtures j_choras_*; band-wise absorption
-> frequency-dependent impedance, i.e. the per-term wall-factor path of the kernels)."""
import json

import numpy as np
import pytest

from conftest import Golden, rel_l2
from deism_b200 import choras, materials
from deism_b200.core_deism import DEISM


def _job(tmp_path, kind):
    g = Golden(f"j_choras_{kind}")
    path = tmp_path / f"{kind}.json"
    path.write_text(str(g.z["job_json"]))
    return g, str(path), json.loads(str(g.z["overrides_json"]))


@pytest.mark.parametrize("kind", ["shoebox", "convex"])
def test_job_host_stages_equal_reference(tmp_path, kind):
    """Up to the first GPU stage: room type, materials (absorption -> impedance -> T60), PCHIP grid."""
    g, path, overrides = _job(tmp_path, kind)
    job = materials.read_choras_json(path)
    assert choras.is_shoebox_corners(job["vertices"]) == (kind == "shoebox")
    d = DEISM("RIR", kind, silent=True)
    d.params.update(overrides)
    dims = job["vertices"].max(axis=0) - job["vertices"].min(axis=0) if kind == "shoebox" else job["vertices"]
    d.update_room(dims, job["wallCenters"], job["roomVolumn"], job["roomAreas"])
    d.update_wall_materials(job["absorption"], job["freqs_bands"], "absorption")
    d.update_freqs()
    assert np.array_equal(np.asarray(d.params["reverberationTime"]), g.z["reverberationTime"])
    assert np.array_equal(np.asarray(d.params["impedance"]), g.z["impedance"])
    assert len(d.params["freqs"]) == len(g.z["RTF"])


def test_cancel_flag_and_missing_path(tmp_path):
    g, path, _ = _job(tmp_path, "shoebox")
    job = json.loads(open(path).read())
    job["should_cancel"] = True
    open(path, "w").write(json.dumps(job))
    assert choras.should_cancel(path) is True
    assert choras.deism_method(path) is None                 # nothing runs, nothing is written
    assert json.loads(open(path).read())["results"][0]["responses"][0]["receiverResults"] == []
    assert choras.deism_method(None) is None
    assert choras.should_cancel(str(tmp_path / "absent.json")) is False


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["shoebox", "convex"])
@pytest.mark.parametrize("device_resident", [True, False])
def test_job_end_to_end(tmp_path, kind, device_resident):
    g, path, overrides = _job(tmp_path, kind)
    rir = choras.deism_method(path, overrides=overrides, device_resident=device_resident)
    assert rir.shape == g.z["RIR"].shape and rel_l2(rir, g.z["RIR"]) < 1e-8
    stored = json.loads(open(path).read())["results"][0]["responses"][0]["receiverResults"]
    assert np.array_equal(np.asarray(stored), rir)
