"""CHORAS job runner: the product-level entry point of the reference
(``examples/DeismInterface.py:113-260`` ``deism_method``) on the B200 engine.

    deism_method(json_file_path)        # reads the job, runs DEISM("RIR", room), writes the RIR back

The job file is the reference's wire format (``examples/exampleInput_Deism.json``): absorption
coefficients per named surface and band, source / receiver positions, and a ``geometry[0]`` block
(vertices, wall centres, wall areas, volume).  The reference fills that block from a Gmsh ``.geo``
file first (``room_check.sync_room_geometry``); Gmsh is not part of this engine, so the block must
be present (it is in every file the reference has processed once).  The room type is decided as
``room_check.is_shoebox_corners`` decides it (two distinct values per axis -> shoebox, else convex).
"""
from __future__ import annotations

import json

import numpy as np

from . import materials
from .core_deism import DEISM


def is_shoebox_corners(corners, tol=1e-6):
    """room_check.py:79-83."""
    corners = np.asarray(corners, dtype=float)
    return all(len(np.unique(np.round(corners[:, ax] / tol) * tol)) == 2 for ax in range(3))


def should_cancel(json_file_path):
    """DeismInterface.py:151-165: the user's cancel flag, False when it cannot be read."""
    try:
        with open(json_file_path, "r", encoding="utf-8") as fh:
            return bool(json.load(fh).get("should_cancel", False))
    except Exception:
        return False


def deism_method(json_file_path=None, overrides=None, device_resident=True, silent=True):
    """Run one CHORAS job.  ``overrides`` (an addition) is written into ``params`` before the
    update_* calls, e.g. ``{"maxReflOrder": 8}``.  Returns the impulse response that was stored
    under ``results[0].responses[0].receiverResults``, or None when there was nothing to do."""
    if json_file_path is None:
        print("No JSON file path provided. Exiting.")
        return None
    if should_cancel(json_file_path):
        return None
    job = materials.read_choras_json(json_file_path)
    vertices = job["vertices"]
    room = "shoebox" if is_shoebox_corners(vertices) else "convex"
    d = DEISM("RIR", room, silent=silent, argv=[], device_resident=device_resident)
    d.params.update(overrides or {})
    if room == "shoebox":
        # update_room takes [length, width, height] for a shoebox, not the mesh corners
        # (DeismInterface.py:216-221)
        room_dimensions = vertices.max(axis=0) - vertices.min(axis=0)
    else:
        room_dimensions = vertices
    d.update_room(room_dimensions, job["wallCenters"], job["roomVolumn"], job["roomAreas"])
    d.update_wall_materials(job["absorption"], job["freqs_bands"], "absorption")
    d.update_freqs()
    d.params["posSource"] = job["posSource"]
    d.params["posReceiver"] = job["posReceiver"]
    d.update_source_receiver()
    d.update_directivities()
    d.run_DEISM(if_clean_up=True)
    rir = d.get_results()
    materials.write_choras_result(json_file_path, rir)
    return rir
