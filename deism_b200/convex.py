"""Convex-room (DEISM-ARG) image enumeration on the GPU: the drop-in for the reference's native
engine ``libroom_deism.Room_deism`` and the Python glue around it.

    Room_deism            ~ libroom.cpp:58-127 binding of Room_deism<3> (room.hpp:124-406):
                            add_mic, n_bands, image_source_model + readonly sources, orders,
                            gen_walls, attenuations, reflection_matrix, visible_mics
    get_R_sI_to_r_from_room, get_ref_paths_ARG   ~ core_deism_arg.py:1188-1262

Walls are passed exactly as the reference passes them: objects exposing ``normal``, ``origin``,
``basis``, ``flat_corners``, ``reflection_matrix`` and ``impedence_bands`` (libroom
``Wall_deism`` objects or anything equivalent) — wall *construction* (ConvexHull face grouping,
SVD basis, core_deism_arg.py:935-983 / wall.cpp:338-397) stays with the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import ConvexWalls, check, lib
from .backend import _device, _ptr, _stream_ptr


def pack_walls(walls):
    """Flat float32 tables for struct deism_convex_walls."""
    W = len(walls)
    mc = max(np.asarray(w.flat_corners).shape[1] for w in walls)
    t = {"normal": np.zeros((W, 3), np.float32), "origin": np.zeros((W, 3), np.float32),
         "basis": np.zeros((W, 3, 2), np.float32), "flat_corners": np.zeros((W, 2, mc), np.float32),
         "n_corners": np.zeros(W, np.int32), "reflection": np.zeros((W, 3, 3), np.float32)}
    imp = []
    for i, w in enumerate(walls):
        t["normal"][i] = np.asarray(w.normal, np.float32).reshape(3)
        t["origin"][i] = np.asarray(w.origin, np.float32).reshape(3)
        t["basis"][i] = np.asarray(w.basis, np.float32)
        fc = np.asarray(w.flat_corners, np.float32)
        t["flat_corners"][i, :, :fc.shape[1]] = fc
        t["n_corners"][i] = fc.shape[1]
        t["reflection"][i] = np.asarray(w.reflection_matrix, np.float32)
        # only the real part of the impedance is used by the native engine (wall.cpp:428,633)
        imp.append(np.real(np.asarray(w.impedence_bands)).astype(np.float32).reshape(-1))
    n_bands = {len(x) for x in imp}
    if len(n_bands) != 1:
        # room.cpp:249-254
        raise RuntimeError("Error: All walls should have the same number of frequency bands")
    t["impedance"] = np.ascontiguousarray(np.stack(imp, axis=0))
    return t


def _walls_struct(t):
    s = ConvexWalls()
    s.n_walls = int(t["normal"].shape[0])
    s.max_corners = int(t["flat_corners"].shape[2])
    s.h_normal = t["normal"].ctypes.data
    s.h_origin = t["origin"].ctypes.data
    s.h_basis = t["basis"].ctypes.data
    s.h_flat_corners = t["flat_corners"].ctypes.data
    s.h_n_corners = t["n_corners"].ctypes.data
    s.h_reflection = t["reflection"].ctypes.data
    s.h_impedance = t["impedance"].ctypes.data
    return s


class Room_deism:
    """GPU reflection-tree engine with the attribute surface of libroom_deism.Room_deism."""

    dim = 3

    def __init__(self, walls, obstructing_walls=(), microphones=(), sound_speed=343.0,
                 ism_order=0, *unused_ray_tracing_args):
        if isinstance(walls, dict):      # pre-packed tables (pack_walls)
            self.walls, self._tables = [], walls
            n_walls = int(walls["normal"].shape[0])
        else:
            self.walls = list(walls)
            n_walls = len(self.walls)
            self._tables = pack_walls(self.walls) if n_walls > 3 else None
        if n_walls <= 3:
            raise RuntimeError("Error: The minimum number of walls is 4")      # room.cpp:262-264
        if len(obstructing_walls) > 0:
            raise NotImplementedError("only convex rooms are supported (no obstructing walls)")
        self.obstructing_walls = list(obstructing_walls)
        self.microphones = [np.asarray(m, np.float32).reshape(3) for m in microphones]
        self.sound_speed = float(sound_speed)
        self.ism_order = int(ism_order)
        self.n_bands = int(self._tables["impedance"].shape[1])
        # results stay on the GPU (self.device_arrays); the reference's readonly NumPy attributes
        # (libroom.cpp:101-127) are copied to the host the first time they are read
        self.device_arrays = None
        self._host = {"sources": np.zeros((3, 0), np.float32), "orders": np.zeros(0, np.int32),
                      "gen_walls": np.zeros(0, np.int32),
                      "attenuations": np.zeros((self.n_bands, 0), np.float32), "reflection_matrix": []}
        self.visible_mics = np.zeros((0, 0), bool)

    def _lazy(self, name):
        if name not in self._host:
            t = self.device_arrays[name].cpu().numpy()
            self._host[name] = list(t) if name == "reflection_matrix" else t
        return self._host[name]

    sources = property(lambda self: self._lazy("sources"))
    orders = property(lambda self: self._lazy("orders"))
    gen_walls = property(lambda self: self._lazy("gen_walls"))
    attenuations = property(lambda self: self._lazy("attenuations"))
    reflection_matrix = property(lambda self: self._lazy("reflection_matrix"))

    def add_mic(self, loc):
        """room.hpp:330-336.  (The reference's wrapper pushes a new microphone on every
        update_images call, core_deism_arg.py:895; a source is listed when ANY microphone sees
        it and its attenuation is the one of the LAST microphone that sees it.)"""
        self.microphones.append(np.asarray(loc, np.float32).reshape(3))

    def image_source_model(self, source_location):
        """room.cpp:271-296.  Returns the number of visible images."""
        if len(self.microphones) != 1:
            raise NotImplementedError("the GPU engine supports exactly one microphone "
                                      f"(got {len(self.microphones)})")
        dev = _device()
        t = self._tables
        if int(self.n_bands) != int(t["impedance"].shape[1]):
            raise RuntimeError("n_bands does not match the walls' impedance bands")
        ws = _walls_struct(t)
        src = np.ascontiguousarray(np.asarray(source_location, np.float32).reshape(3))
        mic = np.ascontiguousarray(self.microphones[0])
        n = C.c_int64(0)
        st = _stream_ptr()
        check(lib().deism_convex_images_count(C.byref(ws), src.ctypes.data, mic.ctypes.data,
                                              self.ism_order, C.byref(n), st),
              "deism_convex_images_count")
        I, K = int(n.value), int(self.n_bands)
        sources = torch.empty((3, I), dtype=torch.float32, device=dev)
        orders = torch.empty(I, dtype=torch.int32, device=dev)
        gen_walls = torch.empty(I, dtype=torch.int32, device=dev)
        atten = torch.empty((K, I), dtype=torch.float32, device=dev)
        refl = torch.empty((I, 3, 3), dtype=torch.float32, device=dev)
        check(lib().deism_convex_images_fill(C.byref(ws), K, _ptr(sources), _ptr(orders),
                                             _ptr(gen_walls), _ptr(atten), _ptr(refl), st),
              "deism_convex_images_fill")
        self.device_arrays = {"sources": sources, "orders": orders, "gen_walls": gen_walls,
                              "attenuations": atten, "reflection_matrix": refl}
        self._host = {}
        self.visible_mics = np.ones((1, I), dtype=bool)
        return I


def cart2sph(x, y, z):
    """utilities.py:95-102."""
    H_xy = np.hypot(x, y)
    r = np.hypot(H_xy, z)
    el = np.arctan2(z, H_xy)
    az = np.arctan2(y, x)
    return az, el, r


def get_R_sI_to_r_from_room(receiver, sources):
    """core_deism_arg.py:1188-1201: (phi, theta, r) of receiver - image, FP64."""
    R = np.asarray(receiver)[:, None] - sources
    phi, el, r = cart2sph(R[0, :], R[1, :], R[2, :])
    return np.asarray([phi, np.pi / 2 - el, r])


def get_ref_paths_ARG(params, room_engine):
    """core_deism_arg.py:1204-1262 with ``room_engine`` = a Room_deism (this module's or the
    reference's): fills params['images'] and params['reflection_matrix']."""
    reflection_matrix = np.array(room_engine.reflection_matrix, dtype=np.float32)
    reflection_matrix = np.moveaxis(reflection_matrix, 0, 2)
    R_sI_r_all = get_R_sI_to_r_from_room(np.asarray(params["posReceiver"]),
                                         room_engine.sources).astype(np.float32)
    atten_all = np.asarray(room_engine.attenuations, dtype=np.complex64)
    if params["ifRemoveDirectPath"]:
        R_sI_r_all = R_sI_r_all[:, 1:]
        reflection_matrix = reflection_matrix[:, :, 1:]
        atten_all = atten_all[:, 1:]
    images = {"R_sI_r_all": R_sI_r_all, "atten_all": atten_all}
    if params["DEISM_method"] == "MIX":
        # NB: computed from the un-sliced orders even when the direct path was removed — the
        # reference's behaviour (core_deism_arg.py:1222-1233), kept for parity
        early = np.where(room_engine.orders <= params["mixEarlyOrder"])[0]
        late = np.setdiff1d(np.arange(R_sI_r_all.shape[1]), early)
        images = {"early_indices": early, "late_indices": late, "R_sI_r_all": R_sI_r_all,
                  "atten_all": atten_all}
    params["images"] = images
    params["reflection_matrix"] = reflection_matrix
    return params


def get_ref_paths_ARG_device(params, room_engine):
    """get_ref_paths_ARG (core_deism_arg.py:1204-1262) without leaving the GPU: the image
    geometry, attenuations and reflection matrices the tree enumeration left in
    ``room_engine.device_arrays`` become the torch tensors ``run_DEISM_ARG`` consumes.  Only the
    reflection orders (4 bytes per image) visit the host, to split early / late images.
    ``params['reflection_matrix']`` keeps the engine's [I, 3, 3] layout (what
    ``directivities.cal_C_nm_s_ARG_vec_device`` takes)."""
    d = room_engine.device_arrays
    if d is None:
        raise RuntimeError("image_source_model() has not run on this engine")
    src = d["sources"].to(torch.float64)
    rcv = torch.as_tensor(np.asarray(params["posReceiver"], dtype=np.float64), device=src.device)
    R = rcv[:, None] - src
    hxy = torch.hypot(R[0], R[1])
    geom = torch.stack((torch.atan2(R[1], R[0]), np.pi / 2 - torch.atan2(R[2], hxy),
                        torch.hypot(hxy, R[2]))).to(torch.float32)
    atten = d["attenuations"].to(torch.complex64)
    refl = d["reflection_matrix"]
    if params["ifRemoveDirectPath"]:
        geom, atten, refl = geom[:, 1:], atten[:, 1:], refl[1:]
    images = {"R_sI_r_all": geom.contiguous(), "atten_all": atten.contiguous()}
    if params["DEISM_method"] == "MIX":
        orders = room_engine.orders          # un-sliced, as the reference does (cda:1222-1233)
        early = np.where(orders <= params["mixEarlyOrder"])[0]
        images["early_indices"] = early
        images["late_indices"] = np.setdiff1d(np.arange(geom.shape[1]), early)
    params["images"] = images
    params["reflection_matrix"] = refl.contiguous()
    return params
