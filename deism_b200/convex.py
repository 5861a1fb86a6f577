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
    dev_arrays = getattr(room_engine, "device_arrays", None)
    if dev_arrays is not None:
        # this module's engine: the [K, I] attenuation was produced on the GPU and is consumed
        # there again; params sees a NumPy-facing handle that only visits the host if somebody
        # reads the values (backend.DeviceArray)
        from .backend import DeviceArray

        att = dev_arrays["attenuations"]
        if params["ifRemoveDirectPath"]:
            att = att[:, 1:]
        atten_all = DeviceArray(att.to(torch.complex64).contiguous(), np.complex64)
    else:
        atten_all = np.asarray(room_engine.attenuations, dtype=np.complex64)
        if params["ifRemoveDirectPath"]:
            atten_all = atten_all[:, 1:]
    if params["ifRemoveDirectPath"]:
        R_sI_r_all = R_sI_r_all[:, 1:]
        reflection_matrix = reflection_matrix[:, :, 1:]
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


# ----------------------------------------------------------------------------------------------
# Wall construction and the room wrapper (host side; a handful of walls per room)
# ----------------------------------------------------------------------------------------------
_F = np.float32
_LIBROOM_EPS = 1e-5          # common.hpp: libroom_eps


def _cross32(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], _F)


def _dot32(a, b):
    # fixed-size 3-vector reduction order of the reference's Eigen build: x + (y + z)
    p = (a * b).astype(_F)
    return _F(p[0] + _F(p[1] + p[2]))


def _area_2d(flat):
    """geometry.cpp area_2d_polygon: signed shoelace sum over the projected corners."""
    a = _F(0)
    n = flat.shape[1]
    for c1 in range(n):
        c2 = 0 if c1 == n - 1 else c1 + 1
        base = _F(0.5) * _F(flat[1, c2] + flat[1, c1])
        height = _F(flat[0, c2] - flat[0, c1])
        a = _F(a - _F(height * base))
    return a


class Wall_deism:
    """Host mirror of the reference's planar 3-D wall (``libroom_deism.Wall_deism``,
    wall.cpp:338-397 + orderPoints wall.cpp:105-143), in float32 like the Eigen original:

        normal   cross(c1 - c0, c2 - c0) normalised, flipped to point away from the room centroid
        origin   the first corner as given
        corners  re-ordered by angle around the wall centre
        basis / flat_corners   an orthonormal basis of the plane and the corners projected on it
                 (any such basis gives the same polygon test; the reference takes it from a
                 float32 Jacobi SVD, here it is the float64 SVD rounded to float32)
        reflection_matrix      I - 2 n n^T

    ``impedence_bands`` [K] is carried along (only its real part enters the attenuation,
    wall.cpp:428,633)."""

    dim = 3

    def __init__(self, points, centroid, impedence, name=""):
        corners = np.ascontiguousarray(np.asarray(points, dtype=np.float64).T.astype(_F))   # [3, n]
        if corners.shape[0] != 3:
            raise TypeError("The first dimension of points should be 2 or 3!")
        if corners.shape[1] < 3:
            raise RuntimeError("a wall needs at least three corners")
        centroid = np.asarray(centroid, dtype=np.float64).reshape(3).astype(_F)
        self.name = name
        self.impedence_bands = np.asarray(impedence)
        v1 = corners[:, 1] - corners[:, 0]
        v2 = corners[:, 2] - corners[:, 0]
        n = _cross32(v1, v2)
        n = (n / np.sqrt(_dot32(n, n))).astype(_F)
        self.origin = corners[:, 0].copy()
        if _dot32(n, centroid - self.origin) > 0:
            n = -n
        self.normal = n
        self.corners = self._order_points(corners)
        rel = (self.corners - self.origin[:, None]).astype(np.float64)
        U, S, _ = np.linalg.svd(rel, full_matrices=False)
        if S[2] > _LIBROOM_EPS:
            raise RuntimeError("The corners of the wall do not lie in a plane")           # wall.cpp:376-379
        basis = np.stack((U[:, 0], -U[:, 1]), axis=1).astype(_F)
        flat = (basis.T @ (self.corners - self.origin[:, None])).astype(_F)
        if _area_2d(flat) < 0:
            basis = np.ascontiguousarray(basis[:, ::-1])
            flat = np.ascontiguousarray(flat[::-1, :])
        self.basis, self.flat_corners = basis, flat
        self.reflection_matrix = (np.eye(3, dtype=_F) - _F(2) * np.outer(n, n).astype(_F)).astype(_F)

    def _order_points(self, pts):
        center = np.zeros(3, _F)
        for i in range(pts.shape[1]):           # sequential float32 sum, then / n (Eigen mean())
            center = (center + pts[:, i]).astype(_F)
        center = (center / _F(pts.shape[1])).astype(_F)
        ref = (pts[:, 0] - center).astype(_F)
        angles = np.zeros(pts.shape[1], _F)
        for i in range(pts.shape[1]):
            v = (pts[:, i] - center).astype(_F)
            c = _cross32(v, ref)
            ang = _F(np.arctan2(np.sqrt(_dot32(c, c)), _dot32(v, ref)))
            if _dot32(self.normal, c) <= 0:
                ang = _F(2 * np.pi - np.float64(ang))
            angles[i] = ang
        return np.ascontiguousarray(pts[:, np.argsort(angles, kind="stable")])


def _faces_by_normal(vertices):
    """Hull facets grouped by their rounded outward normal, in the reference's order
    (core_deism_arg.py:935-950, 1109-1131, 1142-1185 all walk ``list(set(normals))``)."""
    from scipy.spatial import ConvexHull

    hull = ConvexHull(vertices)
    rounded = np.round(hull.equations[:, :3], decimals=5)
    normals = [tuple(face) for face in rounded]
    groups = []
    for normal in list(set(normals)):
        groups.append([i for i in range(len(normals)) if normals[i] == normal])
    return hull, groups


def find_wall_centers(vertices, *choose_wall_centers):
    """core_deism_arg.py:1109-1139."""
    hull, groups = _faces_by_normal(np.asarray(vertices))
    centers = []
    for facets in groups:
        pts = np.unique(np.concatenate([hull.points[hull.simplices[i]] for i in facets]), axis=0)
        if not choose_wall_centers:
            centers.append(np.mean(pts, axis=0))
        else:
            for wc in choose_wall_centers:
                if np.linalg.norm(pts.mean(axis=0) - wc) < 0.0001:
                    centers.append(np.mean(pts, axis=0))
    return np.array(centers)


def convex_room_volume_and_areas(vertices):
    """core_deism_arg.py:1142-1185: hull volume and one area per wall (find_wall_centers order)."""
    vertices = np.asarray(vertices, dtype=float)
    if vertices.ndim != 2 or vertices.shape[1] != 3:
        raise ValueError("vertices must be (N, 3)")
    hull, groups = _faces_by_normal(vertices)
    areas = []
    for facets in groups:
        a = 0.0
        for i in facets:
            tri = hull.points[hull.simplices[i]]
            a += 0.5 * np.linalg.norm(np.cross(tri[1] - tri[0], tri[2] - tri[0]))
        areas.append(a)
    return float(hull.volume), np.array(areas, dtype=float)


def rotate_room_src_rec(params):
    """core_deism_arg.py:1265-1281."""
    from .directivities import rotation_matrix_ZXZ

    rr = np.asarray(params["roomRotation"]) * np.pi / 180
    R = rotation_matrix_ZXZ(rr[0], rr[1], rr[2])
    params["vertices"] = (R @ params["vertices"].T).T
    params["posSource"] = R @ params["posSource"]
    params["posReceiver"] = R @ params["posReceiver"]
    params["wallCenters"] = (R @ params["wallCenters"].T).T
    return params


class Room_deism_cpp:
    """Convex-room wrapper with the surface of the reference's ``Room_deism_cpp``
    (core_deism_arg.py:823-983): walls from the hull of ``params['vertices']``, each taking the
    impedance row of the nearest ``params['wallCenters']`` entry; ``update_images`` runs the GPU
    reflection tree (``room_engine`` = this module's Room_deism)."""

    def __init__(self, params, *choose_wall_centers):
        self.params = params
        self.points = np.asarray(params["vertices"])
        self.centroid = np.mean(self.points, axis=0)
        self.walls = []
        self.source = params["posSource"]
        self.microphones = [params["posReceiver"]]
        self.c = params["soundSpeed"]
        self.ism_order = params["maxReflOrder"]
        self.impedence = params["impedance"]
        self.freqs = params["freqs"]
        self.wall_centers = (params["wallCenters"] if "wallCenters" in params
                             else find_wall_centers(self.points))
        if len(self.freqs) != self.impedence.shape[1]:
            raise ValueError("The number of frequencies in the frequency array and the second "
                             "dimension of the impedance matrix are not the same")
        if params["convexRoom"]:
            self.generate_walls_convex(*choose_wall_centers)
        self.dim = self.points.shape[1]
        if self.dim != 3:
            raise TypeError("The room dimension should only be 2 or 3")     # 2-D rooms: not on the path
        self.room_engine = Room_deism(self.walls, [], [], self.c, self.ism_order)

    def generate_walls_convex(self, *choose_wall_centers):
        hull, groups = _faces_by_normal(self.points)
        for facets in groups:
            pts = np.unique(np.concatenate([hull.points[hull.simplices[i]] for i in facets]), axis=0)
            dist = np.linalg.norm(np.array(self.wall_centers) - np.mean(pts, axis=0), axis=1)
            if np.min(dist) > 0.001:
                raise ValueError("The face center is not close enough to any wall center")
            wall = Wall_deism(pts, self.centroid, self.impedence[np.argmin(dist)])
            if not choose_wall_centers:
                self.walls.append(wall)
            else:
                for wc in choose_wall_centers:
                    if np.linalg.norm(wall.corners.T.mean(axis=0) - wc) < 0.0001:
                        self.walls.append(wall)

    def update_images(self, source=None, receiver=None):
        if source is not None:
            self.source = source
        if receiver is not None:
            self.microphones[0] = receiver
        # the reference appends a microphone per call (core_deism_arg.py:895); the engine here
        # holds the current one
        self.room_engine.microphones = []
        self.room_engine.add_mic(np.asarray(self.microphones[0]))
        self.room_engine.n_bands = len(self.freqs)
        self.room_engine.image_source_model(np.asarray(self.source))
        self.sources = self.room_engine.sources
        self.gen_walls = self.room_engine.gen_walls
