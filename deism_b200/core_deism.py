"""``DEISM`` — the user-facing workflow class, standalone on the B200 engine.

Same constructor, methods, ``params`` keys and errors as the reference's
``deism.core_deism.DEISM`` (core_deism.py:44-769), so a script written against the reference
runs unchanged after ``from deism_b200.core_deism import DEISM``:

    d = DEISM("RIR", "shoebox")            # YAML defaults + command line   (cd:49-86)
    d.update_room(...)                     # cd:176-255
    d.update_wall_materials(...)           # cd:257-405   -> materials.update_wall_materials
    d.update_freqs()                       # cd:407-465   -> materials.update_freqs (+ convex walls)
    d.update_directivities()               # cd:467-492   -> directivities.*, wigner.pre_calc_Wigner
    d.update_source_receiver(...)          # cd:88-162    -> GPU image lattice / GPU reflection tree
    d.run_DEISM()                          # cd:608-628   -> backend.run_DEISM[_ARG] (CUDA kernels)
    p = d.get_results()                    # cd:675-769   -> rir.get_results (cuFFT)

``DEISM(..., device_resident=True)`` (an addition) keeps the bulk intermediates — image lists,
impedance on the frequency grid, directivity coefficients, convex attenuations and reflection
matrices — as torch CUDA tensors inside ``params`` instead of NumPy arrays, so that consecutive
stages hand each other device memory; scalars, the frequency grid and ``params["RTF"]`` stay NumPy.

Every stage lands in the C-ABI library; there is no CPU fallback.  ``shoeboxImageCalcVersion``
values other than the default name the reference's legacy CPU generators and raise here.
"""
from __future__ import annotations

import numpy as np

from . import backend, directivities, images as gimages, materials, rir, wigner
from .data_loader import ConflictChecks, cmdArgsToDict, printDict


class DEISM:
    """User-facing workflow class for shoebox and convex DEISM simulations."""

    def __init__(self, mode, roomtype, silent=False, argv=None, device_resident=False):
        self.device_resident = bool(device_resident)
        self.mode = mode
        self.roomtype = roomtype
        self.silent = silent
        self.auto_update = True
        self.track_updated_where = True
        self._argv = argv
        self.init_params()
        self.update_room()

    # ---------------------------------------------------------------- cd:66-86
    def init_params(self):
        params, cmdArgs = cmdArgsToDict(self.mode, self.roomtype, self._argv)
        params["roomType"] = self.roomtype
        params["mode"] = self.mode
        if cmdArgs.quiet or self.silent:
            params["silentMode"] = 1
        printDict(params)
        if cmdArgs.run or all(value in [None, False] for value in vars(cmdArgs).values()):
            ConflictChecks.check_all_conflicts(params)
            if self.track_updated_where:
                params["track_updated_where"] = True
                params["updated_where"] = {}
                for key in list(params.keys()):
                    params["updated_where"][key] = ["init_params"]
            self.params = params

    def _update_where_tracking(self, parameter_name, function_name):
        if not self.track_updated_where:
            return
        self.params["updated_where"].setdefault(parameter_name, []).append(function_name)

    def print_tracking_summary(self):
        print(self.params["updated_where"])

    # ---------------------------------------------------------------- cd:88-162
    def update_source_receiver(self, source=None, receiver=None):
        p = self.params
        p["posSource"] = p["posSource"] if source is None else source
        p["posReceiver"] = p["posReceiver"] if receiver is None else receiver
        ConflictChecks.distance_spheres_checks(p)
        ConflictChecks.distance_boundaries_checks(p)
        if self.roomtype == "shoebox":
            version = str(p.get("shoeboxImageCalcVersion", "v2-numba")).lower().replace("_", "-")
            if version not in ("numba", "v2-numba", "numba-v2", "b200", "gpu"):
                raise NotImplementedError(f"shoeboxImageCalcVersion={version!r} names one of the reference's "
                                          "legacy CPU generators; this engine enumerates images on the GPU")
            if self.device_resident and not p.get("shoeboxCompactImages") and not p.get("shoeboxMaterializeAtten"):
                dev, n_e, _ = gimages.shoebox_images_device(p)
                p["images"] = {"storage": "fused"}
                for key, t in dev.items():       # early and late are one buffer, back to back
                    p["images"][key + "_early"], p["images"][key + "_late"] = t[:n_e], t[n_e:]
            else:
                p["images"] = gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba(p)
        elif self.roomtype == "convex":
            from .convex import get_ref_paths_ARG, get_ref_paths_ARG_device

            self.room_convex.update_images(np.asarray(p["posSource"]), np.asarray(p["posReceiver"]))
            paths = get_ref_paths_ARG_device if self.device_resident else get_ref_paths_ARG
            self.params = paths(p, self.room_convex.room_engine)
        for key in ("posSource", "posReceiver", "images"):
            self._update_where_tracking(key, "update_source_receiver")
        if self.roomtype == "shoebox" and self.params["DEISM_method"] in ("ORG", "LC"):
            self.params["images"] = gimages.merge_images(self.params["images"])

    # ---------------------------------------------------------------- cd:176-255
    def update_room(self, roomDimensions=None, wallCenters=None, roomVolumn=None, roomAreas=None):
        p = self.params
        if self.roomtype == "shoebox":
            if roomDimensions is not None:
                roomDimensions = np.asarray(roomDimensions)
                if roomDimensions.ndim != 1 or roomDimensions.shape[0] != 3:
                    raise ValueError("roomDimensions must be a size-3 1D array")
                p["roomSize"] = roomDimensions
            length, width, height = p["roomSize"][0], p["roomSize"][1], p["roomSize"][2]
            p["roomVolumn"] = length * width * height
            # x1, x2, y1, y2, z1, z2
            p["roomAreas"] = np.array([width * height, width * height, length * height, length * height,
                                       length * width, length * width])
            for key in ("roomSize", "roomVolumn", "roomAreas"):
                self._update_where_tracking(key, "update_room")
        elif self.roomtype == "convex":
            if roomDimensions is not None:
                p["vertices"] = roomDimensions
                self._update_where_tracking("vertices", "update_room")
            if wallCenters is not None:
                p["wallCenters"] = wallCenters
                self._update_where_tracking("wallCenters", "update_room")
            if (roomVolumn is None or roomAreas is None) and p.get("vertices") is not None:
                from .convex import convex_room_volume_and_areas

                vol, areas = convex_room_volume_and_areas(p["vertices"])
                if roomVolumn is None:
                    p["roomVolumn"] = vol
                    self._update_where_tracking("roomVolumn", "update_room")
                if roomAreas is None:
                    p["roomAreas"] = areas
                    self._update_where_tracking("roomAreas", "update_room")
            if roomVolumn is not None:
                p["roomVolumn"] = roomVolumn
                self._update_where_tracking("roomVolumn", "update_room")
            if roomAreas is not None:
                p["roomAreas"] = roomAreas
                self._update_where_tracking("roomAreas", "update_room")
        else:
            raise ValueError("The room type is not supported")

    # ---------------------------------------------------------------- cd:257-405
    def update_wall_materials(self, datain=None, freqs_bands=None, datatype=None):
        p = self.params
        if datain is not None and datatype is not None:
            if datatype == "reverberationTime" and self.roomtype == "convex":
                raise ValueError("T60 input is not supported for convex room yet, please use impedance or "
                                 "absorption coefficients instead")
            if datatype not in ("impedance", "absorption", "reverberationTime"):
                raise ValueError("The parameter type is not supported")
            p["givenMaterials"] = [datatype]
            p[datatype] = datain
            ConflictChecks.wall_material_checks(p)
        materials.update_wall_materials(p, datain, freqs_bands, datatype, roomtype=self.roomtype)
        keys = ["freqs_bands", "impedance", "absorption", "reverberationTime"]
        if self.roomtype == "shoebox":
            keys += ["n1", "n2", "n3"]
        for key in keys:
            self._update_where_tracking(key, "update_wall_materials")
        if not p["silentMode"]:
            kind = p["givenMaterials"][0]
            print(f"[Data] {kind} -> impedance {p['impedance']}, absorption {p['absorption']}, "
                  f"reverberation time {p['reverberationTime']}\n")

    # ---------------------------------------------------------------- cd:407-465
    def update_freqs(self):
        p = materials.update_freqs(self.params, device_impedance=self.device_resident and self.roomtype == "shoebox")
        for key in ("freqs", "waveNumbers"):
            self._update_where_tracking(key, "update_freqs")
        if p["ifReceiverNormalize"] == 1:
            self._update_where_tracking("pointSrcStrength", "update_freqs")
        self._update_where_tracking("impedance", "interpolate_materials")
        if self.roomtype == "convex":
            from .convex import Room_deism_cpp

            self.room_convex = Room_deism_cpp(p)
            self._update_where_tracking("room_convex", "update_freqs")

    def interpolate_materials(self, freqs_bands, datatype):
        """cd:494-529."""
        if datatype not in ("impedance", "absorption", "reverberationTime"):
            return
        self.params[datatype] = materials.interpolate_functions(self.params[datatype], freqs_bands,
                                                                self.params["freqs"])
        self._update_where_tracking(datatype, "interpolate_materials")

    # ---------------------------------------------------------------- cd:467-492
    def update_directivities(self):
        p = self.params
        if self.roomtype == "shoebox":
            p = directivities.init_receiver_directivities(p)
            p = directivities.init_source_directivities(p)
        elif self.roomtype == "convex":
            p = directivities.init_receiver_directivities_ARG(p)
            p = directivities.init_source_directivities_ARG(p)
        if p["DEISM_method"] in ("LC", "MIX"):
            p = directivities.vectorize_C_nm_s(p) if self.roomtype == "shoebox" \
                else directivities.vectorize_C_nm_s_ARG(p)
            p = directivities.vectorize_C_vu_r(p)
        if p["DEISM_method"] in ("ORG", "MIX"):
            p = wigner.pre_calc_Wigner(p)
        if self.device_resident:
            import torch

            for key in ("C_nm_s", "C_vu_r", "C_nm_s_vec", "C_vu_r_vec", "C_nm_s_ARG", "C_nm_s_ARG_vec"):
                if key in p and not isinstance(p[key], torch.Tensor):
                    p[key] = backend.to_device(np.asarray(p[key], dtype=np.complex64), torch.complex64)
        self.params = p

    # ---------------------------------------------------------------- cd:608-628
    def run_DEISM(self, if_clean_up: bool = True, if_shutdown_ray: bool = True):
        """One process per GPU.  Multi-GPU is opt-in: with ``params["b200Sharding"]`` set to
        "frequency" (contiguous frequency slabs + one all_gather) or "image" (image slabs of a merged
        ORG / LC list + one all_reduce) — or the environment variable DEISM_B200_SHARDING — and
        ``torch.distributed`` initialised with more than one rank (e.g. torchrun, NCCL backend, each
        rank after ``torch.cuda.set_device(LOCAL_RANK)``), every rank computes its slab and the
        collective leaves the full ``params["RTF"]`` on all of them.  The default is "off": a script
        that uses torchrun for data parallelism (one room per rank) keeps computing whole, independent
        simulations.  Before sharding, the ranks verify that they hold the same job."""
        import os

        import torch.distributed as dist

        compute = backend.run_DEISM_device if self.roomtype == "shoebox" else backend.run_DEISM_ARG_device
        mode = self.params.get("b200Sharding") or os.environ.get("DEISM_B200_SHARDING", "off")
        if mode not in ("off", "frequency", "image"):
            raise ValueError(f"b200Sharding must be 'off', 'frequency' or 'image', got {mode!r}")
        if mode != "off" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from . import distributed

            distributed.check_same_job(self.params)
            rtf = distributed.run_sharded(self.params, compute, mode=mode)
        else:
            rtf = compute(self.params)
        self.params["RTF"] = rtf.cpu().numpy()
        if if_clean_up:
            # (the reference also forces a gc.collect() here, cd:625-628; the image arrays carry no
            # reference cycles and a forced collection costs 30-90 ms of host time per call)
            self.params.pop("images", None)

    def run_DEISM_ray(self, if_clean_up: bool = True, if_shutdown_ray: bool = True):
        raise NotImplementedError("the Ray backend is the reference's legacy CPU fan-out; use run_DEISM()")

    # ---------------------------------------------------------------- cd:531-606, 675-769
    def create_bandpass_window(self, freqs, f_low, f_high, transition_width_low=None,
                               transition_width_high=None):
        return rir.create_bandpass_window(freqs, f_low, f_high, transition_width_low, transition_width_high)

    def apply_highpass_filter(self, data, fs, fcut=30.0, zero_phase=True):
        """cd:531-541."""
        return rir.highpass_RIR(data, fs, fcut, zero_phase=zero_phase)

    def get_results(self, highpass_filter: bool = False, bandpass_window: bool = True,
                    cut_freq: float = 30.0, zero_phase: bool = True):
        if highpass_filter and bandpass_window:
            raise ValueError("Only one of highpass_filter or bandpass_window can be True")
        if self.mode == "RTF":
            return self.params["RTF"]
        p = rir.get_results(self.params, mode="RIR", highpass_filter=highpass_filter,
                            bandpass_window=bandpass_window, cut_freq=cut_freq, zero_phase=zero_phase)
        if bandpass_window:
            # the reference leaves the windowed spectrum in params["RTF"] (cd:744)
            fs = self.params["sampleRate"]
            f_low, f_high = 150, int(fs / 2 * 0.7)
            self.params["RTF"] = self.params["RTF"] * rir.create_bandpass_window(
                self.params["freqs"], f_low, f_high, int(f_low * 0.3), int(f_high * 0.15))
        return p
