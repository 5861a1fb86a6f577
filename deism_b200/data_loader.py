"""YAML configuration + command line -> the ``params`` dict of a DEISM object.

Host-side mirror of the reference's loader for the drop-in class (same keys, same defaults, same
errors), so that a ``deism_b200.core_deism.DEISM`` starts from exactly the state the reference's
class starts from:

    ConflictChecks                         ~ data_loader.py:10-175
    readYaml                               ~ data_loader.py:386-425
    parseCmdArgs                           ~ data_loader.py:428-610 (shoebox), 613-770 (convex)
    loadSingleParam / cmdArgsToDict        ~ data_loader.py:773-939, 1057-1096

The four default configurations (``configSingleParam_{RTF,RIR,ARG_RTF,ARG_RIR}.yml``) ship in
``deism_b200/configs``; like the reference, a file of the same name under ``./examples`` or the
working directory wins.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class ConflictChecks:
    """data_loader.py:10-175."""

    @staticmethod
    def directivity_checks(params):
        if params["sourceType"] == "monopole":
            params["sourceOrder"] = 0
        if params["receiverType"] == "monopole":
            params["receiverOrder"] = 0
            params["ifReceiverNormalize"] = 0
        if params["sourceType"] != "monopole" and params["sourceOrder"] == 0:
            print("[Warning] Spherical harmonic order is set to 0 for source, but source type is not monopole! \n")
        if params["receiverType"] != "monopole" and params["receiverOrder"] == 0:
            print("[Warning] Spherical harmonic order is set to 0 for receiver, but receiver type is not monopole! \n")
        if params["receiverType"] != "monopole" and params["ifReceiverNormalize"] == 0:
            print("[Warning] Receiver normalization is set to 0 for non-monopole receiver, make sure you know "
                  "what you are doing! \n")

    @staticmethod
    def distance_spheres_checks(params):
        dist = np.linalg.norm(np.asarray(params["posSource"]) - np.asarray(params["posReceiver"]))
        for which, label in (("Source", "source"), ("Receiver", "receiver")):
            if params[label + "Type"] == "monopole":
                continue
            if params["radius" + which] is None:
                raise ValueError(f"{which} radius is not set for non-monopole {label}")
            if dist < params["radius" + which]:
                print(f"[Warning] Distance between source and receiver is smaller than the {label} radius! \n")
        if params["sourceType"] != "monopole" and params["receiverType"] != "monopole":
            if dist < params["radiusSource"] + params["radiusReceiver"]:
                print("[Warning] Distance between source and receiver is smaller than the sum of the radii "
                      "of the transparent spheres! \n")

    @staticmethod
    def wall_material_checks(params):
        given = params.get("givenMaterials", [])
        if len(given) > 1:
            raise ValueError("The user can only define one of the following three parameters: impedance, "
                             "absorption coefficient, or reverberation time. The following parameters are "
                             f"defined: {', '.join(given)}")

        def shape_ok(name, label):
            if name not in params:
                return
            value = params[name]
            if isinstance(value, (int, float)):
                return
            shoebox = params.get("roomType", "shoebox") == "shoebox"
            arr = np.asarray(value)
            if np.issubdtype(arr.dtype, np.number):
                if arr.ndim == 0:
                    return
                walls_ok = arr.shape[0] == 6 if shoebox else arr.shape[0] > 0
                if arr.ndim in (1, 2) and walls_ok:
                    return
            raise ValueError(f"{label} data must be a scalar, a list/1D array with one value per wall, or a 2D "
                             "array of shape (walls, frequency bands)"
                             + (" with 6 walls for a shoebox room" if shoebox else "") + f", got {value!r}")

        if "impedance" in given:
            shape_ok("impedance", "Impedance")
        elif "absorption" in given:
            shape_ok("absorption", "Absorption coefficient")
        elif "reverberationTime" in given:
            rt = params.get("reverberationTime")
            if rt is not None and not isinstance(rt, (int, float)):
                arr = np.asarray(rt)
                if not (np.issubdtype(arr.dtype, np.number) and arr.size >= 1 and arr.ndim <= 2):
                    raise ValueError("Reverberation time data must be a scalar or a small numeric array, "
                                     f"got {rt!r}")

    @staticmethod
    def distance_boundaries_checks(params):
        if params["roomType"] != "shoebox":
            return
        room = np.asarray(params["roomSize"], dtype=float)
        for label, pos in (("Source", params["posSource"]), ("Receiver", params["posReceiver"])):
            pos = np.asarray(pos, dtype=float)
            if np.any(pos <= 0) or np.any(pos >= room):
                print(f"[Warning] {label} is on the boundaries or outside of the room!")

    @staticmethod
    def check_all_conflicts(params):
        ConflictChecks.directivity_checks(params)
        ConflictChecks.distance_spheres_checks(params)
        ConflictChecks.distance_boundaries_checks(params)
        ConflictChecks.wall_material_checks(params)


def readYaml(filePath):
    """data_loader.py:386-425: ``examples/<name>`` under the working directory first, then the
    path as given; (addition) then the packaged default of the same name."""
    candidates = [os.path.join("examples", filePath), filePath, os.path.join(CONFIG_DIR, os.path.basename(filePath))]
    for path in candidates:
        if os.path.isfile(path):
            break
    else:
        if os.path.exists(filePath):
            raise FileNotFoundError(f"{filePath} is not a file!")
        raise FileNotFoundError(f"{filePath} doesn't exist!")
    with open(path, "r") as stream:
        configs = yaml.safe_load(stream)
    for key, value in configs.items():
        if isinstance(value, list):
            configs[key] = np.array(value)
    return configs


def parseCmdArgs(mode="RTF", roomtype="shoebox", argv=None):
    """The reference's command-line overrides (data_loader.py:428-610, 613-770), same flags.  Like
    the reference nothing is read from ``sys.argv`` under pytest; here the same holds whenever the
    running program is not a DEISM script (``argv`` can be passed explicitly)."""
    p = argparse.ArgumentParser(description="DEISM parameters (override the YAML configuration)")
    p.add_argument("-c", metavar="c0", type=float, help="speed of sound (m/s)")
    p.add_argument("-rho", metavar="rho0", type=float, help="air density")
    if roomtype == "shoebox":
        p.add_argument("-room", metavar=("Lx", "Ly", "Lz"), nargs=3, type=float, help="room size (m)")
    p.add_argument("-nro", type=int, help="maximum reflection order")
    p.add_argument("-zs", nargs=6, type=float, help="wall impedances x1 x2 y1 y2 z1 z2")
    p.add_argument("-absp", nargs=6, type=float, help="wall absorption coefficients")
    if roomtype == "shoebox":
        p.add_argument("-t60", nargs=1, type=float, help="reverberation time (s)")
        p.add_argument("-adrc", type=int, help="angle-dependent reflection coefficients (0/1)")
    p.add_argument("-xs", nargs=3, type=float, help="source position (m)")
    p.add_argument("-xr", nargs=3, type=float, help="receiver position (m)")
    p.add_argument("-so", nargs=3, type=float, help="source orientation (deg, Z-X-Z)")
    p.add_argument("-ro", nargs=3, type=float, help="receiver orientation (deg, Z-X-Z)")
    if mode == "RTF":
        p.add_argument("-fmin", type=float, help="start frequency (Hz)")
        p.add_argument("-fstep", type=float, help="frequency step (Hz)")
        p.add_argument("-fmax", type=float, help="stop frequency (Hz)")
    elif mode == "RIR":
        p.add_argument("-fs", type=int, help="sampling rate")
        p.add_argument("-K", type=float, help="over sampling factor")
        p.add_argument("-rirlen", type=float, help="RIR length (s)")
    p.add_argument("-srctype", type=str, help="source type")
    p.add_argument("-rectype", type=str, help="receiver type")
    p.add_argument("-srcorder", type=int, help="source spherical-harmonic order")
    p.add_argument("-srcr0", type=float, help="radius of the source's transparent sphere (m)")
    p.add_argument("-recorder", type=int, help="receiver spherical-harmonic order")
    p.add_argument("-recr0", type=float, help="radius of the receiver's transparent sphere (m)")
    if roomtype == "convex":
        p.add_argument("-ifconvex", type=int, help="room is convex (0/1)")
    p.add_argument("-ird", type=int, help="remove the direct path (0/1)")
    p.add_argument("-method", type=str, help="ORG, LC or MIX")
    p.add_argument("-meo", type=int, help="highest reflection order handled by ORG in MIX")
    p.add_argument("-npi", type=int, help="images per batch")
    p.add_argument("-irn", type=int, help="normalise the receiver directivity (0/1)")
    p.add_argument("-q", type=float, help="point-source flow strength")
    p.add_argument("--run", action="store_true", help="run after loading")
    p.add_argument("--quiet", action="store_true", help="no console output")
    if argv is None:
        argv = [] if "pytest" in sys.modules else sys.argv[1:]
    args = p.parse_args(args=list(argv))
    for name in ("room", "t60", "adrc", "ifconvex", "fmin", "fstep", "fmax", "fs", "K", "rirlen"):
        if not hasattr(args, name):
            setattr(args, name, None)
    return args


def loadSingleParam(configs, args, mode="RTF", roomtype="shoebox"):
    """data_loader.py:773-939: YAML values, overridden by whatever the command line gave."""
    if not isinstance(args, argparse.Namespace):
        raise TypeError(f"args type:{type(args)} is not of type argparse.Namespace.")
    if not isinstance(configs, dict):
        raise TypeError("configs is not of type dict")
    params = {}
    params["soundSpeed"] = args.c or float(configs["Environment"]["soundSpeed"])
    params["airDensity"] = args.rho or float(configs["Environment"]["airDensity"])
    dims = configs["Dimensions"]
    if roomtype == "shoebox":
        params["roomSize"] = np.array(args.room or list(dims.values()))
    elif roomtype == "convex":
        params["vertices"] = np.array(list(dims["vertices"]))
        params["wallCenters"] = np.array(list(dims["wallCenters"]))
        params["ifRotateRoom"] = dims["ifRotateRoom"]
        params["roomRotation"] = np.array(list(dims["roomRotation"].values()))
        params["convexRoom"] = args.ifconvex if args.ifconvex is not None else dims["ifConvexRoom"]
    else:
        raise ValueError(f"Invalid room type: {roomtype}, must be 'shoebox' or 'convex'")
    refl = configs["Reflections"]
    if args.nro is not None or refl["maxReflectionOrder"] is not None:
        params["maxReflOrder"] = args.nro or refl["maxReflectionOrder"]
    else:
        params["maxReflOrder"] = -1
    given = []
    if args.zs or "impedance" in refl:
        params["impedance"] = args.zs or refl["impedance"]
        given.append("impedance")
    absorption_key = "absorption" if "absorption" in refl else "absorpCoefficient"
    if args.absp or absorption_key in refl:
        params["absorption"] = args.absp or refl[absorption_key]
        given.append("absorption")
    if args.t60 or "reverberationTime" in refl:
        params["reverberationTime"] = args.t60 or refl["reverberationTime"]
        given.append("reverberationTime")
    if len(given) > 1:
        raise ValueError("The user can only define one of the following three parameters: -zs or -absp or "
                         "-t60. The following parameters are defined: " + ", ".join(given))
    params["givenMaterials"] = given
    if args.adrc is not None:
        params["angDepFlag"] = args.adrc
    elif "angleDependentFlag" in refl:
        params["angDepFlag"] = refl["angleDependentFlag"]
    params["posSource"] = np.array(args.xs or list(configs["Positions"]["source"].values()))
    params["posReceiver"] = np.array(args.xr or list(configs["Positions"]["receiver"].values()))
    params["orientSource"] = np.array(args.so or list(configs["Orientations"]["source"].values()))
    params["orientReceiver"] = np.array(args.ro or list(configs["Orientations"]["receiver"].values()))
    if mode == "RTF":
        params["startFreq"] = args.fmin or configs["Frequencies"]["startFrequency"]
        params["freqStep"] = args.fstep or configs["Frequencies"]["frequencyStep"]
        params["endFreq"] = args.fmax or configs["Frequencies"]["endFrequency"]
    elif mode == "RIR":
        params["sampleRate"] = args.fs or configs["Signal"]["samplingRate"]
        params["RIRLength"] = args.rirlen or configs["Signal"]["RIRLength"]
        params["overSamplingFactor"] = args.K or configs["Signal"]["overSamplingFactor"]
    params["mode"] = mode
    params["sourceOrder"] = args.srcorder or configs["MaxSphDirectivityOrder"]["sourceOrder"]
    params["receiverOrder"] = args.recorder or configs["MaxSphDirectivityOrder"]["receiverOrder"]
    params["radiusSource"] = args.srcr0 or configs["Radius"]["source"]
    params["radiusReceiver"] = args.recr0 or configs["Radius"]["receiver"]
    params["sourceType"] = args.srctype or configs["Directivities"]["source"]
    params["receiverType"] = args.rectype or configs["Directivities"]["receiver"]
    specs = configs["DEISM_specs"]
    params["ifRemoveDirectPath"] = args.ird if args.ird is not None else specs["ifRemoveDirect"]
    params["DEISM_method"] = args.method or specs["Method"]
    if args.meo or "mixEarlyOrder" in specs:
        params["mixEarlyOrder"] = args.meo or specs["mixEarlyOrder"]
    params["numParaImages"] = args.npi or specs["numParaImages"]
    params["ifReceiverNormalize"] = args.irn if args.irn is not None else specs["ifReceiverNormalize"]
    params["qFlowStrength"] = args.q or specs["QFlowStrength"]
    params["silentMode"] = args.quiet or configs["SilentMode"]
    return params


_YML = {("shoebox", "RTF"): "configSingleParam_RTF.yml", ("shoebox", "RIR"): "configSingleParam_RIR.yml",
        ("convex", "RTF"): "configSingleParam_ARG_RTF.yml", ("convex", "RIR"): "configSingleParam_ARG_RIR.yml"}


def cmdArgsToDict(mode="RTF", roomtype="shoebox", argv=None):
    """data_loader.py:1057-1096."""
    if roomtype not in ("shoebox", "convex"):
        raise ValueError(f"Invalid room: {roomtype}, must be 'shoebox' or 'convex'")
    if mode not in ("RTF", "RIR"):
        raise ValueError(f"Invalid mode: {mode}")
    configs = readYaml(_YML[(roomtype, mode)])
    args = parseCmdArgs(mode, roomtype, argv)
    return loadSingleParam(configs, args, mode, roomtype), args


def printDict(params):
    """data_loader.py:942-1054 prints a formatted table; here: one line per scalar entry."""
    if params.get("silentMode"):
        return
    print("[Parameters]")
    for key, val in params.items():
        if np.ndim(val) == 0 or np.size(val) <= 12:
            print(f"  {key}: {val}")


def load_directive_pressure(silentMode, src_or_rec, name):
    """data_loader.py:1099-1145: the sampled directional pressure field of a source / receiver profile,
    ``<data>/sampled_directivity/<source|receiver>/<name>.mat`` with the reference's lookup rule
    (``./examples/data`` when an ``examples/data`` directory is visible from the CWD, else ``data``)
    and its errors: (freqs_mesh flattened, Psh [K, n_dir], Dir_all [n_dir, 2], r0)."""
    import os

    import scipy.io as sio

    path = "data/sampled_directivity"
    if os.path.exists("examples/data"):
        path = "./examples/" + path
    data_location = "{}/{}/{}.mat".format(path, src_or_rec, name)
    try:
        with open(data_location, "rb") as file:
            FEM_data = sio.loadmat(file)
    except FileNotFoundError:
        raise FileNotFoundError(f"File {data_location} not found.")
    except Exception as e:
        raise RuntimeError(f"An error occurred while loading the file: {e}")
    freqs_all = FEM_data["freqs_mesh"].flatten()
    pressure = FEM_data["Psh"]
    Dir_all = FEM_data["Dir_all"]
    r0 = FEM_data["r0"]
    if not silentMode:
        print("Load directivity data from {}, ".format(path), end="")
        print(f"sampled on {len(Dir_all)} directions and {len(freqs_all)} frequencies. ", end="")
    return freqs_all, pressure, Dir_all, r0
