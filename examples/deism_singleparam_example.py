"""DEISM with one set of parameters through the DEISM class — the reference's
examples/deism_singleparam_example.py, call for call, on the B200 engine.

    python examples/deism_singleparam_example.py [--device-resident] [--save]

RIR mode, shoebox 10 x 8 x 2.5 m, T60 = 1 s (-> 24 000 frequencies at 48 kHz), reflection order 30,
monopole source and receiver, DEISM-MIX.  Prints the wall time of every stage.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deism_b200.core_deism import DEISM  # noqa: E402


def main(device_resident=False, save=False, argv=()):
    import torch

    def lap(name, t0):
        torch.cuda.synchronize()
        print(f"  {name:28s} {(time.perf_counter() - t0) * 1e3:9.2f} ms")

    t_all = time.perf_counter()
    deism = DEISM("RIR", "shoebox", silent=True, argv=list(argv), device_resident=device_resident)
    room_dims = [10.0, 8.0, 2.5]
    deism.update_room(roomDimensions=np.array(room_dims))
    params_save = deism.params.copy()
    T60 = 1
    t0 = time.perf_counter()
    deism.update_wall_materials(datain=T60, datatype="reverberationTime")
    lap("update_wall_materials", t0)
    deism.params["sampleRate"] = 48000
    deism.params["reverberationTime"] = T60
    t0 = time.perf_counter()
    deism.update_freqs()
    lap("update_freqs", t0)
    t0 = time.perf_counter()
    deism.update_directivities()
    lap("update_directivities", t0)
    deism.params["maxReflOrder"] = 30
    t0 = time.perf_counter()
    deism.update_source_receiver()
    lap("update_source_receiver", t0)
    images = deism.params["images"]
    n_images = len(images["A_early"]) + len(images["A_late"])
    t0 = time.perf_counter()
    deism.run_DEISM(if_clean_up=True, if_shutdown_ray=True)
    lap("run_DEISM", t0)
    P = deism.params["RTF"]
    t0 = time.perf_counter()
    rir = deism.get_results()
    lap("get_results", t0)
    print(f"{n_images} images x {len(P)} frequencies, whole script "
          f"{(time.perf_counter() - t_all) * 1e3:.1f} ms (first call in a process: CUDA context + library load)")
    if save:
        save_path = "./outputs/RTFs"
        os.makedirs(save_path, exist_ok=True)
        np.savez(os.path.join(save_path, f"DEISM_shoebox_RTFs_{time.strftime('%Y%m%d_%H%M%S')}.npz"),
                 P_DEISM=P, RIR=rir, params=params_save)
    return P, rir


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--device-resident", action="store_true")
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--repeat", type=int, default=2, help="run the script body this many times (the second is warm)")
    a = ap.parse_args()
    for i in range(a.repeat):
        print(f"pass {i}:")
        main(a.device_resident, a.save and i == a.repeat - 1)
