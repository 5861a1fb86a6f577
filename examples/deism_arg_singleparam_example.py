"""DEISM-ARG (convex room) with one set of parameters through the DEISM class — the flow of the
reference's examples/deism_arg_singleparam_example.py on the B200 engine.

    python examples/deism_arg_singleparam_example.py [--order 10] [--device-resident]

The room of configSingleParam_ARG_RIR.yml (tilted ceiling), monopole source and receiver, DEISM-MIX.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deism_b200.core_deism import DEISM  # noqa: E402


def main(order=10, device_resident=False):
    import torch

    t_all = time.perf_counter()
    deism = DEISM("RIR", "convex", silent=True, argv=[], device_resident=device_resident)
    deism.params["maxReflOrder"] = order
    deism.update_wall_materials()
    deism.update_freqs()
    deism.update_source_receiver()
    deism.update_directivities()
    n_images = int(deism.params["images"]["R_sI_r_all"].shape[1])
    deism.run_DEISM()
    rir = deism.get_results()
    torch.cuda.synchronize()
    print(f"order {order}: {n_images} images x {len(deism.params['freqs'])} frequencies, "
          f"T60 {deism.params['reverberationTime']:.4f} s, {(time.perf_counter() - t_all) * 1e3:.1f} ms; "
          f"max |RIR| {np.max(np.abs(rir)):.4e}")
    return rir


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=10)
    ap.add_argument("--device-resident", action="store_true")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    for _ in range(a.repeat):
        main(a.order, a.device_resident)
