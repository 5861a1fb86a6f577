"""The DEISM class under torchrun: one process per GPU, run_DEISM() shards the frequencies by itself.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        scripts/class_multi_gpu.py [c5|c2n5|c1]

Prints (rank 0) the warm run_DEISM wall time and the parity of the gathered RTF against the
reference's frozen result for the configuration.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_class  # noqa: E402
from deism_b200 import workloads  # noqa: E402

FIXTURE = {"c5": "c5_full_N5", "c2n5": "c2_full_N5", "c1": "c1_full", "c2": "c2_full_mono"}

if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "c5"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    os.environ["DEISM_B200_SHARDING"] = "frequency"      # multi-GPU is opt-in (core_deism.DEISM.run_DEISM)
    d, rows, n_img, K, _ = bench_class.run(case, None, True)           # cold pass (context, scratch)
    times = []
    for _ in range(5):
        d.update_source_receiver()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        d.run_DEISM()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    ref = workloads.load_fixture(FIXTURE[case])[1]["RTF"]
    rtf = d.params["RTF"]
    err = float(np.linalg.norm(rtf - ref) / np.linalg.norm(ref))
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if rank == 0:
        print(json.dumps({"case": case, "n_gpus": world, "images": n_img, "freqs": K,
                          "run_DEISM_wall_ms_median": round(float(np.median(times)), 3),
                          "parity_rel_l2_per_rank": errs}))
    dist.destroy_process_group()
