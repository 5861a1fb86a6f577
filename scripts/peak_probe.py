import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deism_b200._lib import lib, check
for kind, name in ((0, "DFMA all 8 warps"), (1, "DMMA all 8 warps"), (2, "mixed: 4 warps DMMA + 4 warps DFMA"), (3, "DMMA fed from smem, 16 warps/SM (production occupancy)"),
                   (4, "DMMA fed from smem, 4 warps/SM (one per sub-partition)"), (5, "DMMA fed from smem, 8 warps/SM (two per sub-partition)")):
    tf, ms = C.c_double(0), C.c_double(0)
    check(lib().deism_fp64_peak(kind, 20000, C.byref(tf), C.byref(ms)))
    print(f"{name}: {tf.value:.2f} TFLOP/s, {ms.value:.3f} ms")
