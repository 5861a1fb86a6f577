"""deism_b200 — B200-native (sm_100a) engine for the run_DEISM hot path of audiolabs/DEISM.

Public surface mirrors the reference's plug points:

    backend.run_DEISM / run_DEISM_ARG          <- deism.parallel_backends.run_DEISM[_ARG]
    images.pre_calc_images_src_rec_optimized_nofs_v2_numba, merge_images
                                               <- deism.core_deism (same names)
    wigner.pre_calc_Wigner                     <- deism.core_deism.pre_calc_Wigner

Importing the package does not touch CUDA; the first call does.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def build(verbose: bool = False) -> str:
    """Compile libdeism_b200.so for sm_100a."""
    return _lib.build(verbose)
