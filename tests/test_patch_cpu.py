"""CPU: deism_b200.patch wires this engine behind the UNMODIFIED reference class (only where the
reference tree is importable, i.e. in the build container)."""
import pytest

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available() or ref_import.libroom_path() is None,
                                reason="reference tree / oracle/_ref not available")


def test_install_swaps_plug_points_and_routes_calls():
    import torch

    ref_import.import_reference()
    import deism.core_deism as cd
    import deism.parallel_backends as pb

    import deism_b200.patch as patch
    from deism_b200 import backend, images, wigner

    orig = (pb.run_DEISM, cd.pre_calc_Wigner)
    names = patch.install()
    try:
        assert {"run_DEISM", "run_DEISM_ARG", "pre_calc_Wigner", "merge_images", "_init_room", "cal_C_nm_s_new",
                "pre_calc_images_src_rec_optimized_nofs_v2_numba"} <= set(names)
        assert pb.run_DEISM is backend.run_DEISM and pb.run_DEISM_ARG is backend.run_DEISM_ARG
        assert cd.pre_calc_Wigner is wigner.pre_calc_Wigner
        assert cd.pre_calc_images_src_rec_optimized_nofs_v2_numba is \
            images.pre_calc_images_src_rec_optimized_nofs_v2_numba
        d = ref_import.make_deism("RTF", "shoebox")
        d.params["endFreq"] = 40
        d.params["maxReflOrder"] = 2
        d.update_wall_materials()
        d.update_freqs()
        d.update_directivities()          # native Wigner tables through the reference class
        assert d.params["Wigner"]["W_2_all"].shape == (1, 1, 1, 1, 1)
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                d.update_source_receiver()   # the reference class now calls the GPU generator
    finally:
        patch.uninstall()
    assert (pb.run_DEISM, cd.pre_calc_Wigner) == orig
