"""Install the B200 engine behind an importable reference ``deism`` package.

    import deism_b200.patch; deism_b200.patch.install()      # before building DEISM objects

After that the UNMODIFIED reference class ``deism.core_deism.DEISM`` drives this engine:

  * ``deism.parallel_backends.run_DEISM`` / ``run_DEISM_ARG``  (imported lazily inside
    ``DEISM.run_DEISM``, core_deism.py:617-618)            -> deism_b200.backend
  * ``deism.core_deism.pre_calc_images_src_rec_optimized_nofs_v2_numba`` / ``merge_images``
    (core_deism.py:136-139, 161-162)                       -> deism_b200.images
  * ``deism.core_deism.pre_calc_Wigner`` (core_deism.py:490) -> deism_b200.wigner
  * ``deism.core_deism.cal_C_nm_s_new`` (called by init_source_directivities_ARG,
    core_deism.py:1702)                                    -> deism_b200.directivities
  * ``deism.core_deism_arg.Room_deism_cpp._init_room``'s engine (core_deism_arg.py:906-933)
                                                           -> deism_b200.convex.Room_deism

``uninstall()`` restores the originals.
"""
from __future__ import annotations

_saved = {}


def install(images: bool = True, wigner: bool = True, convex: bool = True, directivities: bool = True):
    import deism.core_deism as cd
    import deism.parallel_backends as pb

    from . import backend, images as gimages, wigner as gwigner

    def swap(mod, name, new):
        _saved.setdefault((mod.__name__, name), (mod, getattr(mod, name)))
        setattr(mod, name, new)

    swap(pb, "run_DEISM", backend.run_DEISM)
    swap(pb, "run_DEISM_ARG", backend.run_DEISM_ARG)
    if images:
        swap(cd, "pre_calc_images_src_rec_optimized_nofs_v2_numba",
             gimages.pre_calc_images_src_rec_optimized_nofs_v2_numba)
        swap(cd, "merge_images", gimages.merge_images)
    if wigner:
        swap(cd, "pre_calc_Wigner", gwigner.pre_calc_Wigner)
    if directivities:
        from . import directivities as gdir

        swap(cd, "cal_C_nm_s_new", gdir.cal_C_nm_s_new)
    if convex:
        import deism.core_deism_arg as cda

        from . import convex as gconvex

        def _init_room(self, *args):
            # same call as core_deism_arg.py:906-933, the engine object is the GPU one; walls
            # stay the reference's libroom Wall_deism objects (host-side construction)
            self.room_engine = gconvex.Room_deism([w.libroom_walls for w in self.walls], [], [],
                                                  self.c, self.ism_order)

        swap(cda.Room_deism_cpp, "_init_room", _init_room)
    return sorted(name for (_, name) in _saved)


def uninstall():
    for (_, name), (mod, old) in list(_saved.items()):
        setattr(mod, name, old)
    _saved.clear()
