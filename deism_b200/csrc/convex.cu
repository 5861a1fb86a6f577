// convex.cu — convex-room reflection tree (placeholder).
#include "common.cuh"
using namespace deism;
extern "C" int deism_convex_images_count(const deism_convex_walls *, const float *, const float *, int,
                                         int64_t *, void *) {
    return set_error(DEISM_ERR_INVALID, "deism_convex_images_count: not implemented yet");
}
extern "C" int deism_convex_images_fill(const deism_convex_walls *, int64_t, float *, int32_t *, int32_t *,
                                        float *, float *, void *) {
    return set_error(DEISM_ERR_INVALID, "deism_convex_images_fill: not implemented yet");
}
