// arg.cu — convex DEISM-ARG LC accumulation (placeholder).
#include "common.cuh"
using namespace deism;
extern "C" int deism_lc_arg(int64_t, int64_t, int, int, const int64_t *, const int64_t *, const int64_t *,
                            const int64_t *, const float *, const float *, const float *, const float *,
                            const double *, const int64_t *, int64_t, double *, int, void *) {
    return set_error(DEISM_ERR_INVALID, "deism_lc_arg: not implemented yet");
}
