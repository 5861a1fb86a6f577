// org.cu — DEISM-ORG accumulation (placeholder until the factorised / direct kernels land).
#include "common.cuh"
using namespace deism;
extern "C" int deism_org_shoebox(int64_t, int64_t, int, int, const float *, const float *, const float *,
                                 const float *, const int32_t *, const float *,
                                 const deism_shoebox_atten *, const double *, double *, int, int, void *) {
    return set_error(DEISM_ERR_INVALID, "deism_org_shoebox: not implemented yet");
}
extern "C" int deism_org_arg(int64_t, int64_t, int, int, const float *, const float *, const float *,
                             const float *, const float *, const float *, const double *,
                             const int64_t *, int64_t, double *, int, void *) {
    return set_error(DEISM_ERR_INVALID, "deism_org_arg: not implemented yet");
}
