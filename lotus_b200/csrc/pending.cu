// pending.cu — entry points declared in include/lotus_b200.h whose kernels land next (dedup, k-means).
#include "common.cuh"
using namespace b2;
extern "C" {
int b2_threshold_pairs(b2_index*, float, int32_t, int32_t, int64_t*, int64_t*, int64_t, int64_t*) {
    set_error("b2_threshold_pairs: not implemented in this build");
    return B2_EINVAL;
}
int b2_connected_components(int64_t, const int64_t*, const int64_t*, int64_t, int32_t, int64_t*) {
    set_error("b2_connected_components: not implemented in this build");
    return B2_EINVAL;
}
int b2_kmeans(b2_index*, const int64_t*, int64_t, int32_t, int32_t, int64_t, int32_t, int64_t*, float*, float*) {
    set_error("b2_kmeans: not implemented in this build");
    return B2_EINVAL;
}
int b2_kmeans_assign(b2_index*, const int64_t*, int64_t, const float*, int32_t, int64_t*, float*) {
    set_error("b2_kmeans_assign: not implemented in this build");
    return B2_EINVAL;
}
}
