// tcgen05 / TMA tensor-core engine (placeholder until the kernels land in this file).
#include "common.cuh"
using namespace rf;

int rf_corr_argmax_tc(const float*, int, const float*, int, int, unsigned long long*, unsigned long long*, cudaStream_t) {
    return fail_msg("rf_corr_mutual_nn: precision=1 (tcgen05) engine is not available in this build");
}
int rf_conv2d_tc(const ImgSet&, const ConvParams&, const float*, cudaStream_t) {
    return fail_msg("rf_conv2d_nhwc: engine=1 (tcgen05) is not available in this build");
}
