// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace murcl {
bool tc_fwd_supported(int64_t, int, int, int, int) { return false; }
bool tc_bwd_input_supported(int64_t, int, int, int) { return false; }
bool tc_bwd_weight_supported(int64_t, int, int, int) { return false; }
int tc_linear_fwd(const void*, const void*, const float*, void*, int64_t, int, int, int, int, cudaStream_t) { return MURCL_EUNSUPPORTED; }
int tc_linear_bwd_input(const void*, const void*, void*, int64_t, int, int, const void*, const float*, const float*, const int32_t*, cudaStream_t) { return MURCL_EUNSUPPORTED; }
int64_t tc_linear_bwd_weight_workspace(int64_t, int, int) { return 0; }
int tc_linear_bwd_weight(const void*, const void*, float*, int64_t, int, int, float*, cudaStream_t) { return MURCL_EUNSUPPORTED; }
}
