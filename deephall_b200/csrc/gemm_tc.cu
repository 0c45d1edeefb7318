// placeholder until the tcgen05 kernel lands
#include "kernels.h"
namespace dh {
int gemm_tc_supported(int, int) { return 0; }
int gemm_tc(const float*, const float*, const float*, const float*, float*, int64_t, int, int, int64_t, int, int, cudaStream_t) { return -2; }
int split_weight_tc(const float*, int64_t, int, int, float*, float*, cudaStream_t) { return -2; }
}
