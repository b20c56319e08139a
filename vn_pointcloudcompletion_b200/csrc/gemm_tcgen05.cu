// gemm_tcgen05.cu -- placeholder until the tensor-core kernels land (returns UNSUPPORTED so callers take the fp32 path)
#include "vnpcc_internal.h"
extern "C" {
int vnpcc_gemm_rows_tf32(const float*, long long, const float*, long long, float*, long long, long long, int, int,
                         const float*, long long, long long, void*) { return VNPCC_ERR_UNSUPPORTED; }
int vnpcc_gemm_wgrad_tf32(const float*, long long, const float*, long long, float*, long long, long long, int, int,
                          float*, size_t, void*) { return VNPCC_ERR_UNSUPPORTED; }
size_t vnpcc_gemm_wgrad_tf32_workspace_bytes(long long, int, int) { return 0; }
}
