// vnpcc_internal.h -- shared helpers for the kernels behind the C-ABI (include/vnpcc.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define VNPCC_OK 0
#define VNPCC_ERR_WORKSPACE 10001
#define VNPCC_ERR_BAD_ARG 10002
#define VNPCC_ERR_UNSUPPORTED 10003
#define VNPCC_ERR_DRIVER 10004

namespace vnpcc {

inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// launch-time errors only (no synchronisation); 0 on success, the cudaError_t value otherwise
inline int last_error() { return (int)cudaGetLastError(); }

// process-wide count of kernel launches enqueued by this library (vnpcc_launch_count(); bench.py's gpu_launches).
// Written as `count_launch(), kernel<<<...>>>(...)` at every launch site.
unsigned long long& launch_counter();
inline int count_launch() {
    ++launch_counter();
    return 0;
}

inline int grid_for(size_t total, int block, int per_sm) {
    size_t g = (total + block - 1) / block;
    size_t cap = (size_t)sm_count() * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace vnpcc
