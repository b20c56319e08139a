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

// index of the current device into per-device caches (function attributes are per device)
inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= 64) ? 0 : dev;
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

// resident CTAs per SM of one kernel at a block size / dynamic shared-memory size (occupancy calculator; cached per kernel).
// Launchers size persistent / grid-stride grids as sm_count() * resident so that the whole grid is ONE wave of equal shares.
int resident_ctas_impl(const void* fn, int threads, size_t smem);
template <class K>
inline int resident_ctas(K* kernel, int threads, size_t smem = 0) {
    return resident_ctas_impl(reinterpret_cast<const void*>(kernel), threads, smem);
}
// development knobs (vnpcc_set_tuning): 0 = default behaviour
enum { TUNE_GRID_LEGACY = 0, TUNE_FOLD_MINB = 1, TUNE_STATS_GEMM = 2, TUNE_STATS_NOMATH = 3, TUNE_CHAMFER_OVERHEAD = 4, TUNE_CHAMFER_CTAS = 5, TUNE_CHAMFER_VARIANT = 6, TUNE_TAIL_WGRAD = 7, TUNE_FOLD_FWD = 8, TUNE_N = 9 };
int tuning(int knob);

// chunk length for kernels whose grid is (groups x chunks) blocks of equal work over N points per group, each block walking its chunk with
// `lanes` point lanes: the chunk count that minimises  waves x (steps per block + a fixed per-block cost)  for `slots` resident CTAs.
inline int plan_chunk_len(long long groups, int N, long long slots, int lanes, int min_chunk, int fixed_steps = 4) {
    if (groups < 1) groups = 1;
    if (slots < 1) slots = 1;
    int best = N > min_chunk ? N : min_chunk;
    long long best_cost = -1;
    const int max_chunks = N / min_chunk > 1 ? N / min_chunk : 1;
    long long hi = (8 * slots + groups - 1) / groups;
    if (hi > max_chunks) hi = max_chunks;
    if (hi < 1) hi = 1;
    for (long long c = 1; c <= hi; ++c) {
        const int len = (int)((N + c - 1) / c);
        const long long chunks = (N + len - 1) / len;
        const long long waves = (groups * chunks + slots - 1) / slots;
        const long long cost = waves * ((len + lanes - 1) / lanes + fixed_steps);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = len;
        }
    }
    return best;
}

inline int grid_for(size_t total, int block, int per_sm) {
    size_t g = (total + block - 1) / block;
    size_t cap = (size_t)sm_count() * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// vectorised (4 channels / thread) streaming kernels, vn_stream.cu: return false when the shape / alignment does not allow them
bool try_norm_stats_v4(const float* p, long long ldp, long long P, int C, double* sums, cudaStream_t st);
bool try_bn_leaky_fwd_v4(const float* p, long long ldp, const float* d, long long ldd, float* out, long long ldo, long long P, int C,
                         const float* stat, const float* gamma, const float* beta, float ns, cudaStream_t st);
bool try_bn_leaky_bwd1_v4(const float* g, long long ldg, const float* p, long long ldp, const float* d, long long ldd, float* gp,
                          long long ldgp, float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma,
                          const float* beta, float ns, double* sums, cudaStream_t st);
bool try_bn_bwd2_v4(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat, const float* gamma,
                    const float* beta, const double* sums, double count, int training, cudaStream_t st);

bool try_bn_bwd2_v4_sbias(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat, const float* gamma,
                          const float* beta, const double* sums, double count, int training, const float* gd, long long ldgd, float* gbias,
                          long long ldgb, long long pts_per_sample, cudaStream_t st);

bool try_bn_leaky_dot_fwd_v4(const float* p, long long ldp, const float* d, long long ldd, long long P, int C, const float* stat,
                             const float* gamma, const float* beta, float ns, const float* w2, const float* res, float* y,
                             cudaStream_t st);
bool try_bn_leaky_dot_bwd1_v4(const float* gy, const float* p, long long ldp, const float* d, long long ldd, float* gp, long long ldgp,
                              float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma, const float* beta,
                              float ns, double* sums, const float* w2, double* gw2, cudaStream_t st);

bool try_bn_leaky_dot_sums_v4(const float* gy, const float* p, long long ldp, const float* d, long long ldd, long long P, int C,
                              const float* stat, const float* gamma, const float* beta, float ns, double* sums, const float* w2, double* gw2,
                              cudaStream_t st);

}  // namespace vnpcc
