// microbench.cu -- measures the FP32 pipe peak the Chamfer roofline is quoted against (BASELINE.md section 5:
// "measure with an FFMA micro-benchmark"): scalar FFMA, packed FFMA2 (two fp32 FMAs per issue slot), and the
// Chamfer instruction mix (3 packed FP32 + 0.5 FMNMX3 per pair) without any memory traffic.
#include <cuda_runtime.h>

#include "vnpcc_internal.h"

namespace vnpcc {

typedef unsigned long long u64;

template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    constexpr int U = 16;
    if (MODE == 0) {
        float acc[U];
#pragma unroll
        for (int i = 0; i < U; ++i) acc[i] = threadIdx.x * 1e-3f + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < U; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < U; ++i) s += acc[i];
        if (s == 12345.678f) out[0] = s;
    } else if (MODE == 1) {
        u64 acc[U];
        const u64 aa = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a);
        const u64 bb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
#pragma unroll
        for (int i = 0; i < U; ++i) acc[i] = ((u64)__float_as_uint(threadIdx.x * 1e-3f + i) << 32) | (u64)(i + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < U; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
        }
        u64 s = 0;
#pragma unroll
        for (int i = 0; i < U; ++i) s ^= acc[i];
        if (s == 0x1234567ull) out[0] = 1.f;
    } else {
        // Chamfer mix per 2 pairs: 3 FADD2, 1 FMUL2, 2 FFMA2, 1 FMNMX3
        u64 c[U];
        float m[U / 4];
        const u64 q = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a);
#pragma unroll
        for (int i = 0; i < U; ++i) c[i] = ((u64)__float_as_uint(threadIdx.x * 1e-3f + i) << 32) | (u64)__float_as_uint(b + i);
#pragma unroll
        for (int i = 0; i < U / 4; ++i) m[i] = 3e38f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < U; i += 4) {
                u64 dx, dy, dz, t;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(c[i]), "l"(q));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(c[i + 1]), "l"(q));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(c[i + 2]), "l"(q));
                asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(t) : "l"(dx));
                asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(t) : "l"(dz));
                m[i / 4] = fminf(fminf(m[i / 4], __uint_as_float((unsigned)t)), __uint_as_float((unsigned)(t >> 32)));
                c[i + 3] ^= (u64)it;   // keep the loop from being hoisted; integer pipe, off the FP32 path
                c[i] += c[i + 3] & 1;
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < U / 4; ++i) s += m[i];
        if (s == 12345.678f) out[0] = s;
    }
}

}  // namespace vnpcc

extern "C" int vnpcc_measure_fp32_peak(int mode, int iters, float* scratch_dev, float* ms_out, double* lane_ops_out,
                                       void* stream) {
    using namespace vnpcc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = sm_count() * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, st);
        if (mode == 0) count_launch(), fp32_peak_kernel<0><<<grid, block, 0, st>>>(scratch_dev, iters, 1.0001f, 0.5f);
        else if (mode == 1) count_launch(), fp32_peak_kernel<1><<<grid, block, 0, st>>>(scratch_dev, iters, 1.0001f, 0.5f);
        else count_launch(), fp32_peak_kernel<2><<<grid, block, 0, st>>>(scratch_dev, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms;
    const double threads = (double)grid * block;
    // fp32 lane-operations issued (an FMA counts as ONE lane-op here, i.e. one FP32-pipe slot)
    if (mode == 0) *lane_ops_out = threads * 16.0 * iters;
    else if (mode == 1) *lane_ops_out = threads * 16.0 * 2.0 * iters;
    else *lane_ops_out = threads * 4.0 * 12.0 * iters;   // 4 pair-pairs per iteration, 6 packed instr = 12 lane-ops each
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return last_error();
}
