// vn_kernels.cu -- HBM-bound Vector-Neuron kernels for sm_100a (norm statistics, BatchNorm-on-norms + leaky
// projection forward/backward, arg-max pooling, per-sample bias / reductions, the C->1 VNLinear).
//
// Physical layout ("channels-last", the same one the reference's nn.Linear calls physically produce):
// a logical VN tensor [B, C, 3, N] is stored as a row-major matrix X[R, C] with R = 3*B*N rows, row index
// r = (b*N + n)*3 + v, channel contiguous, leading dimension `ld` floats (so halves of a stacked [R, 2C] buffer
// can be addressed in place).  A warp reads 32 (or 128 with float4) consecutive channels of one row: every access
// is a full 128-byte line.  The three components of one vector sit in three consecutive rows.
//
// Reference semantics restated (file:line under /root/reference):
//   VNBatchNorm        models/vn_layers.py:116-127   norm = ||x|| + 1e-6 ; x / norm * BN(norm)
//   leaky projection   models/vn_layers.py:39-42, 70-73   (mask = dot >= 0, eps on ||d||^2)
//   VNMaxPool          models/vn_layers.py:162-166   argmax_n <x,d>, first maximum wins
// Backward formulas: SURVEY.md Appendix C (derived from the forward definitions, the reference uses autograd).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vnpcc_internal.h"

namespace vnpcc {

constexpr float VN_EPS = 1e-6f;      // models/vn_layers.py:10
typedef unsigned long long u64;

struct V3 {
    float x, y, z;
};

__device__ __forceinline__ float dot3(const V3& a, const V3& b) {
    // (a*b).sum(2): three rounded products summed left to right (no contraction)
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ V3 ld3(const float* __restrict__ p, size_t ld) {
    V3 r;
    r.x = __ldg(p);
    r.y = __ldg(p + ld);
    r.z = __ldg(p + 2 * ld);
    return r;
}
__device__ __forceinline__ void st3(float* __restrict__ p, size_t ld, const V3& v) {
    p[0] = v.x;
    p[ld] = v.y;
    p[2 * ld] = v.z;
}

// BatchNorm-on-norm for one vector: returns the scale nb/n and the pieces the backward needs
struct BNPiece {
    float r, n, nhat, nb;
};
__device__ __forceinline__ BNPiece bn_piece(const V3& p, float mean, float invstd, float gamma, float beta) {
    BNPiece o;
    o.r = sqrtf(dot3(p, p));
    o.n = o.r + VN_EPS;
    o.nhat = (o.n - mean) * invstd;
    o.nb = o.nhat * gamma + beta;
    return o;
}

// -------------------------------------------------------------------------------------------------------------
// 1. per-channel statistics of the vector norms:  sums[c] += sum n ; sums[C+c] += sum n^2   (double)
//    block = (32 channel lanes) x (8 point lanes); grid.x tiles channels by 32, grid.y strides over points.
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vn_norm_stats_kernel(const float* __restrict__ p, size_t ld, long long P, int C,
                                                             double* __restrict__ sums) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        long long pt = (long long)blockIdx.y * blockDim.y + threadIdx.y;
        const long long stride = (long long)gridDim.y * blockDim.y;
        // accumulate n and n*n in double: var = E[n^2] - mean^2 cancels catastrophically in fp32 when the norms of a
        // channel barely vary (decoder final_conv[0]: the per-sample global feature dominates every point's norm)
#pragma unroll 4
        for (; pt < P; pt += stride) {
            V3 v = ld3(p + (size_t)pt * 3 * ld + c, ld);
            const double n = (double)(sqrtf(dot3(v, v)) + VN_EPS);
            s1 += n;
            s2 = fma(n, n, s2);
        }
    }
    __shared__ double sh1[8][33], sh2[8][33];
    sh1[threadIdx.y][threadIdx.x] = s1;
    sh2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int k = 1; k < 8; ++k) {
            s1 += sh1[k][threadIdx.x];
            s2 += sh2[k][threadIdx.x];
        }
        atomicAdd(sums + c, s1);
        atomicAdd(sums + C + c, s2);
    }
}

// finalize: stat[c] = mean, stat[C+c] = 1/sqrt(var+eps).  training: batch stats (+ running update, unbiased var);
// eval: running stats.  One thread per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float bn_eps, float* __restrict__ stat) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (training) {
        const double mean = sums[c] / count;
        double var = sums[C + c] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[c] = (float)mean;
        stat[C + c] = (float)(1.0 / sqrt(var + (double)bn_eps));
        if (running_mean) {
            const double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
            running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
        }
    } else {
        stat[c] = running_mean[c];
        stat[C + c] = (float)(1.0 / sqrt((double)running_var[c] + (double)bn_eps));
    }
}

// -------------------------------------------------------------------------------------------------------------
// 2. forward apply:  out = leaky( BN(p), d )     (BN optional: stat == NULL ; leaky optional: d == NULL)
//    one thread per (point, channel); consecutive threads -> consecutive channels.
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 leaky_fwd(const V3& p, const V3& d, float ns) {
    // op-by-op rounding of the eager expression  ns*p + (1-ns)*(mask*p + (1-mask)*(p - (dot/(dsq+EPS))*d))
    const float dot = dot3(p, d);
    const float k = 1.f - ns;
    V3 in = p;
    if (!(dot >= 0.f)) {
        const float a = dot / __fadd_rn(dot3(d, d), VN_EPS);
        in.x = __fsub_rn(p.x, __fmul_rn(a, d.x));
        in.y = __fsub_rn(p.y, __fmul_rn(a, d.y));
        in.z = __fsub_rn(p.z, __fmul_rn(a, d.z));
    }
    V3 o;
    o.x = __fadd_rn(__fmul_rn(ns, p.x), __fmul_rn(k, in.x));
    o.y = __fadd_rn(__fmul_rn(ns, p.y), __fmul_rn(k, in.y));
    o.z = __fadd_rn(__fmul_rn(ns, p.z), __fmul_rn(k, in.z));
    return o;
}

__global__ void __launch_bounds__(256) vn_bn_leaky_fwd_kernel(const float* __restrict__ p, size_t ldp,
                                                               const float* __restrict__ d, size_t ldd,
                                                               float* __restrict__ out, size_t ldo, long long P, int C,
                                                               const float* __restrict__ stat,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float ns) {
    const long long total = P * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long pt = t / C;
        const int c = (int)(t - pt * C);
        V3 v = ld3(p + (size_t)pt * 3 * ldp + c, ldp);
        if (stat) {
            BNPiece b = bn_piece(v, __ldg(stat + c), __ldg(stat + C + c), __ldg(gamma + c), __ldg(beta + c));
            v.x = v.x / b.n * b.nb;    // x / norm * norm_bn  (vn_layers.py:125)
            v.y = v.y / b.n * b.nb;
            v.z = v.z / b.n * b.nb;
        }
        if (d) {
            V3 dv = ld3(d + (size_t)pt * 3 * ldd + c, ldd);
            v = leaky_fwd(v, dv, ns);
        }
        st3(out + (size_t)pt * 3 * ldo + c, ldo, v);
    }
}

// -------------------------------------------------------------------------------------------------------------
// 3. backward pass 1: from g = dL/dout recompute BN(p), back through the leaky projection ->
//      gp <- dL/d(BN(p))   (grad w.r.t. the post-BN vector; pass 2 turns it into dL/dp)
//      gd <- dL/dd
//    and, when BN is on, accumulate S1 = sum d_nb and S2 = sum d_nb*nhat per channel (d_nb = <gp,p>/n).
//    Same 32x8 thread shape as the stats kernel so the per-channel sums stay in registers.
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vn_bn_leaky_bwd1_kernel(const float* __restrict__ g, size_t ldg,
                                                                const float* __restrict__ p, size_t ldp,
                                                                const float* __restrict__ d, size_t ldd,
                                                                float* __restrict__ gp, size_t ldgp,
                                                                float* __restrict__ gd, size_t ldgd, long long P, int C,
                                                                const float* __restrict__ stat,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float ns,
                                                                double* __restrict__ sums) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        float mean = 0.f, invstd = 0.f, ga = 0.f, be = 0.f;
        if (stat) {
            mean = __ldg(stat + c);
            invstd = __ldg(stat + C + c);
            ga = __ldg(gamma + c);
            be = __ldg(beta + c);
        }
        const float k = 1.f - ns;
        long long pt = (long long)blockIdx.y * blockDim.y + threadIdx.y;
        const long long stride = (long long)gridDim.y * blockDim.y;
        {
            for (; pt < P; pt += stride) {
                const V3 pr = ld3(p + (size_t)pt * 3 * ldp + c, ldp);
                const V3 gv = ld3(g + (size_t)pt * 3 * ldg + c, ldg);
                V3 pb = pr;
                BNPiece b;
                if (stat) {
                    b = bn_piece(pr, mean, invstd, ga, be);
                    pb.x = pr.x / b.n * b.nb;
                    pb.y = pr.y / b.n * b.nb;
                    pb.z = pr.z / b.n * b.nb;
                }
                V3 gpb = gv;
                if (d) {
                    const V3 dv = ld3(d + (size_t)pt * 3 * ldd + c, ldd);
                    const float s = dot3(pb, dv);
                    V3 gdv = {0.f, 0.f, 0.f};
                    if (s < 0.f) {
                        const float q = dot3(dv, dv) + VN_EPS;
                        const float a = s / q;
                        const float gdq = dot3(gv, dv) / q;
                        gpb.x = gv.x - k * gdq * dv.x;
                        gpb.y = gv.y - k * gdq * dv.y;
                        gpb.z = gv.z - k * gdq * dv.z;
                        gdv.x = -k * (a * gv.x + gdq * pb.x - 2.f * a * gdq * dv.x);
                        gdv.y = -k * (a * gv.y + gdq * pb.y - 2.f * a * gdq * dv.y);
                        gdv.z = -k * (a * gv.z + gdq * pb.z - 2.f * a * gdq * dv.z);
                    }
                    st3(gd + (size_t)pt * 3 * ldgd + c, ldgd, gdv);
                }
                st3(gp + (size_t)pt * 3 * ldgp + c, ldgp, gpb);
                if (stat) {
                    const float dnb = dot3(gpb, pr) / b.n;
                    s1 += (double)dnb;
                    s2 = fma((double)dnb, (double)b.nhat, s2);
                }
            }
        }
    }
    if (sums) {
        __shared__ double sh1[8][33], sh2[8][33];
        sh1[threadIdx.y][threadIdx.x] = s1;
        sh2[threadIdx.y][threadIdx.x] = s2;
        __syncthreads();
        if (threadIdx.y == 0 && c < C) {
            for (int kk = 1; kk < 8; ++kk) {
                s1 += sh1[kk][threadIdx.x];
                s2 += sh2[kk][threadIdx.x];
            }
            atomicAdd(sums + c, s1);
            atomicAdd(sums + C + c, s2);
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// 4. backward pass 2 (BatchNorm-on-norm backward, in place on gp):  gp <- dL/dp
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vn_bn_bwd2_kernel(float* __restrict__ gp, size_t ldgp,
                                                          const float* __restrict__ p, size_t ldp, long long P, int C,
                                                          const float* __restrict__ stat,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const double* __restrict__ sums, double count, int training) {
    const long long total = P * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long pt = t / C;
        const int c = (int)(t - pt * C);
        const V3 pr = ld3(p + (size_t)pt * 3 * ldp + c, ldp);
        float* gptr = gp + (size_t)pt * 3 * ldgp + c;
        V3 gv;
        gv.x = gptr[0];
        gv.y = gptr[ldgp];
        gv.z = gptr[2 * ldgp];
        const float ga = __ldg(gamma + c), invstd = __ldg(stat + C + c);
        const BNPiece b = bn_piece(pr, __ldg(stat + c), invstd, ga, __ldg(beta + c));
        const float gx = dot3(gv, pr);
        const float dnb = gx / b.n;
        float dn = ga * dnb;
        if (training) {
            const float m1 = (float)(sums[c] / count) * ga;
            const float m2 = (float)(sums[C + c] / count) * ga;
            dn = dn - m1 - b.nhat * m2;
        }
        dn = dn * invstd - gx * b.nb / (b.n * b.n);
        const float sc = b.nb / b.n;
        const float ur = b.r > 0.f ? dn / b.r : 0.f;
        V3 o;
        o.x = gv.x * sc + ur * pr.x;
        o.y = gv.y * sc + ur * pr.y;
        o.z = gv.z * sc + ur * pr.z;
        st3(gptr, ldgp, o);
    }
}

__global__ void double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

// -------------------------------------------------------------------------------------------------------------
// 5. VNMaxPool.  score(b,c,n) = <x,d> ; winner = max score, lowest n on ties (torch.max picks the first maximum).
//    Partial winners are merged with a 64-bit atomicMax on (orderable score bits << 32 | ~n).
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned orderable(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256) vn_maxpool_argmax_kernel(const float* __restrict__ x, size_t ldx,
                                                                 const float* __restrict__ d, size_t ldd, int B, int N,
                                                                 int C, int n_chunk, u64* __restrict__ best) {
    // grid: x -> channel tiles of 32, y -> (sample, chunk of points); block (32, 8)
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int chunks_per_b = (N + n_chunk - 1) / n_chunk;
    const int b = blockIdx.y / chunks_per_b;
    const int ck = blockIdx.y - b * chunks_per_b;
    const int n0 = ck * n_chunk, n1 = min(N, n0 + n_chunk);
    u64 mine = 0ull;
    if (c < C) {
        for (int n = n0 + threadIdx.y; n < n1; n += blockDim.y) {
            const size_t row = ((size_t)b * N + n) * 3;
            const V3 xv = ld3(x + row * ldx + c, ldx);
            const V3 dv = ld3(d + row * ldd + c, ldd);
            const float s = dot3(xv, dv);
            const u64 key = ((u64)orderable(s) << 32) | (u64)(0xffffffffu - (unsigned)n);
            mine = key > mine ? key : mine;
        }
    }
    __shared__ u64 sh[8][33];
    sh[threadIdx.y][threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int k = 1; k < 8; ++k) mine = sh[k][threadIdx.x] > mine ? sh[k][threadIdx.x] : mine;
        atomicMax(best + (size_t)b * C + c, mine);
    }
}

__global__ void vn_maxpool_decode_kernel(const u64* __restrict__ best, long long total, long long* __restrict__ idx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) idx[t] = (long long)(0xffffffffu - (unsigned)(best[t] & 0xffffffffu));
}

// out[(b,v), c] = x[(b, idx[b,c], v), c]
__global__ void vn_maxpool_gather_kernel(const float* __restrict__ x, size_t ldx, const long long* __restrict__ idx,
                                         int B, int N, int C, float* __restrict__ out, size_t ldo) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * C) return;
    const int b = (int)(t / C), c = (int)(t - (long long)b * C);
    const long long n = idx[t];
    const V3 v = ld3(x + (((size_t)b * N + n) * 3) * ldx + c, ldx);
    st3(out + ((size_t)b * 3) * ldo + c, ldo, v);
}

// gx[(b, idx[b,c], v), c] += g[(b,v), c]     (one owner per element: plain read-modify-write)
__global__ void vn_maxpool_scatter_kernel(const float* __restrict__ g, size_t ldg, const long long* __restrict__ idx,
                                          int B, int N, int C, float* __restrict__ gx, size_t ldgx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * C) return;
    const int b = (int)(t / C), c = (int)(t - (long long)b * C);
    const long long n = idx[t];
    float* dst = gx + (((size_t)b * N + n) * 3) * ldgx + c;
    const float* src = g + ((size_t)b * 3) * ldg + c;
    dst[0] += src[0];
    dst[ldgx] += src[ldg];
    dst[2 * ldgx] += src[2 * ldg];
}

// -------------------------------------------------------------------------------------------------------------
// 6. per-sample bias (broadcast channels folded out of a GEMM) and its adjoint
//    y[(b,n,v), c] += bias[(b,v), c]            ;   out[(b,v), c] = sum_n g[(b,n,v), c]
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_add_sample_bias_kernel(float* __restrict__ y, size_t ldy,
                                                                    const float* __restrict__ bias, size_t ldb, int B,
                                                                    int N, int C) {
    const long long total = (long long)B * N * 3 * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / C;
        const int c = (int)(t - row * C);
        const int v = (int)(row % 3);
        const int b = (int)(row / (3LL * N));
        y[(size_t)row * ldy + c] += __ldg(bias + ((size_t)b * 3 + v) * ldb + c);
    }
}

__global__ void __launch_bounds__(256) rows_sample_sum_kernel(const float* __restrict__ g, size_t ldg, int B, int N,
                                                               int C, int n_chunk, float* __restrict__ out, size_t ldo) {
    // grid: x -> channel tiles of 32, y -> (sample, chunk); block (32, 8); out must be zeroed
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int chunks_per_b = (N + n_chunk - 1) / n_chunk;
    const int b = blockIdx.y / chunks_per_b;
    const int ck = blockIdx.y - b * chunks_per_b;
    const int n0 = ck * n_chunk, n1 = min(N, n0 + n_chunk);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (c < C) {
        for (int n = n0 + threadIdx.y; n < n1; n += blockDim.y) {
            const V3 v = ld3(g + (((size_t)b * N + n) * 3) * ldg + c, ldg);
            sx += v.x;
            sy += v.y;
            sz += v.z;
        }
    }
    __shared__ float sh[3][8][33];
    sh[0][threadIdx.y][threadIdx.x] = sx;
    sh[1][threadIdx.y][threadIdx.x] = sy;
    sh[2][threadIdx.y][threadIdx.x] = sz;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int k = 1; k < 8; ++k) {
            sx += sh[0][k][threadIdx.x];
            sy += sh[1][k][threadIdx.x];
            sz += sh[2][k][threadIdx.x];
        }
        float* o = out + ((size_t)b * 3) * ldo + c;
        atomicAdd(o, sx);
        atomicAdd(o + ldo, sy);
        atomicAdd(o + 2 * ldo, sz);
    }
}

// -------------------------------------------------------------------------------------------------------------
// 7. VNLinear(C -> 1): y[r] = sum_c x[r,c] w[c] (+ res[r]) ; one warp per row.
//    backward: gx[r,c] = gy[r] w[c] ; gw[c] = sum_r gy[r] x[r,c]
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_dot_kernel(const float* __restrict__ x, size_t ldx,
                                                        const float* __restrict__ w, long long R, int C,
                                                        const float* __restrict__ res, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < R; r += nwarps) {
        const float* xr = x + (size_t)r * ldx;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(xr + c), __ldg(w + c), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[r] = res ? acc + __ldg(res + r) : acc;
    }
}

__global__ void __launch_bounds__(256) rows_outer_kernel(const float* __restrict__ gy, const float* __restrict__ w,
                                                          long long R, int C, float* __restrict__ gx, size_t ldgx) {
    const long long total = R * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / C;
        const int c = (int)(t - r * C);
        gx[(size_t)r * ldgx + c] = __ldg(gy + r) * __ldg(w + c);
    }
}

__global__ void __launch_bounds__(256) rows_wsum_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                         size_t ldx, long long R, int C, float* __restrict__ gw) {
    // block (32, 8): channel tile x row lanes; gw must be zeroed
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < C) {
        for (long long r = (long long)blockIdx.y * blockDim.y + threadIdx.y; r < R; r += (long long)gridDim.y * blockDim.y)
            acc = fmaf(__ldg(gy + r), __ldg(x + (size_t)r * ldx + c), acc);
    }
    __shared__ float sh[8][33];
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int k = 1; k < 8; ++k) acc += sh[k][threadIdx.x];
        atomicAdd(gw + c, acc);
    }
}


// -------------------------------------------------------------------------------------------------------------
// 8. small-K VNLinear (K <= 4 input channels: first_conv[0] has K=1, the decoder's final_conv[0] has K=2 local
//    channels after the broadcast global feature is folded into the per-sample bias).  These are HBM-bound
//    streaming kernels, not GEMMs: the weight fits in registers.
//      fwd  : y[r, o] = bias[(b,v), o] + sum_k x[r, k] W[o, k]
//      dgrad: gx[r, k] = sum_o gy[r, o] W[o, k]
//      wgrad: gW[o, k] = sum_r gy[r, o] x[r, k]   and (optionally)  gbias[(b,v), o] = sum_n gy[(b,n,v), o]
// -------------------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(128) smallk_fwd_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ W,
                                                          size_t ldw, const float* __restrict__ bias, size_t ldb, long long n_per_sample,
                                                          float* __restrict__ y, size_t ldy, long long P, int Cout,
                                                          long long pts_per_block) {
    // thread owns one channel quad (weights in registers) and streams over a contiguous chunk of points (3 rows each);
    // the per-sample bias rows are reloaded only when the sample changes
    const int q = blockIdx.y * 128 + threadIdx.x;
    if (4 * q >= Cout) return;
    float w[4][KS];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < KS; ++k) w[j][k] = __ldg(W + (size_t)(4 * q + j) * ldw + k);
    const long long p0 = (long long)blockIdx.x * pts_per_block;
    const long long p1 = (p0 + pts_per_block < P) ? p0 + pts_per_block : P;
    long long b = bias ? p0 / n_per_sample : 0;
    long long nrem = bias ? p0 - b * n_per_sample : 0;
    float4 bz[3];
    bz[0] = bz[1] = bz[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool reload = bias != nullptr;
#pragma unroll 2
    for (long long pt = p0; pt < p1; ++pt) {
        if (reload) {
#pragma unroll
            for (int v = 0; v < 3; ++v) bz[v] = __ldg(reinterpret_cast<const float4*>(bias + (size_t)(b * 3 + v) * ldb) + q);
            reload = false;
        }
        const size_t row = (size_t)pt * 3;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            float4 acc = bz[v];
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const float xv = __ldg(x + (row + v) * ldx + k);
                acc.x = fmaf(xv, w[0][k], acc.x);
                acc.y = fmaf(xv, w[1][k], acc.y);
                acc.z = fmaf(xv, w[2][k], acc.z);
                acc.w = fmaf(xv, w[3][k], acc.w);
            }
            reinterpret_cast<float4*>(y + (row + v) * ldy)[q] = acc;
        }
        if (bias && ++nrem == n_per_sample) {
            nrem = 0;
            ++b;
            reload = true;
        }
    }
}

template <int KS, int NQL>
__global__ void __launch_bounds__(256) smallk_dgrad_kernel(const float* __restrict__ gy, size_t ldgy, const float* __restrict__ W,
                                                            size_t ldw, float* __restrict__ gx, size_t ldgx, long long R,
                                                            int Cout) {
    // one warp per row; lane owns channel quads lane + 32*i (weights in registers)
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nq = Cout >> 2;
    float w[NQL][4][KS];
#pragma unroll
    for (int i = 0; i < NQL; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const int q = lane + 32 * i;
                w[i][j][k] = q < nq ? __ldg(W + (size_t)(4 * q + j) * ldw + k) : 0.f;
            }
#pragma unroll 2
    for (long long r = warp; r < R; r += nwarps) {
        const float4* g4 = reinterpret_cast<const float4*>(gy + (size_t)r * ldgy);
        float4 g[NQL];
#pragma unroll
        for (int i = 0; i < NQL; ++i) g[i] = (lane + 32 * i < nq) ? __ldg(g4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) acc[k] = 0.f;
#pragma unroll
        for (int i = 0; i < NQL; ++i)
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                acc[k] = fmaf(g[i].x, w[i][0][k], acc[k]);
                acc[k] = fmaf(g[i].y, w[i][1][k], acc[k]);
                acc[k] = fmaf(g[i].z, w[i][2][k], acc[k]);
                acc[k] = fmaf(g[i].w, w[i][3][k], acc[k]);
            }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            if (lane == 0) gx[(size_t)r * ldgx + k] = acc[k];
        }
    }
}

template <int KS>
__global__ void __launch_bounds__(256) smallk_wgrad_kernel(const float* __restrict__ gy, size_t ldgy, const float* __restrict__ x,
                                                            size_t ldx, int B, int N, int Cout, int n_chunk,
                                                            float* __restrict__ gW, size_t ldgw, float* __restrict__ gbias,
                                                            size_t ldgb) {
    // grid: x -> tiles of 32 channel quads, y -> (sample, chunk of points); block (32, 8); gW / gbias zeroed by the launcher
    const int q = blockIdx.x * 32 + threadIdx.x;
    const bool active = 4 * q < Cout;
    const int chunks_per_b = (N + n_chunk - 1) / n_chunk;
    const int b = blockIdx.y / chunks_per_b;
    const int ck = blockIdx.y - b * chunks_per_b;
    const int n0 = ck * n_chunk, n1 = min(N, n0 + n_chunk);
    float sb[3][4], sw[KS][4];
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int l = 0; l < 4; ++l) sb[v][l] = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k)
#pragma unroll
        for (int l = 0; l < 4; ++l) sw[k][l] = 0.f;
    if (active) {
#pragma unroll 2
        for (int n = n0 + threadIdx.y; n < n1; n += 8) {
            const size_t row = ((size_t)b * N + n) * 3;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gy + (row + v) * ldgy) + q);
                sb[v][0] += g.x;
                sb[v][1] += g.y;
                sb[v][2] += g.z;
                sb[v][3] += g.w;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const float xv = __ldg(x + (row + v) * ldx + k);
                    sw[k][0] = fmaf(g.x, xv, sw[k][0]);
                    sw[k][1] = fmaf(g.y, xv, sw[k][1]);
                    sw[k][2] = fmaf(g.z, xv, sw[k][2]);
                    sw[k][3] = fmaf(g.w, xv, sw[k][3]);
                }
            }
        }
    }
    __shared__ float sh[3 + KS][8][32][4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
#pragma unroll
        for (int v = 0; v < 3; ++v) sh[v][threadIdx.y][threadIdx.x][l] = sb[v][l];
#pragma unroll
        for (int k = 0; k < KS; ++k) sh[3 + k][threadIdx.y][threadIdx.x][l] = sw[k][l];
    }
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int c = 4 * q + l;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                float a = 0.f;
                for (int i = 0; i < 8; ++i) a += sh[v][i][threadIdx.x][l];
                if (gbias) atomicAdd(gbias + ((size_t)b * 3 + v) * ldgb + c, a);
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                float a = 0.f;
                for (int i = 0; i < 8; ++i) a += sh[3 + k][i][threadIdx.x][l];
                atomicAdd(gW + (size_t)c * ldgw + k, a);
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// 9. backward of  VNLinear -> VNMaxPool  (f = x W^T, out[b,c,:] = f[b,c,:,idx[b,c]]) without the dense [R, C] gradient:
//    only one point per (sample, channel) receives a gradient, so
//      gx[(b, idx[b,c], v), k] += g[(b,v), c] * W[c, k]                       (scatter, vector red.add)
//      gW[c, k]                 = sum_{b,v} g[(b,v), c] * x[(b, idx[b,c], v), k]   (gather, no atomics)
//    replaces a C x K x R dgrad GEMM and a C x K x R wgrad GEMM (825 GFLOP each for second_conv[1] at B=32).
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_linear_bwd_x_kernel(const float* __restrict__ g, size_t ldg,
                                                                 const long long* __restrict__ idx,
                                                                 const float* __restrict__ W, size_t ldw, int B, int N,
                                                                 int C, int K, float* __restrict__ gx, size_t ldgx) {
    // grid (C / CPB, B); block 256 threads = K/4 float4 lanes (grid-stride over k quads)
    const int b = blockIdx.y;
    const int kq = K >> 2;
    constexpr int CPB = 8;
    const int c0 = blockIdx.x * CPB;
    for (int ci = 0; ci < CPB; ++ci) {
        const int c = c0 + ci;
        if (c >= C) break;
        const long long n = idx[(size_t)b * C + c];
        const float g0 = __ldg(g + ((size_t)b * 3 + 0) * ldg + c);
        const float g1 = __ldg(g + ((size_t)b * 3 + 1) * ldg + c);
        const float g2 = __ldg(g + ((size_t)b * 3 + 2) * ldg + c);
        float* row = gx + (((size_t)b * N + n) * 3) * ldgx;
        const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)c * ldw);
        for (int q = threadIdx.x; q < kq; q += blockDim.x) {
            const float4 w = __ldg(w4 + q);
            atomicAdd(reinterpret_cast<float4*>(row) + q, make_float4(g0 * w.x, g0 * w.y, g0 * w.z, g0 * w.w));
            atomicAdd(reinterpret_cast<float4*>(row + ldgx) + q, make_float4(g1 * w.x, g1 * w.y, g1 * w.z, g1 * w.w));
            atomicAdd(reinterpret_cast<float4*>(row + 2 * ldgx) + q, make_float4(g2 * w.x, g2 * w.y, g2 * w.z, g2 * w.w));
        }
    }
}

__global__ void __launch_bounds__(256) pool_linear_bwd_w_kernel(const float* __restrict__ g, size_t ldg,
                                                                 const long long* __restrict__ idx,
                                                                 const float* __restrict__ x, size_t ldx, int B, int N,
                                                                 int C, int K, float* __restrict__ gW, size_t ldgw) {
    // one block per output channel c; threads over k quads; loop over the B*3 gathered rows
    const int c = blockIdx.x;
    const int kq = K >> 2;
    for (int q = threadIdx.x; q < kq; q += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < B; ++b) {
            const long long n = idx[(size_t)b * C + c];
            const float* row = x + (((size_t)b * N + n) * 3) * ldx;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const float gv = __ldg(g + ((size_t)b * 3 + v) * ldg + c);
                const float4 xv = __ldg(reinterpret_cast<const float4*>(row + v * ldx) + q);
                acc.x = fmaf(gv, xv.x, acc.x);
                acc.y = fmaf(gv, xv.y, acc.y);
                acc.z = fmaf(gv, xv.z, acc.z);
                acc.w = fmaf(gv, xv.w, acc.w);
            }
        }
        reinterpret_cast<float4*>(gW + (size_t)c * ldgw)[q] = acc;
    }
}


// pooled output of  VNLinear -> VNMaxPool  recomputed from the selected input rows (the layer output itself was never
// stored by the fused pooling GEMM):  out[(b,v), c] = sum_k W[c,k] * x[(b, idx[b,c], v), k].  One warp per (b, c).
__global__ void __launch_bounds__(256) pool_linear_gather_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ W,
                                                                  size_t ldw, const long long* __restrict__ idx, int B, int N, int C,
                                                                  int K, float* __restrict__ out, size_t ldo) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= (long long)B * C) return;
    const int b = (int)(warp / C), c = (int)(warp - (long long)b * C);
    const long long n = idx[warp];
    const float* xr = x + (((size_t)b * N + n) * 3) * ldx;
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)c * ldw);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int q = lane; q < (K >> 2); q += 32) {
        const float4 w = __ldg(w4 + q);
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(xr) + q);
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(xr + ldx) + q);
        const float4 x2 = __ldg(reinterpret_cast<const float4*>(xr + 2 * ldx) + q);
        a0 = fmaf(w.x, x0.x, fmaf(w.y, x0.y, fmaf(w.z, x0.z, fmaf(w.w, x0.w, a0))));
        a1 = fmaf(w.x, x1.x, fmaf(w.y, x1.y, fmaf(w.z, x1.z, fmaf(w.w, x1.w, a1))));
        a2 = fmaf(w.x, x2.x, fmaf(w.y, x2.y, fmaf(w.z, x2.z, fmaf(w.w, x2.w, a2))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        out[((size_t)b * 3 + 0) * ldo + c] = a0;
        out[((size_t)b * 3 + 1) * ldo + c] = a1;
        out[((size_t)b * 3 + 2) * ldo + c] = a2;
    }
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

// sums: 2*C doubles, zeroed here.
int vnpcc_vn_norm_stats(const float* p, long long ldp, long long P, int C, double* sums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    if (P > 0 && C > 0 && try_norm_stats_v4(p, ldp, P, C, sums, st)) return last_error();
    if (P > 0 && C > 0) {
        dim3 block(32, 8);
        const int gx = (C + 31) / 32;
        long long gy = (P + 8 * 16 - 1) / (8 * 16);
        const long long cap = ((long long)sm_count() * 16 + gx - 1) / gx;
        if (gy > cap) gy = cap;
        if (gy < 1) gy = 1;
        count_launch(), vn_norm_stats_kernel<<<dim3(gx, (unsigned)gy), block, 0, st>>>(p, (size_t)ldp, P, C, sums);
    }
    return last_error();
}

int vnpcc_bn_finalize(const double* sums, double count, int C, int training, float* running_mean, float* running_var,
                      float momentum, float bn_eps, float* stat, void* stream) {
    if (C <= 0) return 0;
    count_launch(), bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, C, training, running_mean,
                                                                        running_var, momentum, bn_eps, stat);
    return last_error();
}

int vnpcc_vn_bn_leaky_fwd(const float* p, long long ldp, const float* d, long long ldd, float* out, long long ldo,
                          long long P, int C, const float* stat, const float* gamma, const float* beta, float ns,
                          void* stream) {
    if (P <= 0 || C <= 0) return 0;
    if (try_bn_leaky_fwd_v4(p, ldp, d, ldd, out, ldo, P, C, stat, gamma, beta, ns, (cudaStream_t)stream)) return last_error();
    count_launch(), vn_bn_leaky_fwd_kernel<<<grid_for((size_t)P * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        p, (size_t)ldp, d, (size_t)ldd, out, (size_t)ldo, P, C, stat, gamma, beta, ns);
    return last_error();
}

// sums (2*C doubles) is zeroed here when stat != NULL; pass NULL sums/stat for "no BatchNorm".
int vnpcc_vn_bn_leaky_bwd1(const float* g, long long ldg, const float* p, long long ldp, const float* d, long long ldd,
                           float* gp, long long ldgp, float* gd, long long ldgd, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, float ns, double* sums, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (stat && !sums) return VNPCC_ERR_BAD_ARG;
    if (stat) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    if (P > 0 && C > 0 && try_bn_leaky_bwd1_v4(g, ldg, p, ldp, d, ldd, gp, ldgp, gd, ldgd, P, C, stat, gamma, beta, ns,
                                               stat ? sums : nullptr, st))
        return last_error();
    if (P > 0 && C > 0) {
        dim3 block(32, 8);
        const int gx = (C + 31) / 32;
        long long gy = (P + 8 * 16 - 1) / (8 * 16);
        const long long cap = ((long long)sm_count() * 16 + gx - 1) / gx;
        if (gy > cap) gy = cap;
        if (gy < 1) gy = 1;
        count_launch(), vn_bn_leaky_bwd1_kernel<<<dim3(gx, (unsigned)gy), block, 0, st>>>(g, (size_t)ldg, p, (size_t)ldp, d, (size_t)ldd, gp,
                                                                        (size_t)ldgp, gd, (size_t)ldgd, P, C, stat, gamma,
                                                                        beta, ns, stat ? sums : nullptr);
    }
    return last_error();
}

// dgamma = S2, dbeta = S1 are written to gweight / gbias (fp32) when non-NULL.
int vnpcc_vn_bn_bwd2(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat,
                     const float* gamma, const float* beta, const double* sums, double count, int training,
                     float* gweight, float* gbias, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 0) return 0;
    if (P > 0 && try_bn_bwd2_v4(gp, ldgp, p, ldp, P, C, stat, gamma, beta, sums, count, training, st)) {
    } else if (P > 0)
        count_launch(), vn_bn_bwd2_kernel<<<grid_for((size_t)P * C, 256, 16), 256, 0, st>>>(gp, (size_t)ldgp, p, (size_t)ldp, P, C, stat, gamma,
                                                                          beta, sums, count, training);
    if (gbias) count_launch(), double_to_float_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, gbias, C);
    if (gweight) count_launch(), double_to_float_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums + C, gweight, C);
    return last_error();
}

// vnpcc_vn_bn_bwd2 that additionally returns the gradient of the per-sample bias rows of the producing GEMM (the broadcast half of
// cat([global.expand(N), local]), models/pcn.py:172): gbias [B*3, 2C] (leading dimension ldgb, zeroed here) = per-sample column sums of the
// final gp (p half) and of gd (d half, [P*3, C] with pitch ldgd, final since bwd1).  VNPCC_ERR_UNSUPPORTED when the vectorised kernel does
// not take the shape (callers then run vnpcc_vn_bn_bwd2 + vnpcc_rows_sample_sum).
int vnpcc_vn_bn_bwd2_sbias(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, const double* sums, double count, int training, float* gweight,
                           float* gbn_bias, const float* gd, long long ldgd, float* gbias, long long ldgb, long long pts_per_sample,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 0 || P <= 0 || pts_per_sample <= 0 || P % pts_per_sample != 0) return VNPCC_ERR_UNSUPPORTED;
    const long long B = P / pts_per_sample;
    cudaMemset2DAsync(gbias, (size_t)ldgb * sizeof(float), 0, (size_t)2 * C * sizeof(float), (size_t)B * 3, st);
    if (!try_bn_bwd2_v4_sbias(gp, ldgp, p, ldp, P, C, stat, gamma, beta, sums, count, training, gd, ldgd, gbias, ldgb, pts_per_sample, st))
        return VNPCC_ERR_UNSUPPORTED;
    if (gbn_bias) count_launch(), double_to_float_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, gbn_bias, C);
    if (gweight) count_launch(), double_to_float_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums + C, gweight, C);
    return last_error();
}

// ws: B*C u64.  idx (int64 [B,C]) receives the selections.
int vnpcc_vn_maxpool_argmax(const float* x, long long ldx, const float* d, long long ldd, int B, int N, int C,
                            unsigned long long* ws, long long* idx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || C <= 0 || N <= 0) return 0;
    cudaMemsetAsync(ws, 0, sizeof(u64) * (size_t)B * C, st);
    const int gx = (C + 31) / 32;
    // (sample, chunk) blocks in whole waves of the kernel's resident CTAs; chunks of >= 64 points
    const int n_chunk = plan_chunk_len((long long)gx * B, N, (long long)sm_count() * resident_ctas(vn_maxpool_argmax_kernel, 256), 8, 64);
    const int chunks = (N + n_chunk - 1) / n_chunk;
    count_launch(), vn_maxpool_argmax_kernel<<<dim3(gx, (unsigned)(B * chunks)), dim3(32, 8), 0, st>>>(x, (size_t)ldx, d, (size_t)ldd, B, N, C,
                                                                                     n_chunk, ws);
    const long long total = (long long)B * C;
    count_launch(), vn_maxpool_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, total, idx);
    return last_error();
}

int vnpcc_vn_maxpool_gather(const float* x, long long ldx, const long long* idx, int B, int N, int C, float* out,
                            long long ldo, void* stream) {
    const long long total = (long long)B * C;
    if (total <= 0) return 0;
    count_launch(), vn_maxpool_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, idx, B, N, C,
                                                                                              out, (size_t)ldo);
    return last_error();
}

int vnpcc_vn_maxpool_scatter_add(const float* g, long long ldg, const long long* idx, int B, int N, int C, float* gx,
                                 long long ldgx, void* stream) {
    const long long total = (long long)B * C;
    if (total <= 0) return 0;
    count_launch(), vn_maxpool_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, (size_t)ldg, idx, B, N, C,
                                                                                               gx, (size_t)ldgx);
    return last_error();
}

int vnpcc_rows_add_sample_bias(float* y, long long ldy, const float* bias, long long ldb, int B, int N, int C,
                               void* stream) {
    const size_t total = (size_t)B * N * 3 * C;
    if (total == 0) return 0;
    count_launch(), rows_add_sample_bias_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(y, (size_t)ldy, bias, (size_t)ldb,
                                                                                          B, N, C);
    return last_error();
}

// out [B*3, C] (ld ldo) is zeroed here.
int vnpcc_rows_sample_sum(const float* g, long long ldg, int B, int N, int C, float* out, long long ldo, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || C <= 0) return 0;
    cudaMemset2DAsync(out, (size_t)ldo * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)B * 3, st);
    if (N <= 0) return last_error();
    const int gx = (C + 31) / 32;
    const int n_chunk = plan_chunk_len((long long)gx * B, N, (long long)sm_count() * resident_ctas(rows_sample_sum_kernel, 256), 8, 64);
    const int chunks = (N + n_chunk - 1) / n_chunk;
    count_launch(), rows_sample_sum_kernel<<<dim3(gx, (unsigned)(B * chunks)), dim3(32, 8), 0, st>>>(g, (size_t)ldg, B, N, C, n_chunk, out,
                                                                                   (size_t)ldo);
    return last_error();
}

int vnpcc_rows_dot(const float* x, long long ldx, const float* w, long long R, int C, const float* res, float* y,
                   void* stream) {
    if (R <= 0) return 0;
    count_launch(), rows_dot_kernel<<<grid_for((size_t)R * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, w, R, C, res, y);
    return last_error();
}

// gx[r,c] = gy[r]*w[c]  and  gw[c] = sum_r gy[r]*x[r,c]  (gw zeroed here; either output may be NULL)
int vnpcc_rows_dot_bwd(const float* gy, const float* x, long long ldx, const float* w, long long R, int C, float* gx,
                       long long ldgx, float* gw, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (gw) cudaMemsetAsync(gw, 0, sizeof(float) * C, st);
    if (R <= 0 || C <= 0) return last_error();
    if (gx) count_launch(), rows_outer_kernel<<<grid_for((size_t)R * C, 256, 16), 256, 0, st>>>(gy, w, R, C, gx, (size_t)ldgx);
    if (gw) {
        const int gxx = (C + 31) / 32;
        long long gy_ = (R + 8 * 64 - 1) / (8 * 64);
        const long long cap = ((long long)sm_count() * 8 + gxx - 1) / gxx;
        if (gy_ > cap) gy_ = cap;
        if (gy_ < 1) gy_ = 1;
        count_launch(), rows_wsum_kernel<<<dim3(gxx, (unsigned)gy_), dim3(32, 8), 0, st>>>(gy, x, (size_t)ldx, R, C, gw);
    }
    return last_error();
}

// ---- small-K VNLinear (1 <= K <= 4); Cout % 4 == 0 and 16-byte aligned y/gy/bias rows required ----
#define VNPCC_KS_DISPATCH(KS, CALL) \
    switch (KS) {                   \
        case 1: { constexpr int K_ = 1; CALL; } break; \
        case 2: { constexpr int K_ = 2; CALL; } break; \
        case 3: { constexpr int K_ = 3; CALL; } break; \
        default: { constexpr int K_ = 4; CALL; } break; \
    }

static bool smallk_ok(int K, int Cout, const void* a, long long lda, const void* b, long long ldb_) {
    return K >= 1 && K <= 4 && Cout >= 4 && (Cout & 3) == 0 && (lda & 3) == 0 && ((uintptr_t)a & 15) == 0 &&
           (b == nullptr || ((ldb_ & 3) == 0 && ((uintptr_t)b & 15) == 0));
}

int vnpcc_smallk_fwd(const float* x, long long ldx, const float* W, long long ldw, const float* bias, long long ldbias,
                     long long rows_per_sample, float* y, long long ldy, long long R, int K, int Cout, void* stream) {
    if (R <= 0) return 0;
    if (!smallk_ok(K, Cout, y, ldy, bias, ldbias) || (R % 3) != 0) return VNPCC_ERR_UNSUPPORTED;
    if (bias && (rows_per_sample <= 0 || rows_per_sample % 3 != 0)) return VNPCC_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = R / 3;
    const int gy = (Cout / 4 + 127) / 128;
    long long blocks = ((long long)sm_count() * 16 + gy - 1) / gy;
    long long ppb = (P + blocks - 1) / blocks;
    if (ppb < 16) ppb = 16;
    blocks = (P + ppb - 1) / ppb;
    VNPCC_KS_DISPATCH(K, (count_launch(), smallk_fwd_kernel<K_><<<dim3((unsigned)blocks, gy), 128, 0, st>>>(
                             x, (size_t)ldx, W, (size_t)ldw, bias, (size_t)ldbias, bias ? rows_per_sample / 3 : 1, y, (size_t)ldy, P, Cout,
                             ppb)));
    return last_error();
}

int vnpcc_smallk_dgrad(const float* gy, long long ldgy, const float* W, long long ldw, float* gx, long long ldgx, long long R,
                       int K, int Cout, void* stream) {
    if (R <= 0) return 0;
    if (!smallk_ok(K, Cout, gy, ldgy, nullptr, 0) || Cout > 1024) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for((size_t)R * 32, 256, 8);
    const int nql = (Cout / 4 + 31) / 32;
#define VNPCC_DG(NQL_) \
    VNPCC_KS_DISPATCH(K, (count_launch(), smallk_dgrad_kernel<K_, NQL_><<<grid, 256, 0, st>>>(gy, (size_t)ldgy, W, (size_t)ldw, gx, (size_t)ldgx, R, Cout)))
    if (nql <= 1) { VNPCC_DG(1); }
    else if (nql <= 2) { VNPCC_DG(2); }
    else if (nql <= 4) { VNPCC_DG(4); }
    else { VNPCC_DG(8); }
#undef VNPCC_DG
    return last_error();
}

// gW [Cout, K] (pitch ldgw) and, when gbias != NULL, gbias [B*3, Cout] are zeroed here and then accumulated.
// rows are grouped as B samples x N points x 3 (pass B = 1, N = R/3 when there is no bias).
int vnpcc_smallk_wgrad(const float* gy, long long ldgy, const float* x, long long ldx, int B, int N, int K, int Cout, float* gW,
                       long long ldgw, float* gbias, long long ldgb, void* stream) {
    if (K < 1 || K > 4 || Cout <= 0) return VNPCC_ERR_UNSUPPORTED;
    if (!smallk_ok(K, Cout, gy, ldgy, nullptr, 0)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemset2DAsync(gW, (size_t)ldgw * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)Cout, st);
    if (gbias) cudaMemset2DAsync(gbias, (size_t)ldgb * sizeof(float), 0, (size_t)Cout * sizeof(float), (size_t)B * 3, st);
    if (B <= 0 || N <= 0) return last_error();
    const int gx = (Cout / 4 + 31) / 32;
    int chunks = (int)(((long long)sm_count() * 8 + (long long)gx * B - 1) / ((long long)gx * B));
    if (chunks < 1) chunks = 1;
    int n_chunk = (N + chunks - 1) / chunks;
    if (n_chunk < 64) n_chunk = 64;
    chunks = (N + n_chunk - 1) / n_chunk;
    VNPCC_KS_DISPATCH(K, (count_launch(), smallk_wgrad_kernel<K_><<<dim3(gx, (unsigned)(B * chunks)), dim3(32, 8), 0, st>>>(
                             gy, (size_t)ldgy, x, (size_t)ldx, B, N, Cout, n_chunk, gW, (size_t)ldgw, gbias, (size_t)ldgb)));
    return last_error();
}

// backward of VNLinear -> VNMaxPool (see section 9).  g [B*3, C] rows (b,v); idx [B, C] int64; x [R, K]; W [C, K].
// gx [R, K] is zeroed here and then scattered into (pass NULL to skip); gW [C, K] is overwritten (pass NULL to skip).
int vnpcc_pool_linear_bwd(const float* g, long long ldg, const long long* idx, const float* x, long long ldx, const float* W,
                          long long ldw, int B, int N, int C, int K, float* gx, long long ldgx, float* gW, long long ldgw,
                          void* stream) {
    if (B <= 0 || C <= 0 || K <= 0) return 0;
    if ((K & 3) || (ldx & 3) || (ldw & 3) || (ldgx & 3) || (ldgw & 3) || ((uintptr_t)x & 15) || ((uintptr_t)W & 15) ||
        ((uintptr_t)gx & 15) || ((uintptr_t)gW & 15))
        return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (gx) {
        cudaMemset2DAsync(gx, (size_t)ldgx * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)B * N * 3, st);
        count_launch(), pool_linear_bwd_x_kernel<<<dim3((C + 7) / 8, B), 256, 0, st>>>(g, (size_t)ldg, idx, W, (size_t)ldw, B, N, C, K, gx,
                                                                                 (size_t)ldgx);
    }
    if (gW) count_launch(), pool_linear_bwd_w_kernel<<<C, 256, 0, st>>>(g, (size_t)ldg, idx, x, (size_t)ldx, B, N, C, K, gW, (size_t)ldgw);
    return last_error();
}

// Fused tail  VNLinearLeakyReLU -> VNLinear(C,1) (+ residual)  (models/pcn.py:340-345,387): the [R,C] activation of the
// last VNLinearLeakyReLU and its gradient are never written.  Returns VNPCC_ERR_UNSUPPORTED for shapes it does not take.
//   fwd : y[r] = sum_c leaky(BN(p), d)[r,c] * w2[c] (+ res[r])
//   bwd1: like vnpcc_vn_bn_leaky_bwd1 with g[r,c] = gy[r]*w2[c]; additionally gw2 (C doubles, zeroed here) accumulates
//         sum_r gy[r]*out[r,c].  Follow with vnpcc_vn_bn_bwd2 as usual.
int vnpcc_bn_leaky_dot_fwd(const float* p, long long ldp, const float* d, long long ldd, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, float ns, const float* w2, const float* res, float* y,
                           void* stream) {
    if (P <= 0 || C <= 0) return 0;
    if (!try_bn_leaky_dot_fwd_v4(p, ldp, d, ldd, P, C, stat, gamma, beta, ns, w2, res, y, (cudaStream_t)stream))
        return VNPCC_ERR_UNSUPPORTED;
    return last_error();
}

int vnpcc_bn_leaky_dot_bwd1(const float* gy, const float* p, long long ldp, const float* d, long long ldd, float* gp, long long ldgp,
                            float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma, const float* beta,
                            float ns, double* sums, const float* w2, double* gw2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (stat && !sums) return VNPCC_ERR_BAD_ARG;
    if (stat) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    cudaMemsetAsync(gw2, 0, sizeof(double) * C, st);
    if (P <= 0 || C <= 0) return last_error();
    if (!try_bn_leaky_dot_bwd1_v4(gy, p, ldp, d, ldd, gp, ldgp, gd, ldgd, P, C, stat, gamma, beta, ns, stat ? sums : nullptr, w2, gw2, st))
        return VNPCC_ERR_UNSUPPORTED;
    return last_error();
}

int vnpcc_double_to_float(const double* in, float* out, int n, void* stream) {
    if (n <= 0) return 0;
    count_launch(), double_to_float_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, out, n);
    return last_error();
}

int vnpcc_vn_maxpool_decode(const unsigned long long* best, long long total, long long* idx, void* stream) {
    if (total <= 0) return 0;
    count_launch(), vn_maxpool_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(best, total, idx);
    return last_error();
}

int vnpcc_pool_linear_gather(const float* x, long long ldx, const float* W, long long ldw, const long long* idx, int B, int N, int C, int K,
                             float* out, long long ldo, void* stream) {
    if (B <= 0 || C <= 0) return 0;
    if ((K & 3) || (ldx & 3) || (ldw & 3) || ((uintptr_t)x & 15) || ((uintptr_t)W & 15)) return VNPCC_ERR_UNSUPPORTED;
    const long long warps = (long long)B * C;
    count_launch(), pool_linear_gather_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, (size_t)ldx, W, (size_t)ldw, idx, B, N, C, K, out, (size_t)ldo);
    return last_error();
}

}  // extern "C"
