// attention.cu -- kernels of the transformer-refined VN folding decoder for sm_100a:
//   vn_layernorm_fwd / _bwd   VNLayerNorm                                   models/vn_layers.py:129-150
//   vn_attention_fwd / _bwd   the VN multi-head attention core of Attention  models/transformer.py:89-100
//                             (softmax(q k^T * scale) v over tokens, per head, on the Frobenius inner product of the
//                             head's [C/H, 3] vector features), flash-style: the [N, N] score matrix never reaches HBM
//   rows_add                  the residual additions of VN_Block            models/transformer.py:60,68
//
// Everything works on the channels-last ROW layout (include/vnpcc.h): token (b, n) = 3 consecutive rows (v = 0..2), channel
// c of head h at column h*D + c.  A head's 3*D-dimensional feature of a token is therefore three D-float segments; the
// kernels read q / k / v straight out of the stacked projection output  qkv[R, 3C] = (q | k | v)  and write the attention
// output in the same layout, so no head-major permutation (models/transformer.py:89-91,98-99) is ever materialised.
//
// This is the exact-fp32 (parity mode) attention: SIMT FMAs, register-tiled 64 x 64 score tiles, online softmax.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

// -------------------------------------------------------------------------------------------------------------
// 1. VNLayerNorm.  One warp per token: norm[c] = ||x[c, :]|| + eps ; LayerNorm over the C channels (biased variance,
//    eps 1e-5, affine) ; y = x / norm * ln(norm).
// -------------------------------------------------------------------------------------------------------------
constexpr int LN_CPL = 16;          // channels per lane (C <= 512)
constexpr float LN_VEPS = 1e-6f;    // models/vn_layers.py:10

__global__ void __launch_bounds__(256) vn_layernorm_fwd_kernel(const float* __restrict__ x, size_t ldx, long long P, int C,
                                                              const float* __restrict__ w, const float* __restrict__ bvec, float ln_eps,
                                                              float* __restrict__ y, size_t ldy, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < P; t += nwarps) {
        const float* xp = x + (size_t)t * 3 * ldx;
        float xv[LN_CPL][3], nr[LN_CPL];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_CPL; ++i) {
            const int c = lane + 32 * i;
            nr[i] = 0.f;
            if (c < C) {
                xv[i][0] = __ldg(xp + c), xv[i][1] = __ldg(xp + ldx + c), xv[i][2] = __ldg(xp + 2 * ldx + c);
                nr[i] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xv[i][0], xv[i][0]), __fmul_rn(xv[i][1], xv[i][1])), __fmul_rn(xv[i][2], xv[i][2]))) +
                        LN_VEPS;
                sum += nr[i];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)C;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < LN_CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float d = nr[i] - mean;
                var = fmaf(d, d, var);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = 1.0f / sqrtf(var / (float)C + ln_eps);
        if (stats && lane == 0) {
            stats[2 * t] = mean;
            stats[2 * t + 1] = rstd;
        }
        float* yp = y + (size_t)t * 3 * ldy;
#pragma unroll
        for (int i = 0; i < LN_CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float l = (nr[i] - mean) * rstd * __ldg(w + c) + __ldg(bvec + c);
                yp[c] = xv[i][0] / nr[i] * l;
                yp[ldy + c] = xv[i][1] / nr[i] * l;
                yp[2 * ldy + c] = xv[i][2] / nr[i] * l;
            }
        }
    }
}

// backward: a = <g, x> / n ; dl = a ; dnhat = a * w ; dn = rstd * (dnhat - mean_c(dnhat) - nhat * mean_c(dnhat * nhat)) - a * l / n ;
// gx = g * (l / n) + dn * x / r ; gw += a * nhat ; gb += a
__global__ void __launch_bounds__(256) vn_layernorm_bwd_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ x, size_t ldx,
                                                              long long P, int C, const float* __restrict__ w, const float* __restrict__ bvec,
                                                              const float* __restrict__ stats, float* __restrict__ gx, size_t ldgx,
                                                              float* __restrict__ gw, float* __restrict__ gb) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float aw[LN_CPL], ab[LN_CPL];
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) aw[i] = ab[i] = 0.f;
    for (long long t = warp; t < P; t += nwarps) {
        const float* xp = x + (size_t)t * 3 * ldx;
        const float* gp = g + (size_t)t * 3 * ldg;
        const float mean = __ldg(stats + 2 * t), rstd = __ldg(stats + 2 * t + 1);
        float xv[LN_CPL][3], gv[LN_CPL][3], n[LN_CPL], nh[LN_CPL], a[LN_CPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                xv[i][0] = __ldg(xp + c), xv[i][1] = __ldg(xp + ldx + c), xv[i][2] = __ldg(xp + 2 * ldx + c);
                gv[i][0] = __ldg(gp + c), gv[i][1] = __ldg(gp + ldg + c), gv[i][2] = __ldg(gp + 2 * ldg + c);
                const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xv[i][0], xv[i][0]), __fmul_rn(xv[i][1], xv[i][1])), __fmul_rn(xv[i][2], xv[i][2])));
                n[i] = r + LN_VEPS;
                nh[i] = (n[i] - mean) * rstd;
                a[i] = (gv[i][0] * xv[i][0] + gv[i][1] * xv[i][1] + gv[i][2] * xv[i][2]) / n[i];
                const float dnh = a[i] * __ldg(w + c);
                s1 += dnh;
                s2 = fmaf(dnh, nh[i], s2);
                aw[i] = fmaf(a[i], nh[i], aw[i]);
                ab[i] += a[i];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / (float)C, m2 = s2 / (float)C;
        float* op = gx + (size_t)t * 3 * ldgx;
#pragma unroll
        for (int i = 0; i < LN_CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float wc = __ldg(w + c);
                const float l = nh[i] * wc + __ldg(bvec + c);
                const float dn = rstd * (a[i] * wc - m1 - nh[i] * m2) - a[i] * l / n[i];
                const float r = n[i] - LN_VEPS;
                const float dr = r > 0.f ? dn / r : 0.f;
                const float sc = l / n[i];
                op[c] = fmaf(gv[i][0], sc, dr * xv[i][0]);
                op[ldgx + c] = fmaf(gv[i][1], sc, dr * xv[i][1]);
                op[2 * ldgx + c] = fmaf(gv[i][2], sc, dr * xv[i][2]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LN_CPL; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
            atomicAdd(gw + c, aw[i]);
            atomicAdd(gb + c, ab[i]);
        }
    }
}

// ---- vectorised VNLayerNorm (C % 4 == 0, 16-byte aligned rows): a lane owns float4 groups of channels {4 (lane + 32 i) .. + 3}
constexpr int LN_V4 = 4;            // float4 groups per lane (C <= 512)

__global__ void __launch_bounds__(256) vn_layernorm_fwd_v4_kernel(const float* __restrict__ x, size_t ldx, long long P, int C,
                                                                 const float* __restrict__ w, const float* __restrict__ bvec, float ln_eps,
                                                                 float* __restrict__ y, size_t ldy, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int G = C >> 2;
    float4 wv[LN_V4], bv[LN_V4];
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
        const int gi = lane + 32 * i;
        if (gi < G) {
            wv[i] = __ldg(reinterpret_cast<const float4*>(w) + gi);
            bv[i] = __ldg(reinterpret_cast<const float4*>(bvec) + gi);
        }
    }
    for (long long t = warp; t < P; t += nwarps) {
        const float* xp = x + (size_t)t * 3 * ldx;
        float4 xv[LN_V4][3], nr[LN_V4];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
#pragma unroll
                for (int v = 0; v < 3; ++v) xv[i][v] = __ldg(reinterpret_cast<const float4*>(xp + v * ldx) + gi);
#define LN_NORM(f) (sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xv[i][0].f, xv[i][0].f), __fmul_rn(xv[i][1].f, xv[i][1].f)), __fmul_rn(xv[i][2].f, xv[i][2].f))) + LN_VEPS)
                nr[i] = make_float4(LN_NORM(x), LN_NORM(y), LN_NORM(z), LN_NORM(w));
#undef LN_NORM
                sum += (nr[i].x + nr[i].y) + (nr[i].z + nr[i].w);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)C;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
                const float dx = nr[i].x - mean, dy = nr[i].y - mean, dz = nr[i].z - mean, dw = nr[i].w - mean;
                var = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, var))));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = 1.0f / sqrtf(var / (float)C + ln_eps);
        if (stats && lane == 0) {
            stats[2 * t] = mean;
            stats[2 * t + 1] = rstd;
        }
        float* yp = y + (size_t)t * 3 * ldy;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
                const float4 l = make_float4((nr[i].x - mean) * rstd * wv[i].x + bv[i].x, (nr[i].y - mean) * rstd * wv[i].y + bv[i].y,
                                             (nr[i].z - mean) * rstd * wv[i].z + bv[i].z, (nr[i].w - mean) * rstd * wv[i].w + bv[i].w);
#pragma unroll
                for (int v = 0; v < 3; ++v)
                    reinterpret_cast<float4*>(yp + v * ldy)[gi] = make_float4(xv[i][v].x / nr[i].x * l.x, xv[i][v].y / nr[i].y * l.y,
                                                                              xv[i][v].z / nr[i].z * l.z, xv[i][v].w / nr[i].w * l.w);
            }
        }
    }
}

__global__ void __launch_bounds__(256) vn_layernorm_bwd_v4_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ x, size_t ldx,
                                                                 long long P, int C, const float* __restrict__ w, const float* __restrict__ bvec,
                                                                 const float* __restrict__ stats, float* __restrict__ gx, size_t ldgx,
                                                                 float* __restrict__ gw, float* __restrict__ gb) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int G = C >> 2;
    float wv[LN_V4][4], bv[LN_V4][4], aw[LN_V4][4], ab[LN_V4][4];
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
        const int gi = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < 4; ++e) aw[i][e] = ab[i][e] = wv[i][e] = bv[i][e] = 0.f;
        if (gi < G) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w) + gi), b4 = __ldg(reinterpret_cast<const float4*>(bvec) + gi);
            wv[i][0] = a.x, wv[i][1] = a.y, wv[i][2] = a.z, wv[i][3] = a.w;
            bv[i][0] = b4.x, bv[i][1] = b4.y, bv[i][2] = b4.z, bv[i][3] = b4.w;
        }
    }
    for (long long t = warp; t < P; t += nwarps) {
        const float* xp = x + (size_t)t * 3 * ldx;
        const float* gp = g + (size_t)t * 3 * ldg;
        const float mean = __ldg(stats + 2 * t), rstd = __ldg(stats + 2 * t + 1);
        float xs[LN_V4][3][4], gs[LN_V4][3][4], n[LN_V4][4], nh[LN_V4][4], a[LN_V4][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(xp + v * ldx) + gi), g4 = __ldg(reinterpret_cast<const float4*>(gp + v * ldg) + gi);
                    xs[i][v][0] = a4.x, xs[i][v][1] = a4.y, xs[i][v][2] = a4.z, xs[i][v][3] = a4.w;
                    gs[i][v][0] = g4.x, gs[i][v][1] = g4.y, gs[i][v][2] = g4.z, gs[i][v][3] = g4.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xs[i][0][e], xs[i][0][e]), __fmul_rn(xs[i][1][e], xs[i][1][e])),
                                                    __fmul_rn(xs[i][2][e], xs[i][2][e])));
                    n[i][e] = r + LN_VEPS;
                    nh[i][e] = (n[i][e] - mean) * rstd;
                    a[i][e] = (gs[i][0][e] * xs[i][0][e] + gs[i][1][e] * xs[i][1][e] + gs[i][2][e] * xs[i][2][e]) / n[i][e];
                    const float dnh = a[i][e] * wv[i][e];
                    s1 += dnh;
                    s2 = fmaf(dnh, nh[i][e], s2);
                    aw[i][e] = fmaf(a[i][e], nh[i][e], aw[i][e]);
                    ab[i][e] += a[i][e];
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / (float)C, m2 = s2 / (float)C;
        float* op = gx + (size_t)t * 3 * ldgx;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int gi = lane + 32 * i;
            if (gi < G) {
                float sc[4], dr[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float l = nh[i][e] * wv[i][e] + bv[i][e];
                    const float dn = rstd * (a[i][e] * wv[i][e] - m1 - nh[i][e] * m2) - a[i][e] * l / n[i][e];
                    const float r = n[i][e] - LN_VEPS;
                    dr[e] = r > 0.f ? dn / r : 0.f;
                    sc[e] = l / n[i][e];
                }
#pragma unroll
                for (int v = 0; v < 3; ++v)
                    reinterpret_cast<float4*>(op + v * ldgx)[gi] =
                        make_float4(fmaf(gs[i][v][0], sc[0], dr[0] * xs[i][v][0]), fmaf(gs[i][v][1], sc[1], dr[1] * xs[i][v][1]),
                                    fmaf(gs[i][v][2], sc[2], dr[2] * xs[i][v][2]), fmaf(gs[i][v][3], sc[3], dr[3] * xs[i][v][3]));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
        const int gi = lane + 32 * i;
        if (gi < G) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                atomicAdd(gw + gi * 4 + e, aw[i][e]);
                atomicAdd(gb + gi * 4 + e, ab[i][e]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) rows_add_kernel(const float* __restrict__ a, size_t lda, const float* __restrict__ b, size_t ldb,
                                                      float* __restrict__ out, size_t ldo, long long R, int C) {
    const long long total = R * C;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long long r = t / C;
        const int c = (int)(t - r * C);
        out[(size_t)r * ldo + c] = __ldg(a + (size_t)r * lda + c) + __ldg(b + (size_t)r * ldb + c);
    }
}

// -------------------------------------------------------------------------------------------------------------
// 2. attention.  Tiles of 64 tokens; 256 threads as a 16 x 16 grid (ty, tx); a thread owns score rows ty + 16 i and
//    score columns tx + 16 j (i, j < 4), and columns tx + 16 jj of the [64, 3D] feature tiles (jj < 3D / 16).
//    Feature tiles live in shared memory with a row stride of 3D + 4 floats ((3D + 4) / 4 odd: conflict-free LDS.128
//    for 8 consecutive rows); element (v, c) of a token's head feature is stored at v * D + c.
// -------------------------------------------------------------------------------------------------------------
namespace att {
constexpr int BT = 64;          // tokens per tile
constexpr int NT = 256;
constexpr int PLD = BT + 4;     // row stride of the score tiles

template <int D>
struct Cfg {
    static constexpr int DH = 3 * D;
    static constexpr int LD = DH + 4;
    static constexpr int NC = DH / 16;
    static_assert(D % 16 == 0, "head channels must be a multiple of 16");
    static_assert(((DH + 4) / 4) % 2 == 1, "row stride must be an odd number of float4");
};

// tile of BT tokens starting at n0: s[t * LD + v * D + c] = g[((n0 + t) * 3 + v) * ld + c]   (zeros beyond N)
template <int D>
__device__ __forceinline__ void load_tile(float* __restrict__ s, const float* __restrict__ g, size_t ld, int n0, int N) {
    constexpr int LD = Cfg<D>::LD, Q4 = D / 4;
    for (int i = threadIdx.x; i < BT * 3 * Q4; i += NT) {
        const int t = i / (3 * Q4);
        const int rem = i - t * (3 * Q4);
        const int v = rem / Q4, c4 = rem - v * Q4;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + t < N) val = __ldg(reinterpret_cast<const float4*>(g + ((size_t)(n0 + t) * 3 + v) * ld + c4 * 4));
        *reinterpret_cast<float4*>(s + t * LD + v * D + c4 * 4) = val;
    }
}

// acc[i][j] = sum_d A[(ty + 16 i), d] * B[(tx + 16 j), d]
template <int D>
__device__ __forceinline__ void tile_abt(const float* __restrict__ A, const float* __restrict__ Bm, int ty, int tx, float (&acc)[4][4]) {
    constexpr int LD = Cfg<D>::LD, DH = Cfg<D>::DH;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int d = 0; d < DH; d += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty + 16 * i) * LD + d);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bm + (tx + 16 * j) * LD + d);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
}

// o[i][jj] += sum_k S[(ty + 16 i), k] * Bm[k, tx + 16 jj]      (S: score tile, row stride PLD)
template <int D>
__device__ __forceinline__ void tile_sb(const float* __restrict__ S, const float* __restrict__ Bm, int ty, int tx, float (&o)[4][Cfg<D>::NC]) {
    constexpr int LD = Cfg<D>::LD, NC = Cfg<D>::NC;
#pragma unroll 2
    for (int k = 0; k < BT; k += 4) {
        float4 p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4*>(S + (ty + 16 * i) * PLD + k);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float bv[NC];
#pragma unroll
            for (int jj = 0; jj < NC; ++jj) bv[jj] = Bm[(k + t) * LD + tx + 16 * jj];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float pv = t == 0 ? p[i].x : (t == 1 ? p[i].y : (t == 2 ? p[i].z : p[i].w));
#pragma unroll
                for (int jj = 0; jj < NC; ++jj) o[i][jj] = fmaf(pv, bv[jj], o[i][jj]);
            }
        }
    }
}

// o[i][jj] += sum_q S[q, (ty + 16 i)] * Bm[q, tx + 16 jj]      (transposed use of the score tile)
template <int D>
__device__ __forceinline__ void tile_stb(const float* __restrict__ S, const float* __restrict__ Bm, int ty, int tx, float (&o)[4][Cfg<D>::NC]) {
    constexpr int LD = Cfg<D>::LD, NC = Cfg<D>::NC;
#pragma unroll 4
    for (int q = 0; q < BT; ++q) {
        float sv[4], bv[NC];
#pragma unroll
        for (int i = 0; i < 4; ++i) sv[i] = S[q * PLD + ty + 16 * i];
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) bv[jj] = Bm[q * LD + tx + 16 * jj];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < NC; ++jj) o[i][jj] = fmaf(sv[i], bv[jj], o[i][jj]);
    }
}

template <int D>
__global__ void __launch_bounds__(NT, 1) attn_fwd_kernel(const float* __restrict__ qkv, size_t ld, int N, int H, int C, float scale,
                                                        float* __restrict__ out, size_t ldo, float* __restrict__ lse) {
    constexpr int LD = Cfg<D>::LD, NC = Cfg<D>::NC;
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;
    float* Ks = Qs + BT * LD;
    float* Vs = Ks + BT * LD;
    float* Ps = Vs + BT * LD;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int q0 = blockIdx.x * BT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* base = qkv + (size_t)b * N * 3 * ld + (size_t)h * D;
    load_tile<D>(Qs, base, ld, q0, N);
    float m[4], l[4], o[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.f;
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) o[i][jj] = 0.f;
    }
    for (int k0 = 0; k0 < N; k0 += BT) {
        __syncthreads();
        load_tile<D>(Ks, base + C, ld, k0, N);
        load_tile<D>(Vs, base + 2 * C, ld, k0, N);
        __syncthreads();
        float s[4][4];
        tile_abt<D>(Qs, Ks, ty, tx, s);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[i][j] = (k0 + tx + 16 * j < N) ? s[i][j] * scale : -INFINITY;
                rmax = fmaxf(rmax, s[i][j]);
            }
#pragma unroll
            for (int of = 8; of > 0; of >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, of));
            const float mn = fmaxf(m[i], rmax);
            const float corr = expf(m[i] - mn);
            float rsum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = expf(s[i][j] - mn);
                rsum += p;
                Ps[(ty + 16 * i) * PLD + tx + 16 * j] = p;
            }
#pragma unroll
            for (int of = 8; of > 0; of >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, of);
            l[i] = fmaf(l[i], corr, rsum);
            m[i] = mn;
#pragma unroll
            for (int jj = 0; jj < NC; ++jj) o[i][jj] *= corr;
        }
        __syncwarp();          // a score row is written and read by the same half-warp
        tile_sb<D>(Ps, Vs, ty, tx, o);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = q0 + ty + 16 * i;
        if (n >= N) continue;
        const float inv = 1.0f / l[i];
        float* orow = out + ((size_t)b * N + n) * 3 * ldo + (size_t)h * D;
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
            const int col = tx + 16 * jj;
            const int v = col / D, c = col - v * D;
            orow[(size_t)v * ldo + c] = o[i][jj] * inv;
        }
        if (tx == 0) lse[(size_t)bh * N + n] = m[i] + logf(l[i]);
    }
}

// delta[b, h, n] = sum over the head's features of dO * O
__global__ void __launch_bounds__(256) attn_delta_kernel(const float* __restrict__ dO, size_t lddo, const float* __restrict__ O, size_t ldo, int B,
                                                        int N, int H, int D, float* __restrict__ delta) {
    const long long total = (long long)B * N * H;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int h = (int)(t % H);
    const long long bn = t / H;
    const int b = (int)(bn / N), n = (int)(bn - (long long)b * N);
    float s = 0.f;
    for (int v = 0; v < 3; ++v) {
        const float* a = dO + ((size_t)bn * 3 + v) * lddo + (size_t)h * D;
        const float* o = O + ((size_t)bn * 3 + v) * ldo + (size_t)h * D;
        for (int c = 0; c < D; c += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a + c));
            const float4 y = __ldg(reinterpret_cast<const float4*>(o + c));
            s = fmaf(x.x, y.x, s);
            s = fmaf(x.y, y.y, s);
            s = fmaf(x.z, y.z, s);
            s = fmaf(x.w, y.w, s);
        }
    }
    delta[((size_t)b * H + h) * N + n] = s;
}

// one CTA per (key tile, batch, head): dK, dV accumulate in registers over all query tiles, dQ goes out with red.add
template <int D>
__global__ void __launch_bounds__(NT, 1) attn_bwd_kernel(const float* __restrict__ qkv, size_t ld, const float* __restrict__ dO, size_t lddo,
                                                        const float* __restrict__ lse, const float* __restrict__ delta, int N, int H, int C,
                                                        float scale, float* __restrict__ dqkv, size_t lddq) {
    constexpr int LD = Cfg<D>::LD, NC = Cfg<D>::NC;
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;
    float* Vs = Ks + BT * LD;
    float* Qs = Vs + BT * LD;
    float* dOs = Qs + BT * LD;
    float* Ps = dOs + BT * LD;
    float* dSs = Ps + BT * PLD;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int k0 = blockIdx.x * BT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* base = qkv + (size_t)b * N * 3 * ld + (size_t)h * D;
    const float* dobase = dO + (size_t)b * N * 3 * lddo + (size_t)h * D;
    float* dbase = dqkv + (size_t)b * N * 3 * lddq + (size_t)h * D;
    load_tile<D>(Ks, base + C, ld, k0, N);
    load_tile<D>(Vs, base + 2 * C, ld, k0, N);
    float dk[4][NC], dv[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) dk[i][jj] = dv[i][jj] = 0.f;
    for (int q0 = 0; q0 < N; q0 += BT) {
        __syncthreads();
        load_tile<D>(Qs, base, ld, q0, N);
        load_tile<D>(dOs, dobase, lddo, q0, N);
        __syncthreads();
        {
            float s[4][4], dp[4][4];
            tile_abt<D>(Qs, Ks, ty, tx, s);
            tile_abt<D>(dOs, Vs, ty, tx, dp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = q0 + ty + 16 * i;
                const float L = q < N ? __ldg(lse + (size_t)bh * N + q) : 0.f;
                const float dl = q < N ? __ldg(delta + (size_t)bh * N + q) : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = q < N && (k0 + tx + 16 * j) < N;
                    const float p = ok ? expf(s[i][j] * scale - L) : 0.f;
                    Ps[(ty + 16 * i) * PLD + tx + 16 * j] = p;
                    dSs[(ty + 16 * i) * PLD + tx + 16 * j] = p * (dp[i][j] - dl) * scale;
                }
            }
        }
        __syncthreads();
        tile_stb<D>(Ps, dOs, ty, tx, dv);        // dV[k] += sum_q P[q,k] dO[q]
        tile_stb<D>(dSs, Qs, ty, tx, dk);        // dK[k] += sum_q dS[q,k] Q[q]
        float dq[4][NC];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < NC; ++jj) dq[i][jj] = 0.f;
        tile_sb<D>(dSs, Ks, ty, tx, dq);         // dQ[q] += sum_k dS[q,k] K[k]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + ty + 16 * i;
            if (q >= N) continue;
#pragma unroll
            for (int jj = 0; jj < NC; ++jj) {
                const int col = tx + 16 * jj;
                const int v = col / D, c = col - v * D;
                atomicAdd(dbase + ((size_t)q * 3 + v) * lddq + c, dq[i][jj]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty + 16 * i;
        if (k >= N) continue;
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
            const int col = tx + 16 * jj;
            const int v = col / D, c = col - v * D;
            dbase[((size_t)k * 3 + v) * lddq + C + c] = dk[i][jj];
            dbase[((size_t)k * 3 + v) * lddq + 2 * C + c] = dv[i][jj];
        }
    }
}

template <int D>
int launch_fwd(const float* qkv, size_t ld, int B, int N, int H, int C, float scale, float* out, size_t ldo, float* lse, cudaStream_t st) {
    const size_t smem = (size_t)(3 * BT * Cfg<D>::LD + BT * PLD) * sizeof(float);
    if (cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VNPCC_ERR_DRIVER;
    dim3 grid((unsigned)((N + BT - 1) / BT), (unsigned)(B * H));
    count_launch(), attn_fwd_kernel<D><<<grid, NT, smem, st>>>(qkv, ld, N, H, C, scale, out, ldo, lse);
    return last_error();
}

template <int D>
int launch_bwd(const float* qkv, size_t ld, const float* dO, size_t lddo, const float* lse, const float* delta, int B, int N, int H, int C,
               float scale, float* dqkv, size_t lddq, cudaStream_t st) {
    const size_t smem = (size_t)(4 * BT * Cfg<D>::LD + 2 * BT * PLD) * sizeof(float);
    if (cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return VNPCC_ERR_DRIVER;
    dim3 grid((unsigned)((N + BT - 1) / BT), (unsigned)(B * H));
    count_launch(), attn_bwd_kernel<D><<<grid, NT, smem, st>>>(qkv, ld, dO, lddo, lse, delta, N, H, C, scale, dqkv, lddq);
    return last_error();
}
}  // namespace att

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

int vnpcc_vn_layernorm_fwd(const float* x, long long ldx, long long P, int C, const float* weight, const float* bias, float ln_eps, float* y,
                           long long ldy, float* stats, void* stream) {
    if (C <= 0 || C > 32 * LN_CPL) return VNPCC_ERR_UNSUPPORTED;
    if (P <= 0) return 0;
    const int grid = grid_for((size_t)P * 32, 256, 8);
    const bool v4 = C % 4 == 0 && C <= 128 * LN_V4 && ldx % 4 == 0 && ldy % 4 == 0 &&
                    !(((uintptr_t)x | (uintptr_t)y | (uintptr_t)weight | (uintptr_t)bias) & 15);
    if (v4)
        count_launch(), vn_layernorm_fwd_v4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, P, C, weight, bias, ln_eps, y, (size_t)ldy,
                                                                                        stats);
    else
        count_launch(), vn_layernorm_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, P, C, weight, bias, ln_eps, y, (size_t)ldy,
                                                                                     stats);
    return last_error();
}

// gweight / gbias [C] are zeroed here, then accumulated
int vnpcc_vn_layernorm_bwd(const float* g, long long ldg, const float* x, long long ldx, long long P, int C, const float* weight,
                           const float* bias, const float* stats, float* gx, long long ldgx, float* gweight, float* gbias, void* stream) {
    if (C <= 0 || C > 32 * LN_CPL) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(gweight, 0, sizeof(float) * C, st);
    cudaMemsetAsync(gbias, 0, sizeof(float) * C, st);
    if (P <= 0) return last_error();
    const int grid = grid_for((size_t)P * 32, 256, 4);
    const bool v4 = C % 4 == 0 && C <= 128 * LN_V4 && ldx % 4 == 0 && ldg % 4 == 0 && ldgx % 4 == 0 &&
                    !(((uintptr_t)x | (uintptr_t)g | (uintptr_t)gx | (uintptr_t)weight | (uintptr_t)bias) & 15);
    if (v4)
        count_launch(), vn_layernorm_bwd_v4_kernel<<<grid, 256, 0, st>>>(g, (size_t)ldg, x, (size_t)ldx, P, C, weight, bias, stats, gx, (size_t)ldgx,
                                                                        gweight, gbias);
    else
        count_launch(), vn_layernorm_bwd_kernel<<<grid, 256, 0, st>>>(g, (size_t)ldg, x, (size_t)ldx, P, C, weight, bias, stats, gx, (size_t)ldgx,
                                                                     gweight, gbias);
    return last_error();
}

int vnpcc_rows_add(const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, long long R, int C, void* stream) {
    if (R <= 0 || C <= 0) return 0;
    count_launch(), rows_add_kernel<<<grid_for((size_t)R * C, 256, 16), 256, 0, (cudaStream_t)stream>>>(a, (size_t)lda, b, (size_t)ldb, out, (size_t)ldo,
                                                                                               R, C);
    return last_error();
}

int vnpcc_vn_attention_fwd(const float* qkv, long long ld, int B, int N, int H, int D, float scale, float* out, long long ldo, float* lse,
                           void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (H <= 0 || ld % 4 != 0 || ((uintptr_t)qkv & 15)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = H * D;
    switch (D) {
        case 16: return att::launch_fwd<16>(qkv, (size_t)ld, B, N, H, C, scale, out, (size_t)ldo, lse, st);
        case 32: return att::launch_fwd<32>(qkv, (size_t)ld, B, N, H, C, scale, out, (size_t)ldo, lse, st);
        case 48: return att::launch_fwd<48>(qkv, (size_t)ld, B, N, H, C, scale, out, (size_t)ldo, lse, st);
        default: return VNPCC_ERR_UNSUPPORTED;
    }
}

// dqkv [R, 3C]: the q part is zeroed here and accumulated with fp32 atomics, the k / v parts are plain stores.
// delta: workspace of B*H*N floats.
int vnpcc_vn_attention_bwd(const float* qkv, long long ld, const float* dout, long long lddo, const float* out, long long ldo, const float* lse,
                           int B, int N, int H, int D, float scale, float* dqkv, long long lddq, float* delta, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (H <= 0 || ld % 4 != 0 || lddo % 4 != 0 || ldo % 4 != 0 || ((uintptr_t)qkv & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)out & 15))
        return VNPCC_ERR_UNSUPPORTED;
    if (D != 16 && D != 32 && D != 48) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = H * D;
    cudaMemset2DAsync(dqkv, (size_t)lddq * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)B * N * 3, st);
    const long long total = (long long)B * N * H;
    count_launch(), att::attn_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dout, (size_t)lddo, out, (size_t)ldo, B, N, H, D, delta);
    switch (D) {
        case 16: return att::launch_bwd<16>(qkv, (size_t)ld, dout, (size_t)lddo, lse, delta, B, N, H, C, scale, dqkv, (size_t)lddq, st);
        case 32: return att::launch_bwd<32>(qkv, (size_t)ld, dout, (size_t)lddo, lse, delta, B, N, H, C, scale, dqkv, (size_t)lddq, st);
        default: return att::launch_bwd<48>(qkv, (size_t)ld, dout, (size_t)lddo, lse, delta, B, N, H, C, scale, dqkv, (size_t)lddq, st);
    }
}

// delta[b, h, n] = sum over head h's features of dout * out  (shared by the SIMT and the tensor-core backward)
int vnpcc_vn_attention_delta(const float* dout, long long lddo, const float* out, long long ldo, int B, int N, int H, int D, float* delta,
                             void* stream) {
    const long long total = (long long)B * N * H;
    if (total <= 0) return 0;
    if (D % 4 != 0 || lddo % 4 != 0 || ldo % 4 != 0 || ((uintptr_t)dout & 15) || ((uintptr_t)out & 15)) return VNPCC_ERR_UNSUPPORTED;
    count_launch(), att::attn_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dout, (size_t)lddo, out, (size_t)ldo, B, N,
                                                                                                     H, D, delta);
    return last_error();
}

}  // extern "C"
