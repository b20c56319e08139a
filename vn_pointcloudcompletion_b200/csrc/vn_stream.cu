// vn_stream.cu -- 128-bit vectorised versions of the HBM-bound Vector-Neuron row kernels (norm statistics, BatchNorm
// + leaky projection forward / backward).  Same arithmetic as the scalar kernels in vn_kernels.cu (which remain the
// path for channel counts / pitches that are not multiples of 4); each thread owns FOUR consecutive channels, so a warp
// reads 512 contiguous bytes per row, per-channel parameters live in registers, and up to nine 16-byte loads are in
// flight per thread per point.
//
// Reference semantics: VNBatchNorm models/vn_layers.py:116-127, leaky projection :39-42/:70-73; backward formulas
// SURVEY.md Appendix C.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vnpcc_internal.h"
#include "vn_math.cuh"

namespace vnpcc {

// block (32, 8): lane -> channel quad, y -> point lane; grid.x tiles quads by 32, grid.y strides over points
#define VS_BLOCK_REDUCE2(S1, S2, sums, C, c0)                                   \
    {                                                                           \
        __shared__ double sh[2][8][32][4];                                      \
        _Pragma("unroll") for (int l = 0; l < 4; ++l) {                         \
            sh[0][threadIdx.y][threadIdx.x][l] = S1[l];                         \
            sh[1][threadIdx.y][threadIdx.x][l] = S2[l];                         \
        }                                                                       \
        __syncthreads();                                                        \
        if (threadIdx.y == 0 && active) {                                       \
            _Pragma("unroll") for (int l = 0; l < 4; ++l) {                     \
                double a = 0.0, b = 0.0;                                        \
                for (int k = 0; k < 8; ++k) {                                   \
                    a += sh[0][k][threadIdx.x][l];                              \
                    b += sh[1][k][threadIdx.x][l];                              \
                }                                                               \
                atomicAdd(sums + c0 + l, a);                                    \
                atomicAdd(sums + C + c0 + l, b);                                \
            }                                                                   \
        }                                                                       \
    }

__global__ void __launch_bounds__(256) norm_stats_v4_kernel(const float* __restrict__ p, size_t ld, long long P, int C,
                                                             double* __restrict__ sums) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const bool active = c0 < C;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (active) {
        const long long stride = (long long)gridDim.y * 8;
#pragma unroll 2
        for (long long pt = (long long)blockIdx.y * 8 + threadIdx.y; pt < P; pt += stride) {
            const V4x3 v = ld43(p + (size_t)pt * 3 * ld + c0, ld);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const double n = (double)(sqrtf(dot3l(v, v, l)) + VS_EPS);
                s1[l] += n;
                s2[l] = fma(n, n, s2[l]);
            }
        }
    }
    VS_BLOCK_REDUCE2(s1, s2, sums, C, c0)
}

template <bool HAS_BN, bool HAS_D, bool FAST>
__global__ void __launch_bounds__(256) bn_leaky_fwd_v4_kernel(const float* __restrict__ p, size_t ldp, const float* __restrict__ d,
                                                               size_t ldd, float* __restrict__ out, size_t ldo, long long P, int C,
                                                               const float* __restrict__ stat, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float ns) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (c0 >= C) return;
    const ChanParams cp = load_params(HAS_BN ? stat : nullptr, gamma, beta, C, c0);
    const float k = 1.f - ns;
    const long long stride = (long long)gridDim.y * 8;
#pragma unroll 2
    for (long long pt = (long long)blockIdx.y * 8 + threadIdx.y; pt < P; pt += stride) {
        V4x3 v = ld43(p + (size_t)pt * 3 * ldp + c0, ldp);
        V4x3 dv;
        if (HAS_D) dv = ld43(d + (size_t)pt * 3 * ldd + c0, ldd);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (HAS_BN) {
                float n, nhat, nb;
                bn_apply_lane_t<FAST>(v, l, cp, n, nhat, nb);
            }
            if (HAS_D) leaky_lane_t<FAST>(v, dv, l, ns, k);
        }
        st43(out + (size_t)pt * 3 * ldo + c0, ldo, v);
    }
}

// throughput-mode (fast math) forward with a direction, packed fp32x2: out = t p - (k a) d per lane pair
template <bool HAS_BN>
__global__ void __launch_bounds__(256) bn_leaky_fwd_p2_kernel(const float* __restrict__ p, size_t ldp, const float* __restrict__ d,
                                                               size_t ldd, float* __restrict__ out, size_t ldo, long long P, int C,
                                                               const float* __restrict__ stat, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float ns) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (c0 >= C) return;
    const BNPair bn[2] = {load_bn_pair(HAS_BN ? stat : nullptr, gamma, beta, C, c0), load_bn_pair(HAS_BN ? stat : nullptr, gamma, beta, C, c0 + 2)};
    const float k = 1.f - ns;
    const long long stride = (long long)gridDim.y * 8;
#pragma unroll 2
    for (long long pt = (long long)blockIdx.y * 8 + threadIdx.y; pt < P; pt += stride) {
        float4 p4[3], d4[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            p4[v] = __ldg(reinterpret_cast<const float4*>(p + ((size_t)pt * 3 + v) * ldp + c0));
            d4[v] = __ldg(reinterpret_cast<const float4*>(d + ((size_t)pt * 3 + v) * ldd + c0));
        }
        f2 a[3] = {mk2(p4[0].x, p4[0].y), mk2(p4[1].x, p4[1].y), mk2(p4[2].x, p4[2].y)};
        f2 b[3] = {mk2(p4[0].z, p4[0].w), mk2(p4[1].z, p4[1].w), mk2(p4[2].z, p4[2].w)};
        const f2 da[3] = {mk2(d4[0].x, d4[0].y), mk2(d4[1].x, d4[1].y), mk2(d4[2].x, d4[2].y)};
        const f2 db[3] = {mk2(d4[0].z, d4[0].w), mk2(d4[1].z, d4[1].w), mk2(d4[2].z, d4[2].w)};
        leaky_bn_pair_fwd<HAS_BN>(a, da, bn[0], k);
        leaky_bn_pair_fwd<HAS_BN>(b, db, bn[1], k);
#pragma unroll
        for (int v = 0; v < 3; ++v)
            *reinterpret_cast<float4*>(out + ((size_t)pt * 3 + v) * ldo + c0) = make_float4(a[v].v.x, a[v].v.y, b[v].v.x, b[v].v.y);
    }
}

// TAIL: the consumer of this layer is VNLinear(C,1) (+ residual): its gradient is rank one, g[r,c] = gy[r]*w2[c], so it is
// formed on the fly (g = gy, ldg unused) and the weight gradient gw2[c] = sum_r gy[r]*out[r,c] is accumulated from the
// recomputed layer output -- the [R,C] activation and its gradient never exist in HBM.
template <bool HAS_BN, bool HAS_D, bool TAIL>
__global__ void __launch_bounds__(256, 2) bn_leaky_bwd1_v4_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ p,
                                                                   size_t ldp, const float* __restrict__ d, size_t ldd,
                                                                   float* __restrict__ gp, size_t ldgp, float* __restrict__ gd,
                                                                   size_t ldgd, long long P, int C, const float* __restrict__ stat,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float ns, double* __restrict__ sums,
                                                                   const float* __restrict__ w2, double* __restrict__ gw2) {
    // register diet (two CTAs per SM): the post-BN vector is kept as pr * sc, and gp / gd overwrite g / d in place
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const bool active = c0 < C;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float s3[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        const ChanParams cp = load_params(HAS_BN ? stat : nullptr, gamma, beta, C, c0);
        const float k = 1.f - ns;
        float w2l[4] = {0.f, 0.f, 0.f, 0.f};
        if (TAIL) {
#pragma unroll
            for (int l = 0; l < 4; ++l) w2l[l] = __ldg(w2 + c0 + l);
        }
        const long long stride = (long long)gridDim.y * 8;
        for (long long pt = (long long)blockIdx.y * 8 + threadIdx.y; pt < P; pt += stride) {
            const V4x3 pr = ld43(p + (size_t)pt * 3 * ldp + c0, ldp);
            V4x3 gv;
            float gy3[3] = {0.f, 0.f, 0.f};
            if (TAIL) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    gy3[c] = __ldg(g + (size_t)pt * 3 + c);
#pragma unroll
                    for (int l = 0; l < 4; ++l) gv.v[c][l] = gy3[c] * w2l[l];
                }
            } else {
                gv = ld43(g + (size_t)pt * 3 * ldg + c0, ldg);
            }
            V4x3 dv;
            if (HAS_D) dv = ld43(d + (size_t)pt * 3 * ldd + c0, ldd);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                float n = 1.f, nhat = 0.f;
                float pb0 = pr.v[0][l], pb1 = pr.v[1][l], pb2 = pr.v[2][l];      // BN(p), same op order as the forward
                if (HAS_BN) {
                    n = fsqrt_fast(dot3l(pr, pr, l)) + VS_EPS;
                    nhat = (n - cp.mean[l]) * cp.invstd[l];
                    const float t = (nhat * cp.gamma[l] + cp.beta[l]) * frcp(n);
                    pb0 *= t;
                    pb1 *= t;
                    pb2 *= t;
                }
                if (HAS_D) {
                    const float s = __fadd_rn(__fadd_rn(__fmul_rn(pb0, dv.v[0][l]), __fmul_rn(pb1, dv.v[1][l])), __fmul_rn(pb2, dv.v[2][l]));
                    if (TAIL) {
                        // recomputed layer output (same expression as the forward) for the tail weight gradient
                        const float af = (s < 0.f) ? s / __fadd_rn(dot3l(dv, dv, l), VS_EPS) : 0.f;
                        const float o0 = __fadd_rn(__fmul_rn(ns, pb0), __fmul_rn(k, __fsub_rn(pb0, __fmul_rn(af, dv.v[0][l]))));
                        const float o1 = __fadd_rn(__fmul_rn(ns, pb1), __fmul_rn(k, __fsub_rn(pb1, __fmul_rn(af, dv.v[1][l]))));
                        const float o2 = __fadd_rn(__fmul_rn(ns, pb2), __fmul_rn(k, __fsub_rn(pb2, __fmul_rn(af, dv.v[2][l]))));
                        s3[l] += gy3[0] * o0 + gy3[1] * o1 + gy3[2] * o2;
                    }
                    if (s < 0.f) {
                        const float rq = frcp(dot3l(dv, dv, l) + VS_EPS);
                        const float a = s * rq;
                        const float gdq = dot3l(gv, dv, l) * rq;
                        const float pbv[3] = {pb0, pb1, pb2};
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float gval = gv.v[c][l], dval = dv.v[c][l];
                            gv.v[c][l] = gval - k * gdq * dval;                                        // dL/d BN(p)
                            dv.v[c][l] = -k * (a * gval + gdq * pbv[c] - 2.f * a * gdq * dval);        // dL/d d
                        }
                    } else {
                        dv.v[0][l] = dv.v[1][l] = dv.v[2][l] = 0.f;
                    }
                }
                if (HAS_BN) {
                    const float dnb = dot3l(gv, pr, l) * frcp(n);
                    s1[l] += (double)dnb;
                    s2[l] = fma((double)dnb, (double)nhat, s2[l]);
                }
            }
            st43(gp + (size_t)pt * 3 * ldgp + c0, ldgp, gv);
            if (HAS_D) st43(gd + (size_t)pt * 3 * ldgd + c0, ldgd, dv);
        }
    }
    if (HAS_BN) VS_BLOCK_REDUCE2(s1, s2, sums, C, c0)
    if (TAIL) {
        __shared__ float sh3[8][32][4];
#pragma unroll
        for (int l = 0; l < 4; ++l) sh3[threadIdx.y][threadIdx.x][l] = s3[l];
        __syncthreads();
        if (threadIdx.y == 0 && active) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double a = 0.0;
                for (int i = 0; i < 8; ++i) a += (double)sh3[i][threadIdx.x][l];
                atomicAdd(gw2 + c0 + l, a);
            }
        }
    }
}

// Packed (fp32x2) version of the pass above for the case with a direction (HAS_D): the four channel lanes are two pairs
// and the backward is expressed through the dot products p.p, p.d, d.d, g.p, g.d (see vn_fused.cu for the formulas), so
// the `s < 0` branch only selects two scalar coefficients and every vector update is a packed FMA.
// STORE = false: only the per-channel reductions (BatchNorm backward sums, tail weight gradient) are produced -- the sums pre-pass of
// the fused tail backward (gemm_tcgen05.cu, tail_dgrad_tf32_kernel), which forms gp / gd itself, final, inside the dgrad GEMM.
template <bool HAS_BN, bool TAIL, bool STORE = true>
__global__ void __launch_bounds__(256, 2) bn_leaky_bwd1_p2_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ p,
                                                                   size_t ldp, const float* __restrict__ d, size_t ldd,
                                                                   float* __restrict__ gp, size_t ldgp, float* __restrict__ gd,
                                                                   size_t ldgd, long long P, int C, const float* __restrict__ stat,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float ns, double* __restrict__ sums,
                                                                   const float* __restrict__ w2, double* __restrict__ gw2) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const bool active = c0 < C;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float s3[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        f2 mean[2], invstd[2], ga[2], be[2], w2p[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + 2 * h;
            mean[h] = HAS_BN ? mk2(__ldg(stat + c), __ldg(stat + c + 1)) : bc2(0.f);
            invstd[h] = HAS_BN ? mk2(__ldg(stat + C + c), __ldg(stat + C + c + 1)) : bc2(0.f);
            ga[h] = HAS_BN ? mk2(__ldg(gamma + c), __ldg(gamma + c + 1)) : bc2(0.f);
            be[h] = HAS_BN ? mk2(__ldg(beta + c), __ldg(beta + c + 1)) : bc2(0.f);
            w2p[h] = TAIL ? mk2(__ldg(w2 + c), __ldg(w2 + c + 1)) : bc2(0.f);
        }
        const float k = 1.f - ns;
        float f1[4] = {0, 0, 0, 0}, f2s[4] = {0, 0, 0, 0};
        int since_flush = 0;
        const long long stride = (long long)gridDim.y * 8;
        for (long long pt = (long long)blockIdx.y * 8 + threadIdx.y; pt < P; pt += stride) {
            float4 p4[3], d4[3], g4[3];
            float gy3[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                p4[v] = __ldg(reinterpret_cast<const float4*>(p + ((size_t)pt * 3 + v) * ldp + c0));
                d4[v] = __ldg(reinterpret_cast<const float4*>(d + ((size_t)pt * 3 + v) * ldd + c0));
                if (TAIL) gy3[v] = __ldg(g + (size_t)pt * 3 + v);
                else g4[v] = __ldg(reinterpret_cast<const float4*>(g + ((size_t)pt * 3 + v) * ldg + c0));
            }
            float4 ogp[3], ogd[3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                f2 pp_[3], dd_[3], gg[3];
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    pp_[v] = h == 0 ? mk2(p4[v].x, p4[v].y) : mk2(p4[v].z, p4[v].w);
                    dd_[v] = h == 0 ? mk2(d4[v].x, d4[v].y) : mk2(d4[v].z, d4[v].w);
                    if (TAIL) gg[v] = bc2(gy3[v]) * w2p[h];
                    else gg[v] = h == 0 ? mk2(g4[v].x, g4[v].y) : mk2(g4[v].z, g4[v].w);
                }
                const f2 pd = dot3p(pp_, dd_), dq = dot3p(dd_, dd_), gpd = dot3p(gg, pp_), gdd = dot3p(gg, dd_);
                f2 t = bc2(1.f), rn = bc2(0.f), nhat = bc2(0.f);
                if (HAS_BN) {
                    const f2 pp = dot3p(pp_, pp_);
                    f2 rs = rsqrt2(pp);
                    rs = mk2(pp.v.x > 0.f ? rs.v.x : 0.f, pp.v.y > 0.f ? rs.v.y : 0.f);
                    const f2 n = fma2p(pp, rs, bc2(VS_EPS));
                    rn = rcp2(n);
                    nhat = (n - mean[h]) * invstd[h];
                    t = fma2p(nhat, ga[h], be[h]) * rn;
                }
                const f2 sdot = t * pd;
                const f2 rq = rcp2(dq + bc2(VS_EPS));
                const bool mx = sdot.v.x < 0.f, my = sdot.v.y < 0.f;
                const f2 a = sel0(mx, my, sdot * rq);
                const f2 c1 = sel0(mx, my, bc2(k) * (gdd * rq));
                const f2 ca = neg2(bc2(k) * a), cb = neg2(c1 * t), cc = bc2(2.f) * (a * c1);
                f2 ogpv[3], ogdv[3];
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    ogpv[v] = gg[v] - c1 * dd_[v];                                          // dL/d BN(p)
                    ogdv[v] = fma2p(cc, dd_[v], fma2p(cb, pp_[v], ca * gg[v]));             // dL/d d
                    if (h == 0) {
                        ogp[v].x = ogpv[v].v.x; ogp[v].y = ogpv[v].v.y;
                        ogd[v].x = ogdv[v].v.x; ogd[v].y = ogdv[v].v.y;
                    } else {
                        ogp[v].z = ogpv[v].v.x; ogp[v].w = ogpv[v].v.y;
                        ogd[v].z = ogdv[v].v.x; ogd[v].w = ogdv[v].v.y;
                    }
                }
                if (HAS_BN) {
                    const f2 dnb = (gpd - c1 * pd) * rn;
                    const f2 dn2 = dnb * nhat;
                    f1[2 * h] += dnb.v.x;
                    f1[2 * h + 1] += dnb.v.y;
                    f2s[2 * h] += dn2.v.x;
                    f2s[2 * h + 1] += dn2.v.y;
                }
                if (TAIL) {
                    // sum_v gy[v] out[v] with out = t p - k a d (the recomputed layer output)
                    const f2 gyb[3] = {bc2(gy3[0]), bc2(gy3[1]), bc2(gy3[2])};
                    const f2 q = t * dot3p(gyb, pp_) - (bc2(k) * a) * dot3p(gyb, dd_);
                    s3[2 * h] += q.v.x;
                    s3[2 * h + 1] += q.v.y;
                }
            }
            if (STORE) {
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    *reinterpret_cast<float4*>(gp + ((size_t)pt * 3 + v) * ldgp + c0) = ogp[v];
                    *reinterpret_cast<float4*>(gd + ((size_t)pt * 3 + v) * ldgd + c0) = ogd[v];
                }
            }
            if (HAS_BN && ++since_flush == 32) {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    s1[l] += (double)f1[l];
                    s2[l] += (double)f2s[l];
                    f1[l] = f2s[l] = 0.f;
                }
                since_flush = 0;
            }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            s1[l] += (double)f1[l];
            s2[l] += (double)f2s[l];
        }
    }
    if (HAS_BN) VS_BLOCK_REDUCE2(s1, s2, sums, C, c0)
    if (TAIL) {
        __shared__ float sh3[8][32][4];
#pragma unroll
        for (int l = 0; l < 4; ++l) sh3[threadIdx.y][threadIdx.x][l] = s3[l];
        __syncthreads();
        if (threadIdx.y == 0 && active) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double a = 0.0;
                for (int i = 0; i < 8; ++i) a += (double)sh3[i][threadIdx.x][l];
                atomicAdd(gw2 + c0 + l, a);
            }
        }
    }
}

// per-sample column sums for the per-sample-bias gradient (the broadcast half of cat([global.expand(N), local]), models/pcn.py:172):
// a thread that walks a CONTIGUOUS range of points keeps the running sums of its channels for the current sample in registers and
// flushes them with fp32 atomics when the sample changes / at the end (1-2 flushes per thread).
template <int NCH, bool ON = true>
struct SampleAcc {
    float a[3][NCH];
    long long cur;
    __device__ __forceinline__ SampleAcc() : cur(-1) {
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < NCH; ++l) a[v][l] = 0.f;
    }
    __device__ __forceinline__ void flush(float* __restrict__ gb, size_t ldgb, int c0) {
        if (cur < 0) return;
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < NCH; ++l) {
                atomicAdd(gb + (size_t)(cur * 3 + v) * ldgb + c0 + l, a[v][l]);
                a[v][l] = 0.f;
            }
    }
    __device__ __forceinline__ void at(long long sample, float* __restrict__ gb, size_t ldgb, int c0) {
        if (sample != cur) {
            flush(gb, ldgb, c0);
            cur = sample;
        }
    }
};

template <int NCH>
struct SampleAcc<NCH, false> {      // disabled: no state, no code
    float a[3][NCH];                // (never read; lets the call sites compile unchanged)
    __device__ __forceinline__ void flush(float*, size_t, int) {}
    __device__ __forceinline__ void at(long long, float*, size_t, int) {}
};

// fused forward tail: y[r] = sum_c leaky(BN(p), d)[r, c] * w2[c] (+ res[r]); block (C/4, 256/(C/4)): one point per block row
template <bool HAS_BN, bool FAST>
__global__ void __launch_bounds__(256) bn_leaky_dot_fwd_v4_kernel(const float* __restrict__ p, size_t ldp, const float* __restrict__ d,
                                                                   size_t ldd, long long P, int C, const float* __restrict__ stat,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float ns, const float* __restrict__ w2,
                                                                   const float* __restrict__ res, float* __restrict__ y) {
    __shared__ float part[8][8][3];     // [block row][warp within row][component]
    const int c0 = threadIdx.x * 4;
    const ChanParams cp = load_params(HAS_BN ? stat : nullptr, gamma, beta, C, c0);
    const BNPair bnp[2] = {load_bn_pair(HAS_BN ? stat : nullptr, gamma, beta, C, c0), load_bn_pair(HAS_BN ? stat : nullptr, gamma, beta, C, c0 + 2)};
    float w2l[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) w2l[l] = __ldg(w2 + c0 + l);
    const float k = 1.f - ns;
    const int wx = threadIdx.x >> 5, nwx = blockDim.x >> 5, lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.y;
    const long long iters = (P + stride - 1) / stride;
    for (long long it = 0; it < iters; ++it) {
        const long long pt = it * stride + (long long)blockIdx.x * blockDim.y + threadIdx.y;
        float acc[3] = {0.f, 0.f, 0.f};
        if (pt < P) {
            V4x3 v = ld43(p + (size_t)pt * 3 * ldp + c0, ldp);
            const V4x3 dv = ld43(d + (size_t)pt * 3 * ldd + c0, ldd);
            if (FAST) {
                f2 a[3] = {mk2(v.v[0][0], v.v[0][1]), mk2(v.v[1][0], v.v[1][1]), mk2(v.v[2][0], v.v[2][1])};
                f2 b[3] = {mk2(v.v[0][2], v.v[0][3]), mk2(v.v[1][2], v.v[1][3]), mk2(v.v[2][2], v.v[2][3])};
                const f2 da[3] = {mk2(dv.v[0][0], dv.v[0][1]), mk2(dv.v[1][0], dv.v[1][1]), mk2(dv.v[2][0], dv.v[2][1])};
                const f2 db[3] = {mk2(dv.v[0][2], dv.v[0][3]), mk2(dv.v[1][2], dv.v[1][3]), mk2(dv.v[2][2], dv.v[2][3])};
                leaky_bn_pair_fwd<HAS_BN>(a, da, bnp[0], k);
                leaky_bn_pair_fwd<HAS_BN>(b, db, bnp[1], k);
                const f2 wa = mk2(w2l[0], w2l[1]), wb = mk2(w2l[2], w2l[3]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const f2 q = fma2p(b[c], wb, a[c] * wa);
                    acc[c] = q.v.x + q.v.y;
                }
            } else {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (HAS_BN) {
                        float n, nhat, nb;
                        bn_apply_lane_t<FAST>(v, l, cp, n, nhat, nb);
                    }
                    leaky_lane_t<FAST>(v, dv, l, ns, k);
                    acc[0] = fmaf(v.v[0][l], w2l[l], acc[0]);
                    acc[1] = fmaf(v.v[1][l], w2l[l], acc[1]);
                    acc[2] = fmaf(v.v[2][l], w2l[l], acc[2]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0) {
            part[threadIdx.y][wx][0] = acc[0];
            part[threadIdx.y][wx][1] = acc[1];
            part[threadIdx.y][wx][2] = acc[2];
        }
        __syncthreads();
        if (threadIdx.x < 3 && pt < P) {
            float t = 0.f;
            for (int w = 0; w < nwx; ++w) t += part[threadIdx.y][w][threadIdx.x];
            const size_t r = (size_t)pt * 3 + threadIdx.x;
            y[r] = res ? t + __ldg(res + r) : t;
        }
        __syncthreads();
    }
}

// SBIAS: each block row walks a contiguous range of points and the thread also accumulates the per-sample column sums of the final gp
// (its own write) and of the matching channels of gd (final since bwd1; one extra read) = the gradient of the per-sample bias rows
// [B*3, 2C] = (p half | d half).  Replaces the rows_sample_sum pass that re-read the whole stacked gradient.
template <bool SBIAS>
__global__ void __launch_bounds__(256, SBIAS ? 2 : 0) bn_bwd2_v4_kernel(float* __restrict__ gp, size_t ldgp, const float* __restrict__ p, size_t ldp,
                                                          long long P, int C, const float* __restrict__ stat,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const double* __restrict__ sums, double count, int training,
                                                          const float* __restrict__ gd, size_t ldgd, float* __restrict__ gbias,
                                                          size_t ldgb, long long pts_per_sample) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (c0 >= C) return;
    const ChanParams cp = load_params(stat, gamma, beta, C, c0);
    float m1[4], m2[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        m1[l] = training ? (float)(sums[c0 + l] / count) * cp.gamma[l] : 0.f;
        m2[l] = training ? (float)(sums[C + c0 + l] / count) * cp.gamma[l] : 0.f;
    }
    SampleAcc<4, SBIAS> acc, accd;
    const long long per = SBIAS ? (P + gridDim.y - 1) / gridDim.y : 0;
    const long long pt_begin = SBIAS ? (long long)blockIdx.y * per + threadIdx.y : (long long)blockIdx.y * 8 + threadIdx.y;
    const long long pt_end = SBIAS ? (((long long)blockIdx.y + 1) * per < P ? ((long long)blockIdx.y + 1) * per : P) : P;
    const long long stride = SBIAS ? 8 : (long long)gridDim.y * 8;
#pragma unroll 2
    for (long long pt = pt_begin; pt < pt_end; pt += stride) {
        const V4x3 pr = ld43(p + (size_t)pt * 3 * ldp + c0, ldp);
        float* gptr = gp + (size_t)pt * 3 * ldgp + c0;
        V4x3 gv = ld43_rw(gptr, ldgp);
        if (SBIAS) {
            const long long smp = pt / pts_per_sample;
            acc.at(smp, gbias, ldgb, c0);
            accd.at(smp, gbias, ldgb, C + c0);
            const V4x3 dv = ld43(gd + (size_t)pt * 3 * ldgd + c0, ldgd);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int l = 0; l < 4; ++l) accd.a[c][l] += dv.v[c][l];
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const float r = fsqrt_fast(dot3l(pr, pr, l));
            const float n = r + VS_EPS;
            const float rn = frcp(n);
            const float nhat = (n - cp.mean[l]) * cp.invstd[l];
            const float nb = nhat * cp.gamma[l] + cp.beta[l];
            const float gx = dot3l(gv, pr, l);
            const float dnb = gx * rn;
            float dn = cp.gamma[l] * dnb;
            if (training) dn = dn - m1[l] - nhat * m2[l];
            dn = dn * cp.invstd[l] - gx * nb * rn * rn;
            const float sc = nb * rn;
            const float ur = r > 0.f ? dn * frcp(r) : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gv.v[c][l] = gv.v[c][l] * sc + ur * pr.v[c][l];
                if (SBIAS) acc.a[c][l] += gv.v[c][l];
            }
        }
        st43(gptr, ldgp, gv);
    }
    if (SBIAS) {
        acc.flush(gbias, ldgb, c0);
        accd.flush(gbias, ldgb, C + c0);
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool ok4(const void* p, long long ld) { return p == nullptr || (al16(p) && (ld & 3) == 0); }

// grid-stride over points: ONE wave of sm_count() * resident CTAs (resident = what fits of this kernel on an SM), so that every
// CTA streams an equal share and no partial last wave runs at reduced bandwidth
static dim3 stream_grid(long long P, int C, int resident) {
    const int gx = (C / 4 + 31) / 32;
    long long gy = (P + 7) / 8;
    const long long slots = (long long)sm_count() * (tuning(TUNE_GRID_LEGACY) ? 8 : resident);
    long long cap = tuning(TUNE_GRID_LEGACY) ? (slots + gx - 1) / gx : slots / gx;
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}
#define STREAM_GRID(KERNEL) stream_grid(P, C, resident_ctas(KERNEL, 256))

bool try_norm_stats_v4(const float* p, long long ldp, long long P, int C, double* sums, cudaStream_t st) {
    if ((C & 3) || !ok4(p, ldp)) return false;
    count_launch(), norm_stats_v4_kernel<<<STREAM_GRID(norm_stats_v4_kernel), dim3(32, 8), 0, st>>>(p, (size_t)ldp, P, C, sums);
    return true;
}

bool try_bn_leaky_fwd_v4(const float* p, long long ldp, const float* d, long long ldd, float* out, long long ldo, long long P, int C,
                         const float* stat, const float* gamma, const float* beta, float ns, cudaStream_t st) {
    if ((C & 3) || !ok4(p, ldp) || !ok4(d, ldd) || !ok4(out, ldo)) return false;
    const dim3 block(32, 8);
    const bool fast = fast_math_enabled();
#define VS_FWD(BN_, D_)                                                                                                                   \
    {                                                                                                                                     \
        if (fast)                                                                                                                         \
            count_launch(), bn_leaky_fwd_v4_kernel<BN_, D_, true><<<STREAM_GRID((bn_leaky_fwd_v4_kernel<BN_, D_, true>)), block, 0, st>>>(p, (size_t)ldp, d, (size_t)ldd, out, (size_t)ldo, P, \
                                                                                           C, stat, gamma, beta, ns);                        \
        else                                                                                                                              \
            count_launch(), bn_leaky_fwd_v4_kernel<BN_, D_, false><<<STREAM_GRID((bn_leaky_fwd_v4_kernel<BN_, D_, false>)), block, 0, st>>>(p, (size_t)ldp, d, (size_t)ldd, out, (size_t)ldo, P, \
                                                                                            C, stat, gamma, beta, ns);                       \
    }
    if (fast && d) {
        if (stat)
            count_launch(), bn_leaky_fwd_p2_kernel<true><<<STREAM_GRID(bn_leaky_fwd_p2_kernel<true>), block, 0, st>>>(p, (size_t)ldp, d, (size_t)ldd, out, (size_t)ldo, P, C, stat, gamma,
                                                                             beta, ns);
        else
            count_launch(), bn_leaky_fwd_p2_kernel<false><<<STREAM_GRID(bn_leaky_fwd_p2_kernel<false>), block, 0, st>>>(p, (size_t)ldp, d, (size_t)ldd, out, (size_t)ldo, P, C, stat, gamma,
                                                                              beta, ns);
        return true;
    }
    if (stat && d) VS_FWD(true, true)
    else if (stat) VS_FWD(true, false)
    else if (d) VS_FWD(false, true)
    else VS_FWD(false, false)
#undef VS_FWD
    return true;
}

bool try_bn_leaky_bwd1_v4(const float* g, long long ldg, const float* p, long long ldp, const float* d, long long ldd, float* gp,
                          long long ldgp, float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma,
                          const float* beta, float ns, double* sums, cudaStream_t st) {
    if ((C & 3) || !ok4(g, ldg) || !ok4(p, ldp) || !ok4(d, ldd) || !ok4(gp, ldgp) || !ok4(gd, ldgd)) return false;
    const dim3 block(32, 8);
#define VS_BWD(BN_, D_)                                                                                                            \
    count_launch(), bn_leaky_bwd1_v4_kernel<BN_, D_, false><<<STREAM_GRID((bn_leaky_bwd1_v4_kernel<BN_, D_, false>)), block, 0, st>>>(g, (size_t)ldg, p, (size_t)ldp, d, (size_t)ldd, gp,    \
                                                                                     (size_t)ldgp, gd, (size_t)ldgd, P, C, stat, gamma, beta, \
                                                                                     ns, sums, nullptr, nullptr)
    if (stat && d)
        count_launch(), bn_leaky_bwd1_p2_kernel<true, false><<<STREAM_GRID((bn_leaky_bwd1_p2_kernel<true, false>)), block, 0, st>>>(g, (size_t)ldg, p, (size_t)ldp, d, (size_t)ldd, gp, (size_t)ldgp,
                                                                                 gd, (size_t)ldgd, P, C, stat, gamma, beta, ns, sums, nullptr, nullptr);
    else if (stat) VS_BWD(true, false);
    else if (d)
        count_launch(), bn_leaky_bwd1_p2_kernel<false, false><<<STREAM_GRID((bn_leaky_bwd1_p2_kernel<false, false>)), block, 0, st>>>(g, (size_t)ldg, p, (size_t)ldp, d, (size_t)ldd, gp, (size_t)ldgp,
                                                                                  gd, (size_t)ldgd, P, C, stat, gamma, beta, ns, sums, nullptr, nullptr);
    else VS_BWD(false, false);
#undef VS_BWD
    return true;
}

bool try_bn_bwd2_v4(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat, const float* gamma,
                    const float* beta, const double* sums, double count, int training, cudaStream_t st) {
    if ((C & 3) || !ok4(gp, ldgp) || !ok4(p, ldp)) return false;
    count_launch(), bn_bwd2_v4_kernel<false><<<STREAM_GRID(bn_bwd2_v4_kernel<false>), dim3(32, 8), 0, st>>>(gp, (size_t)ldgp, p, (size_t)ldp, P, C, stat, gamma, beta,
                                                                             sums, count, training, nullptr, 0, nullptr, 0, 1);
    return true;
}

// bwd2 that also produces the per-sample bias gradient gbias rows (sample, v) x 2C = (sums of gp | sums of gd), zeroed by the caller.
// Returns false when the shape is not taken.
bool try_bn_bwd2_v4_sbias(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat, const float* gamma,
                          const float* beta, const double* sums, double count, int training, const float* gd, long long ldgd, float* gbias,
                          long long ldgb, long long pts_per_sample, cudaStream_t st) {
    if ((C & 3) || !ok4(gp, ldgp) || !ok4(p, ldp) || !gd || !ok4(gd, ldgd) || !gbias || pts_per_sample <= 0) return false;
    count_launch(), bn_bwd2_v4_kernel<true><<<STREAM_GRID(bn_bwd2_v4_kernel<true>), dim3(32, 8), 0, st>>>(gp, (size_t)ldgp, p, (size_t)ldp, P, C, stat, gamma, beta,
                                                                           sums, count, training, gd, (size_t)ldgd, gbias, (size_t)ldgb,
                                                                           pts_per_sample);
    return true;
}

bool try_bn_leaky_dot_fwd_v4(const float* p, long long ldp, const float* d, long long ldd, long long P, int C, const float* stat,
                             const float* gamma, const float* beta, float ns, const float* w2, const float* res, float* y,
                             cudaStream_t st) {
    if ((C & 127) || C > 1024 || !ok4(p, ldp) || !ok4(d, ldd) || d == nullptr) return false;
    const int bx = C / 4, by = 256 / bx;
    const bool fast = fast_math_enabled();
    auto dot_grid = [&](int resident) {
        long long g = (P + by - 1) / by;
        const long long cap = (long long)sm_count() * (tuning(TUNE_GRID_LEGACY) ? 8 : resident);
        return (unsigned)(g > cap ? cap : g);
    };
#define VS_DOT(BN_, F_)                                                                                                                       \
    count_launch(), bn_leaky_dot_fwd_v4_kernel<BN_, F_><<<dot_grid(resident_ctas(bn_leaky_dot_fwd_v4_kernel<BN_, F_>, 256)), dim3(bx, by), 0, st>>>(p, (size_t)ldp, d, (size_t)ldd, P, C, stat, gamma, \
                                                                                              beta, ns, w2, res, y)
    if (stat && fast) VS_DOT(true, true);
    else if (stat) VS_DOT(true, false);
    else if (fast) VS_DOT(false, true);
    else VS_DOT(false, false);
#undef VS_DOT
    return true;
}

bool try_bn_leaky_dot_bwd1_v4(const float* gy, const float* p, long long ldp, const float* d, long long ldd, float* gp, long long ldgp,
                              float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma, const float* beta,
                              float ns, double* sums, const float* w2, double* gw2, cudaStream_t st) {
    if ((C & 3) || !ok4(p, ldp) || !ok4(d, ldd) || !ok4(gp, ldgp) || !ok4(gd, ldgd) || d == nullptr) return false;
    const dim3 block(32, 8);
    if (stat)
        count_launch(), bn_leaky_bwd1_p2_kernel<true, true><<<STREAM_GRID((bn_leaky_bwd1_p2_kernel<true, true>)), block, 0, st>>>(gy, 0, p, (size_t)ldp, d, (size_t)ldd, gp, (size_t)ldgp, gd,
                                                                                (size_t)ldgd, P, C, stat, gamma, beta, ns, sums, w2, gw2);
    else
        count_launch(), bn_leaky_bwd1_p2_kernel<false, true><<<STREAM_GRID((bn_leaky_bwd1_p2_kernel<false, true>)), block, 0, st>>>(gy, 0, p, (size_t)ldp, d, (size_t)ldd, gp, (size_t)ldgp, gd,
                                                                                 (size_t)ldgd, P, C, stat, gamma, beta, ns, sums, w2, gw2);
    return true;
}

// sums pre-pass of the fused tail backward: sums [2C] (sum dnb | sum dnb * nhat) and gw2 [C] accumulated (zeroed by the caller)
bool try_bn_leaky_dot_sums_v4(const float* gy, const float* p, long long ldp, const float* d, long long ldd, long long P, int C,
                              const float* stat, const float* gamma, const float* beta, float ns, double* sums, const float* w2, double* gw2,
                              cudaStream_t st) {
    if ((C & 3) || !ok4(p, ldp) || !ok4(d, ldd) || d == nullptr || stat == nullptr) return false;
    count_launch(), bn_leaky_bwd1_p2_kernel<true, true, false><<<STREAM_GRID((bn_leaky_bwd1_p2_kernel<true, true, false>)), dim3(32, 8), 0, st>>>(
        gy, 0, p, (size_t)ldp, d, (size_t)ldd, nullptr, 0, nullptr, 0, P, C, stat, gamma, beta, ns, sums, w2, gw2);
    return true;
}

}  // namespace vnpcc
