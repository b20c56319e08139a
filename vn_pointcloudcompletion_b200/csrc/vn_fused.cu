// vn_fused.cu -- VNLinearLeakyReLU whose GEMM input has at most four channels plus a per-sample bias, executed WITHOUT
// materialising the linear outputs p = W_feat x + b_p and d = W_dir x + b_d.
//
// This is the decoder's first layer (models/pcn.py:336, :383-387): of its 2050 input channels 2048 are the broadcast
// global feature (folded into the per-sample bias rows b_p | b_d, see pcn.py in this package) and only {seed,
// point_feat} vary per point.  p and d are two FMAs per component, so every pass recomputes them from the [R, 2] local
// rows instead of reading 2 x 1.6 GB from HBM:
//   stats : per-channel sum ||p||, sum ||p||^2          (no large tensor touched)
//   fwd   : out = leaky(BN(p), d)                       (writes out only)
//   bwd A : per-channel S1 = sum d_nb, S2 = sum d_nb nhat   (reads g = dL/dout)
//   bwd B : reads g again, forms dL/dp, dL/dd in registers and reduces them on the fly into
//             gW[2C, K], gbias[B*3, 2C] (per-channel reductions) and gx[R, K] (per-row reduction over channels)
// Reference semantics: VNLinearLeakyReLU models/vn_layers.py:60-74, VNBatchNorm :116-127; backward SURVEY.md App. C.
//
// Thread layout: block (C/4, 256/(C/4)); threadIdx.x owns 4 consecutive channels (weights + the sample's bias rows in
// registers), each block row walks over a chunk of points of ONE sample.  grid = (chunks per sample, B), the chunk count chosen so that
// the grid is a whole number of waves of the kernel's resident CTAs (fold_geometry).  The two backward kernels default to TWO channels per
// thread (block (C/2, 256/(C/2)), template parameter NP = 1): half the accumulators and context fit 2 CTAs / SM without spills.
// Row mode (many small samples, e.g. the folding MLPs of Attention_VN_FoldingNet: 32768 tokens x 16 points): every block ROW owns whole
// samples and loops over them, so the per-block prologue, shared-memory reductions and atomics are paid once per block instead of once
// per 16 points, and the per-sample bias gradient is a plain store from registers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vn_math.cuh"
#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

template <int KS>
struct FoldCtx {
    float wf[4][KS], wd[4][KS];   // [lane][k]
    float bp[3][4], bd[3][4];     // per-sample bias rows [component][lane]
};

template <int KS>
__device__ __forceinline__ void fold_load_ctx(FoldCtx<KS>& cx, const float* __restrict__ w, size_t ldw, const float* __restrict__ bias,
                                              size_t ldb, int b, int C, int c0) {
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            cx.wf[l][k] = __ldg(w + (size_t)(c0 + l) * ldw + k);
            cx.wd[l][k] = __ldg(w + (size_t)(C + c0 + l) * ldw + k);
        }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const float4 p4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + (size_t)(b * 3 + v) * ldb + c0)) : make_float4(0, 0, 0, 0);
        const float4 d4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + (size_t)(b * 3 + v) * ldb + C + c0)) : make_float4(0, 0, 0, 0);
        cx.bp[v][0] = p4.x; cx.bp[v][1] = p4.y; cx.bp[v][2] = p4.z; cx.bp[v][3] = p4.w;
        cx.bd[v][0] = d4.x; cx.bd[v][1] = d4.y; cx.bd[v][2] = d4.z; cx.bd[v][3] = d4.w;
    }
}

// p (and d) of one point for this thread's 4 channels; xv[v][k] = x[(pt*3+v), k]
template <int KS, bool WITH_D>
__device__ __forceinline__ void fold_pd(const FoldCtx<KS>& cx, const float (&xv)[3][KS], V4x3& p, V4x3& d) {
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float a = cx.bp[v][l], e = cx.bd[v][l];
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                a = fmaf(xv[v][k], cx.wf[l][k], a);
                if (WITH_D) e = fmaf(xv[v][k], cx.wd[l][k], e);
            }
            p.v[v][l] = a;
            if (WITH_D) d.v[v][l] = e;
        }
}

template <int KS>
__device__ __forceinline__ void fold_load_x(const float* __restrict__ x, size_t ldx, size_t row, float (&xv)[3][KS]) {
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int k = 0; k < KS; ++k) xv[v][k] = __ldg(x + (row + v) * ldx + k);
}

// sample loop shared by all fold kernels.  Block mode (row_mode == 0): one sample per blockIdx.y, the rows split its chunk of points.
// Row mode: row threadIdx.y of block blockIdx.y owns samples b = blockIdx.y * blockDim.y + threadIdx.y, + gridDim.y * blockDim.y, ...
#define FOLD_SAMPLES_BEGIN                                                                                   \
    const int fs_step = row_mode ? (int)(gridDim.y * blockDim.y) : B;                                        \
    for (int b = row_mode ? (int)(blockIdx.y * blockDim.y + threadIdx.y) : (int)blockIdx.y; b < B; b += fs_step) { \
        const int n0 = row_mode ? 0 : (int)blockIdx.x * n_chunk;                                             \
        const int n1 = row_mode ? N : min(N, n0 + n_chunk);                                                  \
        const int nbeg = row_mode ? 0 : n0 + (int)threadIdx.y;                                               \
        const int nstep = row_mode ? 1 : (int)blockDim.y;
#define FOLD_SAMPLES_END }

// per-channel reduction of NRED double values per lane over the block rows, then one atomicAdd per channel
template <int NRED, int NL = 4>
__device__ __forceinline__ void fold_reduce_channels(double (&acc)[NRED][NL], double* __restrict__ out, int C, int c0, double* sh) {
    // sh: [blockDim.y][blockDim.x][NL] doubles, reused NRED times
    for (int i = 0; i < NRED; ++i) {
        __syncthreads();
#pragma unroll
        for (int l = 0; l < NL; ++l) sh[((size_t)threadIdx.y * blockDim.x + threadIdx.x) * NL + l] = acc[i][l];
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                double a = 0.0;
                for (int y = 0; y < (int)blockDim.y; ++y) a += sh[((size_t)y * blockDim.x + threadIdx.x) * NL + l];
                atomicAdd(out + (size_t)i * C + c0 + l, a);
            }
        }
    }
}

template <int KS, bool FAST = false>
__global__ void __launch_bounds__(256) fold_stats_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ w, size_t ldw,
                                                          const float* __restrict__ bias, size_t ldb, int B, int N, int C, int n_chunk,
                                                          int row_mode, double* __restrict__ sums) {
    extern __shared__ double fold_sh[];
    const int c0 = threadIdx.x * 4;
    double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    FOLD_SAMPLES_BEGIN
    FoldCtx<KS> cx;
    fold_load_ctx<KS>(cx, w, ldw, bias, ldb, b, C, c0);
    if (FAST) {
        // throughput mode: fp32 partial sums over 16 points, flushed to the fp64 totals (the totals feed a mean / variance)
        float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
        int since_flush = 0;
#pragma unroll 2
        for (int n = nbeg; n < n1; n += nstep) {
            float xv[3][KS];
            fold_load_x<KS>(x, ldx, ((size_t)b * N + n) * 3, xv);
            V4x3 p, d;
            fold_pd<KS, false>(cx, xv, p, d);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float pp = dot3l(p, p, l);
                const float nn = (pp > 0.f ? pp * mufu_rsqrt(pp) : 0.f) + VS_EPS;
                f1[l] += nn;
                f2[l] = fmaf(nn, nn, f2[l]);
            }
            if (++since_flush == 16) {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    acc[0][l] += (double)f1[l];
                    acc[1][l] += (double)f2[l];
                    f1[l] = f2[l] = 0.f;
                }
                since_flush = 0;
            }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            acc[0][l] += (double)f1[l];
            acc[1][l] += (double)f2[l];
        }
    } else {
        for (int n = nbeg; n < n1; n += nstep) {
            float xv[3][KS];
            fold_load_x<KS>(x, ldx, ((size_t)b * N + n) * 3, xv);
            V4x3 p, d;
            fold_pd<KS, false>(cx, xv, p, d);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const double nn = (double)(sqrtf(dot3l(p, p, l)) + VS_EPS);
                acc[0][l] += nn;
                acc[1][l] = fma(nn, nn, acc[1][l]);
            }
        }
    }
    FOLD_SAMPLES_END
    fold_reduce_channels<2>(acc, sums, C, c0, fold_sh);
}

template <int KS, bool FAST>
__global__ void __launch_bounds__(256) fold_fwd_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ w, size_t ldw,
                                                        const float* __restrict__ bias, size_t ldb, int B, int N, int C, int n_chunk, int row_mode,
                                                        const float* __restrict__ stat, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float ns, float* __restrict__ out, size_t ldo) {
    const int c0 = threadIdx.x * 4;
    const ChanParams cp = load_params(stat, gamma, beta, C, c0);
    const float k1 = 1.f - ns;
    FOLD_SAMPLES_BEGIN
    FoldCtx<KS> cx;
    fold_load_ctx<KS>(cx, w, ldw, bias, ldb, b, C, c0);
#pragma unroll 2
    for (int n = nbeg; n < n1; n += nstep) {
        const size_t row = ((size_t)b * N + n) * 3;
        float xv[3][KS];
        fold_load_x<KS>(x, ldx, row, xv);
        V4x3 v, dv;
        fold_pd<KS, true>(cx, xv, v, dv);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            if (stat) {
                float nn, nhat, nb;
                bn_apply_lane_t<FAST>(v, l, cp, nn, nhat, nb);
            }
            leaky_lane_t<FAST>(v, dv, l, ns, k1);
        }
        st43(out + row * ldo + c0, ldo, v);
    }
    FOLD_SAMPLES_END
}

// shared lane math of the two backward passes: given raw p (pr), d (dv) and g (gv, dL/dout) of one lane, turn gv into
// dL/dBN(p) and dv into dL/dd in place; returns n, nhat, nb of the BatchNorm-on-norm
__device__ __forceinline__ void fold_lane_bwd(const V4x3& pr, V4x3& dv, V4x3& gv, int l, const ChanParams& cp, bool has_bn, float k1,
                                              float& n, float& nhat, float& nb) {
    n = 1.f;
    nhat = 0.f;
    nb = 1.f;
    float pb0 = pr.v[0][l], pb1 = pr.v[1][l], pb2 = pr.v[2][l];
    if (has_bn) {
        n = fsqrt_fast(dot3l(pr, pr, l)) + VS_EPS;
        nhat = (n - cp.mean[l]) * cp.invstd[l];
        nb = nhat * cp.gamma[l] + cp.beta[l];
        const float t = nb * frcp(n);
        pb0 *= t;
        pb1 *= t;
        pb2 *= t;
    }
    const float s = pb0 * dv.v[0][l] + pb1 * dv.v[1][l] + pb2 * dv.v[2][l];
    if (s < 0.f) {
        const float rq = frcp(dot3l(dv, dv, l) + VS_EPS);
        const float a = s * rq;
        const float gdq = dot3l(gv, dv, l) * rq;
        const float pbv[3] = {pb0, pb1, pb2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float gval = gv.v[c][l], dval = dv.v[c][l];
            gv.v[c][l] = gval - k1 * gdq * dval;
            dv.v[c][l] = -k1 * (a * gval + gdq * pbv[c] - 2.f * a * gdq * dval);
        }
    } else {
        dv.v[0][l] = dv.v[1][l] = dv.v[2][l] = 0.f;
    }
}

// ---- packed (fp32x2) backward: the four channel lanes of a thread are two pairs; the leaky / BatchNorm backward is
// written in terms of five dot products per lane (p.p, p.d, d.d, g.p, g.d) so that no per-lane vector temporaries and
// no divergent branches remain (the `s < 0` case only selects scalar coefficients):
//     t = nb/n,  s = t (p.d);   if s < 0:  c1 = k (g.d)/q,  a = s/q  (q = d.d + eps)   else c1 = a = 0
//     dL/dBN(p) = g - c1 d                      dL/dd = -k a g - c1 t p + 2 a c1 d
//     <dL/dBN(p), p> = g.p - c1 (p.d)           d_nb = that / n
//     dL/dp = t (g - c1 d) + (dn / r) p,  dn = (gamma d_nb - m1 - nhat m2) invstd - <.,p> nb / n^2       
// NP = channel pairs per thread: 2 (four channels, float4 gradient loads) or 1 (two channels: half the accumulators and context,
// so that the backward kernels fit twice as many warps on an SM)
template <int KS, int NP = 2>
struct FoldCtx2 {
    f2 wf[NP][KS], wd[NP][KS];    // [pair][k]
    f2 bp[3][NP], bd[3][NP];      // [component][pair]
};

// NP consecutive channel pairs starting at p (8- or 16-byte aligned)
template <int NP>
__device__ __forceinline__ void ld_pairs(const float* __restrict__ p, f2 (&o)[NP]) {
    if constexpr (NP == 2) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        o[0] = mk2(t.x, t.y);
        o[1] = mk2(t.z, t.w);
    } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        o[0] = mk2(t.x, t.y);
    }
}
template <int NP>
__device__ __forceinline__ void st_pairs(float* __restrict__ p, const f2 (&o)[NP]) {
    if constexpr (NP == 2) *reinterpret_cast<float4*>(p) = make_float4(o[0].v.x, o[0].v.y, o[1].v.x, o[1].v.y);
    else *reinterpret_cast<float2*>(p) = make_float2(o[0].v.x, o[0].v.y);
}

template <int KS, int NP>
__device__ __forceinline__ void fold_load_ctx2(FoldCtx2<KS, NP>& cx, const float* __restrict__ w, size_t ldw, const float* __restrict__ bias,
                                               size_t ldb, int b, int C, int c0) {
#pragma unroll
    for (int h = 0; h < NP; ++h)
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            cx.wf[h][k] = mk2(__ldg(w + (size_t)(c0 + 2 * h) * ldw + k), __ldg(w + (size_t)(c0 + 2 * h + 1) * ldw + k));
            cx.wd[h][k] = mk2(__ldg(w + (size_t)(C + c0 + 2 * h) * ldw + k), __ldg(w + (size_t)(C + c0 + 2 * h + 1) * ldw + k));
        }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        if (bias) {
            ld_pairs<NP>(bias + (size_t)(b * 3 + v) * ldb + c0, cx.bp[v]);
            ld_pairs<NP>(bias + (size_t)(b * 3 + v) * ldb + C + c0, cx.bd[v]);
        } else {
#pragma unroll
            for (int h = 0; h < NP; ++h) cx.bp[v][h] = cx.bd[v][h] = bc2(0.f);
        }
    }
}

template <int NP = 2>
struct ChanParams2 {
    f2 mean[NP], invstd[NP], gamma[NP], beta[NP];
};
template <int NP = 2>
__device__ __forceinline__ ChanParams2<NP> load_params2(const float* stat, const float* gamma, const float* beta, int C, int c0) {
    ChanParams2<NP> p;
#pragma unroll
    for (int h = 0; h < NP; ++h) {
        const int c = c0 + 2 * h;
        p.mean[h] = stat ? mk2(__ldg(stat + c), __ldg(stat + c + 1)) : bc2(0.f);
        p.invstd[h] = stat ? mk2(__ldg(stat + C + c), __ldg(stat + C + c + 1)) : bc2(0.f);
        p.gamma[h] = stat ? mk2(__ldg(gamma + c), __ldg(gamma + c + 1)) : bc2(0.f);
        p.beta[h] = stat ? mk2(__ldg(beta + c), __ldg(beta + c + 1)) : bc2(0.f);
    }
    return p;
}

// everything the two backward passes share for one lane pair
struct PairBwd {
    f2 p[3], d[3];       // raw linear outputs
    f2 t;                // nb / n  (1 without BatchNorm)
    f2 nhat, nb, rn, rs; // BatchNorm-on-norm pieces: rn = 1/n, rs = 1/r (0 where r == 0)
    f2 c1, a;            // leaky coefficients (0 where <BN(p), d> >= 0)
    f2 gxd;              // <dL/dBN(p), p>
};

template <int KS, bool HAS_BN, int NP>
__device__ __forceinline__ void fold_pair_bwd(PairBwd& o, const FoldCtx2<KS, NP>& cx, int h, const float (&xv)[3][KS], const f2 (&g)[3],
                                              const ChanParams2<NP>& cp, float k1) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        f2 a = cx.bp[v][h], e = cx.bd[v][h];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const f2 xb = bc2(xv[v][k]);
            a = fma2p(xb, cx.wf[h][k], a);
            e = fma2p(xb, cx.wd[h][k], e);
        }
        o.p[v] = a;
        o.d[v] = e;
    }
    const f2 pd = dot3p(o.p, o.d), dd = dot3p(o.d, o.d), gp = dot3p(g, o.p), gd = dot3p(g, o.d);
    if (HAS_BN) {
        const f2 pp = dot3p(o.p, o.p);
        const f2 rs = rsqrt2(pp);
        o.rs = mk2(pp.v.x > 0.f ? rs.v.x : 0.f, pp.v.y > 0.f ? rs.v.y : 0.f);
        const f2 n = fma2p(pp, o.rs, bc2(VS_EPS));            // r + eps
        o.rn = rcp2(n);
        o.nhat = (n - cp.mean[h]) * cp.invstd[h];
        o.nb = fma2p(o.nhat, cp.gamma[h], cp.beta[h]);
        o.t = o.nb * o.rn;
    } else {
        o.t = bc2(1.f);
        o.nhat = o.nb = o.rn = o.rs = bc2(0.f);
    }
    const f2 s = o.t * pd;
    const f2 rq = rcp2(dd + bc2(VS_EPS));
    const bool mx = s.v.x < 0.f, my = s.v.y < 0.f;
    o.a = sel0(mx, my, s * rq);
    o.c1 = sel0(mx, my, bc2(k1) * (gd * rq));
    o.gxd = gp - o.c1 * pd;
}

// per-channel reduction of fp32 lane values over the block rows, then one atomicAdd per channel
template <int NL>
__device__ __forceinline__ void fold_reduce_store(const float (&a)[NL], float* red, float* dst, size_t stride_lane) {
    __syncthreads();
    float* mine = red + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * NL;
#pragma unroll
    for (int l = 0; l < NL; ++l) mine[l] = a[l];
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            float t = 0.f;
            for (int y = 0; y < (int)blockDim.y; ++y) t += red[((size_t)y * blockDim.x + threadIdx.x) * NL + l];
            atomicAdd(dst + (size_t)l * stride_lane, t);
        }
    }
}

template <int KS, int NP, int MINB>
__global__ void __launch_bounds__(256, MINB) fold_bwd_sums_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ x, size_t ldx,
                                                                   const float* __restrict__ w, size_t ldw, const float* __restrict__ bias,
                                                                   size_t ldb, int B, int N, int C, int n_chunk, int row_mode,
                                                                   const float* __restrict__ stat, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float ns, double* __restrict__ sums) {
    extern __shared__ double fold_sh[];
    constexpr int NL = 2 * NP;
    const int c0 = threadIdx.x * NL;
    const ChanParams2<NP> cp = load_params2<NP>(stat, gamma, beta, C, c0);
    const float k1 = 1.f - ns;
    double acc[2][NL];
    float f1[NL], f2s[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        acc[0][l] = acc[1][l] = 0.0;
        f1[l] = f2s[l] = 0.f;
    }
    int since_flush = 0;
    FOLD_SAMPLES_BEGIN
    FoldCtx2<KS, NP> cx;
    fold_load_ctx2<KS, NP>(cx, w, ldw, bias, ldb, b, C, c0);
    // software pipeline: the gradient rows of the next point are in flight while this one is processed
    f2 gnext[3][NP];
    float xnext[3][KS];
    {
        const int n = nbeg;
        if (n < n1) {
            const size_t row = ((size_t)b * N + n) * 3;
#pragma unroll
            for (int v = 0; v < 3; ++v) ld_pairs<NP>(g + (row + v) * ldg + c0, gnext[v]);
            fold_load_x<KS>(x, ldx, row, xnext);
        }
    }
#pragma unroll 1
    for (int n = nbeg; n < n1; n += nstep) {
        f2 g4[3][NP];
        float xv[3][KS];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
#pragma unroll
            for (int h = 0; h < NP; ++h) g4[v][h] = gnext[v][h];
#pragma unroll
            for (int k = 0; k < KS; ++k) xv[v][k] = xnext[v][k];
        }
        if (n + nstep < n1) {
            const size_t rown = ((size_t)b * N + n + nstep) * 3;
#pragma unroll
            for (int v = 0; v < 3; ++v) ld_pairs<NP>(g + (rown + v) * ldg + c0, gnext[v]);
            fold_load_x<KS>(x, ldx, rown, xnext);
        }
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            const f2 gg[3] = {g4[0][h], g4[1][h], g4[2][h]};
            PairBwd o;
            fold_pair_bwd<KS, true, NP>(o, cx, h, xv, gg, cp, k1);
            const f2 dnb = o.gxd * o.rn;
            const f2 dn2 = dnb * o.nhat;
            f1[2 * h] += dnb.v.x;
            f1[2 * h + 1] += dnb.v.y;
            f2s[2 * h] += dn2.v.x;
            f2s[2 * h + 1] += dn2.v.y;
        }
        if (++since_flush == 32) {         // short fp32 partial sums, flushed to fp64 (the totals feed a mean subtraction)
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                acc[0][l] += (double)f1[l];
                acc[1][l] += (double)f2s[l];
                f1[l] = f2s[l] = 0.f;
            }
            since_flush = 0;
        }
    }
    FOLD_SAMPLES_END
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        acc[0][l] += (double)f1[l];
        acc[1][l] += (double)f2s[l];
    }
    fold_reduce_channels<2, NL>(acc, sums, C, c0, fold_sh);
}

// pass B: gW (2C x KS, fp32 atomics), gbias ([B*3, 2C], fp32 atomics), gx ([R, KS], red.add; only columns >= gx_k0)
template <int KS, int NP, int MINB>
__global__ void __launch_bounds__(256, MINB) fold_bwd_main_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ x, size_t ldx,
                                                                   const float* __restrict__ w, size_t ldw, const float* __restrict__ bias,
                                                                   size_t ldb, int B, int N, int C, int n_chunk, int row_mode,
                                                                   const float* __restrict__ stat, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float ns, const double* __restrict__ sums,
                                                                   double count, int training,
                                                                   float* __restrict__ gx, size_t ldgx, int gx_k0, float* __restrict__ gw,
                                                                   size_t ldgw, float* __restrict__ gbias, size_t ldgb) {
    extern __shared__ double fold_sh[];
    float* shf = reinterpret_cast<float*>(fold_sh);
    constexpr int NL = 2 * NP;
    const int c0 = threadIdx.x * NL;
    const bool has_bn = stat != nullptr;
    const ChanParams2<NP> cp = load_params2<NP>(stat, gamma, beta, C, c0);
    const float k1 = 1.f - ns;
    f2 m1[NP], m2[NP];
#pragma unroll
    for (int h = 0; h < NP; ++h) {
        m1[h] = m2[h] = bc2(0.f);
        if (has_bn && training) {
            const int c = c0 + 2 * h;
            m1[h] = mk2((float)(sums[c] / count), (float)(sums[c + 1] / count)) * cp.gamma[h];
            m2[h] = mk2((float)(sums[C + c] / count), (float)(sums[C + c + 1] / count)) * cp.gamma[h];
        }
    }
    f2 awf[KS][NP], awd[KS][NP], abp[3][NP], abd[3][NP];
#pragma unroll
    for (int h = 0; h < NP; ++h) {
#pragma unroll
        for (int k = 0; k < KS; ++k) awf[k][h] = awd[k][h] = bc2(0.f);
#pragma unroll
        for (int v = 0; v < 3; ++v) abp[v][h] = abd[v][h] = bc2(0.f);
    }
    const int lane = threadIdx.x & 31;
    FOLD_SAMPLES_BEGIN
    FoldCtx2<KS, NP> cx;
    fold_load_ctx2<KS, NP>(cx, w, ldw, bias, ldb, b, C, c0);
    f2 gnext[3][NP];
    float xnext[3][KS];
    {
        const int n = nbeg;
        if (n < n1) {
            const size_t row = ((size_t)b * N + n) * 3;
#pragma unroll
            for (int v = 0; v < 3; ++v) ld_pairs<NP>(g + (row + v) * ldg + c0, gnext[v]);
            fold_load_x<KS>(x, ldx, row, xnext);
        }
    }
#pragma unroll 1
    for (int n = nbeg; n < n1; n += nstep) {
        const size_t row = ((size_t)b * N + n) * 3;
        f2 g4[3][NP];
        float xv[3][KS];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
#pragma unroll
            for (int h = 0; h < NP; ++h) g4[v][h] = gnext[v][h];
#pragma unroll
            for (int k = 0; k < KS; ++k) xv[v][k] = xnext[v][k];
        }
        if (n + nstep < n1) {
            const size_t rown = ((size_t)b * N + n + nstep) * 3;
#pragma unroll
            for (int v = 0; v < 3; ++v) ld_pairs<NP>(g + (rown + v) * ldg + c0, gnext[v]);
            fold_load_x<KS>(x, ldx, rown, xnext);
        }
        f2 gxp[3][KS];
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int k = 0; k < KS; ++k) gxp[v][k] = bc2(0.f);
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            const f2 gg[3] = {g4[0][h], g4[1][h], g4[2][h]};
            PairBwd o;
            if (has_bn) fold_pair_bwd<KS, true, NP>(o, cx, h, xv, gg, cp, k1);
            else fold_pair_bwd<KS, false, NP>(o, cx, h, xv, gg, cp, k1);
            // dL/dd = (-k a) g + (-c1 t) p + (2 a c1) d ;  dL/dp = t (g - c1 d) + ur p
            const f2 ca = neg2(bc2(k1) * o.a), cb = neg2(o.c1 * o.t), cc = bc2(2.f) * (o.a * o.c1);
            f2 ur = bc2(0.f);
            if (has_bn) {
                const f2 dnb = o.gxd * o.rn;
                f2 dn = cp.gamma[h] * dnb;
                if (training) dn = dn - m1[h] - o.nhat * m2[h];
                dn = dn * cp.invstd[h] - o.gxd * (o.nb * (o.rn * o.rn));
                ur = dn * o.rs;
            }
            const f2 tc1 = o.t * o.c1;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const f2 gdv = fma2p(cc, o.d[v], fma2p(cb, o.p[v], ca * gg[v]));
                const f2 gpv = fma2p(ur, o.p[v], o.t * gg[v] - tc1 * o.d[v]);
                abp[v][h] = abp[v][h] + gpv;
                abd[v][h] = abd[v][h] + gdv;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const f2 xb = bc2(xv[v][k]);
                    awf[k][h] = fma2p(gpv, xb, awf[k][h]);
                    awd[k][h] = fma2p(gdv, xb, awd[k][h]);
                    if (k >= gx_k0) gxp[v][k] = fma2p(gpv, cx.wf[h][k], fma2p(gdv, cx.wd[h][k], gxp[v][k]));
                }
            }
        }
        if (gx) {
            // per-row reduction over channels: pair halves, warp shuffle, then one red.add per warp (gx zeroed by the launcher)
#pragma unroll
            for (int v = 0; v < 3; ++v)
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    if (k < gx_k0) continue;
                    float t = gxp[v][k].v.x + gxp[v][k].v.y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == 0) atomicAdd(gx + (row + v) * ldgx + k, t);
                }
        }
    }
    if (row_mode && gbias) {      // this row owned the whole sample: its bias gradient is complete in registers
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            st_pairs<NP>(gbias + (size_t)(b * 3 + v) * ldgb + c0, abp[v]);
            st_pairs<NP>(gbias + (size_t)(b * 3 + v) * ldgb + C + c0, abd[v]);
#pragma unroll
            for (int h = 0; h < NP; ++h) abp[v][h] = abd[v][h] = bc2(0.f);
        }
    }
    FOLD_SAMPLES_END
    // per-channel reductions over the block rows
    __syncthreads();
    float* red = shf;   // [blockDim.y][blockDim.x][NL]
    auto lanes = [](const f2 (&a)[NP], float (&o)[NL]) {
#pragma unroll
        for (int h = 0; h < NP; ++h) {
            o[2 * h] = a[h].v.x;
            o[2 * h + 1] = a[h].v.y;
        }
    };
    float tmp[NL];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        lanes(awf[k], tmp);
        fold_reduce_store<NL>(tmp, red, gw + (size_t)c0 * ldgw + k, ldgw);
        lanes(awd[k], tmp);
        fold_reduce_store<NL>(tmp, red, gw + (size_t)(C + c0) * ldgw + k, ldgw);
    }
    if (gbias && !row_mode) {
        const int b = blockIdx.y;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            lanes(abp[v], tmp);
            fold_reduce_store<NL>(tmp, red, gbias + (size_t)(b * 3 + v) * ldgb + c0, 1);
            lanes(abd[v], tmp);
            fold_reduce_store<NL>(tmp, red, gbias + (size_t)(b * 3 + v) * ldgb + C + c0, 1);
        }
    }
}

// throughput-mode forward of the fused small-K layer, packed fp32x2
template <int KS, int MINB = 1>
__global__ void __launch_bounds__(256, MINB) fold_fwd_p2_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ w, size_t ldw,
                                                           const float* __restrict__ bias, size_t ldb, int B, int N, int C, int n_chunk,
                                                           int row_mode, const float* __restrict__ stat, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float ns, float* __restrict__ out, size_t ldo) {
    const int c0 = threadIdx.x * 4;
    const BNPair bn[2] = {load_bn_pair(stat, gamma, beta, C, c0), load_bn_pair(stat, gamma, beta, C, c0 + 2)};
    const float k1 = 1.f - ns;
    FOLD_SAMPLES_BEGIN
    FoldCtx2<KS> cx;
    fold_load_ctx2<KS>(cx, w, ldw, bias, ldb, b, C, c0);
#pragma unroll 2
    for (int n = nbeg; n < n1; n += nstep) {
        const size_t row = ((size_t)b * N + n) * 3;
        float xv[3][KS];
        fold_load_x<KS>(x, ldx, row, xv);
        f2 o[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            f2 d[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                f2 a = cx.bp[v][h], e = cx.bd[v][h];
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const f2 xb = bc2(xv[v][k]);
                    a = fma2p(xb, cx.wf[h][k], a);
                    e = fma2p(xb, cx.wd[h][k], e);
                }
                o[h][v] = a;
                d[v] = e;
            }
            if (stat) leaky_bn_pair_fwd<true>(o[h], d, bn[h], k1);
            else leaky_bn_pair_fwd<false>(o[h], d, bn[h], k1);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v)
            *reinterpret_cast<float4*>(out + (row + v) * ldo + c0) = make_float4(o[0][v].v.x, o[0][v].v.y, o[1][v].v.x, o[1][v].v.y);
    }
    FOLD_SAMPLES_END
}

static bool fold_ok(int KS, int C, const void* bias, long long ldb, const void* big, long long ldbig) {
    return KS >= 1 && KS <= 4 && (C & 127) == 0 && C <= 1024 && (bias == nullptr || ((ldb & 3) == 0 && ((uintptr_t)bias & 15) == 0)) &&
           (big == nullptr || ((ldbig & 3) == 0 && ((uintptr_t)big & 15) == 0));
}

// `resident` = CTAs of THIS kernel that fit on one SM (resident_ctas): in block mode the chunk count is chosen so that the grid is a
// whole number of waves of sm_count() * resident CTAs (a 608-CTA grid on 148 one-CTA SMs would run 5 waves for 4.1 waves of work).
static void fold_geometry(int B, int N, int C, int resident, dim3& grid, dim3& block, int& n_chunk, size_t& smem, int& row_mode, int lanes = 4) {
    const int bx = C / lanes, by = 256 / bx;
    row_mode = 0;
    block = dim3(bx, by);
    smem = sizeof(double) * 256 * 4;
    if (N <= 64 && B >= 4 * by) {      // many small samples: every block row owns whole samples and loops over them
        row_mode = 1;
        long long blocks = ((long long)B + by - 1) / by;
        const long long cap = (long long)sm_count() * (tuning(TUNE_GRID_LEGACY) ? 8 : resident);
        if (blocks > cap) blocks = cap;
        grid = dim3(1, (unsigned)blocks);
        n_chunk = N;
        return;
    }
    if (tuning(TUNE_GRID_LEGACY)) {
        int chunks = (int)(((long long)sm_count() * 4 + B - 1) / B);
        if (chunks < 1) chunks = 1;
        n_chunk = (N + chunks - 1) / chunks;
        if (n_chunk < by * 8) n_chunk = by * 8;
        chunks = (N + n_chunk - 1) / n_chunk;
        grid = dim3((unsigned)chunks, (unsigned)B);
        return;
    }
    // cost model: waves x (points per block row + a fixed per-block prologue / reduction cost of ~6 points per row)
    const long long slots = (long long)sm_count() * resident;
    const int min_chunk = by * 8;
    const int max_chunks = N / min_chunk > 1 ? N / min_chunk : 1;
    long long lo = (slots + B - 1) / B, hi = (8 * slots + B - 1) / B;
    if (lo > max_chunks) lo = max_chunks;
    if (hi > max_chunks) hi = max_chunks;
    if (lo < 1) lo = 1;
    int best_chunks = (int)lo;
    long long best_cost = -1;
    for (long long c = lo; c <= hi; ++c) {
        const int nc = (int)((N + c - 1) / c);
        const long long chunks = (N + nc - 1) / nc;
        const long long waves = (chunks * B + slots - 1) / slots;
        const long long cost = waves * ((nc + by - 1) / by + 6);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best_chunks = (int)chunks;
        }
    }
    n_chunk = (N + best_chunks - 1) / best_chunks;
    grid = dim3((unsigned)((N + n_chunk - 1) / n_chunk), (unsigned)B);
}

}  // namespace vnpcc

using namespace vnpcc;

static constexpr size_t FOLD_SMEM = sizeof(double) * 256 * 4;
static constexpr int FOLD_BWD_DEFAULT_VARIANT = 4;   // two channels per thread; sums pass at 3 CTAs / SM (80 registers), main pass at 2 (tools/fold_bench.py: backward 1.65 -> 1.47 ms)

#define FOLD_KS_DISPATCH(KS, ...)                                \
    switch (KS) {                                                \
        case 1: { constexpr int K_ = 1; __VA_ARGS__; } break;    \
        case 2: { constexpr int K_ = 2; __VA_ARGS__; } break;    \
        case 3: { constexpr int K_ = 3; __VA_ARGS__; } break;    \
        default: { constexpr int K_ = 4; __VA_ARGS__; } break;   \
    }

extern "C" {

// geometry of one launch of KERNEL (a concrete instantiation): dynamic shared memory SMEM_ bytes
#define FOLD_GEOM_L(KERNEL, SMEM_, LANES_)                                                                        \
    dim3 grid, block;                                                                                             \
    int n_chunk, row_mode;                                                                                        \
    size_t smem;                                                                                                  \
    fold_geometry(B, N, C, resident_ctas(KERNEL, 256, (SMEM_)), grid, block, n_chunk, smem, row_mode, (LANES_)); \
    if ((SMEM_) == 0) smem = 0;
#define FOLD_GEOM(KERNEL, SMEM_) FOLD_GEOM_L(KERNEL, SMEM_, 4)

// host-logic introspection (tests/test_planners_cpu.py): out = {grid.x, grid.y, block.x, block.y, points per chunk, row mode}
void vnpcc_debug_fold_geometry(int B, int N, int C, int resident, int lanes, int* out) {
    dim3 grid, block;
    int n_chunk, row_mode;
    size_t smem;
    fold_geometry(B, N, C, resident, grid, block, n_chunk, smem, row_mode, lanes);
    out[0] = (int)grid.x;
    out[1] = (int)grid.y;
    out[2] = (int)block.x;
    out[3] = (int)block.y;
    out[4] = n_chunk;
    out[5] = row_mode;
}

// x [B*N*3, K] local rows, w [2C, K] stacked (feat | dir) weights of the local channels, bias [B*3, 2C] per-sample rows
// (may be NULL).  sums: 2C doubles (zeroed here).
int vnpcc_fold_stats(const float* x, long long ldx, const float* w, long long ldw, const float* bias, long long ldb, int B, int N,
                     int K, int C, double* sums, void* stream) {
    if (!fold_ok(K, C, bias, ldb, nullptr, 0)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    if (B <= 0 || N <= 0) return last_error();
    FOLD_KS_DISPATCH(K, {
        if (fast_math_enabled() && tuning(TUNE_FOLD_FWD) != 1) {
            auto kern = fold_stats_kernel<K_, true>;
            FOLD_GEOM(kern, FOLD_SMEM);
            count_launch(), kern<<<grid, block, smem, st>>>(x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, sums);
        } else {
            auto kern = fold_stats_kernel<K_, false>;
            FOLD_GEOM(kern, FOLD_SMEM);
            count_launch(), kern<<<grid, block, smem, st>>>(x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, sums);
        }
    });
    return last_error();
}

int vnpcc_fold_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, long long ldb, int B, int N,
                   int K, int C, const float* stat, const float* gamma, const float* beta, float ns, float* out, long long ldo,
                   void* stream) {
    if (!fold_ok(K, C, bias, ldb, out, ldo)) return VNPCC_ERR_UNSUPPORTED;
    if (B <= 0 || N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast_math_enabled()) {
        const int fv = tuning(TUNE_FOLD_FWD);
#define FOLD_FWD_LAUNCH(MINB_)                                                                                                                  \
    {                                                                                                                                           \
        auto kern = fold_fwd_p2_kernel<K_, MINB_>;                                                                                              \
        FOLD_GEOM(kern, 0);                                                                                                                     \
        count_launch(), kern<<<grid, block, smem, st>>>(x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, stat,  \
                                                       gamma, beta, ns, out, (size_t)ldo);                                                     \
    }
        FOLD_KS_DISPATCH(K, {
            if (fv == 3) FOLD_FWD_LAUNCH(3)
            else if (fv == 4) FOLD_FWD_LAUNCH(4)
            else FOLD_FWD_LAUNCH(1)
        });
#undef FOLD_FWD_LAUNCH
    } else {
        FOLD_KS_DISPATCH(K, {
            auto kern = fold_fwd_kernel<K_, false>;
            FOLD_GEOM(kern, 0);
            count_launch(), kern<<<grid, block, smem, st>>>(x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, stat, gamma, beta,
                                                           ns, out, (size_t)ldo);
        });
    }
    return last_error();
}

// g [R, C] = dL/dout.  Outputs: gx [R, K] (may be NULL; columns < gx_first_col are left zero: inputs that need no
// gradient, e.g. the constant folding seed), gw [2C, K] and gbias [B*3, 2C] (zeroed here, gbias may be NULL),
// ggamma / gbeta [C] (written when stat != NULL).  sums: workspace of 2C doubles.
int vnpcc_fold_bwd(const float* g, long long ldg, const float* x, long long ldx, const float* w, long long ldw, const float* bias,
                   long long ldb, int B, int N, int K, int C, const float* stat, const float* gamma, const float* beta, float ns,
                   int training, double* sums, float* gx, long long ldgx, int gx_first_col, float* gw, long long ldgw, float* gbias,
                   long long ldgb, float* ggamma, float* gbeta, void* stream) {
    if (!fold_ok(K, C, bias, ldb, g, ldg) || (gbias && ((ldgb & 3) || ((uintptr_t)gbias & 15)))) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemset2DAsync(gw, (size_t)ldgw * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)2 * C, st);
    if (gbias) cudaMemset2DAsync(gbias, (size_t)ldgb * sizeof(float), 0, (size_t)2 * C * sizeof(float), (size_t)B * 3, st);
    if (stat) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    if (gx) cudaMemset2DAsync(gx, (size_t)ldgx * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)B * N * 3, st);
    if (B <= 0 || N <= 0) return last_error();
    const double count = (double)B * (double)N;
    // register budget / channels per thread of the two backward kernels (vnpcc_set_tuning knob 1):
    //   1 = four channels per thread, 1 CTA / SM (no spills)      2 = four channels, 2 CTAs / SM (<= 128 registers)
    //   3 = two channels per thread, 2 CTAs / SM                   0 = default
    int variant = tuning(TUNE_FOLD_MINB);
    if (variant == 0) variant = FOLD_BWD_DEFAULT_VARIANT;
    if (variant >= 3 && C > 512) variant = 1;
#define FOLD_SUMS_LAUNCH(NP_, MINB_)                                                                                                             \
    {                                                                                                                                            \
        auto kern = fold_bwd_sums_kernel<K_, NP_, MINB_>;                                                                                       \
        FOLD_GEOM_L(kern, FOLD_SMEM, 2 * NP_);                                                                                                   \
        count_launch(), kern<<<grid, block, smem, st>>>(g, (size_t)ldg, x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, \
                                                       stat, gamma, beta, ns, sums);                                                            \
    }
#define FOLD_MAIN_LAUNCH(NP_, MINB_)                                                                                                             \
    {                                                                                                                                            \
        auto kern = fold_bwd_main_kernel<K_, NP_, MINB_>;                                                                                       \
        FOLD_GEOM_L(kern, FOLD_SMEM, 2 * NP_);                                                                                                   \
        count_launch(), kern<<<grid, block, smem, st>>>(g, (size_t)ldg, x, (size_t)ldx, w, (size_t)ldw, bias, (size_t)ldb, B, N, C, n_chunk, row_mode, \
                                                       stat, gamma, beta, ns, sums, count, training, gx, (size_t)ldgx, gx_first_col, gw,        \
                                                       (size_t)ldgw, gbias, (size_t)ldgb);                                                      \
    }
    if (stat) {
        FOLD_KS_DISPATCH(K, {
            if (variant == 3) FOLD_SUMS_LAUNCH(1, 2)
            else if (variant == 4 || variant == 5) FOLD_SUMS_LAUNCH(1, 3)
            else if (variant == 6 || variant == 7) FOLD_SUMS_LAUNCH(1, 4)
            else if (variant == 2) FOLD_SUMS_LAUNCH(2, 2)
            else FOLD_SUMS_LAUNCH(2, 1)
        });
    }
    FOLD_KS_DISPATCH(K, {
        if (variant == 3 || variant == 4 || variant == 6) FOLD_MAIN_LAUNCH(1, 2)
        else if (variant == 5 || variant == 7) FOLD_MAIN_LAUNCH(1, 3)
        else if (variant == 2) FOLD_MAIN_LAUNCH(2, 2)
        else FOLD_MAIN_LAUNCH(2, 1)
    });
#undef FOLD_SUMS_LAUNCH
#undef FOLD_MAIN_LAUNCH
    if (stat && gbeta) vnpcc_double_to_float(sums, gbeta, C, stream);
    if (stat && ggamma) vnpcc_double_to_float(sums + C, ggamma, C, stream);
    return last_error();
}

}  // extern "C"
