// edge_conv.cu -- the edge convolution of VN_DGCNN_fps (models/dgcnn.py:251-278 vn_get_graph_feature -> VNLinearLeakyReLU(dim=5)
// -> mean_pool over the k neighbours, :282-311) WITHOUT the edge tensor.
//
// The reference builds e = cat(x_j - x_i, x_i) ([B, 2C_in, 3, N, k]) and runs the two linear maps of VNLinearLeakyReLU over
// all N*k edges.  Because the maps are linear,
//        W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i                      (W = [W1 | W2], column blocks of width C_in)
// so ONE GEMM over the N points produces  UW[point] = (U_p | U_d | W_p | W_d)  with  U = W1 x,  W = (W2 - W1) x  (feat and dir
// stacked), and an edge's pre-activations are a gather-add:  p = U_p[j] + W_p[i],  d = U_d[j] + W_d[i].  The GEMM shrinks k-fold
// (k = 16) and neither e nor (p, d) ever reaches HBM; UW (<= 50 MB per layer) is L2-resident for the gathers.  Passes:
//   stats : per-channel sum ||p||, sum ||p||^2 over all edges  (BatchNorm2d statistics over B*N*k, models/vn_layers.py:116-127)
//   fwd   : out[i] = mean_j leaky(BN(p_ij), d_ij)
//   bwd A : per-channel S1 = sum d_nb, S2 = sum d_nb nhat
//   bwd B : dL/dp, dL/dd per edge in registers -> gUW: the (W_p | W_d) half of point i is a plain store of the sum over its k
//           edges, the (U_p | U_d) half of neighbour j is accumulated with vector red.add (fp32 atomics: summation order is not
//           deterministic, like the reference's index_put_(accumulate=True) backward of x[idx, :]).
// Thread layout: block (C/4, 256/(C/4)); threadIdx.x owns 4 consecutive output channels, each block row walks over points.
#include <cuda_runtime.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"
#include "vn_math.cuh"

namespace vnpcc {

struct EdgeGeo {
    int c0;            // first of this thread's 4 channels
    long long p0;      // first point of this block row
    long long pstride;
};

__device__ __forceinline__ EdgeGeo edge_geo() {
    EdgeGeo g;
    g.c0 = threadIdx.x * 4;
    g.p0 = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    g.pstride = (long long)gridDim.x * blockDim.y;
    return g;
}

__device__ __forceinline__ void add43(V4x3& a, const V4x3& b) {
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int l = 0; l < 4; ++l) a.v[v][l] += b.v[v][l];
}

// per-channel reduction of NRED doubles per lane over the block rows, then one atomicAdd per channel and block
template <int NRED>
__device__ __forceinline__ void edge_reduce_channels(double (&acc)[NRED][4], double* __restrict__ out, int C, int c0, double* sh) {
    for (int i = 0; i < NRED; ++i) {
        __syncthreads();
#pragma unroll
        for (int l = 0; l < 4; ++l) sh[((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 4 + l] = acc[i][l];
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                double a = 0.0;
                for (int y = 0; y < (int)blockDim.y; ++y) a += sh[((size_t)y * blockDim.x + threadIdx.x) * 4 + l];
                atomicAdd(out + (size_t)i * C + c0 + l, a);
            }
        }
    }
}

__global__ void __launch_bounds__(256) edge_stats_kernel(const float* __restrict__ uw, size_t ld, const long long* __restrict__ idx, long long P,
                                                        int N, int k, int C, double* __restrict__ sums) {
    extern __shared__ double edge_sh[];
    const EdgeGeo ge = edge_geo();
    double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (long long pt = ge.p0; pt < P; pt += ge.pstride) {
        const long long b = pt / N;
        const int n = (int)(pt - b * N);
        const V4x3 wp = ld43(uw + (size_t)pt * 3 * ld + 2 * C + ge.c0, ld);
        for (int j = 0; j < k; ++j) {
            const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
            V4x3 p = ld43(uw + (size_t)(b * N + nb) * 3 * ld + ge.c0, ld);
            add43(p, wp);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const double nn = (double)(sqrtf(dot3l(p, p, l)) + VS_EPS);
                acc[0][l] += nn;
                acc[1][l] = fma(nn, nn, acc[1][l]);
            }
        }
    }
    edge_reduce_channels<2>(acc, sums, C, ge.c0, edge_sh);
}

template <bool FAST>
__global__ void __launch_bounds__(256) edge_fwd_kernel(const float* __restrict__ uw, size_t ld, const long long* __restrict__ idx, long long P, int N,
                                                      int k, int C, const float* __restrict__ stat, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float ns, float* __restrict__ out, size_t ldo) {
    const EdgeGeo ge = edge_geo();
    const ChanParams cp = load_params(stat, gamma, beta, C, ge.c0);
    const float k1 = 1.f - ns;
    const float inv = 1.0f / (float)k;
    for (long long pt = ge.p0; pt < P; pt += ge.pstride) {
        const long long b = pt / N;
        const int n = (int)(pt - b * N);
        const float* own = uw + (size_t)pt * 3 * ld + 2 * C + ge.c0;
        const V4x3 wp = ld43(own, ld), wd = ld43(own + C, ld);
        V4x3 acc;
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) acc.v[v][l] = 0.f;
        for (int j = 0; j < k; ++j) {
            const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
            const float* nbp = uw + (size_t)(b * N + nb) * 3 * ld + ge.c0;
            V4x3 p = ld43(nbp, ld), d = ld43(nbp + C, ld);
            add43(p, wp);
            add43(d, wd);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                float nn, nhat, nbv;
                bn_apply_lane_t<FAST>(p, l, cp, nn, nhat, nbv);
                leaky_lane_t<FAST>(p, d, l, ns, k1);
            }
            add43(acc, p);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) acc.v[v][l] *= inv;
        st43(out + (size_t)pt * 3 * ldo + ge.c0, ldo, acc);
    }
}

// lane math shared by the two backward passes.  In: raw p, d and g = dL/dout of one edge.  Out: gv <- dL/dBN(p), dv <- dL/dd,
// n / nhat / nb of the BatchNorm-on-norm, and gx_dot = <dL/dBN(p), p>.
__device__ __forceinline__ void edge_lane_bwd(const V4x3& pr, V4x3& dv, V4x3& gv, int l, const ChanParams& cp, float k1, float& n, float& nhat,
                                              float& nb, float& gx_dot) {
    // MUFU reciprocal / rsqrt like the other backward kernels (vn_math.cuh): gradients do not need op-by-op IEEE rounding
    n = fsqrt_fast(dot3l(pr, pr, l)) + VS_EPS;
    nhat = (n - cp.mean[l]) * cp.invstd[l];
    nb = nhat * cp.gamma[l] + cp.beta[l];
    const float t = nb * frcp(n);
    const float pb[3] = {pr.v[0][l] * t, pr.v[1][l] * t, pr.v[2][l] * t};
    const float s = pb[0] * dv.v[0][l] + pb[1] * dv.v[1][l] + pb[2] * dv.v[2][l];
    if (s < 0.f) {
        const float rq = frcp(dot3l(dv, dv, l) + VS_EPS);
        const float a = s * rq;
        const float gdq = dot3l(gv, dv, l) * rq;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float gval = gv.v[c][l], dval = dv.v[c][l];
            gv.v[c][l] = gval - k1 * gdq * dval;
            dv.v[c][l] = -k1 * (a * gval + gdq * pb[c] - 2.f * a * gdq * dval);
        }
    } else {
        dv.v[0][l] = dv.v[1][l] = dv.v[2][l] = 0.f;
    }
    gx_dot = gv.v[0][l] * pr.v[0][l] + gv.v[1][l] * pr.v[1][l] + gv.v[2][l] * pr.v[2][l];
}

__global__ void __launch_bounds__(256) edge_bwd_sums_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ uw, size_t ld,
                                                           const long long* __restrict__ idx, long long P, int N, int k, int C,
                                                           const float* __restrict__ stat, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float ns, double* __restrict__ sums) {
    extern __shared__ double edge_sh[];
    const EdgeGeo ge = edge_geo();
    const ChanParams cp = load_params(stat, gamma, beta, C, ge.c0);
    const float k1 = 1.f - ns;
    const float inv = 1.0f / (float)k;
    double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    for (long long pt = ge.p0; pt < P; pt += ge.pstride) {
        const long long b = pt / N;
        const int n = (int)(pt - b * N);
        const float* own = uw + (size_t)pt * 3 * ld + 2 * C + ge.c0;
        const V4x3 wp = ld43(own, ld), wd = ld43(own + C, ld);
        V4x3 g0 = ld43(g + (size_t)pt * 3 * ldg + ge.c0, ldg);
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) g0.v[v][l] *= inv;
        for (int j = 0; j < k; ++j) {
            const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
            const float* nbp = uw + (size_t)(b * N + nb) * 3 * ld + ge.c0;
            V4x3 p = ld43(nbp, ld), d = ld43(nbp + C, ld);
            add43(p, wp);
            add43(d, wd);
            V4x3 gv = g0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                float nn, nhat, nbv, gxd;
                edge_lane_bwd(p, d, gv, l, cp, k1, nn, nhat, nbv, gxd);
                const double dnb = (double)(gxd * frcp(nn));
                acc[0][l] += dnb;
                acc[1][l] = fma(dnb, (double)nhat, acc[1][l]);
            }
        }
    }
    edge_reduce_channels<2>(acc, sums, C, ge.c0, edge_sh);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256) edge_bwd_main_kernel(const float* __restrict__ g, size_t ldg, const float* __restrict__ uw, size_t ld,
                                                           const long long* __restrict__ idx, long long P, int N, int k, int C,
                                                           const float* __restrict__ stat, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float ns, const double* __restrict__ sums, double count,
                                                           int training, float* __restrict__ guw, size_t ldgu, float* __restrict__ ggamma,
                                                           float* __restrict__ gbeta) {
    const EdgeGeo ge = edge_geo();
    const ChanParams cp = load_params(stat, gamma, beta, C, ge.c0);
    const float k1 = 1.f - ns;
    const float inv = 1.0f / (float)k;
    float m1[4], m2[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const double s1 = sums[ge.c0 + l], s2 = sums[C + ge.c0 + l];
        m1[l] = training ? (float)((double)cp.gamma[l] * s1 / count) : 0.f;
        m2[l] = training ? (float)((double)cp.gamma[l] * s2 / count) : 0.f;
        if (blockIdx.x == 0 && threadIdx.y == 0) {
            ggamma[ge.c0 + l] = (float)s2;
            gbeta[ge.c0 + l] = (float)s1;
        }
    }
    for (long long pt = ge.p0; pt < P; pt += ge.pstride) {
        const long long b = pt / N;
        const int n = (int)(pt - b * N);
        const float* own = uw + (size_t)pt * 3 * ld + 2 * C + ge.c0;
        const V4x3 wp = ld43(own, ld), wd = ld43(own + C, ld);
        V4x3 g0 = ld43(g + (size_t)pt * 3 * ldg + ge.c0, ldg);
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) g0.v[v][l] *= inv;
        V4x3 sp, sd;
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int l = 0; l < 4; ++l) sp.v[v][l] = sd.v[v][l] = 0.f;
        for (int j = 0; j < k; ++j) {
            const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
            const size_t nrow = (size_t)(b * N + nb) * 3;
            const float* nbp = uw + nrow * ld + ge.c0;
            V4x3 p = ld43(nbp, ld), d = ld43(nbp + C, ld);
            add43(p, wp);
            add43(d, wd);
            V4x3 gv = g0;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                float nn, nhat, nbv, gxd;
                edge_lane_bwd(p, d, gv, l, cp, k1, nn, nhat, nbv, gxd);
                // BatchNorm-on-norm backward: gp = gpost * nb/n + dn * p / r
                const float rn = frcp(nn);
                const float dnb = gxd * rn;
                const float dn = (cp.gamma[l] * dnb - m1[l] - nhat * m2[l]) * cp.invstd[l] - gxd * nbv * rn * rn;
                const float r = nn - VS_EPS;
                const float dr = r > 0.f ? dn * frcp(r) : 0.f;
                const float t = nbv * rn;
#pragma unroll
                for (int c = 0; c < 3; ++c) gv.v[c][l] = fmaf(gv.v[c][l], t, dr * p.v[c][l]);
            }
            add43(sp, gv);
            add43(sd, d);
            float* dst = guw + nrow * ldgu + ge.c0;
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                red_add_v4(dst + v * ldgu, gv.v[v][0], gv.v[v][1], gv.v[v][2], gv.v[v][3]);
                red_add_v4(dst + v * ldgu + C, d.v[v][0], d.v[v][1], d.v[v][2], d.v[v][3]);
            }
        }
        float* od = guw + (size_t)pt * 3 * ldgu + 2 * C + ge.c0;
        st43(od, ldgu, sp);
        st43(od + C, ldgu, sd);
    }
}

static bool edge_ok(int C, const void* uw, long long ld) {
    return C >= 4 && C % 4 == 0 && C <= 1024 && ld % 4 == 0 && (((uintptr_t)uw) & 15) == 0 && 256 % (C / 4) == 0;
}

static void edge_launch_geo(long long P, int C, dim3& grid, dim3& block, size_t& smem) {
    block = dim3((unsigned)(C / 4), (unsigned)(256 / (C / 4)));
    long long nb = (P + block.y - 1) / block.y;
    const long long cap = (long long)sm_count() * 8;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    grid = dim3((unsigned)nb);
    smem = (size_t)256 * 4 * sizeof(double);
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

// sums: 2C doubles (zeroed here)
int vnpcc_edge_conv_stats(const float* uw, long long ld, const long long* idx, int B, int N, int k, int C, double* sums, void* stream) {
    if (!edge_ok(C, uw, ld)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    const long long P = (long long)B * N;
    if (P <= 0 || k <= 0) return last_error();
    dim3 grid, block;
    size_t smem;
    edge_launch_geo(P, C, grid, block, smem);
    count_launch(), edge_stats_kernel<<<grid, block, smem, st>>>(uw, (size_t)ld, idx, P, N, k, C, sums);
    return last_error();
}

int vnpcc_edge_conv_fwd(const float* uw, long long ld, const long long* idx, int B, int N, int k, int C, const float* stat, const float* gamma,
                        const float* beta, float ns, float* out, long long ldo, void* stream) {
    if (!edge_ok(C, uw, ld) || ldo % 4 != 0 || (((uintptr_t)out) & 15) || stat == nullptr) return VNPCC_ERR_UNSUPPORTED;
    const long long P = (long long)B * N;
    if (P <= 0) return 0;
    if (k <= 0) return VNPCC_ERR_BAD_ARG;
    dim3 grid, block;
    size_t smem;
    edge_launch_geo(P, C, grid, block, smem);
    if (fast_math_enabled())
        count_launch(), edge_fwd_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(uw, (size_t)ld, idx, P, N, k, C, stat, gamma, beta, ns, out,
                                                                                    (size_t)ldo);
    else
        count_launch(), edge_fwd_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(uw, (size_t)ld, idx, P, N, k, C, stat, gamma, beta, ns, out,
                                                                                     (size_t)ldo);
    return last_error();
}

// g [P*3, C] = dL/dout.  guw [P*3, 4C] is fully written: the (U_p | U_d) half is zeroed here and accumulated with red.add, the
// (W_p | W_d) half is stored.  sums: workspace of 2C doubles.  ggamma / gbeta [C] are written.
int vnpcc_edge_conv_bwd(const float* g, long long ldg, const float* uw, long long ld, const long long* idx, int B, int N, int k, int C,
                        const float* stat, const float* gamma, const float* beta, float ns, int training, double* sums, float* guw,
                        long long ldgu, float* ggamma, float* gbeta, void* stream) {
    if (!edge_ok(C, uw, ld) || ldg % 4 != 0 || ldgu % 4 != 0 || (((uintptr_t)g) & 15) || (((uintptr_t)guw) & 15) || stat == nullptr)
        return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)B * N;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    if (P <= 0) return last_error();
    if (k <= 0) return VNPCC_ERR_BAD_ARG;
    cudaMemset2DAsync(guw, (size_t)ldgu * sizeof(float), 0, (size_t)2 * C * sizeof(float), (size_t)P * 3, st);
    dim3 grid, block;
    size_t smem;
    edge_launch_geo(P, C, grid, block, smem);
    count_launch(), edge_bwd_sums_kernel<<<grid, block, smem, st>>>(g, (size_t)ldg, uw, (size_t)ld, idx, P, N, k, C, stat, gamma, beta, ns, sums);
    count_launch(), edge_bwd_main_kernel<<<grid, block, 0, st>>>(g, (size_t)ldg, uw, (size_t)ld, idx, P, N, k, C, stat, gamma, beta, ns, sums,
                                                                (double)P * k, training, guw, (size_t)ldgu, ggamma, gbeta);
    return last_error();
}

}  // extern "C"
