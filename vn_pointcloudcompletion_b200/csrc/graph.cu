// graph.cu -- point-set graph kernels of the VN_DGCNN_fps encoder for sm_100a:
//   knn3d            replaces knn_cuda.KNN(k, transpose_mode=False) as called at models/dgcnn.py:11,236,257-259 (always on
//                    3-D coordinates on this path); third-party wheel KNN_CUDA 0.2 (README.md:31), source not vendored
//   fps              replaces pointnet2_ops.pointnet2_utils.furthest_point_sample (models/dgcnn.py:15,210)
//   points_gather    replaces pointnet2_utils.gather_operation (models/dgcnn.py:16,215-219) on the row layout (+ adjoint)
//   edge_feature     replaces VN_DGCNN_fps.vn_get_graph_feature's index / cat(x_j - x_i, x_i) (models/dgcnn.py:251-278) (+ adjoint)
//   rows_group_mean  replaces mean_pool over the k neighbours (models/vn_layers.py:170-171, dgcnn.py:291,298,304,311) (+ adjoint)
//
// All of them are HBM / latency bound integer-and-gather work: coalesced channel-fastest accesses, candidates staged in
// shared memory, no tensor cores.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

// -------------------------------------------------------------------------------------------------------------
// 1. k nearest neighbours in 3-D.  One thread per query, candidates staged in shared memory in tiles, the running
//    top-K kept sorted in registers.  Distance = fma(dz,dz, fma(dy,dy, dx*dx)) of fp32 differences; the result is
//    ordered by (distance, index): candidates are visited in index order and only a strictly smaller distance moves
//    an entry forward, so equal distances keep the lower index first.
// -------------------------------------------------------------------------------------------------------------
constexpr int KNN_TILE = 1024;
constexpr int KNN_BLOCK = 128;

template <int KMAX>
__global__ void __launch_bounds__(KNN_BLOCK) knn3d_kernel(const float* __restrict__ ref, const float* __restrict__ query, int Nr,
                                                         int Nq, int k, long long* __restrict__ idx, float* __restrict__ dist) {
    __shared__ float4 tile[KNN_TILE];
    const int b = blockIdx.y;
    const int q = blockIdx.x * KNN_BLOCK + threadIdx.x;
    const float* rb = ref + (size_t)b * Nr * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (q < Nq) {
        const float* qp = query + ((size_t)b * Nq + q) * 3;
        qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
    }
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
        bd[t] = FLT_MAX;
        bi[t] = 0;
    }
    for (int base = 0; base < Nr; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, Nr - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += KNN_BLOCK) {
            const float* p = rb + (size_t)(base + t) * 3;
            tile[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (q < Nq) {
#pragma unroll 4
            for (int t = 0; t < cnt; ++t) {
                const float4 c = tile[t];
                const float dx = c.x - qx, dy = c.y - qy, dz = c.z - qz;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d < bd[KMAX - 1]) {
                    bd[KMAX - 1] = d;
                    bi[KMAX - 1] = base + t;
#pragma unroll
                    for (int s = KMAX - 1; s > 0; --s) {
                        if (bd[s] < bd[s - 1]) {
                            const float td = bd[s];
                            bd[s] = bd[s - 1];
                            bd[s - 1] = td;
                            const int ti = bi[s];
                            bi[s] = bi[s - 1];
                            bi[s - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (q < Nq) {
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            if (t < k) {
                const size_t o = ((size_t)b * k + t) * Nq + q;
                idx[o] = bi[t];
                if (dist) dist[o] = sqrtf(bd[t]);     // knn_cuda returns Euclidean (not squared) distances
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// 2. furthest point sampling.  One CTA per sample (the selection is sequential), every thread keeps PER points and
//    their running minimum distance in registers.  Semantics of pointnet2_ops' kernel: start at point 0, running
//    distance initialised to 1e10, points with |p|^2 <= 1e-3 never compete, next = arg-max of the running distance
//    (lowest index among exact ties).
// -------------------------------------------------------------------------------------------------------------
constexpr int FPS_BLOCK = 1024;

template <int PER>
__global__ void __launch_bounds__(FPS_BLOCK) fps_kernel(const float* __restrict__ xyz, int N, int M, int* __restrict__ out) {
    __shared__ unsigned wv[32];
    __shared__ unsigned wi[32];
    __shared__ int s_old;
    const int b = blockIdx.x;
    const float* pb = xyz + (size_t)b * N * 3;
    int* ob = out + (size_t)b * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float px[PER], py[PER], pz[PER], md[PER];
    bool live[PER];
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int n = tid + t * FPS_BLOCK;
        live[t] = false;
        px[t] = py[t] = pz[t] = 0.f;
        md[t] = 1e10f;
        if (n < N) {
            px[t] = __ldg(pb + (size_t)n * 3), py[t] = __ldg(pb + (size_t)n * 3 + 1), pz[t] = __ldg(pb + (size_t)n * 3 + 2);
            const float mag = fmaf(pz[t], pz[t], fmaf(py[t], py[t], px[t] * px[t]));
            live[t] = !((double)mag <= 1e-3);
        }
    }
    int old = 0;
    if (tid == 0) ob[0] = 0;
    for (int j = 1; j < M; ++j) {
        const float ox = __ldg(pb + (size_t)old * 3), oy = __ldg(pb + (size_t)old * 3 + 1), oz = __ldg(pb + (size_t)old * 3 + 2);
        unsigned bestv = 0u, besti = 0xffffffffu;      // bestv = float bits of the best distance (+1 so that "none" = 0)
#pragma unroll
        for (int t = 0; t < PER; ++t) {
            if (live[t]) {
                const float dx = px[t] - ox, dy = py[t] - oy, dz = pz[t] - oz;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                md[t] = fminf(d, md[t]);
                const unsigned v = __float_as_uint(md[t]) + 1u;       // md >= 0: bit pattern is monotone
                if (v > bestv) {
                    bestv = v;
                    besti = (unsigned)(tid + t * FPS_BLOCK);
                }
            }
        }
        // warp arg-max: maximum value, then the lowest index holding it
        unsigned wmax = __reduce_max_sync(0xffffffffu, bestv);
        unsigned wmin = __reduce_min_sync(0xffffffffu, bestv == wmax ? besti : 0xffffffffu);
        if (lane == 0) {
            wv[warp] = wmax;
            wi[warp] = wmin;
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned v = wv[lane], i = wi[lane];
            const unsigned m = __reduce_max_sync(0xffffffffu, v);
            const unsigned mi = __reduce_min_sync(0xffffffffu, v == m ? i : 0xffffffffu);
            if (lane == 0) {
                const int nxt = (m == 0u || mi == 0xffffffffu) ? 0 : (int)mi;      // no live point: the reference keeps index 0
                s_old = nxt;
                ob[j] = nxt;
            }
        }
        __syncthreads();
        old = s_old;
    }
}

// -------------------------------------------------------------------------------------------------------------
// 3. gather of M points per sample (rows (b,n,v) x C  ->  rows (b,m,v) x C) and its adjoint
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) points_gather_kernel(const float* __restrict__ x, size_t ldx, const int* __restrict__ idx, int N,
                                                           int M, int C, long long total, float* __restrict__ out, size_t ldo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;            // output row (b, m, v)
        const int v = (int)(r % 3);
        const long long pm = r / 3;           // b*M + m
        const long long b = pm / M;
        const int n = __ldg(idx + pm);
        out[(size_t)r * ldo + c] = __ldg(x + (size_t)((b * N + n) * 3 + v) * ldx + c);
    }
}

__global__ void __launch_bounds__(256) points_scatter_kernel(const float* __restrict__ g, size_t ldg, const int* __restrict__ idx, int N,
                                                            int M, int C, long long total, float* __restrict__ gx, size_t ldgx) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;
        const int v = (int)(r % 3);
        const long long pm = r / 3;
        const long long b = pm / M;
        const int n = __ldg(idx + pm);
        atomicAdd(gx + (size_t)((b * N + n) * 3 + v) * ldgx + c, __ldg(g + (size_t)r * ldg + c));
    }
}

// -------------------------------------------------------------------------------------------------------------
// 4. edge features.  x rows (b,n,v) x C, idx [B,k,N] (knn layout) -> out rows ((b,n,j),v) x 2C:
//        out[.., c] = x[(b, idx[b,j,n]), v, c] - x[(b,n), v, c]        out[.., C + c] = x[(b,n), v, c]
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_feature_fwd_kernel(const float* __restrict__ x, size_t ldx, const long long* __restrict__ idx,
                                                              int N, int k, int C, long long total, float* __restrict__ out, size_t ldo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;            // output row ((b,n,j), v)
        const int v = (int)(r % 3);
        const long long e = r / 3;            // (b*N + n)*k + j
        const int j = (int)(e % k);
        const long long pn = e / k;           // b*N + n
        const long long b = pn / N;
        const int n = (int)(pn - b * N);
        const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
        const float xi = __ldg(x + (size_t)(pn * 3 + v) * ldx + c);
        const float xj = __ldg(x + (size_t)((b * N + nb) * 3 + v) * ldx + c);
        out[(size_t)r * ldo + c] = xj - xi;
        out[(size_t)r * ldo + C + c] = xi;
    }
}

// adjoint: one thread per (b, n, v, c) walks its k edges: gx[(b,n)] += sum_j (g2 - g1) ; gx[(b, nbr_j)] += g1
__global__ void __launch_bounds__(256) edge_feature_bwd_kernel(const float* __restrict__ g, size_t ldg, const long long* __restrict__ idx,
                                                              int N, int k, int C, long long total, float* __restrict__ gx, size_t ldgx) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;            // input-point row (b, n, v)
        const int v = (int)(r % 3);
        const long long pn = r / 3;
        const long long b = pn / N;
        const int n = (int)(pn - b * N);
        float own = 0.f;
        for (int j = 0; j < k; ++j) {
            const size_t er = (size_t)((pn * k + j) * 3 + v);
            const float g1 = __ldg(g + er * ldg + c);
            const float g2 = __ldg(g + er * ldg + C + c);
            own += g2 - g1;
            const long long nb = __ldg(idx + ((size_t)b * k + j) * N + n);
            atomicAdd(gx + (size_t)((b * N + nb) * 3 + v) * ldgx + c, g1);
        }
        atomicAdd(gx + (size_t)r * ldgx + c, own);
    }
}

// -------------------------------------------------------------------------------------------------------------
// 5. mean over groups of k consecutive points (rows ((g,j),v) x C -> rows (g,v) x C) and its adjoint
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_mean_fwd_kernel(const float* __restrict__ x, size_t ldx, int k, int C, long long total,
                                                            float* __restrict__ out, size_t ldo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.0f / (float)k;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;            // output row (g, v)
        const int v = (int)(r % 3);
        const long long gp = r / 3;
        float s = 0.f;
        for (int j = 0; j < k; ++j) s += __ldg(x + (size_t)((gp * k + j) * 3 + v) * ldx + c);
        out[(size_t)r * ldo + c] = s * inv;
    }
}

__global__ void __launch_bounds__(256) group_mean_bwd_kernel(const float* __restrict__ g, size_t ldg, int k, int C, long long total,
                                                            float* __restrict__ gx, size_t ldgx) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float inv = 1.0f / (float)k;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int c = (int)(t % C);
        const long long r = t / C;            // gx row ((g,j), v)
        const int v = (int)(r % 3);
        const long long gp = (r / 3) / k;
        gx[(size_t)r * ldgx + c] = __ldg(g + (size_t)(gp * 3 + v) * ldg + c) * inv;
    }
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

int vnpcc_knn3d(const float* ref, const float* query, int B, int Nr, int Nq, int k, long long* idx, float* dist, void* stream) {
    if (B < 0 || Nr < 0 || Nq < 0 || k <= 0 || k > 32 || k > Nr) return VNPCC_ERR_BAD_ARG;
    if (B == 0 || Nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((Nq + KNN_BLOCK - 1) / KNN_BLOCK), (unsigned)B);
    if (k <= 8)
        count_launch(), knn3d_kernel<8><<<grid, KNN_BLOCK, 0, st>>>(ref, query, Nr, Nq, k, idx, dist);
    else if (k <= 16)
        count_launch(), knn3d_kernel<16><<<grid, KNN_BLOCK, 0, st>>>(ref, query, Nr, Nq, k, idx, dist);
    else
        count_launch(), knn3d_kernel<32><<<grid, KNN_BLOCK, 0, st>>>(ref, query, Nr, Nq, k, idx, dist);
    return last_error();
}

int vnpcc_fps(const float* xyz, int B, int N, int M, int* idx, void* stream) {
    if (B < 0 || N <= 0 || M < 0) return VNPCC_ERR_BAD_ARG;
    if (B == 0 || M == 0) return 0;
    if (N > 16 * FPS_BLOCK) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int per = (N + FPS_BLOCK - 1) / FPS_BLOCK;
    if (per <= 1)
        count_launch(), fps_kernel<1><<<B, FPS_BLOCK, 0, st>>>(xyz, N, M, idx);
    else if (per <= 2)
        count_launch(), fps_kernel<2><<<B, FPS_BLOCK, 0, st>>>(xyz, N, M, idx);
    else if (per <= 4)
        count_launch(), fps_kernel<4><<<B, FPS_BLOCK, 0, st>>>(xyz, N, M, idx);
    else if (per <= 8)
        count_launch(), fps_kernel<8><<<B, FPS_BLOCK, 0, st>>>(xyz, N, M, idx);
    else
        count_launch(), fps_kernel<16><<<B, FPS_BLOCK, 0, st>>>(xyz, N, M, idx);
    return last_error();
}

int vnpcc_points_gather(const float* x, long long ldx, const int* idx, int B, int N, int M, int C, float* out, long long ldo,
                        void* stream) {
    const long long total = (long long)B * M * 3 * C;
    if (total <= 0) return 0;
    count_launch(), points_gather_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, idx, N, M, C, total,
                                                                                                    out, (size_t)ldo);
    return last_error();
}

// gx [B*N*3, C] is zeroed here, then accumulated.
int vnpcc_points_scatter_add(const float* g, long long ldg, const int* idx, int B, int N, int M, int C, float* gx, long long ldgx,
                             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || C <= 0) return 0;
    cudaMemset2DAsync(gx, (size_t)ldgx * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)B * N * 3, st);
    const long long total = (long long)B * M * 3 * C;
    if (total <= 0) return last_error();
    count_launch(), points_scatter_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, st>>>(g, (size_t)ldg, idx, N, M, C, total, gx, (size_t)ldgx);
    return last_error();
}

int vnpcc_edge_feature_fwd(const float* x, long long ldx, const long long* idx, int B, int N, int k, int C, float* out, long long ldo,
                           void* stream) {
    const long long total = (long long)B * N * k * 3 * C;
    if (total <= 0) return 0;
    count_launch(), edge_feature_fwd_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, idx, N, k, C, total,
                                                                                                       out, (size_t)ldo);
    return last_error();
}

// gx [B*N*3, C] is zeroed here, then accumulated (fp32 atomics: summation order is not deterministic, like the
// reference's index_put_(accumulate=True) backward of x[idx, :]).
int vnpcc_edge_feature_bwd(const float* g, long long ldg, const long long* idx, int B, int N, int k, int C, float* gx, long long ldgx,
                           void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 0 || N <= 0 || C <= 0) return 0;
    cudaMemset2DAsync(gx, (size_t)ldgx * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)B * N * 3, st);
    const long long total = (long long)B * N * 3 * C;
    if (k <= 0) return last_error();
    count_launch(), edge_feature_bwd_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, st>>>(g, (size_t)ldg, idx, N, k, C, total, gx, (size_t)ldgx);
    return last_error();
}

int vnpcc_rows_group_mean(const float* x, long long ldx, long long G, int k, int C, float* out, long long ldo, void* stream) {
    const long long total = G * 3 * C;
    if (total <= 0 || k <= 0) return total <= 0 ? 0 : VNPCC_ERR_BAD_ARG;
    count_launch(), group_mean_fwd_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, (size_t)ldx, k, C, total, out,
                                                                                                     (size_t)ldo);
    return last_error();
}

int vnpcc_rows_group_mean_bwd(const float* g, long long ldg, long long G, int k, int C, float* gx, long long ldgx, void* stream) {
    const long long total = G * k * 3 * C;
    if (total <= 0) return 0;
    count_launch(), group_mean_bwd_kernel<<<grid_for((size_t)total, 256, 16), 256, 0, (cudaStream_t)stream>>>(g, (size_t)ldg, k, C, total, gx,
                                                                                                     (size_t)ldgx);
    return last_error();
}

}  // extern "C"
