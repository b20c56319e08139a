// loss_variants.cu -- the tails of the Chamfer-based loss variants as kernels on the search's (dist, idx) outputs:
//
//   cd_persample_{fwd,bwd}   utils/loss.py:14-31 calc_cd: per-sample  cd_p = (mean sqrt d1 + mean sqrt d2) / 2,  cd_t = mean d1 + mean d2
//                            and the four directed means of its `separate` form
//   fscore_sq                extensions/ChamferDistancePytorch/fscore.py:3-16: precision_k = mean(dist_k < threshold) on SQUARED
//                            distances, f = 2 p1 p2 / (p1 + p2) (0 where that is 0 / 0)
//   nn_counts                the torch.bincount of utils/loss.py:57,62: count[b, k] = #{j : idx[b, j] == k}
//   dcd_{fwd,bwd}            utils/loss.py:33-74 density-aware Chamfer distance:
//                              loss_b = 1/2 [ mean_j (1 - exp(-alpha d1[b,j]) frac_21 / (count1[b, idx1[b,j]]^lambda + 1e-6))
//                                           + mean_k (1 - exp(-alpha d2[b,k]) frac_12 / (count2[b, idx2[b,k]]^lambda + 1e-6)) ]
//                            (the counts are detached in the reference: they carry no gradient)
//
// One CTA per (sample, direction); per-thread fp64 partial sums, deterministic in-block tree (no atomics on the results).
#include <cuda_runtime.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += sh[i];
    return r;
}

// grid (B, 2): direction 0 reduces dist1 [B,N], direction 1 dist2 [B,M]; part[b][dir] = (mean sqrt, mean)
__global__ void __launch_bounds__(256) cd_persample_part_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int N, int M,
                                                                 float* __restrict__ part) {
    __shared__ double sh[8];
    const int b = blockIdx.x, dir = blockIdx.y;
    const int n = dir ? M : N;
    const float* d = (dir ? d2 : d1) + (size_t)b * n;
    double ss = 0.0, sm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = __ldg(d + i);
        ss += (double)sqrtf(v);
        sm += (double)v;
    }
    ss = block_sum(ss, sh);
    sm = block_sum(sm, sh);
    if (threadIdx.x == 0) {
        // torch.mean returns fp32: each directed mean is rounded to fp32 before the reference combines them
        part[(b * 2 + dir) * 2 + 0] = n > 0 ? (float)(ss / n) : 0.f;
        part[(b * 2 + dir) * 2 + 1] = n > 0 ? (float)(sm / n) : 0.f;
    }
}

// out[b] = (cd_p, cd_t, mean sqrt d1, mean sqrt d2, mean d1, mean d2)
__global__ void cd_persample_finish_kernel(const float* __restrict__ part, int B, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float s1 = part[b * 4 + 0], m1 = part[b * 4 + 1], s2 = part[b * 4 + 2], m2 = part[b * 4 + 3];
    out[b * 6 + 0] = (s1 + s2) / 2;
    out[b * 6 + 1] = m1 + m2;
    out[b * 6 + 2] = s1;
    out[b * 6 + 3] = s2;
    out[b * 6 + 4] = m1;
    out[b * 6 + 5] = m2;
}

// gdist_k[b, j] = (gout[b,0]/2 + gout[b,2+k]) / (n 2 sqrt d) + (gout[b,1] + gout[b,4+k]) / n
__global__ void __launch_bounds__(256) cd_persample_bwd_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int N, int M,
                                                                const float* __restrict__ gout, float* __restrict__ g1,
                                                                float* __restrict__ g2) {
    const int b = blockIdx.x, dir = blockIdx.y;
    const int n = dir ? M : N;
    const float* d = (dir ? d2 : d1) + (size_t)b * n;
    float* g = (dir ? g2 : g1) + (size_t)b * n;
    const float* go = gout + (size_t)b * 6;
    const float cs = (0.5f * __ldg(go) + __ldg(go + 2 + dir)) / (float)n;
    const float cm = (__ldg(go + 1) + __ldg(go + 4 + dir)) / (float)n;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x)
        g[i] = cs * (0.5f / sqrtf(__ldg(d + i))) + cm;      // d == 0 -> inf, like autograd through torch.sqrt (0 * inf = nan only if cs == 0)
}

// out[b] = (f, precision_1, precision_2) on squared distances
__global__ void __launch_bounds__(256) fscore_sq_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int N, int M, float th,
                                                        float* __restrict__ out) {
    __shared__ unsigned cnt[2];
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.x;
    unsigned c1 = 0, c2 = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) c1 += __ldg(d1 + (size_t)b * N + i) < th;
    for (int i = threadIdx.x; i < M; i += blockDim.x) c2 += __ldg(d2 + (size_t)b * M + i) < th;
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&cnt[0], c1);
        atomicAdd(&cnt[1], c2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float p1 = (float)cnt[0] / (float)N, p2 = (float)cnt[1] / (float)M;      // N == 0 -> nan, like torch.mean of an empty row
        float f = 2.f * p1 * p2 / (p1 + p2);
        if (f != f) f = 0.f;                                                            // fscore[isnan] = 0
        out[b * 3] = f;
        out[b * 3 + 1] = p1;
        out[b * 3 + 2] = p2;
    }
}

__global__ void __launch_bounds__(256) nn_counts_kernel(const int* __restrict__ idx, long long total, int N, int K, int* __restrict__ counts) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = __ldg(idx + t);
        if ((unsigned)k < (unsigned)K) atomicAdd(counts + (t / N) * K + k, 1);
    }
}

__device__ __forceinline__ float dcd_weight(int count, float n_lambda, float frac) {
    const float c = (float)count;
    const float cl = n_lambda == 1.f ? c : powf(c, n_lambda);
    return frac / (cl + 1e-6f);      // (count^lambda + 1e-6)^-1 * frac
}

// grid (B, 2); part[b][dir] = mean_j (1 - exp(-alpha d) w)
__global__ void __launch_bounds__(256) dcd_part_kernel(const float* __restrict__ d1, const float* __restrict__ d2, const int* __restrict__ i1,
                                                        const int* __restrict__ i2, const int* __restrict__ c1, const int* __restrict__ c2,
                                                        int N, int M, float alpha, float n_lambda, float frac_21, float frac_12,
                                                        float* __restrict__ part) {
    __shared__ double sh[8];
    const int b = blockIdx.x, dir = blockIdx.y;
    const int n = dir ? M : N, K = dir ? N : M;      // dist1 [B,N] indexes the other cloud (M points) and vice versa
    const float* d = (dir ? d2 : d1) + (size_t)b * n;
    const int* ix = (dir ? i2 : i1) + (size_t)b * n;
    const int* cn = (dir ? c2 : c1) + (size_t)b * K;
    const float frac = dir ? frac_12 : frac_21;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float w = dcd_weight(__ldg(cn + __ldg(ix + i)), n_lambda, frac);
        s += (double)(-expf(-__ldg(d + i) * alpha) * w + 1.f);
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[b * 2 + dir] = n > 0 ? (float)(s / n) : 0.f;
}

__global__ void dcd_finish_kernel(const float* __restrict__ part, int B, float* __restrict__ loss) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) loss[b] = (part[b * 2] + part[b * 2 + 1]) / 2;
}

// d loss_b / d dist_k[b, j] = gloss[b] / 2 / n * alpha exp(-alpha d) w
__global__ void __launch_bounds__(256) dcd_bwd_kernel(const float* __restrict__ d1, const float* __restrict__ d2, const int* __restrict__ i1,
                                                       const int* __restrict__ i2, const int* __restrict__ c1, const int* __restrict__ c2,
                                                       int N, int M, float alpha, float n_lambda, float frac_21, float frac_12,
                                                       const float* __restrict__ gloss, float* __restrict__ g1, float* __restrict__ g2) {
    const int b = blockIdx.x, dir = blockIdx.y;
    const int n = dir ? M : N, K = dir ? N : M;
    const float* d = (dir ? d2 : d1) + (size_t)b * n;
    const int* ix = (dir ? i2 : i1) + (size_t)b * n;
    const int* cn = (dir ? c2 : c1) + (size_t)b * K;
    float* g = (dir ? g2 : g1) + (size_t)b * n;
    const float frac = dir ? frac_12 : frac_21;
    const float s = __ldg(gloss + b) * 0.5f / (float)n * alpha;
    for (int i = blockIdx.z * blockDim.x + threadIdx.x; i < n; i += gridDim.z * blockDim.x)
        g[i] = s * expf(-__ldg(d + i) * alpha) * dcd_weight(__ldg(cn + __ldg(ix + i)), n_lambda, frac);
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

int vnpcc_cd_persample_fwd(const float* dist1, const float* dist2, int B, int N, int M, float* part, float* out, void* stream) {
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    count_launch(), cd_persample_part_kernel<<<dim3(B, 2), 256, 0, st>>>(dist1, dist2, N, M, part);
    count_launch(), cd_persample_finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(part, B, out);
    return last_error();
}

int vnpcc_cd_persample_bwd(const float* dist1, const float* dist2, int B, int N, int M, const float* gout, float* gdist1, float* gdist2,
                           void* stream) {
    if (B <= 0 || (N <= 0 && M <= 0)) return 0;
    int z = (sm_count() * 4 + 2 * B - 1) / (2 * B);
    const int maxz = ((N > M ? N : M) + 255) / 256;
    if (z > maxz) z = maxz;
    if (z < 1) z = 1;
    count_launch(), cd_persample_bwd_kernel<<<dim3(B, 2, z), 256, 0, (cudaStream_t)stream>>>(dist1, dist2, N, M, gout, gdist1, gdist2);
    return last_error();
}

int vnpcc_fscore_sq(const float* dist1, const float* dist2, int B, int N, int M, float threshold, float* out, void* stream) {
    if (B <= 0) return 0;
    count_launch(), fscore_sq_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dist1, dist2, N, M, threshold, out);
    return last_error();
}

int vnpcc_nn_counts(const int* idx, int B, int N, int K, int* counts, void* stream) {
    if (B <= 0 || K <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)B * K, st);
    if (N <= 0) return last_error();
    const long long total = (long long)B * N;
    count_launch(), nn_counts_kernel<<<grid_for((size_t)total, 256, 8), 256, 0, st>>>(idx, total, N, K, counts);
    return last_error();
}

int vnpcc_dcd_fwd(const float* dist1, const float* dist2, const int* idx1, const int* idx2, const int* count1, const int* count2, int B,
                  int N, int M, float alpha, float n_lambda, float frac_21, float frac_12, float* part, float* loss, void* stream) {
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    count_launch(), dcd_part_kernel<<<dim3(B, 2), 256, 0, st>>>(dist1, dist2, idx1, idx2, count1, count2, N, M, alpha, n_lambda, frac_21,
                                                                frac_12, part);
    count_launch(), dcd_finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(part, B, loss);
    return last_error();
}

int vnpcc_dcd_bwd(const float* dist1, const float* dist2, const int* idx1, const int* idx2, const int* count1, const int* count2, int B,
                  int N, int M, float alpha, float n_lambda, float frac_21, float frac_12, const float* gloss, float* gdist1,
                  float* gdist2, void* stream) {
    if (B <= 0 || (N <= 0 && M <= 0)) return 0;
    int z = (sm_count() * 4 + 2 * B - 1) / (2 * B);
    const int maxz = ((N > M ? N : M) + 255) / 256;
    if (z > maxz) z = maxz;
    if (z < 1) z = 1;
    count_launch(), dcd_bwd_kernel<<<dim3(B, 2, z), 256, 0, (cudaStream_t)stream>>>(dist1, dist2, idx1, idx2, count1, count2, N, M, alpha,
                                                                                    n_lambda, frac_21, frac_12, gloss, gdist1, gdist2);
    return last_error();
}

}  // extern "C"
