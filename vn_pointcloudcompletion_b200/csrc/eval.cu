// eval.cu -- evaluation extras of test.py:73-78 on the GPU, so that an 8-GPU evaluation is not bottlenecked
// by the reference's per-sample CPU loops:
//   fscore_counts     metrics/metric.py:31-48 f_score (open3d compute_point_cloud_distance = Euclidean NN distances, threshold th):
//                     precision / recall / F from the Chamfer search's squared distances
//   voxel_occupancy   utils/voxel_util.py:89-105 points_to_voxels (pyntcloud VoxelGrid(n_x=n_y=n_z=size_grid): the cloud's own bounding
//                     box made a cube, np.linspace segments, searchsorted - 1) as a bit mask
//   voxel_iou         utils/voxel_util.py:5-14 iou = |A and B| / |A or B|
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : fminf(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : fminf(r, sh[i]);
    return r;
}

// one block per sample: out[b] = (precision, recall, f)
__global__ void __launch_bounds__(256) fscore_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int N, int M, float th,
                                                    float* __restrict__ out) {
    __shared__ unsigned cnt[2];
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.x;
    unsigned c1 = 0, c2 = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) c1 += sqrtf(__ldg(d1 + (size_t)b * N + i)) < th;
    for (int i = threadIdx.x; i < M; i += blockDim.x) c2 += sqrtf(__ldg(d2 + (size_t)b * M + i)) < th;
    c1 = __reduce_add_sync(0xffffffffu, c1);
    c2 = __reduce_add_sync(0xffffffffu, c2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&cnt[0], c1);
        atomicAdd(&cnt[1], c2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float precision = N > 0 ? (float)cnt[0] / (float)N : 0.f;
        const float recall = M > 0 ? (float)cnt[1] / (float)M : 0.f;
        out[b * 3] = precision;
        out[b * 3 + 1] = recall;
        out[b * 3 + 2] = (precision + recall) > 0.f ? 2.f * recall * precision / (recall + precision) : 0.f;
    }
}

// searchsorted(segments, x, side='left') - 1 clipped to [0, n-1], segments[i] = lo + i * step (double), step = (hi - lo) / n
__device__ __forceinline__ int voxel_index(double x, double lo, double hi, int n) {
    const double step = (hi - lo) / (double)n;
    if (!(step > 0.0)) return 0;
    int i = (int)ceil((x - lo) / step);
    if (i < 0) i = 0;
    if (i > n) i = n;
    // smallest i with seg[i] >= x  (seg[n] is exactly hi in np.linspace)
    while (i > 0 && ((i - 1 == n) ? hi : lo + (i - 1) * step) >= x) --i;
    while (i < n && ((i == n) ? hi : lo + i * step) < x) ++i;
    int v = i - 1;
    return v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
}

// one block per sample; bits [B, words] pre-zeroed, words = ceil(n^3 / 32); voxel (x, y, z) -> bit (x*n + y)*n + z
__global__ void __launch_bounds__(256) voxel_occupancy_kernel(const float* __restrict__ xyz, int N, int n, unsigned* __restrict__ bits, int words) {
    __shared__ float sh[8];
    const int b = blockIdx.x;
    const float* p = xyz + (size_t)b * N * 3;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < N; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(p + (size_t)i * 3 + a);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    double dlo[3], dhi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        dlo[a] = (double)block_reduce(lo[a], false, sh);
        dhi[a] = (double)block_reduce(hi[a], true, sh);
    }
    // regular_bounding_box: grow the shorter axes symmetrically to the longest extent
    double ext = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) ext = fmax(ext, dhi[a] - dlo[a]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double margin = ext - (dhi[a] - dlo[a]);
        dlo[a] -= margin / 2;
        dhi[a] += margin / 2;
    }
    unsigned* bb = bits + (size_t)b * words;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int vx = voxel_index((double)__ldg(p + (size_t)i * 3), dlo[0], dhi[0], n);
        const int vy = voxel_index((double)__ldg(p + (size_t)i * 3 + 1), dlo[1], dhi[1], n);
        const int vz = voxel_index((double)__ldg(p + (size_t)i * 3 + 2), dlo[2], dhi[2], n);
        const unsigned lin = (unsigned)((vx * n + vy) * n + vz);
        atomicOr(bb + (lin >> 5), 1u << (lin & 31));
    }
}

__global__ void __launch_bounds__(256) voxel_iou_kernel(const unsigned* __restrict__ a, const unsigned* __restrict__ bmask, int words,
                                                       float* __restrict__ out) {
    __shared__ unsigned cnt[2];
    if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.x;
    unsigned ci = 0, cu = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        const unsigned x = __ldg(a + (size_t)b * words + i), y = __ldg(bmask + (size_t)b * words + i);
        ci += __popc(x & y);
        cu += __popc(x | y);
    }
    ci = __reduce_add_sync(0xffffffffu, ci);
    cu = __reduce_add_sync(0xffffffffu, cu);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&cnt[0], ci);
        atomicAdd(&cnt[1], cu);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[b] = cnt[1] ? (float)cnt[0] / (float)cnt[1] : 0.f;
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

int vnpcc_fscore(const float* dist1, const float* dist2, int B, int N, int M, float th, float* out, void* stream) {
    if (B <= 0) return 0;
    count_launch(), fscore_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(dist1, dist2, N, M, th, out);
    return last_error();
}

int vnpcc_voxel_occupancy(const float* xyz, int B, int N, int size_grid, unsigned* bits, void* stream) {
    if (size_grid <= 0 || size_grid > 1024) return VNPCC_ERR_BAD_ARG;
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long cells = (long long)size_grid * size_grid * size_grid;
    const int words = (int)((cells + 31) / 32);
    cudaMemsetAsync(bits, 0, (size_t)B * words * sizeof(unsigned), st);
    if (N <= 0) return last_error();
    count_launch(), voxel_occupancy_kernel<<<B, 256, 0, st>>>(xyz, N, size_grid, bits, words);
    return last_error();
}

int vnpcc_voxel_iou(const unsigned* bits_a, const unsigned* bits_b, int B, int words, float* out, void* stream) {
    if (B <= 0) return 0;
    count_launch(), voxel_iou_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(bits_a, bits_b, words, out);
    return last_error();
}

}  // extern "C"
