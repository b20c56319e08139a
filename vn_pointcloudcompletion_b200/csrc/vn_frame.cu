// vn_frame.cu -- the invariant-feature step of VNStdFeature (models/vn_layers.py:197-219) on the channels-last row layout.
//
// After vn1 / vn2 / vn_lin the layer holds, per point, J = 3 (or, with normalize_frame, 2) equivariant 3-vectors z_j.  It turns them into
// a 3 x 3 frame F (rows f_0, f_1, f_2) and expresses every channel's vector in that frame:
//     normalize_frame = False :  f_k = z_k
//     normalize_frame = True  :  f_0 = z_0 / (|z_0| + eps),  w = z_1 - <z_1, f_0> f_0,  f_1 = w / (|w| + eps),  f_2 = f_0 x f_1   (eps = 1e-6)
//     x_std[point, c, k] = <x[point, c, :], f_k>                      (the reference's einsum 'bijm,bjkm->bikm' with z0 transposed)
// and returns x_std together with the frame as a [.., 3 (component), 3 (k), ..] tensor.
// Row layout: x rows (point, v) x C; z rows (point, v) x J (channel j = vector index, the layout vn_lin's GEMM produces); outputs
// x_std rows (point, k) x C and frame rows (point, k) x 3 (channel = component).  <= 40 FMAs per (point, channel): one warp per point, lanes
// over channels, the frame rebuilt in registers by every lane; the backward reduces the 9 frame gradients over the channels with shuffles
// and runs the Gram-Schmidt / cross-product adjoint on lane 0.
#include <cuda_runtime.h>
#include <stdint.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {

constexpr float FR_EPS = 1e-6f;      // models/vn_layers.py:10

struct Frame {
    float f[3][3];      // f[k][v]
    // intermediates of the normalised construction (for the adjoint)
    float v1[3], v2[3], w[3], n1, n2, s;
};

__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

template <int J>
__device__ __forceinline__ void build_frame(const float* __restrict__ z, size_t ldz, Frame& F) {
    if (J == 3) {
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int k = 0; k < 3; ++k) F.f[k][v] = __ldg(z + v * ldz + k);
        return;
    }
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        F.v1[v] = __ldg(z + v * ldz);
        F.v2[v] = __ldg(z + v * ldz + 1);
    }
    F.n1 = sqrtf(dot3(F.v1, F.v1));
#pragma unroll
    for (int v = 0; v < 3; ++v) F.f[0][v] = F.v1[v] / (F.n1 + FR_EPS);
    F.s = dot3(F.v2, F.f[0]);
#pragma unroll
    for (int v = 0; v < 3; ++v) F.w[v] = F.v2[v] - F.s * F.f[0][v];
    F.n2 = sqrtf(dot3(F.w, F.w));
#pragma unroll
    for (int v = 0; v < 3; ++v) F.f[1][v] = F.w[v] / (F.n2 + FR_EPS);
    cross3(F.f[0], F.f[1], F.f[2]);
}

template <int J>
__global__ void __launch_bounds__(256) vn_frame_fwd_kernel(const float* __restrict__ x, size_t ldx, const float* __restrict__ z, size_t ldz,
                                                           long long P, int C, float* __restrict__ out, size_t ldo,
                                                           float* __restrict__ zout) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pt = warp; pt < P; pt += nwarps) {
        Frame F;
        build_frame<J>(z + (size_t)pt * 3 * ldz, ldz, F);
        if (lane < 9) zout[(size_t)pt * 9 + lane] = F.f[lane / 3][lane % 3];      // rows (point, k) x component
        const float* xp = x + (size_t)pt * 3 * ldx;
        float* op = out + (size_t)pt * 3 * ldo;
        for (int c = lane; c < C; c += 32) {
            const float x0 = __ldg(xp + c), x1 = __ldg(xp + ldx + c), x2 = __ldg(xp + 2 * ldx + c);
#pragma unroll
            for (int k = 0; k < 3; ++k) op[k * ldo + c] = x0 * F.f[k][0] + x1 * F.f[k][1] + x2 * F.f[k][2];
        }
    }
}

template <int J>
__global__ void __launch_bounds__(256) vn_frame_bwd_kernel(const float* __restrict__ go, size_t ldgo, const float* __restrict__ gzo,
                                                           const float* __restrict__ x, size_t ldx, const float* __restrict__ z, size_t ldz,
                                                           long long P, int C, float* __restrict__ gx, size_t ldgx,
                                                           float* __restrict__ gz, size_t ldgz) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pt = warp; pt < P; pt += nwarps) {
        Frame F;
        build_frame<J>(z + (size_t)pt * 3 * ldz, ldz, F);
        const float* xp = x + (size_t)pt * 3 * ldx;
        const float* gp = go + (size_t)pt * 3 * ldgo;
        float* gxp = gx + (size_t)pt * 3 * ldgx;
        float gF[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};      // gF[k][v] = sum_c g[k, c] x[v, c]
        for (int c = lane; c < C; c += 32) {
            const float xv[3] = {__ldg(xp + c), __ldg(xp + ldx + c), __ldg(xp + 2 * ldx + c)};
            const float gk[3] = {__ldg(gp + c), __ldg(gp + ldgo + c), __ldg(gp + 2 * ldgo + c)};
#pragma unroll
            for (int v = 0; v < 3; ++v) gxp[v * ldgx + c] = gk[0] * F.f[0][v] + gk[1] * F.f[1][v] + gk[2] * F.f[2][v];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int v = 0; v < 3; ++v) gF[k][v] = fmaf(gk[k], xv[v], gF[k][v]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int v = 0; v < 3; ++v) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gF[k][v] += __shfl_xor_sync(0xffffffffu, gF[k][v], o);
            }
        if (lane == 0) {
            if (gzo) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int v = 0; v < 3; ++v) gF[k][v] += __ldg(gzo + (size_t)pt * 9 + k * 3 + v);
            }
            float* gzp = gz + (size_t)pt * 3 * ldgz;
            if (J == 3) {
#pragma unroll
                for (int v = 0; v < 3; ++v)
#pragma unroll
                    for (int k = 0; k < 3; ++k) gzp[v * ldgz + k] = gF[k][v];
            } else {
                float gu1[3] = {gF[0][0], gF[0][1], gF[0][2]}, gu2[3] = {gF[1][0], gF[1][1], gF[1][2]};
                const float* gu3 = gF[2];
                float t[3];
                cross3(F.f[1], gu3, t);      // d <f0 x f1, g> / d f0 = f1 x g
#pragma unroll
                for (int v = 0; v < 3; ++v) gu1[v] += t[v];
                cross3(gu3, F.f[0], t);      // d / d f1 = g x f0
#pragma unroll
                for (int v = 0; v < 3; ++v) gu2[v] += t[v];
                // f1 = w / (n2 + eps),  n2 = |w|
                const float d2 = F.n2 + FR_EPS;
                const float c2 = F.n2 > 0.f ? dot3(gu2, F.w) / (d2 * d2 * F.n2) : 0.f;
                float gw[3], gv2[3], gv1[3];
#pragma unroll
                for (int v = 0; v < 3; ++v) gw[v] = gu2[v] / d2 - c2 * F.w[v];
                // w = v2 - s f0,  s = <v2, f0>
                const float gs = -dot3(gw, F.f[0]);
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    gv2[v] = gw[v] + gs * F.f[0][v];
                    gu1[v] += -F.s * gw[v] + gs * F.v2[v];
                }
                // f0 = v1 / (n1 + eps),  n1 = |v1|
                const float d1 = F.n1 + FR_EPS;
                const float c1 = F.n1 > 0.f ? dot3(gu1, F.v1) / (d1 * d1 * F.n1) : 0.f;
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    gv1[v] = gu1[v] / d1 - c1 * F.v1[v];
                    gzp[v * ldgz] = gv1[v];
                    gzp[v * ldgz + 1] = gv2[v];
                }
            }
        }
    }
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

int vnpcc_vn_frame_fwd(const float* x, long long ldx, const float* z, long long ldz, long long P, int C, int J, float* out, long long ldo,
                       float* zout, void* stream) {
    if (P <= 0) return 0;
    if (J != 2 && J != 3) return VNPCC_ERR_BAD_ARG;
    const int grid = grid_for((size_t)P * 32, 256, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (J == 3)
        count_launch(), vn_frame_fwd_kernel<3><<<grid, 256, 0, st>>>(x, (size_t)ldx, z, (size_t)ldz, P, C, out, (size_t)ldo, zout);
    else
        count_launch(), vn_frame_fwd_kernel<2><<<grid, 256, 0, st>>>(x, (size_t)ldx, z, (size_t)ldz, P, C, out, (size_t)ldo, zout);
    return last_error();
}

int vnpcc_vn_frame_bwd(const float* gout, long long ldgo, const float* gzout, const float* x, long long ldx, const float* z, long long ldz,
                       long long P, int C, int J, float* gx, long long ldgx, float* gz, long long ldgz, void* stream) {
    if (P <= 0) return 0;
    if (J != 2 && J != 3) return VNPCC_ERR_BAD_ARG;
    const int grid = grid_for((size_t)P * 32, 256, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (J == 3)
        count_launch(), vn_frame_bwd_kernel<3><<<grid, 256, 0, st>>>(gout, (size_t)ldgo, gzout, x, (size_t)ldx, z, (size_t)ldz, P, C, gx,
                                                                     (size_t)ldgx, gz, (size_t)ldgz);
    else
        count_launch(), vn_frame_bwd_kernel<2><<<grid, 256, 0, st>>>(gout, (size_t)ldgo, gzout, x, (size_t)ldx, z, (size_t)ldz, P, C, gx,
                                                                     (size_t)ldgx, gz, (size_t)ldgz);
    return last_error();
}

}  // extern "C"
