// gemm_simt.cu -- fp32-exact (no TF32 rounding) SIMT GEMMs over the channels-last row layout.
//
// This is the PARITY-mode contraction for VNLinear-family layers (models/vn_layers.py:21,38,65,69,162,194 all
// reduce to  Y[r, o] = sum_k X[r, k] * W[o, k]  on rows r = (b, n, v)), and the only path for shapes the
// tensor-core kernel (gemm_tcgen05.cu) does not take (K not a multiple of 8, tiny K such as the 1- and 2-channel
// inputs of first_conv[0] / final_conv[0]).  Every product is an fp32 FMA, accumulated in k order per output,
// so results match a cuBLAS SGEMM / CPU matmul to ~1e-6 relative; selections of VNMaxPool are compared in this mode.
//
// Three operand arrangements, one kernel template  C[M,N] (+)= A(M,K) * B(K,N):
//   rows   : Y = X W^T   A = X  (K contiguous)   B(k,n) = W[n*ldw + k]  (K contiguous)      forward / dgrad(W^T given)
//   rows_t : Y = X W     A = X  (K contiguous)   B(k,n) = W[k*ldw + n]  (N contiguous)      dgrad without a transpose
//   wgrad  : G = dY^T X  A(m,k) = dY[k*ld + m]   B(k,n) = X[k*ld + n]   (both MN contiguous), reduction over rows,
//            split across gridDim.z with fp32 atomicAdd into a zeroed / accumulated G.
#include <cuda_runtime.h>
#include <stdint.h>

#include "vnpcc_internal.h"

namespace vnpcc {

constexpr int GS_BM = 128, GS_BN = 128, GS_BK = 8, GS_T = 256;

// load 4 consecutive elements starting at p (element i valid iff i < nvalid), vectorised when aligned
__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ p, int nvalid, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid >= 4 && vec_ok) return __ldg(reinterpret_cast<const float4*>(p));
    if (nvalid > 0) v.x = __ldg(p);
    if (nvalid > 1) v.y = __ldg(p + 1);
    if (nvalid > 2) v.z = __ldg(p + 2);
    if (nvalid > 3) v.w = __ldg(p + 3);
    return v;
}

// A_KC: A element (m,k) at A[m*lda + k] (K contiguous) else at A[k*lda + m] (M contiguous); same for B with n.
// bias (optional): bias[((m / rows_per_sample) * 3 + m % 3) * ldb + n]  -- the per-sample broadcast term.
template <bool A_KC, bool B_KC, bool ATOMIC>
__global__ void __launch_bounds__(GS_T) sgemm_kernel(const float* __restrict__ A, size_t lda, const float* __restrict__ B,
                                                      size_t ldb_, float* __restrict__ C, size_t ldc, long long M, int N,
                                                      long long K, const float* __restrict__ bias, size_t ldbias,
                                                      long long rows_per_sample, long long k_chunk, int accumulate) {
    __shared__ float As[2][GS_BK][GS_BM + 4];
    __shared__ float Bs[2][GS_BK][GS_BN + 4];
    const int t = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * GS_BM;
    const int n0 = blockIdx.y * GS_BN;
    const long long kb = (long long)blockIdx.z * k_chunk;
    const long long ke = (kb + k_chunk < K) ? kb + k_chunk : K;
    if (kb >= ke && ATOMIC) return;

    const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((ldb_ & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // loader coordinates
    // K-contiguous operand: thread -> (row = t/2, 4 k's at (t%2)*4) ; MN-contiguous: (k = t/32, 4 cols at (t%32)*4)
    const int kc_row = t >> 1, kc_k = (t & 1) * 4;
    const int mc_k = t >> 5, mc_col = (t & 31) * 4;

    auto load_a = [&](long long k0) -> float4 {
        if (A_KC) {
            const long long m = m0 + kc_row;
            const long long k = k0 + kc_k;
            if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
            long long nv = ke - k;
            return ld4_guard(A + (size_t)m * lda + k, (int)(nv > 4 ? 4 : (nv < 0 ? 0 : nv)), a_vec && ((k & 3) == 0));
        } else {
            const long long k = k0 + mc_k;
            const long long m = m0 + mc_col;
            if (k >= ke) return make_float4(0.f, 0.f, 0.f, 0.f);
            long long nv = M - m;
            return ld4_guard(A + (size_t)k * lda + m, (int)(nv > 4 ? 4 : (nv < 0 ? 0 : nv)), a_vec && ((m & 3) == 0));
        }
    };
    auto load_b = [&](long long k0) -> float4 {
        if (B_KC) {
            const int n = n0 + kc_row;
            const long long k = k0 + kc_k;
            if (n >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
            long long nv = ke - k;
            return ld4_guard(B + (size_t)n * ldb_ + k, (int)(nv > 4 ? 4 : (nv < 0 ? 0 : nv)), b_vec && ((k & 3) == 0));
        } else {
            const long long k = k0 + mc_k;
            const int n = n0 + mc_col;
            if (k >= ke) return make_float4(0.f, 0.f, 0.f, 0.f);
            int nv = N - n;
            return ld4_guard(B + (size_t)k * ldb_ + n, nv > 4 ? 4 : (nv < 0 ? 0 : nv), b_vec && ((n & 3) == 0));
        }
    };
    auto store_a = [&](int buf, const float4& v) {
        if (A_KC) {
            As[buf][kc_k + 0][kc_row] = v.x;
            As[buf][kc_k + 1][kc_row] = v.y;
            As[buf][kc_k + 2][kc_row] = v.z;
            As[buf][kc_k + 3][kc_row] = v.w;
        } else {
            *reinterpret_cast<float4*>(&As[buf][mc_k][mc_col]) = v;
        }
    };
    auto store_b = [&](int buf, const float4& v) {
        if (B_KC) {
            Bs[buf][kc_k + 0][kc_row] = v.x;
            Bs[buf][kc_k + 1][kc_row] = v.y;
            Bs[buf][kc_k + 2][kc_row] = v.z;
            Bs[buf][kc_k + 3][kc_row] = v.w;
        } else {
            *reinterpret_cast<float4*>(&Bs[buf][mc_k][mc_col]) = v;
        }
    };

    const int ty = t >> 4, tx = t & 15;   // thread owns rows {ty*4..+3, 64+ty*4..+3} x cols {tx*4..+3, 64+tx*4..+3}

    float4 ra = load_a(kb), rb = load_b(kb);
    store_a(0, ra);
    store_b(0, rb);
    __syncthreads();
    int buf = 0;
    for (long long k0 = kb; k0 < ke; k0 += GS_BK) {
        const bool more = k0 + GS_BK < ke;
        if (more) {
            ra = load_a(k0 + GS_BK);
            rb = load_b(k0 + GS_BK);
        }
#pragma unroll
        for (int kk = 0; kk < GS_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            store_a(buf ^ 1, ra);
            store_b(buf ^ 1, rb);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        const float* brow = nullptr;
        if (bias) brow = bias + (size_t)((m / rows_per_sample) * 3 + (m % 3)) * ldbias;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j];
            if (brow) v += __ldg(brow + n);
            float* dst = C + (size_t)m * ldc + n;
            if (ATOMIC) atomicAdd(dst, v);
            else *dst = accumulate ? (*dst + v) : v;
        }
    }
}

__global__ void zero_rows_kernel(float* __restrict__ p, size_t ld, long long rows, int cols) {
    const long long total = rows * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols;
        p[(size_t)r * ld + (t - r * cols)] = 0.f;
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, size_t ldi, float* __restrict__ out, size_t ldo, int rows,
                                 int cols) {
    __shared__ float tile[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = blockIdx.y * 32 + i;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ldi + c] : 0.f;
    }
    __syncthreads();
    const int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c2 = blockIdx.x * 32 + i;
        if (r2 < rows && c2 < cols) out[(size_t)c2 * ldo + r2] = tile[threadIdx.x][i];
    }
}

}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

// Y[r, o] (+)= sum_k X[r, k] * Wop[o, k]  (+ bias[(r / rows_per_sample)*3 + r%3, o])
//   trans_w == 0:  W is [Cout, K] row-major (nn.Linear weight layout)    trans_w == 1:  W is [K, Cout] row-major
int vnpcc_gemm_rows_fp32(const float* X, long long ldx, const float* W, long long ldw, int trans_w, float* Y,
                         long long ldy, long long R, int K, int Cout, const float* bias, long long ldbias,
                         long long rows_per_sample, int accumulate, void* stream) {
    if (R <= 0 || Cout <= 0) return 0;
    if (K < 0 || (bias && rows_per_sample <= 0)) return VNPCC_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((R + GS_BM - 1) / GS_BM), (unsigned)((Cout + GS_BN - 1) / GS_BN), 1);
    if (trans_w)
        count_launch(), sgemm_kernel<true, false, false><<<grid, GS_T, 0, st>>>(X, (size_t)ldx, W, (size_t)ldw, Y, (size_t)ldy, R, Cout, K, bias,
                                                               (size_t)ldbias, rows_per_sample, K > 0 ? K : 1, accumulate);
    else
        count_launch(), sgemm_kernel<true, true, false><<<grid, GS_T, 0, st>>>(X, (size_t)ldx, W, (size_t)ldw, Y, (size_t)ldy, R, Cout, K, bias,
                                                              (size_t)ldbias, rows_per_sample, K > 0 ? K : 1, accumulate);
    return last_error();
}

// G[o, k] (+)= sum_r dY[r, o] * X[r, k]     (weight gradient; reduction over the R rows, split over CTAs)
int vnpcc_gemm_wgrad_fp32(const float* dY, long long lddy, const float* X, long long ldx, float* G, long long ldg,
                          long long R, int Cout, int K, int accumulate, void* stream) {
    if (Cout <= 0 || K <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) count_launch(), zero_rows_kernel<<<grid_for((size_t)Cout * K, 256, 4), 256, 0, st>>>(G, (size_t)ldg, Cout, K);
    if (R <= 0) return last_error();
    const int gm = (Cout + GS_BM - 1) / GS_BM, gn = (K + GS_BN - 1) / GS_BN;
    long long splits = ((long long)sm_count() * 2 + (long long)gm * gn - 1) / ((long long)gm * gn);
    long long max_splits = (R + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long chunk = ((R + splits - 1) / splits + GS_BK - 1) / GS_BK * GS_BK;
    splits = (R + chunk - 1) / chunk;
    dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
    count_launch(), sgemm_kernel<false, false, true><<<grid, GS_T, 0, st>>>(dY, (size_t)lddy, X, (size_t)ldx, G, (size_t)ldg, Cout, K, R, nullptr,
                                                          0, 1, chunk, 1);
    return last_error();
}

// out[c, r] = in[r, c]
int vnpcc_transpose(const float* in, long long ldi, float* out, long long ldo, int rows, int cols, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    count_launch(), transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, (size_t)ldi, out, (size_t)ldo, rows, cols);
    return last_error();
}

}  // extern "C"
