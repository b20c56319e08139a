/*
 * oracle/graph_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the two third-party point-set searches the reference's VN_DGCNN_fps encoder calls
 * (SURVEY.md 8f row f1).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 *
 * PARITY UNPINNED against the third-party binaries: neither package is vendored under /root/reference, and the
 * reference holds no test or golden vector for them.
 *   knn_cuda.KNN (wheel KNN_CUDA 0.2, README.md:31; call sites models/dgcnn.py:11,236,257-259): brute-force k nearest
 *     neighbours, neighbours returned in ascending distance.  Restated as an exact search ordered by
 *     (distance, index) with distance = fma(dz,dz, fma(dy,dy, dx*dx)) of fp32 differences (the published algorithm
 *     expands |r|^2 + |q|^2 - 2 r.q through cuBLAS, whose rounding is not reproducible; neighbour SETS agree except at
 *     near-ties of that rounding).
 *   pointnet2_ops.pointnet2_utils.furthest_point_sample (unpinned git master, README.md:29; call sites
 *     models/dgcnn.py:15,210): published algorithm (furthest_point_sampling_kernel): start at index 0, running minimum
 *     distance initialised to 1e10, points with |p|^2 <= 1e-3 are skipped, next = arg-max of the running minimum.
 *     Exact ties: lowest index (the published kernel's tie order depends on its block size).
 * Anchored instead on the reference's own call sites: tests/golden/make_golden.py runs the UNMODIFIED VN_DGCNN_fps
 * of the reference with these two searches plugged into its knn_cuda / pointnet2_ops imports.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

static inline float sqdist(const float* c, const float* q) {
    const float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* ref [B,Nr,3], query [B,Nq,3] -> idx [B,k,Nq] int64, dist [B,k,Nq] (Euclidean, may be NULL) */
int oracle_knn3d(int B, int Nr, int Nq, int k, const float* ref, const float* query, int64_t* idx, float* dist) {
    if (k <= 0 || k > Nr) return 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int q = 0; q < Nq; ++q) {
            float bd[64];
            int bi[64];
            int kk = k > 64 ? 64 : k;
            for (int t = 0; t < kk; ++t) {
                bd[t] = FLT_MAX;
                bi[t] = 0;
            }
            const float* qp = query + ((size_t)b * Nq + q) * 3;
            for (int c = 0; c < Nr; ++c) {
                const float d = sqdist(ref + ((size_t)b * Nr + c) * 3, qp);
                if (d < bd[kk - 1]) {
                    int s = kk - 1;
                    while (s > 0 && d < bd[s - 1]) {
                        bd[s] = bd[s - 1];
                        bi[s] = bi[s - 1];
                        --s;
                    }
                    bd[s] = d;
                    bi[s] = c;
                }
            }
            for (int t = 0; t < kk; ++t) {
                idx[((size_t)b * k + t) * Nq + q] = bi[t];
                if (dist) dist[((size_t)b * k + t) * Nq + q] = sqrtf(bd[t]);
            }
        }
    }
    return 0;
}

/* xyz [B,N,3] -> idx [B,M] int32 */
int oracle_fps(int B, int N, int M, const float* xyz, int32_t* idx) {
    if (N <= 0) return 1;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const float* p = xyz + (size_t)b * N * 3;
        float* temp = (float*)malloc(sizeof(float) * (size_t)N);
        for (int n = 0; n < N; ++n) temp[n] = 1e10f;
        int old = 0;
        if (M > 0) idx[(size_t)b * M] = 0;
        for (int j = 1; j < M; ++j) {
            int besti = 0;
            float best = -1.0f;
            for (int n = 0; n < N; ++n) {
                const float* c = p + (size_t)n * 3;
                const float mag = fmaf(c[2], c[2], fmaf(c[1], c[1], c[0] * c[0]));
                if ((double)mag <= 1e-3) continue;
                const float d = sqdist(c, p + (size_t)old * 3);
                const float d2 = d < temp[n] ? d : temp[n];
                temp[n] = d2;
                if (d2 > best) {
                    best = d2;
                    besti = n;
                }
            }
            old = besti;
            idx[(size_t)b * M + j] = old;
        }
        free(temp);
    }
    return 0;
}
