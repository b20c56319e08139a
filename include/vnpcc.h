/*
 * vnpcc.h -- C ABI of libvnpcc.so: the B200 (sm_100a) kernels behind the Vector-Neuron point-cloud-completion hot
 * path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host; `stream`
 * is a cudaStream_t passed as void*.  Every function only enqueues work (no synchronisation) and returns 0 on
 * success, a cudaError_t value if a launch failed, or one of the VNPCC_ERR_* codes below.  Callers must raise on a
 * non-zero return (the reference prints and ignores errors: extensions/chamfer_distance/chamfer3D.cu:145-151).
 *
 * Row layout ("channels-last", the physical layout the reference's nn.Linear calls produce): a logical
 * VN tensor [B, C, 3, N] is a row-major matrix X[R, C], R = 3*B*N, row r = (b*N + n)*3 + v, leading dimension `ld`
 * (floats).  "P" below is the number of 3-vectors per channel (P = R / 3).
 *
 * What each entry point replaces in the reference (paths under /root/reference):
 *   vnpcc_chamfer_forward / _backward   extensions/chamfer_distance/chamfer_cuda.cpp:17-27 (pybind `forward` /
 *                                       `backward`), kernels chamfer3D.cu:12-134 and :155-174
 *   vnpcc_cd_reduce / _bwd              metrics/loss.py:20-43, metrics/metric.py:12-23 (sqrt / mean tails)
 *   vnpcc_gemm_*                        nn.Linear(bias=False) inside VNLinear & friends: models/vn_layers.py:21,38,65,69,162,194
 *   vnpcc_vn_norm_stats / bn_finalize / vn_bn_leaky_*   VNBatchNorm models/vn_layers.py:116-127 + the leaky projection
 *                                       models/vn_layers.py:39-42,70-73 (forward) and their autograd
 *   vnpcc_vn_maxpool_*                  VNMaxPool models/vn_layers.py:158-167
 *   vnpcc_rows_*                        cat/expand of a broadcast global feature models/pcn.py:172,383-385 (folded into a
 *                                       per-sample bias), VNLinear(256,1) + residual models/pcn.py:345,387
 */
#ifndef VNPCC_H_
#define VNPCC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNPCC_OK 0
#define VNPCC_ERR_WORKSPACE 10001
#define VNPCC_ERR_BAD_ARG 10002
#define VNPCC_ERR_UNSUPPORTED 10003
#define VNPCC_ERR_DRIVER 10004

#define VNPCC_ABI_VERSION 1
int vnpcc_abi_version(void);
/* number of kernel launches enqueued by this library in this process (bench.py's gpu_launches) */
unsigned long long vnpcc_launch_count(void);
/* forward elementwise kernels: 0 (default) = IEEE sqrt / division with the reference's op-by-op rounding (parity mode),
 * 1 = MUFU reciprocal / rsqrt (throughput mode; the host layer switches it together with the TF32 GEMMs) */
void vnpcc_set_fast_math(int on);
/* ---------------------------------------------------------------- Chamfer ---------------------------------------- */
size_t vnpcc_chamfer_workspace_bytes(int B, int N, int M);
/* xyz1 [B,N,3], xyz2 [B,M,3] contiguous fp32 -> dist1 [B,N], dist2 [B,M] (squared), idx1 [B,N], idx2 [B,M] int32.
 * Bit-identical to the reference kernel for finite inputs; lowest index wins exact ties.  N==0 or M==0: outputs untouched. */
int vnpcc_chamfer_forward(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2,
                          int* idx1, int* idx2, void* workspace, size_t workspace_bytes, void* stream);
/* gradxyz1 [B,N,3] / gradxyz2 [B,M,3] are fully overwritten (no pre-zeroing); either may be NULL. */
int vnpcc_chamfer_backward(const float* xyz1, const float* xyz2, int B, int N, int M, const float* graddist1,
                           const float* graddist2, const int* idx1, const int* idx2, float* gradxyz1, float* gradxyz2,
                           void* stream);
/* reductions of the CD entry points.  mode 0: cd_loss_L1 (mean sqrt, /2)  1: cd_loss_L2 (mean)  2: l1_cd  3: l2_cd
 * (per-sample means summed over the batch).  out: 1 float, written (not accumulated).  scratch: 2 doubles. */
int vnpcc_cd_reduce(const float* dist1, const float* dist2, int B, int N, int M, int mode, double* scratch, float* out,
                    void* stream);
/* gradient of the CD entry point w.r.t. dist1/dist2 for upstream scalar *gout (device) : writes graddist1/2. */
int vnpcc_cd_reduce_bwd(const float* dist1, const float* dist2, int B, int N, int M, int mode, const float* gout,
                        float* graddist1, float* graddist2, void* stream);

/* Chamfer-based loss variants (utils/loss.py:14-74, extensions/ChamferDistancePytorch/fscore.py:3-16) on the search's outputs.
 * _cd_persample: out [B,6] = per sample (cd_p = (mean sqrt d1 + mean sqrt d2)/2, cd_t = mean d1 + mean d2, mean sqrt d1, mean sqrt d2,
 * mean d1, mean d2); part: workspace of 4*B floats; _bwd: gradient of any function of those six w.r.t. dist1 / dist2.
 * _fscore_sq: out [B,3] = (f, precision_1, precision_2), precision_k = mean(dist_k < threshold) on SQUARED distances.
 * _nn_counts: counts [B,K] (zeroed here) = histogram of idx [B,N] (the bincount of calc_dcd).
 * _dcd: density-aware CD, loss [B]; dist1/idx1 [B,N] index the M points of the other cloud (count1 [B,M]), dist2/idx2 [B,M] the N
 * points (count2 [B,N]); part: workspace of 2*B floats; counts carry no gradient. */
int vnpcc_cd_persample_fwd(const float* dist1, const float* dist2, int B, int N, int M, float* part, float* out, void* stream);
int vnpcc_cd_persample_bwd(const float* dist1, const float* dist2, int B, int N, int M, const float* gout, float* gdist1,
                           float* gdist2, void* stream);
int vnpcc_fscore_sq(const float* dist1, const float* dist2, int B, int N, int M, float threshold, float* out, void* stream);
int vnpcc_nn_counts(const int* idx, int B, int N, int K, int* counts, void* stream);
int vnpcc_dcd_fwd(const float* dist1, const float* dist2, const int* idx1, const int* idx2, const int* count1, const int* count2,
                  int B, int N, int M, float alpha, float n_lambda, float frac_21, float frac_12, float* part, float* loss,
                  void* stream);
int vnpcc_dcd_bwd(const float* dist1, const float* dist2, const int* idx1, const int* idx2, const int* count1, const int* count2,
                  int B, int N, int M, float alpha, float n_lambda, float frac_21, float frac_12, const float* gloss,
                  float* gdist1, float* gdist2, void* stream);

/* ---------------------------------------------------------------- GEMMs ------------------------------------------ */
/* Y[r,o] (+)= sum_k X[r,k] * Wop[o,k] (+ bias[(r / rows_per_sample)*3 + r%3, o]);  trans_w: 0 -> W [Cout,K], 1 -> W [K,Cout] */
int vnpcc_gemm_rows_fp32(const float* X, long long ldx, const float* W, long long ldw, int trans_w, float* Y,
                         long long ldy, long long R, int K, int Cout, const float* bias, long long ldbias,
                         long long rows_per_sample, int accumulate, void* stream);
/* G[o,k] (+)= sum_r dY[r,o] * X[r,k] */
int vnpcc_gemm_wgrad_fp32(const float* dY, long long lddy, const float* X, long long ldx, float* G, long long ldg,
                          long long R, int Cout, int K, int accumulate, void* stream);
int vnpcc_transpose(const float* in, long long ldi, float* out, long long ldo, int rows, int cols, void* stream);

/* tensor-core (tcgen05 / TMEM / TMA) versions, TF32 operands, fp32 accumulation.  Same contracts as the _fp32 ones;
 * return VNPCC_ERR_UNSUPPORTED for shapes/alignments they do not take (callers then use the _fp32 entry points). */
int vnpcc_gemm_rows_tf32(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy,
                         long long R, int K, int Cout, const float* bias, long long ldbias, long long rows_per_sample,
                         void* stream);
/* vnpcc_gemm_rows_tf32 for a training-mode VNLinearLeakyReLU (models/vn_layers.py:60-74,116-127): the epilogue also accumulates the
 * BatchNorm-on-norm batch statistics of the first Cstat output channels -- sums[c] = sum over points of (||Y[point, c]|| + 1e-6),
 * sums[Cstat + c] = the same squared, fp64, zeroed here -- so no separate pass re-reads Y.  R % 3 == 0, Cstat % 32 == 0,
 * Cstat <= min(Cout, 1024); VNPCC_ERR_UNSUPPORTED otherwise (callers then use vnpcc_gemm_rows_* + vnpcc_vn_norm_stats). */
int vnpcc_gemm_rows_tf32_stats(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy,
                               long long R, int K, int Cout, const float* bias, long long ldbias, long long rows_per_sample,
                               double* sums, int Cstat, void* stream);
int vnpcc_gemm_wgrad_tf32(const float* dY, long long lddy, const float* X, long long ldx, float* G, long long ldg,
                          long long R, int Cout, int K, float* workspace, size_t workspace_bytes, void* stream);
size_t vnpcc_gemm_wgrad_tf32_workspace_bytes(long long R, int Cout, int K);
/* fp32-accurate GEMMs on the TF32 tensor cores ("3xTF32"): x = hi + lo, hi = tf32(x), lo = tf32(x - hi); the three products hi.hi +
 * lo.hi + hi.lo are ONE vnpcc_gemm_rows_tf32 / vnpcc_gemm_wgrad_tf32 over operands whose contraction axis is tripled.  This writes the
 * tripled operand: layout 0 = columns [hi | lo | hi] (out [R, 3K]), 1 = columns [hi | hi | lo], 2 = rows [hi ; lo ; hi] (out [3R, K]),
 * 3 = rows [hi ; hi ; lo].  Pair an activation split 0 (2) with a weight split 1 (3).  Relative error of a product ~2^-21. */
int vnpcc_split_tf32(const float* x, long long ldx, long long R, int K, float* out, long long ldo, int layout, void* stream);
/* VNLinearLeakyReLU (models/vn_layers.py:60-74) with BatchNorm-on-norm + leaky projection fused into the tcgen05 GEMM epilogue
 * (no-grad / inference forward: the linear outputs p, d never reach HBM).  Wcat [2C, K] = (W_feat ; W_dir); C % 128 == 0.
 * _stats: per-channel (sum ||p||, sum ||p||^2) in fp64 for training-mode statistics; _apply: out [R, C]. */
int vnpcc_gemm_vn_stats(const float* X, long long ldx, const float* Wcat, long long ldw, long long R, int K, int C, const float* bias,
                        long long ldbias, long long rows_per_sample, double* sums, void* stream);
int vnpcc_gemm_vn_apply(const float* X, long long ldx, const float* Wcat, long long ldw, float* out, long long ldo, long long R, int K,
                        int C, const float* bias, long long ldbias, long long rows_per_sample, const float* stat, const float* gamma,
                        const float* beta, float ns, void* stream);

/* ---------------------------------------------------------------- VN elementwise / reductions ------------------- */
int vnpcc_vn_norm_stats(const float* p, long long ldp, long long P, int C, double* sums, void* stream);
int vnpcc_bn_finalize(const double* sums, double count, int C, int training, float* running_mean, float* running_var,
                      float momentum, float bn_eps, float* stat, void* stream);
int vnpcc_vn_bn_leaky_fwd(const float* p, long long ldp, const float* d, long long ldd, float* out, long long ldo,
                          long long P, int C, const float* stat, const float* gamma, const float* beta, float ns,
                          void* stream);
int vnpcc_vn_bn_leaky_bwd1(const float* g, long long ldg, const float* p, long long ldp, const float* d, long long ldd,
                           float* gp, long long ldgp, float* gd, long long ldgd, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, float ns, double* sums, void* stream);
int vnpcc_vn_bn_bwd2(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat,
                     const float* gamma, const float* beta, const double* sums, double count, int training,
                     float* gweight, float* gbias, void* stream);
/* vnpcc_vn_bn_bwd2 fused with the per-sample bias gradient of the producing GEMM (models/pcn.py:172: the broadcast half of the
 * concatenation): gbias [B*3, 2C] (zeroed here) = per-sample column sums of the final gp | of gd ([P*3, C], pitch ldgd).
 * VNPCC_ERR_UNSUPPORTED for shapes the vectorised kernel does not take. */
int vnpcc_vn_bn_bwd2_sbias(float* gp, long long ldgp, const float* p, long long ldp, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, const double* sums, double count, int training, float* gweight,
                           float* gbn_bias, const float* gd, long long ldgd, float* gbias, long long ldgb, long long pts_per_sample,
                           void* stream);
/* VNMaxPool over groups of N consecutive points: ws = B*C u64, idx int64 [B,C] */
int vnpcc_vn_maxpool_argmax(const float* x, long long ldx, const float* d, long long ldd, int B, int N, int C,
                            unsigned long long* ws, long long* idx, void* stream);
int vnpcc_vn_maxpool_gather(const float* x, long long ldx, const long long* idx, int B, int N, int C, float* out,
                            long long ldo, void* stream);
int vnpcc_vn_maxpool_scatter_add(const float* g, long long ldg, const long long* idx, int B, int N, int C, float* gx,
                                 long long ldgx, void* stream);
int vnpcc_rows_add_sample_bias(float* y, long long ldy, const float* bias, long long ldb, int B, int N, int C,
                               void* stream);
int vnpcc_rows_sample_sum(const float* g, long long ldg, int B, int N, int C, float* out, long long ldo, void* stream);
int vnpcc_rows_dot(const float* x, long long ldx, const float* w, long long R, int C, const float* res, float* y,
                   void* stream);
int vnpcc_rows_dot_bwd(const float* gy, const float* x, long long ldx, const float* w, long long R, int C, float* gx,
                       long long ldgx, float* gw, void* stream);

/* VNStdFeature's invariant-feature step (models/vn_layers.py:197-219): z rows (point, v) x J are the J = 3 (or 2: normalize_frame, Gram-
 * Schmidt + cross product, eps 1e-6) frame vectors from vn_lin; out rows (point, k) x C = <x[point, c, :], f_k>; zout [P*3, 3] = the frame,
 * rows (point, k) x component.  _bwd: gx rows (point, v) x C and gz rows (point, v) x J from gout and (optionally) gzout. */
int vnpcc_vn_frame_fwd(const float* x, long long ldx, const float* z, long long ldz, long long P, int C, int J, float* out,
                       long long ldo, float* zout, void* stream);
int vnpcc_vn_frame_bwd(const float* gout, long long ldgo, const float* gzout, const float* x, long long ldx, const float* z,
                       long long ldz, long long P, int C, int J, float* gx, long long ldgx, float* gz, long long ldgz, void* stream);

/* fused tail VNLinearLeakyReLU -> VNLinear(C,1) (+ residual), models/pcn.py:340-345,387: no [R,C] activation / gradient in HBM */
int vnpcc_bn_leaky_dot_fwd(const float* p, long long ldp, const float* d, long long ldd, long long P, int C, const float* stat,
                           const float* gamma, const float* beta, float ns, const float* w2, const float* res, float* y,
                           void* stream);
int vnpcc_bn_leaky_dot_bwd1(const float* gy, const float* p, long long ldp, const float* d, long long ldd, float* gp, long long ldgp,
                            float* gd, long long ldgd, long long P, int C, const float* stat, const float* gamma, const float* beta,
                            float ns, double* sums, const float* w2, double* gw2, void* stream);
int vnpcc_double_to_float(const double* in, float* out, int n, void* stream);
/* the whole backward of that tail, TF32 mode (csrc/gemm_tcgen05.cu): a sums-only pre-pass (BatchNorm backward sums -> sums [2C], tail
 * weight gradient -> gw2 [C]; fp64, zeroed here), then tail_dgrad_tf32_kernel: producer warps form the FINAL gradient gpd of the stacked
 * linear output (p | d) from (pd, gy) in shared memory and feed it to the dgrad MMA: gh [P*3, Cin] = gpd Wcat (Wt [Cin, 2C] = Wcat^T).
 * The weight gradient of the stacked weight: either gW != NULL (needs h [P*3, Cin] and C % 128 == 0): tail_wgrad_tf32_kernel forms gpd
 * on the fly again and accumulates gW [2C, Cin] (zeroed here) -- gpd never exists in HBM, pass gpd = NULL; or gW == NULL: gpd [P*3, 2C]
 * is written once for vnpcc_gemm_wgrad_tf32.  C % 32 == 0, C <= 256, Cin in {128, 256}; VNPCC_ERR_UNSUPPORTED otherwise (callers then
 * run vnpcc_bn_leaky_dot_bwd1 + vnpcc_vn_bn_bwd2 + vnpcc_gemm_rows_* + vnpcc_gemm_wgrad_*). */
int vnpcc_tail_bwd_tf32(const float* gy, const float* pd, long long ldpd, long long P, int C, const float* stat, const float* gamma,
                        const float* beta, float ns, const float* w2, const float* Wt, long long ldwt, int Cin, int training,
                        double* sums, double* gw2, float* gpd, long long ldgpd, float* gh, long long ldgh, const float* h,
                        long long ldh, float* gW, long long ldgw, void* stream);
/* VNLinearLeakyReLU with <= 4 local input channels + per-sample bias, never materialising p / d (decoder final_conv[0]):
 * x [B*N*3, K], w [2C, K] (feat | dir), bias [B*3, 2C] or NULL.  See csrc/vn_fused.cu. */
int vnpcc_fold_stats(const float* x, long long ldx, const float* w, long long ldw, const float* bias, long long ldb, int B, int N,
                     int K, int C, double* sums, void* stream);
int vnpcc_fold_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, long long ldb, int B, int N,
                   int K, int C, const float* stat, const float* gamma, const float* beta, float ns, float* out, long long ldo,
                   void* stream);
int vnpcc_fold_bwd(const float* g, long long ldg, const float* x, long long ldx, const float* w, long long ldw, const float* bias,
                   long long ldb, int B, int N, int K, int C, const float* stat, const float* gamma, const float* beta, float ns,
                   int training, double* sums, float* gx, long long ldgx, int gx_first_col, float* gw, long long ldgw, float* gbias,
                   long long ldgb, float* ggamma, float* gbeta, void* stream);
/* fused VNLinear -> VNMaxPool forward (TF32): arg-max inside the tcgen05 GEMM epilogue, pooled rows recomputed from the
 * selected inputs; the [R, C] layer output and its direction are never stored */
int vnpcc_gemm_vn_pool(const float* X, long long ldx, const float* Wcat, long long ldw, long long R, int K, int C, long long N,
                       unsigned long long* best, void* stream);
int vnpcc_vn_maxpool_decode(const unsigned long long* best, long long total, long long* idx, void* stream);
int vnpcc_pool_linear_gather(const float* x, long long ldx, const float* W, long long ldw, const long long* idx, int B, int N, int C, int K,
                             float* out, long long ldo, void* stream);
/* backward of VNLinear -> VNMaxPool without the dense gradient (gx zeroed + scattered, gW gathered; either may be NULL) */
int vnpcc_pool_linear_bwd(const float* g, long long ldg, const long long* idx, const float* x, long long ldx, const float* W,
                          long long ldw, int B, int N, int C, int K, float* gx, long long ldgx, float* gW, long long ldgw,
                          void* stream);
/* small-K VNLinear (1 <= K <= 4 input channels): HBM-bound streaming kernels (first_conv[0], the local channels of final_conv[0]) */
int vnpcc_smallk_fwd(const float* x, long long ldx, const float* W, long long ldw, const float* bias, long long ldbias,
                     long long rows_per_sample, float* y, long long ldy, long long R, int K, int Cout, void* stream);
int vnpcc_smallk_dgrad(const float* gy, long long ldgy, const float* W, long long ldw, float* gx, long long ldgx, long long R,
                       int K, int Cout, void* stream);
int vnpcc_smallk_wgrad(const float* gy, long long ldgy, const float* x, long long ldx, int B, int N, int K, int Cout, float* gW,
                       long long ldgw, float* gbias, long long ldgb, void* stream);

/* ---------------------------------------------------------------- point-set graph ops (VN_DGCNN_fps) -- */
/* k nearest neighbours of every query among the reference points of the same sample, 3-D (replaces knn_cuda.KNN(k,
 * transpose_mode=False) at models/dgcnn.py:11,236,257-259).  ref [B,Nr,3], query [B,Nq,3] contiguous fp32 ->
 * idx [B,k,Nq] int64 (knn_cuda's layout), dist [B,k,Nq] Euclidean distances (may be NULL).  Ordered by (distance, index);
 * distance = fma(dz,dz,fma(dy,dy,dx*dx)) of fp32 differences.  1 <= k <= min(32, Nr). */
int vnpcc_knn3d(const float* ref, const float* query, int B, int Nr, int Nq, int k, long long* idx, float* dist, void* stream);
/* furthest point sampling (replaces pointnet2_utils.furthest_point_sample, models/dgcnn.py:15,210): xyz [B,N,3] ->
 * idx [B,M] int32; starts at point 0, points with |p|^2 <= 1e-3 never compete, lowest index wins exact ties.  N <= 16384. */
int vnpcc_fps(const float* xyz, int B, int N, int M, int* idx, void* stream);
/* pointnet2_utils.gather_operation on the row layout: rows (b,n,v) x C -> rows (b,m,v) x C with n = idx[b,m]; and its adjoint
 * (gx is zeroed, then accumulated) */
int vnpcc_points_gather(const float* x, long long ldx, const int* idx, int B, int N, int M, int C, float* out, long long ldo,
                        void* stream);
int vnpcc_points_scatter_add(const float* g, long long ldg, const int* idx, int B, int N, int M, int C, float* gx, long long ldgx,
                             void* stream);
/* VN_DGCNN_fps.vn_get_graph_feature (models/dgcnn.py:251-278): x rows (b,n,v) x C, idx [B,k,N] -> rows ((b,n,j),v) x 2C =
 * (x_j - x_i | x_i); adjoint: gx zeroed, then accumulated with fp32 atomics */
int vnpcc_edge_feature_fwd(const float* x, long long ldx, const long long* idx, int B, int N, int k, int C, float* out, long long ldo,
                           void* stream);
int vnpcc_edge_feature_bwd(const float* g, long long ldg, const long long* idx, int B, int N, int k, int C, float* gx, long long ldgx,
                           void* stream);
/* mean_pool over the k neighbours (models/vn_layers.py:170-171): rows ((g,j),v) x C -> rows (g,v) x C; adjoint */
int vnpcc_rows_group_mean(const float* x, long long ldx, long long G, int k, int C, float* out, long long ldo, void* stream);
int vnpcc_rows_group_mean_bwd(const float* g, long long ldg, long long G, int k, int C, float* gx, long long ldgx, void* stream);

/* edge convolution without the edge tensor (csrc/edge_conv.cu): uw [B*N*3, 4C] = (U_p | U_d | W_p | W_d) with U = W1 x, W = (W2 - W1) x of
 * the point GEMM; edge (i, j = idx[b, j, i]): p = U_p[j] + W_p[i], d = U_d[j] + W_d[i]; out[i] = mean_j leaky(BN(p), d).  BatchNorm
 * statistics over the B*N*k edges.  C % 4 == 0, 256 % (C/4) == 0, C <= 1024.  _bwd writes guw [B*N*3, 4C] completely (U half zeroed +
 * red.add, W half stored) and ggamma / gbeta [C]; sums: 2C doubles of workspace. */
int vnpcc_edge_conv_stats(const float* uw, long long ld, const long long* idx, int B, int N, int k, int C, double* sums, void* stream);
int vnpcc_edge_conv_fwd(const float* uw, long long ld, const long long* idx, int B, int N, int k, int C, const float* stat, const float* gamma,
                        const float* beta, float ns, float* out, long long ldo, void* stream);
int vnpcc_edge_conv_bwd(const float* g, long long ldg, const float* uw, long long ld, const long long* idx, int B, int N, int k, int C,
                        const float* stat, const float* gamma, const float* beta, float ns, int training, double* sums, float* guw,
                        long long ldgu, float* ggamma, float* gbeta, void* stream);

/* ---------------------------------------------------------------- transformer-refined decoder (Attention_VN_FoldingNet) -- */
/* VNLayerNorm (models/vn_layers.py:129-150): per token (3 rows) norm[c] = ||x[c,:]|| + 1e-6, nn.LayerNorm over the C channels
 * (weight, bias, eps = ln_eps), y = x / norm * ln(norm).  P tokens, C <= 512.  stats [P,2] = (mean, rstd) for the backward
 * (may be NULL in no-grad forwards).  _bwd zeroes gweight / gbias [C] and accumulates them with fp32 atomics. */
int vnpcc_vn_layernorm_fwd(const float* x, long long ldx, long long P, int C, const float* weight, const float* bias, float ln_eps, float* y,
                           long long ldy, float* stats, void* stream);
int vnpcc_vn_layernorm_bwd(const float* g, long long ldg, const float* x, long long ldx, long long P, int C, const float* weight,
                           const float* bias, const float* stats, float* gx, long long ldgx, float* gweight, float* gbias, void* stream);
/* out = a + b on rows (residual additions of VN_Block, models/transformer.py:60,68) */
int vnpcc_rows_add(const float* a, long long lda, const float* b, long long ldb, float* out, long long ldo, long long R, int C, void* stream);
/* VN multi-head attention core (models/transformer.py:89-100): qkv [B*N*3, 3C] = (q | k | v) rows, C = H*D channels, head h =
 * channels [h*D, (h+1)*D) of the 3 rows of a token (a 3D-dim feature); out[token] = softmax_m(scale * <q_n, k_m>) v_m per head, written
 * in the same row layout [B*N*3, C]; lse [B,H,N] = log-sum-exp of the scaled scores (saved for the backward).  Exact fp32,
 * flash-style (no [N,N] matrix in HBM).  D in {16, 32, 48}; ld % 4 == 0. */
int vnpcc_vn_attention_fwd(const float* qkv, long long ld, int B, int N, int H, int D, float scale, float* out, long long ldo, float* lse,
                           void* stream);
/* tensor-core twin of the forward (csrc/attention_tc.cu): tcgen05 / TMEM, TF32 operands, fp32 accumulation, two-pass softmax with the
 * output accumulator resident in TMEM.  Same arguments; D == 48 only, returns VNPCC_ERR_UNSUPPORTED otherwise. */
int vnpcc_vn_attention_fwd_tf32(const float* qkv, long long ld, int B, int N, int H, int D, float scale, float* out, long long ldo, float* lse,
                                float* p_out, void* stream);
/* p_out (optional, B*H*N*N floats, N % 32 == 0): the normalised attention weights, stored for the GEMM-shaped backward */
/* dqkv [B*N*3, 3C] fully written (q part zeroed then accumulated with fp32 atomics); delta: workspace of B*H*N floats */
int vnpcc_vn_attention_bwd(const float* qkv, long long ld, const float* dout, long long lddo, const float* out, long long ldo, const float* lse,
                           int B, int N, int H, int D, float scale, float* dqkv, long long lddq, float* delta, void* stream);

/* delta[b,h,n] = <dout, out> over head h's features (first step of both attention backwards); and the tensor-core twin of the backward
 * (csrc/attention_tc.cu): dV / dQ / dK by tcgen05 kernels with the accumulators resident in TMEM, dqkv fully written with plain stores
 * (no atomics).  D == 48 only, VNPCC_ERR_UNSUPPORTED otherwise. */
int vnpcc_vn_attention_delta(const float* dout, long long lddo, const float* out, long long ldo, int B, int N, int H, int D, float* delta,
                             void* stream);
int vnpcc_vn_attention_bwd_tf32(const float* qkv, long long ld, const float* dout, long long lddo, const float* out, long long ldo,
                                const float* lse, int B, int N, int H, int D, float scale, float* dqkv, long long lddq, float* delta,
                                float* ds_workspace, size_t ds_workspace_bytes, float* p_buf, void* stream);
/* p_buf (optional: the P the forward stored; DESTROYED, it is turned into dS in place): dV = P^T dO, dQ = dS K and dK = dS^T Q all run as
 * streaming tcgen05 GEMMs and only dP = dO V^T is recomputed.  Else ds_workspace (optional, B*H*N*N floats, N % 32 == 0): the dQ kernel
 * writes dS there and dK = dS^T Q runs as a streaming GEMM; without either dK recomputes S and dP in the key orientation. */

/* ---------------------------------------------------------------- evaluation extras (test.py:73-78) ---------- */
/* metrics/metric.py:31-48 f_score from the Chamfer search's SQUARED distances: out [B,3] = (precision, recall, F) with
 * precision = #{sqrt(dist1) < th} / N, recall = #{sqrt(dist2) < th} / M */
int vnpcc_fscore(const float* dist1, const float* dist2, int B, int N, int M, float th, float* out, void* stream);
/* utils/voxel_util.py:89-105 points_to_voxels (pyntcloud VoxelGrid, regular bounding box of the cloud itself, size_grid^3 cells) as a bit
 * mask: bits [B, ceil(size_grid^3 / 32)] (zeroed here), voxel (x,y,z) -> bit (x*n + y)*n + z; and utils/voxel_util.py:5-14 iou -> out [B] */
int vnpcc_voxel_occupancy(const float* xyz, int B, int N, int size_grid, unsigned* bits, void* stream);
int vnpcc_voxel_iou(const unsigned* bits_a, const unsigned* bits_b, int B, int words, float* out, void* stream);

/* ---------------------------------------------------------------- optimiser / misc ------------------------------ */
/* fused Adam over a flat fp32 buffer (torch.optim.Adam semantics, train.py:70): p,g,m,v length n */
int vnpcc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, void* stream);
/* the same update with lr and the step counter in device memory (state = {lr, 1 - beta1^t, sqrt(1 - beta2^t), t}: the call increments t and
 * refreshes the two corrections on the device), so that a train step captured in a CUDA graph can be replayed (train.py:70, :165-173) */
int vnpcc_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* state, float beta1, float beta2, float eps,
                        float weight_decay, float grad_scale, void* stream);
#ifdef __cplusplus
}
#endif
#endif /* VNPCC_H_ */
