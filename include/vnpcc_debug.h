/*
 * vnpcc_debug.h -- development / test-only entry points of libvnpcc.so.  NOT part of the drop-in boundary (include/vnpcc.h): nothing
 * in the product path depends on them; tools/, tests/ and bench.py's roofline leg use them for A/B measurements, planner
 * introspection without a GPU and the FP32-pipe peak micro-benchmark.
 */
#ifndef VNPCC_DEBUG_H_
#define VNPCC_DEBUG_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* development knobs for A/B measurements (tools/*_bench.py, tools/step_ab.py); every knob defaults to 0 = the shipped behaviour.
 *   0  1 = legacy fixed grids instead of occupancy-sized single-wave grids; also disables split-K for the few-row GEMMs
 *   1  fused small-K backward: 1 = four channels per thread at 1 CTA/SM, 2 = four channels at 2 CTAs/SM, 3 = two channels at 2 CTAs/SM,
 *      0 / 4 = two channels, sums pass at 3 CTAs/SM (shipped), 5-7 = further occupancy variants (measured slower)
 *   2  rows GEMM tiling: 0 = CTA pairs (tcgen05 cta_group::2) where eligible, 1 = one SM per tile (4 stages), 2 = one SM, 3 stages,
 *      3 = one SM with the TMA-store statistics epilogue, 4 = CTA pairs also for the fused VN-epilogue kernels
 *   3  1 = statistics epilogue without its arithmetic (timing experiments only: wrong statistics)
 *   4  Chamfer planner: per-item overhead of the cost model in candidates (0 = 256)      5  search CTAs per SM the planner sizes for (0 = 8)
 *   6  1 = scalar-FFMA ranking loop in the pre-filtered search                            7  1 = never use the fused tail weight-gradient kernel, 3 = tail dgrad with one SM per tile instead of CTA pairs
 *   8  fused small-K forward: 1 = fp64 statistics pass, 3 / 4 = forward at 3 / 4 CTAs per SM (measured slower) */
void vnpcc_set_tuning(int knob, int value);

/* host-logic introspection: the launch planners (work-item splits, chunk lengths, grids) as pure functions of the problem size, so that
 * tests can check them without a GPU (tests/test_planners_cpu.py).  No device work; sm counts are arguments or default to 148. */
void vnpcc_debug_chamfer_plan(int B, int N, int M, int* out4);                       /* {query blocks, splits, split length, queries/block} */
void vnpcc_debug_fold_geometry(int B, int N, int C, int resident, int lanes, int* out6); /* {grid.x, grid.y, block.x, block.y, chunk, row mode} */
void vnpcc_debug_wgrad_plan(long long R, int Cout, int K, int sms, long long* out4);  /* {grid.x, grid.y, splits, rows per split} */
int vnpcc_debug_plan_chunk_len(long long groups, int N, long long slots, int lanes, int min_chunk);
/* rows GEMM form for a problem: {variant (0 one SM per tile, 1 CTA pairs, 2 split-K), K splits, rows per tile, grid, tiles} */
void vnpcc_debug_rows_plan(long long R, int K, int Cout, int has_bias, int stats, int sms, long long* out5);

/* Chamfer search variant: 0 = exact scalar search, 1 = exact packed-fp32 search, 2 (default, any other value) = pre-filtered search with
 * exact resolve.  All three give bit-identical results; tests/test_gpu_chamfer.py runs every case under each of them. */
void vnpcc_chamfer_set_packed_math(int mode);
/* how many queries of the most recent vnpcc_chamfer_forward (pre-filtered mode, same workspace and sizes) failed the pre-filter's margin
 * test and were re-searched exactly: out2_host = {pass xyz1->xyz2, pass xyz2->xyz1}.  Synchronises `stream`. */
int vnpcc_debug_chamfer_slow_counts(const void* workspace, int B, int N, int M, int* out2_host, void* stream);

/* FP32-pipe peak micro-benchmark used for the Chamfer roofline (mode 0: FFMA, 1: FFMA2, 2: Chamfer mix) */
int vnpcc_measure_fp32_peak(int mode, int iters, float* scratch_dev, float* ms_out_host, double* lane_ops_out_host,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VNPCC_DEBUG_H_ */
