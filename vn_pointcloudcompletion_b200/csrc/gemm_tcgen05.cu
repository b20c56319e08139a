// gemm_tcgen05.cu -- the VNLinear contraction on Blackwell tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) -> tcgen05.ld epilogue.  Hand-written PTX, no CUTLASS.
//
// Replaces the cuBLAS SGEMM behind nn.Linear(bias=False) in models/vn_layers.py:21,38,65,69,162,194.
//
// Orientation ("channels on lanes"): for  Y[r, o] = sum_k X[r, k] W[o, k]  the WEIGHT tile is the MMA A operand
// (M = 128 output channels = 128 TMEM lanes) and the activation tile is the B operand (N = up to 256 rows r).  Both
// are K-major in memory (channels-last rows / nn.Linear weight), so both load with plain 2-D TMA boxes of 32 floats
// (128 bytes, SWIZZLE_128B).  An epilogue thread therefore owns ONE output channel and walks over rows: for every
// row the 32 lanes of a warp store 32 consecutive channels = one 128-byte line of the channels-last output, the three
// components of a 3-vector sit in three consecutive TMEM columns of the same thread, and per-channel reductions
// (BatchNorm-on-norm statistics, arg-max pooling) are thread-local -- the layout the fused VN epilogues need.
//
// Pipeline per CTA (persistent over output tiles, m-tiles fastest so concurrently running CTAs share the activation
// tile through L2):  warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator, warps 4..7 =
// epilogue (TMEM lane quadrant = warp % 4).  Shared-memory ring of STAGES x {A 128x32, B BNx32} fp32 tiles with
// full/empty mbarriers; the 512 TMEM columns hold two BN-column accumulators so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// The weight-gradient kernel (G[o,k] = sum_r dY[r,o] X[r,k]) uses the same machinery with both operands MN-major
// (the reduction index r is the slow axis of both row matrices) and the reduction split over CTAs.
//
// Every mbarrier wait is bounded: a protocol error traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "vnpcc_internal.h"

namespace vnpcc {
bool fast_math_enabled();
namespace tc {

constexpr int BM = 128;            // output channels per tile (TMEM lanes)
constexpr int BK = 32;             // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;          // tf32: 32 bytes of K per instruction
constexpr int NUM_THREADS = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 22;   // x (try_wait latency ~0.1-1 us): a protocol error traps within seconds

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// approximate reciprocal / reciprocal square root as ONE MUFU instruction each (the plain intrinsics add denormal range-scaling code;
// every operand here is either >= 1e-6 or guarded by the caller)
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();   // protocol error: never hang the device
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// bulk tensor STORE shared -> global (clipped at the tensor bounds), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {      // at most N groups still READING their shared-memory source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- CTA pair (cta_group::2): two SMs of a cluster work on one 256-channel tile; rank 0 issues the MMAs and owns the "full" barriers
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs: the bytes land in the issuing CTA's shared memory, the transaction count on rank 0's barrier
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 32 lanes x 32 consecutive columns: thread t of the warp gets lane (quadrant*32 + t), v[j] = column (col0 + j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B)
// layout: 2 = SWIZZLE_128B (16-byte chunks XOR row%8, K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR
// row%4 -- the only swizzled layout tcgen05 accepts for MN-major 32-bit (tf32) operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a/b format TF32=2 [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct PipeState {
    int stage = 0;
    uint32_t phase = 0;
    template <int STAGES>
    __device__ __forceinline__ void advance() {
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// rows GEMM:  Y[r, o] = sum_k X[r, k] W[o, k] (+ bias[(r / rps)*3 + r%3, o])
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES, int GAP = 0>
struct RowsSmem {
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + GAP;
    static constexpr int STATS_OFFSET = BAR_OFFSET + 256;   // STATS kernels: 2 * MAX_STAT_C doubles (per-CTA partial sums)
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + alignment slack
    // STATS kernels: + output staging for the TMA-store epilogue, per epilogue warp 2 buffers of 48 rows x 32 channels
    static constexpr int OUT_OFFSET = STATS_OFFSET + 2 * 1024 * 8;
    static constexpr int OUT_BUF_BYTES = 48 * 32 * 4;
    static constexpr int TOTAL_STATS = OUT_OFFSET + 4 * 2 * OUT_BUF_BYTES + 1024;
};
constexpr int MAX_STAT_C = 1024;
constexpr int ACC_STRIDE = 256;    // TMEM columns between the two accumulator stages (BN <= 256)

// STATS = true (training-mode VNLinearLeakyReLU, models/vn_layers.py:60-74 + :116-127): the tile is BN = 240 rows = 80 whole points, and
// while an epilogue thread stores its channel's rows it also accumulates  sum ||p||, sum ||p||^2  (p = 3 consecutive TMEM columns, norm +
// 1e-6 as VNBatchNorm adds it) for the first Cstat output channels (the W_feat half of the stacked weight) in fp64: per-CTA partials in
// shared memory (a channel has exactly one owner thread per CTA, so plain read-modify-write), one fp64 atomicAdd per (CTA, channel) at
// the end.  The separate statistics pass over the freshly written activation (one full HBM read) disappears.
// In this variant the output leaves through shared memory and BULK TENSOR STORES (cp.async.bulk.tensor ... global.shared::cta, two
// 48-row x 32-channel staging buffers per epilogue warp): when HBM back-pressures the stores the warp does not stall on them, so the
// statistics arithmetic overlaps the draining stores instead of queueing behind them.
// SPLITK = true (few rows, R <= BN: the per-sample MLPs of the encoder / decoder heads, 96 rows against 1024 x 1024 weights): a tile is
// (output-channel tile, K range) and its partial product is added to the zero-initialised Y with red.add -- 8 CTAs streaming 512 KB of
// weights each become 64 CTAs streaming 64 KB.  kb_per = K blocks per split.
// CTA2 = true: launched as clusters of two CTAs (one SM pair).  The pair computes a 256-channel x BN-row tile with
// tcgen05.mma.cta_group::2 (M = 256): each CTA stages ITS 128 channels of W and HALF of the row block of X (BN / 2 rows), the tensor
// core reads both halves, each CTA's TMEM receives its 128 channels x BN rows and each CTA runs its own epilogue.  Per CTA the bytes
// fetched per tile drop from 16 + 32 KB to 16 + 16 KB per K block: X crosses the L2 -> SM fabric once per 256 channels instead of once
// per 128.  Rank 0 issues the MMAs; both producers signal rank 0's "full" barriers, the MMA commits free the stage in both CTAs.
// num_m = number of 256-channel tile PAIRS, num_tiles = pairs x row blocks.
template <int BN, int STAGES, bool HAS_BIAS, bool STATS, bool TMA_OUT = false, bool SPLITK = false, bool CTA2 = false, int GAP = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_rows_tf32_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_y, float* __restrict__ Y,
                      size_t ldy, long long R, int K, int Cout, const float* __restrict__ bias, size_t ldbias,
                      long long rows_per_sample, int num_m, long long num_tiles, double* __restrict__ sums, int Cstat, int kb_per) {
    static_assert(!SPLITK || (!HAS_BIAS && !STATS), "split-K tiles only add partial products");
    static_assert(!CTA2 || (!SPLITK && !TMA_OUT && BN % 16 == 0), "CTA pairs: plain tiles, N a multiple of 16");
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0;
    const long long tile_first = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
    const long long tile_step = CTA2 ? (gridDim.x >> 1) : gridDim.x;
    using L = RowsSmem<CTA2 ? BN / 2 : BN, STAGES, GAP>;      // a CTA of a pair stages half of the row block
    static_assert(!STATS || BN % 48 == 0, "the statistics epilogue walks whole points, 16 at a time");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    double* s_stats = reinterpret_cast<double*>(smem + L::STATS_OFFSET);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
        if (STATS && TMA_OUT) tma_prefetch_desc(&map_y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], CTA2 ? 8 : 4);      // rank 0's barrier collects the epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (STATS)
        for (int i = threadIdx.x; i < 2 * Cstat; i += NUM_THREADS) s_stats[i] = 0.0;
    if (warp == 2) {
        if (CTA2) tmem_alloc_2sm(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();      // the peer's barriers are initialised before anything is signalled across the pair
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PipeState ps;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int m0 = (int)((tile % num_m) * (CTA2 ? 2 : 1) + cta_rank) * BM;
                const long long n0 = SPLITK ? 0 : (tile / num_m) * BN;
                const int kb0 = SPLITK ? (int)(tile / num_m) * kb_per : 0;
                const int kb1 = SPLITK ? min(num_kb, kb0 + kb_per) : num_kb;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
                    uint8_t* sa = smem + ps.stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    if (CTA2) {
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[ps.stage], 2 * L::STAGE_BYTES);
                        tma_load_2d_2sm(&map_w, &full_bar[ps.stage], sa, kb * BK, m0);
                        tma_load_2d_2sm(&map_x, &full_bar[ps.stage], sb, kb * BK, (int)n0 + (int)cta_rank * (BN / 2));
                    } else {
                        mbar_expect_tx(&full_bar[ps.stage], L::STAGE_BYTES);
                        tma_load_2d(&map_w, &full_bar[ps.stage], sa, kb * BK, m0);
                        tma_load_2d(&map_x, &full_bar[ps.stage], sb, kb * BK, (int)n0);
                    }
                    ps.advance<STAGES>();
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(CTA2 ? 2 * BM : BM, BN, 0, 0);
            PipeState ps;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                const int kb0 = SPLITK ? (int)(tile / num_m) * kb_per : 0;
                const int kb1 = SPLITK ? min(num_kb, kb0 + kb_per) : num_kb;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[ps.stage], ps.phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + ps.stage * L::STAGE_BYTES);
                    uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t ad = make_desc(sa + k * UMMA_K * 4, 16, 1024);
                        const uint64_t bd = make_desc(sb + k * UMMA_K * 4, 16, 1024);
                        if (CTA2) umma_tf32_2sm(d_tmem, ad, bd, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        else umma_tf32(d_tmem, ad, bd, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
                    if (CTA2) umma_commit_2sm(&empty_bar[ps.stage]);
                    else umma_commit(&empty_bar[ps.stage]);
                    ps.advance<STAGES>();
                }
                // accumulator complete -> epilogue (of both CTAs of a pair)
                if (CTA2) umma_commit_2sm(&tfull_bar[acc]);
                else umma_commit(&tfull_bar[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        int out_buf = 0;
        for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
            const int m0 = (int)((tile % num_m) * (CTA2 ? 2 : 1) + cta_rank) * BM;
            const long long n0 = SPLITK ? 0 : (tile / num_m) * BN;
            const int o = m0 + quad * 32 + lane;
            const bool o_ok = o < Cout;
            // per-sample bias row (b, v) of output row r = (b*N + n)*3 + v.  A tile of BN rows touches at most two samples when
            // rows_per_sample >= BN: the 3 bias rows of the tile's first sample (za) and of the next one (zb) are fetched here, BEFORE
            // the accumulator wait, so their latency and the 64-bit division are paid once per tile and hidden behind the MMAs.
            long long tile_b = 0, tile_rem = 0;
            float za[3] = {0.f, 0.f, 0.f}, zb[3] = {0.f, 0.f, 0.f};
            if (HAS_BIAS) {
                tile_b = n0 / rows_per_sample;
                tile_rem = n0 - tile_b * rows_per_sample;
                if (o_ok) {
                    const float* bp = bias + (size_t)(tile_b * 3) * ldbias + o;
                    const bool next_ok = (tile_b + 1) * rows_per_sample < R;
#pragma unroll
                    for (int v3 = 0; v3 < 3; ++v3) {
                        za[v3] = __ldg(bp + (size_t)v3 * ldbias);
                        zb[v3] = next_ok ? __ldg(bp + (size_t)(3 + v3) * ldbias) : 0.f;
                    }
                }
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
            if (STATS) {
                // 48 columns = 16 whole points per pass (n0 and the pass offsets are multiples of 3: column j holds component j % 3)
                const bool do_stat = o < Cstat;      // warp-uniform (Cstat % 32 == 0)
                float* s_out = reinterpret_cast<float*>(smem + L::OUT_OFFSET) + (size_t)quad * 2 * (L::OUT_BUF_BYTES / 4);
                double s1 = 0.0, s2 = 0.0;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 48) {
                    const long long r0 = n0 + c0;
                    if (r0 >= R) break;      // warp-uniform
                    float v[48];
                    tmem_ld16(t_base + c0, v);
                    tmem_ld16(t_base + c0 + 16, v + 16);
                    tmem_ld16(t_base + c0 + 32, v + 32);
                    tmem_ld_wait();
                    if (HAS_BIAS && o_ok) {
                        const long long remc = tile_rem + c0;
                        const bool in_a = remc + 48 <= rows_per_sample;
                        const bool in_b = remc >= rows_per_sample && remc + 48 <= 2 * rows_per_sample;
                        if (in_a || in_b) {
                            const float t0 = in_a ? za[0] : zb[0], t1 = in_a ? za[1] : zb[1], t2 = in_a ? za[2] : zb[2];
#pragma unroll
                            for (int j = 0; j < 48; ++j) v[j] += (j % 3 == 0) ? t0 : ((j % 3 == 1) ? t1 : t2);
                        } else {        // the pass straddles a sample boundary, or rows_per_sample < BN: walk point by point
                            long long bc = r0 / rows_per_sample;
                            long long rc = r0 - bc * rows_per_sample;
#pragma unroll
                            for (int jp = 0; jp < 16; ++jp) {
                                if (r0 + 3 * jp < R) {
                                    const float* bp = bias + (size_t)(bc * 3) * ldbias + o;
                                    v[3 * jp] += __ldg(bp);
                                    v[3 * jp + 1] += __ldg(bp + ldbias);
                                    v[3 * jp + 2] += __ldg(bp + 2 * ldbias);
                                }
                                rc += 3;
                                if (rc >= rows_per_sample) {
                                    rc = 0;
                                    ++bc;
                                }
                            }
                        }
                    }
                    // stage the pass (row-major 48 x 32: a warp writes one 128-byte row per instruction, conflict-free) and hand it to
                    // the TMA engine; rows >= R and channels >= Cout are clipped by the tensor map
                    if (TMA_OUT) {
                        if (lane == 0) tma_store_wait_read<1>();      // the buffer used two passes ago has been read
                        __syncwarp();
                        float* sb = s_out + (size_t)out_buf * (L::OUT_BUF_BYTES / 4);
#pragma unroll
                        for (int j = 0; j < 48; ++j) sb[j * 32 + lane] = v[j];
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&map_y, sb, m0 + quad * 32, (int)r0);
                            tma_store_commit();
                        }
                        out_buf ^= 1;
                    } else if (o_ok) {
                        float* dst = Y + (size_t)r0 * ldy + o;
                        if (r0 + 48 <= R) {
#pragma unroll
                            for (int j = 0; j < 48; ++j) dst[(size_t)j * ldy] = v[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 48; ++j)
                                if (r0 + j < R) dst[(size_t)j * ldy] = v[j];
                        }
                    }
                    if (do_stat) {
                        // 16 norms per pass: MUFU rsqrt (this kernel only runs in the throughput mode), fp32 pairwise trees for the
                        // pass's two partial sums (16 terms: relative error ~1e-7, unbiased, averaged over >= 1e4 passes per channel),
                        // ONE fp64 add per sum per pass -- a per-point fp64 chain (cvt + DADD + DFMA, one warp per scheduler, no other
                        // warp to hide its latency) cost 0.77 ms on the decoder GEMM, more than the pass it replaces
                        float nn[16], qq[16];
#pragma unroll
                        for (int jp = 0; jp < 16; ++jp) {
                            const float n2 = fmaf(v[3 * jp + 2], v[3 * jp + 2], fmaf(v[3 * jp + 1], v[3 * jp + 1], v[3 * jp] * v[3 * jp]));
                            float n = (n2 > 0.f ? n2 * rsqrt_fast(fmaxf(n2, 1.17549435e-38f)) : 0.f) + 1e-6f;
                            n = (r0 + 3 * jp < R) ? n : 0.f;      // R is a multiple of 3: whole points only
                            nn[jp] = n;
                            qq[jp] = n * n;
                        }
#pragma unroll
                        for (int w_ = 8; w_ > 0; w_ >>= 1)
#pragma unroll
                            for (int jp = 0; jp < w_; ++jp) {
                                nn[jp] += nn[jp + w_];
                                qq[jp] += qq[jp + w_];
                            }
                        s1 += (double)nn[0];
                        s2 += (double)qq[0];
                    }
                }
                if (do_stat && o_ok) {      // this thread is the only one in the CTA that ever touches channel o
                    s_stats[o] += s1;
                    s_stats[Cstat + o] += s2;
                }
            } else
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const long long r0 = n0 + c0;
                if (r0 >= R) break;      // warp-uniform
                float v[32];
                tmem_ld32(t_base + c0, v);
                if (!o_ok) continue;
                float* dst = Y + (size_t)r0 * ldy + o;
                if (HAS_BIAS) {
                    // a 32-row chunk lies inside one sample except at sample boundaries: rotate that sample's 3 bias values by the
                    // chunk's first component and add with compile-time indices
                    const long long remc = tile_rem + c0;
                    const bool in_a = remc + 32 <= rows_per_sample;
                    const bool in_b = remc >= rows_per_sample && remc + 32 <= 2 * rows_per_sample;
                    if (in_a || in_b) {
                        const int vv = (int)((in_a ? remc : remc - rows_per_sample) % 3);
                        const float z0 = in_a ? za[0] : zb[0], z1 = in_a ? za[1] : zb[1], z2 = in_a ? za[2] : zb[2];
                        const float t0 = vv == 0 ? z0 : (vv == 1 ? z1 : z2);
                        const float t1 = vv == 0 ? z1 : (vv == 1 ? z2 : z0);
                        const float t2 = vv == 0 ? z2 : (vv == 1 ? z0 : z1);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += (j % 3 == 0) ? t0 : ((j % 3 == 1) ? t1 : t2);
                    } else {        // the chunk straddles a sample boundary, or rows_per_sample < BN: walk (sample, component) row by row
                        long long bc = r0 / rows_per_sample;                 // (compile-time indices into v: it stays in registers)
                        long long rc = r0 - bc * rows_per_sample;
                        int vc = (int)(rc % 3);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (r0 + j < R) v[j] += __ldg(bias + (size_t)(bc * 3 + vc) * ldbias + o);
                            vc = vc == 2 ? 0 : vc + 1;
                            if (++rc == rows_per_sample) {
                                rc = 0;
                                ++bc;
                            }
                        }
                    }
                }
                if (SPLITK) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (r0 + j < R) atomicAdd(dst + (size_t)j * ldy, v[j]);
                } else if (r0 + 32 <= R) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(size_t)j * ldy] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (r0 + j < R) dst[(size_t)j * ldy] = v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) mbar_arrive_rank0(&tempty_bar[acc]);
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }
    if (STATS && TMA_OUT && warp >= 4 && lane == 0) tma_store_wait_read<0>();      // staging buffers must outlive their bulk stores
    tc_fence_before();
    if (CTA2) cluster_sync_all();      // neither CTA's shared memory / TMEM disappears while the pair still uses it
    else __syncthreads();
    if (STATS)
        for (int i = threadIdx.x; i < 2 * Cstat; i += NUM_THREADS) {
            const double v = s_stats[i];
            if (v != 0.0) atomicAdd(sums + i, v);
        }
    if (warp == 2) {
        if (CTA2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// VNLinearLeakyReLU with the Vector-Neuron tail fused into the GEMM epilogue (no-grad / inference forward):
//     p = W_feat x (+ b_p),  d = W_dir x (+ b_d),  out = leaky(BN_vn(p), d)         models/vn_layers.py:60-74
// One CTA tile = 128 output channels x 96 rows (= 32 points x 3 components).  Two accumulators per tile live in TMEM
// (P for the feat rows of the stacked weight, D for the dir rows, same activation tile as B operand), double-buffered
// (4 x 96 = 384 columns).  An epilogue thread owns one channel; the three components of a point are three consecutive
// TMEM columns of the same thread, so the norm, the BatchNorm-on-norm scale, <p,d> and the projection are thread-local
// and only `out` ever reaches HBM: the linear outputs make no HBM round trip at all.
//   MODE_STATS : feat accumulator only; per-channel sum ||p||, sum ||p||^2 (fp64) for training-mode batch statistics
//   MODE_APPLY : feat + dir accumulators; writes out [R, C]
//   MODE_POOL  : VNLinear -> VNMaxPool (models/vn_layers.py:158-167): P = W x (the pooled layer), D = W_dir' x (its direction);
//                the epilogue forms score = <p, d> per point and keeps the per-channel arg-max (64-bit atomicMax on
//                (orderable score << 32 | ~n), first maximum wins); neither P nor D is written
// ---------------------------------------------------------------------------------------------------------------
constexpr int MODE_STATS = 0, MODE_APPLY = 1, MODE_POOL = 2;

// FBN rows per tile: a multiple of 3 (whole points) and of 16 (UMMA N granularity).  96 with two accumulator stages
// (epilogue of tile i overlaps the MMAs of tile i+1: the store-heavy APPLY mode), 192 with one stage (light epilogues:
// STATS / POOL; N = 192 keeps the shared-memory operand traffic per MMA cycle under the 128 B/clk limit).
template <int STAGES, int MODE, int FBN>
struct FusedSmem {
    static constexpr int A_BYTES = BM * BK * 4;                       // one weight tile (feat or dir)
    static constexpr int NA = MODE == MODE_STATS ? 1 : 2;
    static constexpr int B_BYTES = FBN * BK * 4;
    static constexpr int STAGE_BYTES = NA * A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};


// CTA2 = true: CTA pairs (tcgen05 cta_group::2), as in gemm_rows_tf32_kernel: the pair owns 256 channels (each CTA the feat and dir weight
// tiles of ITS 128 channels) and each CTA stages half of the row block; num_m then counts 256-channel pairs.
// Eight epilogue warps (two per TMEM lane quadrant, alternating 48-column passes): with a single accumulator stage (STATS / POOL) the
// epilogue is exposed between two tiles' MMAs, so it is spread over twice the warps.
constexpr int FUSED_THREADS = 32 * 12;
template <int STAGES, int MODE, bool FAST, int FBN, int NACC, bool CTA2 = false>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
gemm_vn_fused_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, float* __restrict__ out,
                     size_t ldo, long long R, int K, int C, const float* __restrict__ bias, size_t ldbias, long long rows_per_sample,
                     const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta, float ns,
                     double* __restrict__ sums, int num_m, long long num_tiles) {
    using L = FusedSmem<STAGES, MODE, CTA2 ? FBN / 2 : FBN>;      // a CTA of a pair stages half of the row block
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0;
    const long long tile_first = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
    const long long tile_step = CTA2 ? (gridDim.x >> 1) : gridDim.x;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BK - 1) / BK;
    constexpr int ACC_COLS = 2 * FBN;   // P | D per accumulator stage

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], CTA2 ? 16 : 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CTA2) tmem_alloc_2sm(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PipeState ps;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int m0 = (int)((tile % num_m) * (CTA2 ? 2 : 1) + cta_rank) * BM;
                const long long n0 = (tile / num_m) * FBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
                    uint8_t* sa = smem + ps.stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::NA * L::A_BYTES;
                    if (CTA2) {
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[ps.stage], 2 * L::STAGE_BYTES);
                        tma_load_2d_2sm(&map_w, &full_bar[ps.stage], sa, kb * BK, m0);
                        if (MODE != MODE_STATS) tma_load_2d_2sm(&map_w, &full_bar[ps.stage], sa + L::A_BYTES, kb * BK, C + m0);
                        tma_load_2d_2sm(&map_x, &full_bar[ps.stage], sb, kb * BK, (int)n0 + (int)cta_rank * (FBN / 2));
                    } else {
                        mbar_expect_tx(&full_bar[ps.stage], L::STAGE_BYTES);
                        tma_load_2d(&map_w, &full_bar[ps.stage], sa, kb * BK, m0);
                        if (MODE != MODE_STATS) tma_load_2d(&map_w, &full_bar[ps.stage], sa + L::A_BYTES, kb * BK, C + m0);
                        tma_load_2d(&map_x, &full_bar[ps.stage], sb, kb * BK, (int)n0);
                    }
                    ps.advance<STAGES>();
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(CTA2 ? 2 * BM : BM, FBN, 0, 0);
            PipeState ps;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t p_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
                const uint32_t d_tmem = p_tmem + FBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[ps.stage], ps.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + ps.stage * L::STAGE_BYTES);
                    const uint32_t sb = sa + L::NA * L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t bd = make_desc(sb + k * UMMA_K * 4, 16, 1024);
                        if (CTA2) {
                            umma_tf32_2sm(p_tmem, make_desc(sa + k * UMMA_K * 4, 16, 1024), bd, idesc, (kb | k) != 0 ? 1u : 0u);
                            if (MODE != MODE_STATS)
                                umma_tf32_2sm(d_tmem, make_desc(sa + L::A_BYTES + k * UMMA_K * 4, 16, 1024), bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        } else {
                            umma_tf32(p_tmem, make_desc(sa + k * UMMA_K * 4, 16, 1024), bd, idesc, (kb | k) != 0 ? 1u : 0u);
                            if (MODE != MODE_STATS)
                                umma_tf32(d_tmem, make_desc(sa + L::A_BYTES + k * UMMA_K * 4, 16, 1024), bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (CTA2) umma_commit_2sm(&empty_bar[ps.stage]);
                    else umma_commit(&empty_bar[ps.stage]);
                    ps.advance<STAGES>();
                }
                if (CTA2) umma_commit_2sm(&tfull_bar[acc]);
                else umma_commit(&tfull_bar[acc]);
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        double s1 = 0.0, s2 = 0.0;
        int stat_c = -1;                 // channel whose statistics s1/s2 currently hold
        long long pool_g = -1;           // MODE_POOL: group (sample) of the running winner
        unsigned long long pool_key = 0;
        unsigned long long* pool_best = reinterpret_cast<unsigned long long*>(sums);
        const float k1 = 1.f - ns;
        const long long pts_per_sample = rows_per_sample > 0 ? rows_per_sample / 3 : 1;
        for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
            const int m0 = (int)((tile % num_m) * (CTA2 ? 2 : 1) + cta_rank) * BM;
            const long long n0 = (tile / num_m) * FBN;
            const int c = m0 + quad * 32 + lane;             // C is a multiple of 128: always a valid channel
            if (MODE == MODE_STATS && c != stat_c) {
                if (stat_c >= 0) {
                    atomicAdd(sums + stat_c, s1);
                    atomicAdd(sums + C + stat_c, s2);
                }
                s1 = s2 = 0.0;
                stat_c = c;
            }
            float mean = 0.f, invstd = 0.f, ga = 0.f, be = 0.f;
            if (MODE == MODE_APPLY && stat) {
                mean = __ldg(stat + c);
                invstd = __ldg(stat + C + c);
                ga = __ldg(gamma + c);
                be = __ldg(beta + c);
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS);
#pragma unroll 1
            for (int h = warp >= 8 ? 1 : 0; h < FBN / 48; h += 2) {      // 48 columns = 16 points per pass; the quadrant's two warps alternate
                const long long r0 = n0 + h * 48;
                if (r0 >= R) break;                          // warp-uniform
                float pv[48], dv[48];
                tmem_ld16(t_base + h * 48 + 0, pv);
                tmem_ld16(t_base + h * 48 + 16, pv + 16);
                tmem_ld16(t_base + h * 48 + 32, pv + 32);
                if (MODE != MODE_STATS) {
                    tmem_ld16(t_base + FBN + h * 48 + 0, dv);
                    tmem_ld16(t_base + FBN + h * 48 + 16, dv + 16);
                    tmem_ld16(t_base + FBN + h * 48 + 32, dv + 32);
                }
                tmem_ld_wait();
                // per-sample bias rows of this pass (a pass of 16 points lies inside one sample unless it straddles a boundary)
                float bp[3] = {0.f, 0.f, 0.f}, bd[3] = {0.f, 0.f, 0.f};
                long long pt0 = r0 / 3;
                const long long pool_g0 = MODE == MODE_POOL ? pt0 / pts_per_sample : 0;
                const long long pool_n0 = MODE == MODE_POOL ? pt0 - pool_g0 * pts_per_sample : 0;
                long long b0 = 0;
                bool uniform_sample = true;
                if (bias) {
                    b0 = pt0 / pts_per_sample;
                    uniform_sample = (pt0 + 15) / pts_per_sample == b0;
#pragma unroll
                    for (int v = 0; v < 3; ++v) {
                        bp[v] = __ldg(bias + (size_t)(b0 * 3 + v) * ldbias + c);
                        if (MODE == MODE_APPLY) bd[v] = __ldg(bias + (size_t)(b0 * 3 + v) * ldbias + C + c);
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const long long r = r0 + 3 * j;
                    if (r >= R) break;
                    if (bias && !uniform_sample) {
                        const long long bb = (pt0 + j) / pts_per_sample;
#pragma unroll
                        for (int v = 0; v < 3; ++v) {
                            bp[v] = __ldg(bias + (size_t)(bb * 3 + v) * ldbias + c);
                            if (MODE == MODE_APPLY) bd[v] = __ldg(bias + (size_t)(bb * 3 + v) * ldbias + C + c);
                        }
                    }
                    if (MODE == MODE_POOL) {
                        // score = (x*d).sum(2): three rounded products summed left to right (vn_layers.py:163)
                        const float sc = __fadd_rn(__fadd_rn(__fmul_rn(pv[3 * j], dv[3 * j]), __fmul_rn(pv[3 * j + 1], dv[3 * j + 1])),
                                                   __fmul_rn(pv[3 * j + 2], dv[3 * j + 2]));
                        // group / index-in-group of point pt0 + j without a division per point
                        long long g = pool_g0;
                        long long nn = pool_n0 + j;
                        while (nn >= pts_per_sample) {
                            nn -= pts_per_sample;
                            ++g;
                        }
                        const unsigned n = (unsigned)nn;
                        unsigned u = __float_as_uint(sc);
                        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                        const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - n);
                        if (g != pool_g) {
                            if (pool_g >= 0) atomicMax(pool_best + (size_t)pool_g * C + c, pool_key);
                            pool_g = g;
                            pool_key = key;
                        } else {
                            pool_key = key > pool_key ? key : pool_key;
                        }
                        continue;
                    }
                    float p0 = pv[3 * j] + bp[0], p1 = pv[3 * j + 1] + bp[1], p2 = pv[3 * j + 2] + bp[2];
                    const float nrm2 = __fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2));
                    if (MODE == MODE_STATS) {
                        const double n = (double)(sqrtf(nrm2) + 1e-6f);
                        s1 += n;
                        s2 = fma(n, n, s2);
                    } else {
                        if (stat) {
                            const float n = (FAST ? (nrm2 > 0.f ? nrm2 * rsqrt_fast(fmaxf(nrm2, 1.17549435e-38f)) : 0.f) : sqrtf(nrm2)) + 1e-6f;
                            const float nb = ((n - mean) * invstd) * ga + be;
                            if (FAST) {
                                const float t = nb * rcp_fast(n);
                                p0 *= t;
                                p1 *= t;
                                p2 *= t;
                            } else {
                                p0 = p0 / n * nb;
                                p1 = p1 / n * nb;
                                p2 = p2 / n * nb;
                            }
                        }
                        const float d0 = dv[3 * j] + bd[0], d1 = dv[3 * j + 1] + bd[1], d2 = dv[3 * j + 2] + bd[2];
                        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(p0, d0), __fmul_rn(p1, d1)), __fmul_rn(p2, d2));
                        float i0 = p0, i1 = p1, i2 = p2;
                        if (!(dot >= 0.f)) {
                            const float dsq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), 1e-6f);
                            const float a = FAST ? dot * rcp_fast(dsq) : dot / dsq;
                            i0 = __fsub_rn(p0, __fmul_rn(a, d0));
                            i1 = __fsub_rn(p1, __fmul_rn(a, d1));
                            i2 = __fsub_rn(p2, __fmul_rn(a, d2));
                        }
                        float* dst = out + (size_t)r * ldo + c;
                        dst[0] = __fadd_rn(__fmul_rn(ns, p0), __fmul_rn(k1, i0));
                        dst[ldo] = __fadd_rn(__fmul_rn(ns, p1), __fmul_rn(k1, i1));
                        dst[2 * ldo] = __fadd_rn(__fmul_rn(ns, p2), __fmul_rn(k1, i2));
                    }
                }
            }
            if (MODE == MODE_POOL && pool_g >= 0) {      // flush this tile's winner (the next tile may belong to other channels)
                atomicMax(pool_best + (size_t)pool_g * C + c, pool_key);
                pool_g = -1;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) mbar_arrive_rank0(&tempty_bar[acc]);
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == NACC) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (MODE == MODE_STATS && stat_c >= 0) {
            atomicAdd(sums + stat_c, s1);
            atomicAdd(sums + C + stat_c, s2);
        }
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if (CTA2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused backward of the decoder tail  VNLinearLeakyReLU(Cin -> C) -> VNLinear(C, 1) (+ residual)  (models/pcn.py:340-345,387), dgrad half:
//     gh[r, k] = sum_c gpd[r, c] Wcat[c, k],   gpd = (gp | gd) = gradient of the stacked linear output (p | d)
// Unfused, gpd makes three HBM trips between four kernels (bwd1 writes it, bwd2 rewrites its gp half, the dgrad GEMM and the
// weight-gradient GEMM read it).  Here the gradient of the layer output is rank one (g[r, c] = gy[r] w2[c]), the BatchNorm-backward sums
// come from a sums-only pre-pass (vn_stream.cu, bn_leaky_bwd1_p2_kernel<.., STORE = false>), so FINAL gp and gd of a (point, channel)
// are a function of that point's p, d, gy alone: eight producer warps compute them for 32 points x 32 channels at a time (lane = channel:
// every global access is a full 128-byte row segment), write them ONCE to HBM (for the weight-gradient GEMM) and, in the K-major
// SWIZZLE_128B operand layout, to shared memory, from where tcgen05.mma contracts them against Wcat^T (TMA, from L2) into TMEM:
//     D[k (Cin lanes, MT = Cin / 128 tiles), rows (96 = 32 points)] += Wt[k, c-block] . gpd[rows, c-block]^T     (p half, then d half)
// HBM traffic of the tail backward: (pd read) + (pd read, gpd written, gh written) instead of 2 pd + 5 gpd-sized transfers + gh.
// Warp roles: 0 TMA producer (weights), 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue (gh), 8-15 gradient producers.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TB_BN = 96;           // rows per tile = 32 points
constexpr int TB_PW = 8;            // producer warps (4 points each)
constexpr int TB_THREADS = 32 * (8 + TB_PW);
constexpr int TB_SA = 2, TB_SB = 3;
constexpr int TB_A_HALF = 256 * BK * 4;          // one half (p or d) of a weight stage: up to 256 output rows x 32 channels
constexpr int TB_B_HALF = TB_BN * BK * 4;        // 96 rows x 32 channels
constexpr int TB_MAX_C = 256;
template <int A_HALF_, int SA_, int SB_>
struct TailSmemT {
    static constexpr int A_HALF = A_HALF_, SA = SA_, SB = SB_;
    static constexpr int A_STAGE = 2 * A_HALF;
    static constexpr int B_STAGE = 2 * TB_B_HALF;
    static constexpr int B_OFFSET = SA * A_STAGE;
    static constexpr int PAR_OFFSET = B_OFFSET + SB * B_STAGE;      // 7 per-channel parameter rows of TB_MAX_C floats
    static constexpr int BAR_OFFSET = PAR_OFFSET + 7 * TB_MAX_C * 4;
    static constexpr int NUM_BARS = 2 * SA + 5 * SB + 4;      // a_full/a_empty, b_loaded/b_full/b_empty/b_stored/b_pair, t_full/t_empty x 2
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;     // barriers + TMEM slot, alignment slack
    static_assert(NUM_BARS * 8 + 4 <= 512, "barrier block");
};
using TailSmem = TailSmemT<TB_A_HALF, TB_SA, TB_SB>;
// CTA pairs: a CTA stages only its 128 output rows of Wcat^T (16 KB per half) -- the shared memory that frees deepens both rings
using TailSmemPair = TailSmemT<BM * BK * 4, 3, 5>;
static_assert(TailSmemPair::TOTAL <= 232448, "pair layout exceeds the 227 KB of one CTA");


__device__ __forceinline__ uint32_t sw128_off(int r, int kappa) {      // element (row r, k index kappa < 32) of a K-major SWIZZLE_128B k-block
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((kappa >> 2) ^ (r & 7))) << 4) + (kappa & 3) * 4);
}

// Backward of  y = w2 . leaky(BN(p), d)  for ONE (point, channel):  (p, d) -> (dL/dp, dL/dd) in place, given the point's output gradient
// gy (3 components), the channel's BatchNorm parameters / statistics and the batch means m1, m2 of the pre-pass (zero in eval mode).
// Branch-free (the negative-side terms are selected, not jumped over) so that the unrolled callers interleave several points.
// Reference: models/vn_layers.py:60-74 (VNLinearLeakyReLU.forward), :116-127 (VNBatchNorm.forward), differentiated.
struct TailChan {
    float mean, invstd, ga, be, w2c, m1, m2;
};
__device__ __forceinline__ void tail_grad_point(const TailChan& ch, float k1, bool live, float gy0, float gy1, float gy2, float& p0, float& p1,
                                                float& p2, float& d0, float& d1, float& d2) {
    const float pp = fmaf(p2, p2, fmaf(p1, p1, p0 * p0));
    const bool nz = live && pp >= 1.17549435e-38f;
    const float rs = rsqrt_fast(nz ? pp : 1.f);       // 1 / |p|
    const float r = nz ? pp * rs : 0.f;
    const float n = r + 1e-6f;
    const float rn = rcp_fast(n);
    const float nhat = (n - ch.mean) * ch.invstd;
    const float nb = fmaf(nhat, ch.ga, ch.be);
    const float t = nb * rn;
    const float s = t * fmaf(p2, d2, fmaf(p1, d1, p0 * d0));            // <BN(p), d>
    const float g0 = gy0 * ch.w2c, g1 = gy1 * ch.w2c, g2 = gy2 * ch.w2c;      // rank-one gradient of the layer output
    const float rq = rcp_fast(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)) + 1e-6f);
    const bool neg = s < 0.f;
    const float a = s * rq;
    const float c1 = neg ? k1 * (fmaf(g2, d2, fmaf(g1, d1, g0 * d0)) * rq) : 0.f;
    const float e0 = fmaf(-c1, d0, g0), e1 = fmaf(-c1, d1, g1), e2 = fmaf(-c1, d2, g2);      // dL/d BN(p)
    const float ca = neg ? -k1 * a : 0.f, cb2 = -c1 * t, cc = 2.f * a * c1;
    const float q0 = fmaf(cc, d0, fmaf(cb2, p0, ca * g0));                                    // dL/d d
    const float q1 = fmaf(cc, d1, fmaf(cb2, p1, ca * g1));
    const float q2 = fmaf(cc, d2, fmaf(cb2, p2, ca * g2));
    // BatchNorm-on-norm backward with the batch sums of the pre-pass: final gradient w.r.t. the linear output p
    const float gx = fmaf(e2, p2, fmaf(e1, p1, e0 * p0));
    float dn = fmaf(-nhat, ch.m2, ch.ga * (gx * rn) - ch.m1);
    dn = dn * ch.invstd - gx * nb * rn * rn;
    const float ur = nz ? dn * rs : 0.f;
    p0 = fmaf(e0, t, ur * p0);
    p1 = fmaf(e1, t, ur * p1);
    p2 = fmaf(e2, t, ur * p2);
    d0 = live ? q0 : 0.f;
    d1 = live ? q1 : 0.f;
    d2 = live ? q2 : 0.f;
}

// Data path of one (tile, channel block): TMA brings the (p | d) blocks [96 rows x 32 channels] of pd into a B stage (SWIZZLE_128B, the
// MMA operand layout) -- asynchronously, TB_SB stages deep, so HBM latency is covered by the ring and not by registers --, the producer warps
// turn them IN PLACE into (gp | gd), the MMA contracts them and a store warp sends the same shared-memory blocks to gpd by bulk tensor stores.
// Warp roles: 0 TMA loads (weights + pd), 1 MMA issuer, 2 TMEM allocator, 3 TMA stores (gpd), 4-7 epilogue (gh), 8-15 gradient producers.
// CTA2 = true (Cin = 256): CTA pairs.  The pair owns 192 rows (each CTA produces the gradients of ITS 96 rows) and the 256 output rows of
// Wcat^T (each CTA stages the 128 rows it will store: half of the weight bytes per CTA); rank 0 issues tcgen05.mma.cta_group::2 (M = 256,
// N = 192) once both CTAs' producers have arrived on its pair barrier.  num_tiles counts 192-row pair tiles.
template <bool STORE_GPD, bool CTA2 = false>      // STORE_GPD false: gpd is not written (tail_wgrad_tf32_kernel forms the gradient itself)
__global__ void __launch_bounds__(TB_THREADS, 1)
tail_dgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_pd,
                       const __grid_constant__ CUtensorMap map_gpd, const float* __restrict__ gy, long long P, int C, int Cin,
                       const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ beta, float ns,
                       const float* __restrict__ w2, const double* __restrict__ sums, double count, int training, float* __restrict__ gh,
                       size_t ldgh, long long num_tiles) {
    using L = typename std::conditional<CTA2, TailSmemPair, TailSmem>::type;
    constexpr int SA = L::SA, SB = L::SB, A_HALF = L::A_HALF;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    float* s_par = reinterpret_cast<float*>(smem + L::PAR_OFFSET);      // [7][TB_MAX_C]: mean, invstd, gamma, beta, w2, m1, m2
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* a_empty = a_full + SA;
    uint64_t* b_loaded = a_empty + SA;      // TMA: the pd blocks of the stage have landed
    uint64_t* b_full = b_loaded + SB;       // producers: (gp | gd) written (count TB_PW); waited on by the MMA warp AND the store warp
    uint64_t* b_empty = b_full + SB;        // MMA: the stage has been read by the tensor core
    uint64_t* b_stored = b_empty + SB;      // store warp: the stage has been read by its bulk stores
    uint64_t* t_full = b_stored + SB;
    uint64_t* t_empty = t_full + 2;
    uint64_t* b_pair = t_empty + 2;            // CTA2: rank 0's copy collects the producers of both CTAs (count 2 * TB_PW) for the MMA thread
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_pair + SB);
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0;
    const long long tile_first = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;
    const long long tile_step = CTA2 ? (gridDim.x >> 1) : gridDim.x;
    constexpr int TILE_ROWS = CTA2 ? 2 * TB_BN : TB_BN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncb = C / 32;            // channel blocks of the p (and of the d) half
    const int MT = CTA2 ? 1 : Cin / 128;      // output-row tiles of 128 TMEM lanes per CTA (a pair splits Cin = 256 between its CTAs)
    const long long R = P * 3;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_wt);
        tma_prefetch_desc(&map_pd);
        tma_prefetch_desc(&map_gpd);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(&b_loaded[s], 1);
            mbar_init(&b_full[s], TB_PW);
            mbar_init(&b_empty[s], 1);
            mbar_init(&b_stored[s], 1);
            mbar_init(&b_pair[s], 2 * TB_PW);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&t_full[a], 1);
            mbar_init(&t_empty[a], CTA2 ? 8 : 4);
        }
        fence_barrier_init();
    }
    for (int c = threadIdx.x; c < C; c += TB_THREADS) {
        const float ga = __ldg(gamma + c);
        s_par[0 * TB_MAX_C + c] = __ldg(stat + c);
        s_par[1 * TB_MAX_C + c] = __ldg(stat + C + c);
        s_par[2 * TB_MAX_C + c] = ga;
        s_par[3 * TB_MAX_C + c] = __ldg(beta + c);
        s_par[4 * TB_MAX_C + c] = __ldg(w2 + c);
        s_par[5 * TB_MAX_C + c] = training ? (float)(sums[c] / count) * ga : 0.f;
        s_par[6 * TB_MAX_C + c] = training ? (float)(sums[C + c] / count) * ga : 0.f;
    }
    if (warp == 2) {
        if (CTA2) tmem_alloc_2sm(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PipeState pa, pb;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int row0 = (int)(tile * TILE_ROWS + cta_rank * TB_BN);
                for (int cb = 0; cb < ncb; ++cb) {
                    // pd blocks first (HBM latency), then this step's weights (L2)
                    mbar_wait(&b_empty[pb.stage], pb.phase ^ 1);
                    if (STORE_GPD) mbar_wait(&b_stored[pb.stage], pb.phase ^ 1);
                    uint8_t* sb = smem + L::B_OFFSET + pb.stage * L::B_STAGE;
                    mbar_expect_tx(&b_loaded[pb.stage], (uint32_t)L::B_STAGE);
                    tma_load_2d(&map_pd, &b_loaded[pb.stage], sb, cb * 32, row0);
                    tma_load_2d(&map_pd, &b_loaded[pb.stage], sb + TB_B_HALF, C + cb * 32, row0);
                    pb.advance<SB>();
                    mbar_wait(&a_empty[pa.stage], pa.phase ^ 1);
                    uint8_t* sa = smem + pa.stage * L::A_STAGE;
                    if (CTA2) {      // each CTA its 128 output rows, at the same offsets in both; bytes counted on rank 0's barrier
                        if (cta_rank == 0) mbar_expect_tx(&a_full[pa.stage], (uint32_t)(2 * 2 * BM * BK * 4));
                        tma_load_2d_2sm(&map_wt, &a_full[pa.stage], sa, cb * 32, (int)cta_rank * BM);
                        tma_load_2d_2sm(&map_wt, &a_full[pa.stage], sa + A_HALF, C + cb * 32, (int)cta_rank * BM);
                    } else {
                        mbar_expect_tx(&a_full[pa.stage], (uint32_t)(2 * MT * BM * BK * 4));
                        for (int m = 0; m < MT; ++m) {
                            tma_load_2d(&map_wt, &a_full[pa.stage], sa + m * (BM * BK * 4), cb * 32, m * BM);                      // p-half columns
                            tma_load_2d(&map_wt, &a_full[pa.stage], sa + A_HALF + m * (BM * BK * 4), C + cb * 32, m * BM);      // d-half columns
                        }
                    }
                    pa.advance<SA>();
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(CTA2 ? 2 * BM : BM, TILE_ROWS, 0, 0);
            PipeState pa, pb;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                mbar_wait(&t_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                for (int cb = 0; cb < ncb; ++cb) {
                    mbar_wait(&a_full[pa.stage], pa.phase);
                    mbar_wait(CTA2 ? &b_pair[pb.stage] : &b_full[pb.stage], pb.phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + pa.stage * L::A_STAGE);
                    const uint32_t sb = smem_u32(smem + L::B_OFFSET + pb.stage * L::B_STAGE);
                    for (int m = 0; m < MT; ++m) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256 + m * 128);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                const uint64_t ad = make_desc(sa + half * A_HALF + m * (BM * BK * 4) + k * UMMA_K * 4, 16, 1024);
                                const uint64_t bd = make_desc(sb + half * TB_B_HALF + k * UMMA_K * 4, 16, 1024);
                                if (CTA2) umma_tf32_2sm(d_tmem, ad, bd, idesc, (cb | half | k) != 0 ? 1u : 0u);
                                else umma_tf32(d_tmem, ad, bd, idesc, (cb | half | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    if (CTA2) {
                        umma_commit_2sm(&a_empty[pa.stage]);
                        umma_commit_2sm(&b_empty[pb.stage]);
                    } else {
                        umma_commit(&a_empty[pa.stage]);
                        umma_commit(&b_empty[pb.stage]);
                    }
                    pa.advance<SA>();
                    pb.advance<SB>();
                }
                if (CTA2) umma_commit_2sm(&t_full[acc]);
                else umma_commit(&t_full[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp == 3) {
        // gpd leaves from the same shared-memory blocks the MMA reads: two bulk tensor stores per stage (clipped at the tensor bounds)
        if (lane == 0 && STORE_GPD) {
            PipeState pb;
            for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int row0 = (int)(tile * TILE_ROWS + cta_rank * TB_BN);
                for (int cb = 0; cb < ncb; ++cb) {
                    mbar_wait(&b_full[pb.stage], pb.phase);
                    const uint8_t* sb = smem + L::B_OFFSET + pb.stage * L::B_STAGE;
                    tma_store_2d(&map_gpd, sb, cb * 32, row0);
                    tma_store_2d(&map_gpd, sb + TB_B_HALF, C + cb * 32, row0);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&b_stored[pb.stage]);
                    pb.advance<SB>();
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // epilogue: thread = output channel k of gh (TMEM lane), 32 rows per tcgen05.ld
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
            const long long n0 = tile * TILE_ROWS;
            mbar_wait(&t_full[acc], acc_phase);
            tc_fence_after();
            for (int m = 0; m < MT; ++m) {
                const int kch = (CTA2 ? (int)cta_rank : m) * BM + quad * 32 + lane;
                const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + m * 128);
#pragma unroll 1
                for (int c0 = 0; c0 < TILE_ROWS; c0 += 32) {
                    const long long r0 = n0 + c0;
                    if (r0 >= R) break;
                    float v[32];
                    tmem_ld32(t_base + c0, v);
                    float* dst = gh + (size_t)r0 * ldgh + kch;
                    if (r0 + 32 <= R) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) dst[(size_t)j * ldgh] = v[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (r0 + j < R) dst[(size_t)j * ldgh] = v[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) mbar_arrive_rank0(&t_empty[acc]);
                else mbar_arrive(&t_empty[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 8) {
        // gradient producers: warp pw owns points 4 pw .. 4 pw + 3 of the tile, lane = channel inside the 32-channel block
        const int pw = warp - 8;
        const float k1 = 1.f - ns;
        PipeState pb;
        for (long long tile = tile_first; tile < num_tiles; tile += tile_step) {
            const long long pt0 = tile * (TILE_ROWS / 3) + cta_rank * 32 + pw * 4;
            float gyv[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int v = 0; v < 3; ++v) gyv[i][v] = (pt0 + i < P) ? __ldg(gy + (size_t)(pt0 + i) * 3 + v) : 0.f;
            for (int cb = 0; cb < ncb; ++cb) {
                const int c = cb * 32 + lane;
                const TailChan ch = {s_par[c], s_par[TB_MAX_C + c], s_par[2 * TB_MAX_C + c], s_par[3 * TB_MAX_C + c], s_par[4 * TB_MAX_C + c],
                                     s_par[5 * TB_MAX_C + c], s_par[6 * TB_MAX_C + c]};
                mbar_wait(&b_loaded[pb.stage], pb.phase);      // the (p | d) blocks of this stage are in shared memory
                uint8_t* sbp = smem + L::B_OFFSET + pb.stage * L::B_STAGE;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = (pw * 4 + i) * 3;
                    float* ap0 = reinterpret_cast<float*>(sbp + sw128_off(rr + 0, lane));
                    float* ap1 = reinterpret_cast<float*>(sbp + sw128_off(rr + 1, lane));
                    float* ap2 = reinterpret_cast<float*>(sbp + sw128_off(rr + 2, lane));
                    float p0 = *ap0, p1 = *ap1, p2 = *ap2;
                    float d0 = ap0[TB_B_HALF / 4], d1 = ap1[TB_B_HALF / 4], d2 = ap2[TB_B_HALF / 4];
                    tail_grad_point(ch, k1, true, gyv[i][0], gyv[i][1], gyv[i][2], p0, p1, p2, d0, d1, d2);
                    *ap0 = p0;
                    *ap1 = p1;
                    *ap2 = p2;
                    ap0[TB_B_HALF / 4] = d0;
                    ap1[TB_B_HALF / 4] = d1;
                    ap2[TB_B_HALF / 4] = d2;
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the async proxy (tensor core, bulk stores)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&b_full[pb.stage]);                      // this CTA's store warp (and, one SM per tile, its MMA thread)
                    if (CTA2) mbar_arrive_rank0(&b_pair[pb.stage]);      // the pair's MMA thread
                }
                pb.advance<SB>();
            }
        }
    }
    tc_fence_before();
    if (CTA2) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        if (CTA2) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused backward of the decoder tail, weight-gradient half:  gW[c, k] = sum_r gpd[r, c] h[r, k]  with gpd formed on the fly, so that the
// stacked gradient (p | d) never exists in HBM.  Same producers as tail_dgrad_tf32_kernel, MN-major operands as gemm_wgrad_tf32_kernel:
// a stage is 24 rows (8 whole points); TMA (swizzle 128B_ATOM_32B) brings the (p | d) slabs [24 rows x 32 channels] x 4 of this CTA's 128
// channels and the eight 32-channel slabs of h; the producer warps turn (p | d) into (gp | gd) IN PLACE (UMMA layout SWIZZLE_128B_BASE32B:
// 32-byte chunk index XOR (row mod 4)); tcgen05.mma accumulates gp^T h and gd^T h (M = 128 channels each, N = Cin, K = 24 rows) in the 512
// TMEM columns over the CTA's whole row range; the epilogue adds the two tiles into gW with fp32 red.add.
// Grid: (C / 128 channel groups, row splits).  Warp roles: 0 TMA, 1 MMA, 2 TMEM allocator, 4-7 epilogue, 8-15 producers (warp = point).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TW_BR = 24;                        // rows per stage = 8 points
constexpr int TW_SLAB = TW_BR * 128;             // one 32-channel slab of a stage
constexpr int TW_STAGES = 4;
struct TailWSmem {
    static constexpr int A_HALF = 4 * TW_SLAB;             // 128 channels of p (or d)
    static constexpr int B_BYTES = 8 * TW_SLAB;            // up to 256 channels of h
    static constexpr int STAGE = 2 * A_HALF + B_BYTES;     // 48 KB
    static constexpr int PAR_OFFSET = TW_STAGES * STAGE;
    static constexpr int BAR_OFFSET = PAR_OFFSET + 7 * 128 * 4;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

__device__ __forceinline__ uint32_t sw32_off(int r, int c) {      // element (row r, channel c < 32) of an MN-major SWIZZLE_128B_BASE32B slab
    return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 3))) << 5) + (c & 7) * 4);
}

__global__ void __launch_bounds__(TB_THREADS, 1)
tail_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_pd, const __grid_constant__ CUtensorMap map_h, const float* __restrict__ gy,
                       long long P, int C, int Cin, const float* __restrict__ stat, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float ns, const float* __restrict__ w2, const double* __restrict__ sums, double count,
                       int training, float* __restrict__ gW, size_t ldgw, long long pts_per_split) {
    using L = TailWSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    float* s_par = reinterpret_cast<float*>(smem + L::PAR_OFFSET);      // [7][128]
    uint64_t* loaded = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* full = loaded + TW_STAGES;
    uint64_t* empty = full + TW_STAGES;
    uint64_t* t_full = empty + TW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = blockIdx.x;                    // channels [cg * 128, cg * 128 + 128) of both halves
    const long long pt_begin = (long long)blockIdx.y * pts_per_split;
    const long long pt_end = pt_begin + pts_per_split < P ? pt_begin + pts_per_split : P;
    const int num_st = pt_end > pt_begin ? (int)((pt_end - pt_begin + 7) / 8) : 0;
    const int nslab_h = Cin / 32;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_pd);
        tma_prefetch_desc(&map_h);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TW_STAGES; ++s) {
            mbar_init(&loaded[s], 1);
            mbar_init(&full[s], TB_PW);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&t_full[0], 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 128; i += TB_THREADS) {
        const int c = cg * 128 + i;
        const float ga = __ldg(gamma + c);
        s_par[0 * 128 + i] = __ldg(stat + c);
        s_par[1 * 128 + i] = __ldg(stat + C + c);
        s_par[2 * 128 + i] = ga;
        s_par[3 * 128 + i] = __ldg(beta + c);
        s_par[4 * 128 + i] = __ldg(w2 + c);
        s_par[5 * 128 + i] = training ? (float)(sums[c] / count) * ga : 0.f;
        s_par[6 * 128 + i] = training ? (float)(sums[C + c] / count) * ga : 0.f;
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PipeState ps;
            for (int st = 0; st < num_st; ++st) {
                mbar_wait(&empty[ps.stage], ps.phase ^ 1);
                uint8_t* sa = smem + ps.stage * L::STAGE;
                const int row0 = (int)((pt_begin + (long long)st * 8) * 3);
                // rows past pt_end * 3 (but < R) inside the last box of a split belong to the next split: the producers zero them;
                // rows >= R are zero-filled by TMA
                mbar_expect_tx(&loaded[ps.stage], (uint32_t)(2 * L::A_HALF + nslab_h * TW_SLAB));
                // three bulk copies per stage (slab maps): 4 slabs of p, 4 of d, all of h
                tma_load_3d(&map_pd, &loaded[ps.stage], sa, 0, row0, cg * 4);
                tma_load_3d(&map_pd, &loaded[ps.stage], sa + L::A_HALF, 0, row0, (C >> 5) + cg * 4);
                tma_load_3d(&map_h, &loaded[ps.stage], sa + 2 * L::A_HALF, 0, row0, 0);
                ps.advance<TW_STAGES>();
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, Cin, 1, 1);
            PipeState ps;
            for (int st = 0; st < num_st; ++st) {
                mbar_wait(&full[ps.stage], ps.phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + ps.stage * L::STAGE);
                const uint32_t sb = sa + 2 * L::A_HALF;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int k = 0; k < TW_BR / UMMA_K; ++k) {
                        const uint64_t ad = make_desc(sa + half * L::A_HALF + k * 1024, TW_SLAB, 512, 1);
                        const uint64_t bd = make_desc(sb + k * 1024, TW_SLAB, 512, 1);
                        umma_tf32(tmem_base + (uint32_t)(half * 256), ad, bd, idesc, (st | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty[ps.stage]);
                ps.advance<TW_STAGES>();
            }
            umma_commit(&t_full[0]);
        }
    } else if (warp >= 4 && warp < 8) {
        const int quad = warp & 3;
        if (num_st > 0) {
            mbar_wait(&t_full[0], 0);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int o = half * C + cg * 128 + quad * 32 + lane;      // row of gW = channel of the stacked weight
                const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 256);
#pragma unroll 1
                for (int c0 = 0; c0 < Cin; c0 += 32) {
                    float v[32];
                    tmem_ld32(t_base + c0, v);
                    float* dst = gW + (size_t)o * ldgw + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                }
            }
        }
    } else if (warp >= 8) {
        const int pw = warp - 8;                  // point of the stage
        const float k1 = 1.f - ns;
        PipeState ps;
        for (int st = 0; st < num_st; ++st) {
            const long long pt = pt_begin + (long long)st * 8 + pw;
            const bool live = pt < pt_end;
            float gyv[3];
#pragma unroll
            for (int v = 0; v < 3; ++v) gyv[v] = live ? __ldg(gy + (size_t)pt * 3 + v) : 0.f;
            mbar_wait(&loaded[ps.stage], ps.phase);
            uint8_t* sa = smem + ps.stage * L::STAGE;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = j * 32 + lane;
                const TailChan ch = {s_par[ci], s_par[128 + ci], s_par[256 + ci], s_par[384 + ci], s_par[512 + ci], s_par[640 + ci], s_par[768 + ci]};
                uint8_t* sp = sa + j * TW_SLAB;
                float* ap0 = reinterpret_cast<float*>(sp + sw32_off(pw * 3 + 0, lane));
                float* ap1 = reinterpret_cast<float*>(sp + sw32_off(pw * 3 + 1, lane));
                float* ap2 = reinterpret_cast<float*>(sp + sw32_off(pw * 3 + 2, lane));
                float p0 = *ap0, p1 = *ap1, p2 = *ap2;
                float d0 = ap0[L::A_HALF / 4], d1 = ap1[L::A_HALF / 4], d2 = ap2[L::A_HALF / 4];
                tail_grad_point(ch, k1, live, gyv[0], gyv[1], gyv[2], p0, p1, p2, d0, d1, d2);
                *ap0 = p0;
                *ap1 = p1;
                *ap2 = p2;
                ap0[L::A_HALF / 4] = d0;
                ap1[L::A_HALF / 4] = d1;
                ap2[L::A_HALF / 4] = d2;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[ps.stage]);
            ps.advance<TW_STAGES>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient:  G[o, k] += sum_{r in chunk} dY[r, o] X[r, k]      (both operands MN-major, split over r)
//   A = dY^T : MN-major, 4 slabs of {32 o} x BR rows ; B = X : MN-major, BNW/32 slabs of {32 k} x BR rows
//   (TMA swizzle mode 128B_ATOM_32B <-> UMMA layout SWIZZLE_128B_BASE32B, the pairing tf32 MN-major requires)
// ---------------------------------------------------------------------------------------------------------------
constexpr int BR = 32;   // reduction rows per stage (4 MMAs of K = 8)

template <int BNW, int STAGES>
struct WgradSmem {
    static constexpr int A_BYTES = BM * BR * 4;
    static constexpr int B_BYTES = BNW * BR * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BNW, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, float* __restrict__ G,
                       size_t ldg, long long R, int Cout, int K, long long rows_per_split) {
    using L = WgradSmem<BNW, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;        // output-channel tile
    const int k0 = blockIdx.y * BNW;       // input-channel tile
    const long long r_begin = (long long)blockIdx.z * rows_per_split;
    const long long r_end = (r_begin + rows_per_split < R) ? r_begin + rows_per_split : R;
    const int num_rb = (int)((r_end - r_begin + BR - 1) / BR);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tfull_bar[0], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BNW < 32 ? 32 : BNW);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PipeState ps;
            for (int rb = 0; rb < num_rb; ++rb) {
                mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
                uint8_t* sa = smem + ps.stage * L::STAGE_BYTES;
                uint8_t* sb = sa + L::A_BYTES;
                mbar_expect_tx(&full_bar[ps.stage], L::STAGE_BYTES);
                const int r0 = (int)(r_begin + (long long)rb * BR);
                // NOTE: rows past r_end (but < R) inside the last box of a split belong to the next split: they are
                // excluded by loading them into the box only when r_end is BR-aligned (the host aligns splits to BR);
                // rows >= R are zero-filled by TMA.
#pragma unroll
                for (int s = 0; s < BM / 32; ++s) tma_load_2d(&map_dy, &full_bar[ps.stage], sa + s * (BR * 128), m0 + s * 32, r0);
#pragma unroll
                for (int s = 0; s < BNW / 32; ++s) tma_load_2d(&map_x, &full_bar[ps.stage], sb + s * (BR * 128), k0 + s * 32, r0);
                ps.advance<STAGES>();
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BNW, 1, 1);
            PipeState ps;
            for (int rb = 0; rb < num_rb; ++rb) {
                mbar_wait(&full_bar[ps.stage], ps.phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + ps.stage * L::STAGE_BYTES);
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BR / UMMA_K; ++k) {
                    // MN-major, SWIZZLE_128B_BASE32B: a slab is BR rows (reduction index) x 128 bytes (32 channels); the
                    // swizzle atom is 4 rows, so one K=8 instruction spans two atoms: SBO = 512 B between them,
                    // LBO = BR*128 B between the 32-channel slabs along M / N
                    const uint64_t ad = make_desc(sa + k * 1024, BR * 128, 512, 1);
                    const uint64_t bd = make_desc(sb + k * 1024, BR * 128, 512, 1);
                    umma_tf32(tmem_base, ad, bd, idesc, (rb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[ps.stage]);
                ps.advance<STAGES>();
            }
            umma_commit(&tfull_bar[0]);
        }
    } else if (warp >= 4) {
        const int quad = warp & 3;
        const int o = m0 + quad * 32 + lane;
        if (num_rb > 0) {
            mbar_wait(&tfull_bar[0], 0);
            tc_fence_after();
            const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BNW; c0 += 32) {
                if (k0 + c0 >= K) break;
                float v[32];
                tmem_ld32(t_base + c0, v);
                if (o < Cout) {
                    float* dst = G + (size_t)o * ldg + k0 + c0;
                    if (k0 + c0 + 32 <= K) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (k0 + c0 + j < K) atomicAdd(dst + j, v[j]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, BNW < 32 ? 32 : BNW);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major matrix [rows, cols] with leading dimension ld; box = {box_cols (inner), box_rows}; 128B swizzle
static bool make_map(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// The same matrix seen as [rows][cols / 32 slabs][32]: ONE bulk copy of box {32, box_rows, nslabs} lands as `nslabs` consecutive MN-major
// slabs [box_rows x 128 bytes] in shared memory (slab outermost), instead of one TMA instruction per slab -- a stage of the weight-gradient
// kernels is 12-16 slabs of 3-4 KB, and at ~50 ns per TMA instruction their issue alone was most of a stage's time.
static bool make_map_slabs(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows, int nslabs,
                           CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || (cols & 31)) return false;
    cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)nslabs};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, int STAGES, bool STATS, bool TMA_OUT = false>
static int launch_rows(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy, long long R, int K,
                       int Cout, const float* bias, long long ldbias, long long rps, double* sums, int Cstat, cudaStream_t st) {
    using L = RowsSmem<BN, STAGES>;
    constexpr int SMEM = STATS ? (TMA_OUT ? L::TOTAL_STATS : L::OUT_OFFSET + 1024) : L::TOTAL;
    CUtensorMap mw, mx, my;
    if (!make_map(&mw, W, Cout, K, ldw, BK, BM)) return VNPCC_ERR_DRIVER;
    if (!make_map(&mx, X, R, K, ldx, BK, BN)) return VNPCC_ERR_DRIVER;
    if (STATS && TMA_OUT) {      // output boxes of 32 channels x 48 rows, plain row-major staging (no swizzle)
        if (!make_map(&my, Y, R, Cout, ldy, 32, 48, CU_TENSOR_MAP_SWIZZLE_NONE)) return VNPCC_ERR_DRIVER;
    } else {
        my = mx;
    }
    static bool attr_done_dev[64] = {false};      // the attribute is per device
    bool& attr_done = attr_done_dev[current_device_slot()];
    auto kern = bias ? gemm_rows_tf32_kernel<BN, STAGES, true, STATS, TMA_OUT> : gemm_rows_tf32_kernel<BN, STAGES, false, STATS, TMA_OUT>;
    if (!attr_done) {
        if (cudaFuncSetAttribute(gemm_rows_tf32_kernel<BN, STAGES, true, STATS, TMA_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
                cudaSuccess ||
            cudaFuncSetAttribute(gemm_rows_tf32_kernel<BN, STAGES, false, STATS, TMA_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
                cudaSuccess)
            return last_error();
        attr_done = true;
    }
    const int num_m = (Cout + BM - 1) / BM;
    const long long num_n = (R + BN - 1) / BN;
    const long long num_tiles = num_m * num_n;
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    count_launch(), kern<<<grid, NUM_THREADS, SMEM, st>>>(mw, mx, my, Y, (size_t)ldy, R, K, Cout, bias, (size_t)ldbias,
                                                         rps > 0 ? rps : 1, num_m, num_tiles, sums, Cstat, 0);
    return last_error();
}

constexpr int PAIR_LAUNCH_REFUSED = -77;
// CTA pairs (see the kernel's CTA2 note): Cout a multiple of 256, many row blocks
template <int BN, int STAGES, bool STATS, int GAP = 0>      // STAGES of 16 KB of W + BN / 2 rows of X
static int launch_rows_pair(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy, long long R, int K, int Cout,
                            const float* bias, long long ldbias, long long rps, double* sums, int Cstat, cudaStream_t st) {
    using L = RowsSmem<BN / 2, STAGES, GAP>;
    constexpr int SMEM = STATS ? L::OUT_OFFSET + 1024 : L::TOTAL;
    static_assert(SMEM <= 232448, "stage ring exceeds the 227 KB of one CTA");
    CUtensorMap mw, mx;
    if (!make_map(&mw, W, Cout, K, ldw, BK, BM)) return VNPCC_ERR_DRIVER;
    if (!make_map(&mx, X, R, K, ldx, BK, BN / 2)) return VNPCC_ERR_DRIVER;      // each CTA of the pair stages half of the row block
    auto kb = gemm_rows_tf32_kernel<BN, STAGES, true, STATS, false, false, true, GAP>;
    auto kn = gemm_rows_tf32_kernel<BN, STAGES, false, STATS, false, false, true, GAP>;
    static bool attr_done_dev[64] = {false};
    bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(kn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
            return last_error();
        attr_done = true;
    }
    const int num_mp = Cout / (2 * BM);
    const long long num_n = (R + BN - 1) / BN;
    const long long num_tiles = (long long)num_mp * num_n;
    long long pairs = sm_count() / 2;
    if (pairs > num_tiles) pairs = num_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    const cudaError_t e = cudaLaunchKernelEx(&cfg, bias ? kb : kn, mw, mx, mx, Y, (size_t)ldy, R, K, Cout, bias, (size_t)ldbias,
                                             (long long)(rps > 0 ? rps : 1), num_mp, num_tiles, sums, Cstat, 0);
    if (e != cudaSuccess) {      // the cluster launch itself was refused (e.g. a partition without SM pairs): the caller uses one SM per tile
        (void)cudaGetLastError();
        return PAIR_LAUNCH_REFUSED;
    }
    return last_error();
}

// few rows (R <= 128), no bias: split the contraction over CTAs (see the kernel's SPLITK note); Y is zeroed here
static int launch_rows_splitk(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy, long long R, int K, int Cout,
                              int ksplit, cudaStream_t st) {
    constexpr int BN = 128, STAGES = 4;
    using L = RowsSmem<BN, STAGES>;
    CUtensorMap mw, mx;
    if (!make_map(&mw, W, Cout, K, ldw, BK, BM)) return VNPCC_ERR_DRIVER;
    if (!make_map(&mx, X, R, K, ldx, BK, BN)) return VNPCC_ERR_DRIVER;
    auto kern = gemm_rows_tf32_kernel<BN, STAGES, false, false, false, true>;
    static bool attr_done_dev[64] = {false};
    bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return last_error();
        attr_done = true;
    }
    const int num_kb = (K + BK - 1) / BK;
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    ksplit = (num_kb + kb_per - 1) / kb_per;      // no empty split
    const int num_m = (Cout + BM - 1) / BM;
    const long long num_tiles = (long long)num_m * ksplit;
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    if (cudaMemset2DAsync(Y, (size_t)ldy * sizeof(float), 0, (size_t)Cout * sizeof(float), (size_t)R, st) != cudaSuccess) return last_error();
    count_launch(), kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(mw, mx, mx, Y, (size_t)ldy, R, K, Cout, nullptr, 0, 1, num_m, num_tiles, nullptr, 0, kb_per);
    return last_error();
}


constexpr int POOL_PAIR_FBN = 192;
template <int MODE, bool FAST>
static int launch_fused(const float* X, long long ldx, const float* Wcat, long long ldw, float* out, long long ldo, long long R, int K,
                        int C, const float* bias, long long ldbias, long long rps, const float* stat, const float* gamma,
                        const float* beta, float ns, double* sums, cudaStream_t st) {
    constexpr int FBN = MODE == MODE_APPLY ? 96 : 192;
    constexpr int NACC = MODE == MODE_APPLY ? 2 : 1;
    constexpr int STAGES = MODE == MODE_POOL ? 3 : 4;
    using L = FusedSmem<STAGES, MODE, FBN>;
    CUtensorMap mw, mx;
    if (!make_map(&mw, Wcat, MODE == MODE_STATS ? C : 2 * C, K, ldw, BK, BM)) return VNPCC_ERR_DRIVER;
    if (!make_map(&mx, X, R, K, ldx, BK, FBN)) return VNPCC_ERR_DRIVER;
    auto kern = gemm_vn_fused_kernel<STAGES, MODE, FAST, FBN, NACC>;
    static bool attr_done_dev[64] = {false};      // the attribute is per device
    bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return last_error();
        attr_done = true;
    }
    // CTA pairs are built and tested for this kernel too, but measured no faster (fused conv -> pool GEMM 2.50 vs 2.47 ms: it is bound by
    // the tensor pipe and its exposed epilogue, not by operand traffic): one SM per tile unless vnpcc_set_tuning(2, 4)
    if (tuning(TUNE_STATS_GEMM) == 4 && C % (2 * BM) == 0 && R >= 8192 && K >= 256) {
        // CTA pairs: more stages fit because each CTA stages half of the row block.  POOL: with two SMs feeding one MMA the 96-row tile's
        // operand traffic fits the shared-memory port (A 4 KB + half of B 1.5 KB per 48-cycle instruction), so the tile can be
        // double-buffered in TMEM (2 x (P | D) x 96 columns) and the arg-max epilogue overlaps the next tile's MMAs
        constexpr int PFBN = MODE == MODE_POOL ? (POOL_PAIR_FBN) : FBN;
        constexpr int PNACC = MODE == MODE_POOL ? (PFBN == 96 ? 2 : 1) : NACC;
        constexpr int PSTAGES = MODE == MODE_STATS ? 6 : (MODE == MODE_POOL ? (PFBN == 96 ? 5 : 5) : 5);
        using LP = FusedSmem<PSTAGES, MODE, PFBN / 2>;
        static_assert(LP::TOTAL <= 232448, "stage ring exceeds the 227 KB of one CTA");
        CUtensorMap mxh;
        if (!make_map(&mxh, X, R, K, ldx, BK, PFBN / 2)) return VNPCC_ERR_DRIVER;
        auto kp = gemm_vn_fused_kernel<PSTAGES, MODE, FAST, PFBN, PNACC, true>;
        static bool attr_done_p[64] = {false};
        bool& done_p = attr_done_p[current_device_slot()];
        if (!done_p) {
            if (cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::TOTAL) != cudaSuccess) return last_error();
            done_p = true;
        }
        const long long num_n = (R + PFBN - 1) / PFBN;
        const int num_mp = C / (2 * BM);
        const long long tiles = (long long)num_mp * num_n;
        long long pairs = sm_count() / 2;
        if (pairs > tiles) pairs = tiles;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(FUSED_THREADS);
        cfg.dynamicSmemBytes = LP::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        count_launch();
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kp, mw, mxh, out, (size_t)ldo, R, K, C, bias, (size_t)ldbias, rps, stat, gamma, beta, ns,
                                                 sums, num_mp, tiles);
        return e == cudaSuccess ? last_error() : (int)e;
    }
    const int num_m = C / BM;
    const long long num_n = (R + FBN - 1) / FBN;
    const long long num_tiles = num_m * num_n;
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    count_launch(), kern<<<grid, FUSED_THREADS, L::TOTAL, st>>>(mw, mx, out, (size_t)ldo, R, K, C, bias, (size_t)ldbias, rps, stat, gamma, beta,
                                                             ns, sums, num_m, num_tiles);
    return last_error();
}

}  // namespace tc
}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

// Which form of the rows GEMM a problem gets -- a pure host function of the sizes (exported as vnpcc_debug_rows_plan for
// tests/test_planners_cpu.py).  variant 0: one SM per 128-channel tile; 1: CTA pairs (256-channel tiles, cta_group::2); 2: split-K (few rows).
struct RowsPlan {
    int variant, ksplit, bn, grid;
    long long tiles;
};
static RowsPlan plan_rows(long long R, int K, int Cout, bool has_bias, bool stats, int sms, int tiling_knob, int legacy_grid) {
    RowsPlan p = {0, 1, stats ? 240 : 256, 0, 0};
    const int num_m = (Cout + tc::BM - 1) / tc::BM, num_kb = (K + tc::BK - 1) / tc::BK;
    if (!stats && R <= 128) {
        // the grid would be Cout / 128 CTAs: split the contraction until about one CTA per SM streams >= 4 K blocks (128 columns of W)
        p.bn = 128;
        int ksplit = sms / num_m < num_kb / 4 ? sms / num_m : num_kb / 4;
        if (legacy_grid) ksplit = 1;
        if (!has_bias && ksplit >= 2) {
            const int kb_per = (num_kb + ksplit - 1) / ksplit;
            p.variant = 2;
            p.ksplit = (num_kb + kb_per - 1) / kb_per;      // no empty split
        }
        p.tiles = (long long)num_m * p.ksplit;
        p.grid = (int)(p.tiles < sms ? p.tiles : sms);
        return p;
    }
    const long long num_n = (R + p.bn - 1) / p.bn;
    // CTA pairs where a 256-channel tile exists and the contraction is long enough to amortise the pair's epilogues: measured 3-12 % faster
    // than one SM per tile at the train step's shapes, 15 % slower at K = 128 (tools/pair_gemm_bench.py)
    if ((tiling_knob == 0 || tiling_knob == 4) && Cout % 256 == 0 && R >= 8192 && K >= 256 && sms >= 2) {
        p.variant = 1;
        p.tiles = (long long)(Cout / 256) * num_n;
        const long long pairs = p.tiles < sms / 2 ? p.tiles : sms / 2;
        p.grid = (int)(2 * pairs);
        return p;
    }
    p.tiles = (long long)num_m * num_n;
    p.grid = (int)(p.tiles < sms ? p.tiles : sms);
    return p;
}

// include/vnpcc_debug.h: out = {variant, K splits, rows per tile, grid, tiles}
void vnpcc_debug_rows_plan(long long R, int K, int Cout, int has_bias, int stats, int sms, long long* out) {
    const RowsPlan p = plan_rows(R, K, Cout, has_bias != 0, stats != 0, sms > 0 ? sms : 148, 0, 0);
    out[0] = p.variant;
    out[1] = p.ksplit;
    out[2] = p.bn;
    out[3] = p.grid;
    out[4] = p.tiles;
}

int vnpcc_gemm_rows_tf32(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy, long long R,
                         int K, int Cout, const float* bias, long long ldbias, long long rows_per_sample, void* stream) {
    if (R <= 0 || Cout <= 0) return 0;
    // TMA needs 16-byte aligned bases and row pitches; tiny K is better served by the SIMT kernel
    if (K < 32 || (K & 3) || (ldx & 3) || (ldw & 3) || !tc::aligned16(X) || !tc::aligned16(W) || R >= (1ll << 31))
        return VNPCC_ERR_UNSUPPORTED;
    if (Cout < 64 || R < 64) return VNPCC_ERR_UNSUPPORTED;
    if (bias && (rows_per_sample <= 0 || rows_per_sample % 3 != 0)) return VNPCC_ERR_BAD_ARG;   // a sample is whole points (3 rows each)
    cudaStream_t st = (cudaStream_t)stream;
    const RowsPlan plan = plan_rows(R, K, Cout, bias != nullptr, false, sm_count(), tuning(TUNE_STATS_GEMM), tuning(TUNE_GRID_LEGACY));
    if (plan.variant == 2) return tc::launch_rows_splitk(X, ldx, W, ldw, Y, ldy, R, K, Cout, plan.ksplit, st);
    if (R <= 128) return tc::launch_rows<128, 4, false>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, nullptr, 0, st);
    if (plan.variant == 1) {
        const int rc = tc::launch_rows_pair<256, 6, false>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, nullptr, 0, st);
        if (rc != tc::PAIR_LAUNCH_REFUSED) return rc;
    }
    return tc::launch_rows<256, 4, false>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, nullptr, 0, st);
}

// the same GEMM for a training-mode VNLinearLeakyReLU (models/vn_layers.py:60-74): additionally returns the BatchNorm-on-norm batch
// statistics of the first Cstat output channels, sums[c] = sum_points (||y[.,c]|| + 1e-6), sums[Cstat + c] = the same squared (fp64;
// zeroed here), accumulated in the epilogue while the rows are stored -- no separate pass over Y (vnpcc_vn_norm_stats) is needed.
// R must be whole points (R % 3 == 0), Cstat % 32 == 0, Cstat <= min(Cout, 1024).
int vnpcc_gemm_rows_tf32_stats(const float* X, long long ldx, const float* W, long long ldw, float* Y, long long ldy, long long R,
                               int K, int Cout, const float* bias, long long ldbias, long long rows_per_sample, double* sums,
                               int Cstat, void* stream) {
    if (R <= 0 || Cout <= 0) return 0;
    if (K < 32 || (K & 3) || (ldx & 3) || (ldw & 3) || !tc::aligned16(X) || !tc::aligned16(W) || R >= (1ll << 31))
        return VNPCC_ERR_UNSUPPORTED;
    if (Cout < 64 || R < 240 || R % 3 != 0 || !sums || Cstat <= 0 || (Cstat & 31) || Cstat > Cout || Cstat > tc::MAX_STAT_C ||
        (ldy & 3) || !tc::aligned16(Y))
        return VNPCC_ERR_UNSUPPORTED;
    if (bias && (rows_per_sample <= 0 || rows_per_sample % 3 != 0)) return VNPCC_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * Cstat, st);
    const int kc = tuning(TUNE_STATS_NOMATH) ? 0 : Cstat;
    if (plan_rows(R, K, Cout, bias != nullptr, true, sm_count(), tuning(TUNE_STATS_GEMM), 0).variant == 1) {
        const int rc = tc::launch_rows_pair<240, 6, true>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, sums, kc, st);
        if (rc != tc::PAIR_LAUNCH_REFUSED) return rc;
    }
    switch (tuning(TUNE_STATS_GEMM)) {
        case 2: return tc::launch_rows<240, 3, true, false>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, sums, kc, st);
        case 3: return tc::launch_rows<240, 3, true, true>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, sums, kc, st);
        default: return tc::launch_rows<240, 4, true, false>(X, ldx, W, ldw, Y, ldy, R, K, Cout, bias, ldbias, rows_per_sample, sums, kc, st);
    }
}

static bool fused_ok(const float* X, long long ldx, const float* W, long long ldw, long long R, int K, int C, const float* bias,
                     long long rps) {
    return R > 0 && R % 3 == 0 && R < (1ll << 31) && K >= 32 && (K & 3) == 0 && (ldx & 3) == 0 && (ldw & 3) == 0 && tc::aligned16(X) &&
           tc::aligned16(W) && C >= 128 && (C % 128) == 0 && (!bias || (rps > 0 && rps % 3 == 0));
}

// VNLinearLeakyReLU forward with the VN tail fused into the tcgen05 GEMM epilogue (no-grad / inference path).
//   Wcat [2C, K] = (W_feat ; W_dir) stacked, bias [B*3, 2C] per-sample rows or NULL, stat [2C] = (mean | invstd) or NULL (no BN).
// vnpcc_gemm_vn_stats: batch statistics of ||W_feat x + b_p|| (sums: 2C doubles, zeroed here) without writing anything else.
int vnpcc_gemm_vn_stats(const float* X, long long ldx, const float* Wcat, long long ldw, long long R, int K, int C, const float* bias,
                        long long ldbias, long long rows_per_sample, double* sums, void* stream) {
    if (!fused_ok(X, ldx, Wcat, ldw, R, K, C, bias, rows_per_sample)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    return tc::launch_fused<tc::MODE_STATS, false>(X, ldx, Wcat, ldw, nullptr, 0, R, K, C, bias, ldbias, rows_per_sample, nullptr, nullptr,
                                                   nullptr, 0.f, sums, st);
}

int vnpcc_gemm_vn_apply(const float* X, long long ldx, const float* Wcat, long long ldw, float* out, long long ldo, long long R, int K,
                        int C, const float* bias, long long ldbias, long long rows_per_sample, const float* stat, const float* gamma,
                        const float* beta, float ns, void* stream) {
    if (!fused_ok(X, ldx, Wcat, ldw, R, K, C, bias, rows_per_sample)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast_math_enabled())
        return tc::launch_fused<tc::MODE_APPLY, true>(X, ldx, Wcat, ldw, out, ldo, R, K, C, bias, ldbias, rows_per_sample, stat, gamma, beta,
                                                      ns, nullptr, st);
    return tc::launch_fused<tc::MODE_APPLY, false>(X, ldx, Wcat, ldw, out, ldo, R, K, C, bias, ldbias, rows_per_sample, stat, gamma, beta, ns,
                                                   nullptr, st);
}

// VNLinear(K -> C) followed by VNMaxPool(C) over groups of N consecutive points, arg-max only: Wcat [2C, K] = (W ; W_dir W),
// best: G*C u64 (zeroed here), decoded by vnpcc_vn_maxpool_decode.  Neither the [R, C] layer output nor its direction is stored.
int vnpcc_gemm_vn_pool(const float* X, long long ldx, const float* Wcat, long long ldw, long long R, int K, int C, long long N,
                       unsigned long long* best, void* stream) {
    if (!fused_ok(X, ldx, Wcat, ldw, R, K, C, nullptr, 0) || N <= 0 || R % (3 * N) != 0 || N >= (1ll << 31)) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)(R / (3 * N)) * C, st);
    return tc::launch_fused<tc::MODE_POOL, false>(X, ldx, Wcat, ldw, nullptr, 0, R, K, C, nullptr, 0, 3 * N, nullptr, nullptr, nullptr, 0.f,
                                                  reinterpret_cast<double*>(best), st);
}

// Backward of the fused tail VNLinearLeakyReLU(Cin -> C) -> VNLinear(C, 1) (+ residual) up to the layer input:
//   sums [2C] / gw2 [C] (fp64, zeroed here): BatchNorm backward sums (-> dgamma = sums[C:], dbeta = sums[:C]) and the tail weight gradient
//   gpd [P*3, 2C]: final gradient of the stacked linear output (p | d)   (input of the weight-gradient GEMM)
//   gh  [P*3, Cin]: gradient of the layer input = gpd Wcat, with Wt [Cin, 2C] = Wcat^T
// pd [P*3, 2C] is the stacked linear output saved by the forward.  C % 32 == 0, Cin in {128, 256}; VNPCC_ERR_UNSUPPORTED otherwise.
int vnpcc_tail_bwd_tf32(const float* gy, const float* pd, long long ldpd, long long P, int C, const float* stat, const float* gamma,
                        const float* beta, float ns, const float* w2, const float* Wt, long long ldwt, int Cin, int training,
                        double* sums, double* gw2, float* gpd, long long ldgpd, float* gh, long long ldgh, const float* h, long long ldh,
                        float* gW, long long ldgw, void* stream) {
    if (P <= 0) return 0;
    if (C <= 0 || (C & 31) || C > tc::TB_MAX_C || (Cin != 128 && Cin != 256) || !stat || !gamma || !beta || !w2 || (ldpd & 3) || (ldwt & 3) ||
        !tc::aligned16(pd) || !tc::aligned16(Wt) || P * 3 >= (1ll << 31))
        return VNPCC_ERR_UNSUPPORTED;
    const bool fused_w = gW != nullptr && tuning(TUNE_TAIL_WGRAD) != 1;
    if (fused_w && ((C & 127) || !h || (ldh & 3) || !tc::aligned16(h))) return VNPCC_ERR_UNSUPPORTED;
    if (!fused_w && (!gpd || (ldgpd & 3) || !tc::aligned16(gpd))) return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    cudaMemsetAsync(gw2, 0, sizeof(double) * C, st);
    if (!try_bn_leaky_dot_sums_v4(gy, pd, ldpd, pd + C, ldpd, P, C, stat, gamma, beta, ns, sums, w2, gw2, st)) return VNPCC_ERR_UNSUPPORTED;
    CUtensorMap mwt, mpd, mgpd;
    if (!tc::make_map(&mwt, Wt, Cin, 2 * C, ldwt, tc::BK, tc::BM)) return VNPCC_ERR_DRIVER;
    if (!tc::make_map(&mpd, pd, P * 3, 2 * C, ldpd, tc::BK, tc::TB_BN)) return VNPCC_ERR_DRIVER;
    if (fused_w) mgpd = mpd;
    else if (!tc::make_map(&mgpd, gpd, P * 3, 2 * C, ldgpd, tc::BK, tc::TB_BN)) return VNPCC_ERR_DRIVER;
    static bool attr_done_dev[64] = {false};
    bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (cudaFuncSetAttribute(tc::tail_dgrad_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailSmem::TOTAL) != cudaSuccess ||
            cudaFuncSetAttribute(tc::tail_dgrad_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailSmem::TOTAL) != cudaSuccess ||
            cudaFuncSetAttribute(tc::tail_wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailWSmem::TOTAL) != cudaSuccess)
            return last_error();
        attr_done = true;
    }
    bool launched = false;
    if (Cin == 256 && tuning(TUNE_TAIL_WGRAD) != 3 && P >= 64 * (long long)(sm_count() / 2)) {
        // CTA pairs (see the kernel's CTA2 note): 192-row pair tiles, one cluster of two CTAs per SM pair
        static bool attr_p[64] = {false};
        bool& done_p = attr_p[current_device_slot()];
        auto k1 = tc::tail_dgrad_tf32_kernel<true, true>;
        auto k0 = tc::tail_dgrad_tf32_kernel<false, true>;
        if (!done_p) {
            if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailSmemPair::TOTAL) != cudaSuccess ||
                cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailSmemPair::TOTAL) != cudaSuccess)
                return last_error();
            done_p = true;
        }
        const long long tiles = (P + 63) / 64;
        long long pairs = sm_count() / 2;
        if (pairs > tiles) pairs = tiles;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(tc::TB_THREADS);
        cfg.dynamicSmemBytes = tc::TailSmemPair::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        count_launch();
        const cudaError_t e = cudaLaunchKernelEx(&cfg, fused_w ? k0 : k1, mwt, mpd, mgpd, gy, P, C, Cin, stat, gamma, beta, ns, w2,
                                                 (const double*)sums, (double)P, training, gh, (size_t)ldgh, tiles);
        if (e == cudaSuccess) launched = true;
        else (void)cudaGetLastError();
    }
    const long long num_tiles = (P + 31) / 32;
    const int grid = (int)(num_tiles < sm_count() ? num_tiles : sm_count());
    if (launched) {
    } else if (fused_w)
        count_launch(), tc::tail_dgrad_tf32_kernel<false><<<grid, tc::TB_THREADS, tc::TailSmem::TOTAL, st>>>(
            mwt, mpd, mgpd, gy, P, C, Cin, stat, gamma, beta, ns, w2, sums, (double)P, training, gh, (size_t)ldgh, num_tiles);
    else
        count_launch(), tc::tail_dgrad_tf32_kernel<true><<<grid, tc::TB_THREADS, tc::TailSmem::TOTAL, st>>>(
            mwt, mpd, mgpd, gy, P, C, Cin, stat, gamma, beta, ns, w2, sums, (double)P, training, gh, (size_t)ldgh, num_tiles);
    if (fused_w) {
        // weight gradient of the stacked weight [2C, Cin] with the gradient formed on the fly: (C / 128) channel groups x row splits = one wave
        CUtensorMap mpd2, mh;
        if (!tc::make_map_slabs(&mpd2, pd, P * 3, 2 * C, ldpd, tc::TW_BR, 4, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return VNPCC_ERR_DRIVER;
        if (!tc::make_map_slabs(&mh, h, P * 3, Cin, ldh, tc::TW_BR, Cin / 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return VNPCC_ERR_DRIVER;
        const int groups = C / 128;
        long long splits = sm_count() / groups;
        if (splits < 1) splits = 1;
        long long pps = ((P + splits - 1) / splits + 7) / 8 * 8;      // whole stages of 8 points
        splits = (P + pps - 1) / pps;
        cudaMemset2DAsync(gW, (size_t)ldgw * sizeof(float), 0, (size_t)Cin * sizeof(float), (size_t)2 * C, st);
        count_launch(), tc::tail_wgrad_tf32_kernel<<<dim3(groups, (unsigned)splits), tc::TB_THREADS, tc::TailWSmem::TOTAL, st>>>(
            mpd2, mh, gy, P, C, Cin, stat, gamma, beta, ns, w2, sums, (double)P, training, gW, (size_t)ldgw, pps);
    }
    return last_error();
}

size_t vnpcc_gemm_wgrad_tf32_workspace_bytes(long long, int, int) { return 0; }

// split of the weight-gradient reduction over CTAs: one CTA per SM is resident (192 KB of operand stages), so the grid should be a whole
// number of waves of `slots` CTAs -- 32 output tiles x 10 splits = 320 CTAs would run 3 waves for 2.16 waves of work.  Cost model per
// candidate split count: waves x (rows per split + ~512 rows' worth of prologue / red.add epilogue).
static void plan_wgrad_splits(long long R, long long tiles, long long slots, long long* splits_out, long long* rows_per_split_out) {
    const long long max_splits = (R + 1023) / 1024;
    long long splits = 1, best_cost = -1;
    if (tuning(TUNE_GRID_LEGACY)) {
        splits = (slots * 2 + tiles - 1) / tiles;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    } else {
        long long hi = (4 * slots + tiles - 1) / tiles;
        if (hi > max_splits) hi = max_splits;
        if (hi < 1) hi = 1;
        for (long long c = 1; c <= hi; ++c) {
            const long long rps = ((R + c - 1) / c + tc::BR - 1) / tc::BR * tc::BR;
            const long long actual = (R + rps - 1) / rps;
            const long long waves = (tiles * actual + slots - 1) / slots;
            const long long cost = waves * (rps + 512);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                splits = actual;
            }
        }
    }
    const long long rows_per_split = ((R + splits - 1) / splits + tc::BR - 1) / tc::BR * tc::BR;
    *rows_per_split_out = rows_per_split;
    *splits_out = (R + rows_per_split - 1) / rows_per_split;
}

// host-logic introspection (tests/test_planners_cpu.py): out = {grid.x, grid.y, splits, rows per split}
void vnpcc_debug_wgrad_plan(long long R, int Cout, int K, int sms, long long* out) {
    const int gm = (Cout + tc::BM - 1) / tc::BM, gn = (K + 255) / 256;
    long long splits, rps;
    plan_wgrad_splits(R, (long long)gm * gn, sms, &splits, &rps);
    out[0] = gm;
    out[1] = gn;
    out[2] = splits;
    out[3] = rps;
}

int vnpcc_gemm_wgrad_tf32(const float* dY, long long lddy, const float* X, long long ldx, float* G, long long ldg, long long R,
                          int Cout, int K, float* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace;
    (void)workspace_bytes;
    if (Cout <= 0 || K <= 0) return 0;
    // R < 256 (the per-sample heads, 96 rows): refused -- measured 23-26 us here (the tile leaves through per-element red.add) against
    // 17 + 5 us for the SIMT kernel and its clear (tools/small_gemm_bench.py under ncu)
    if ((lddy & 3) || (ldx & 3) || !tc::aligned16(dY) || !tc::aligned16(X) || R >= (1ll << 31) || R < 256 || Cout < 32 || K < 32 ||
        (Cout & 3) || (K & 3))
        return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BNW = 256, STAGES = 4;
    using L = tc::WgradSmem<BNW, STAGES>;
    CUtensorMap mdy, mx;
    // boxes of {32 channels (inner, 128 bytes), BR rows}
    if (!tc::make_map(&mdy, dY, R, Cout, lddy, 32, tc::BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return VNPCC_ERR_DRIVER;
    if (!tc::make_map(&mx, X, R, K, ldx, 32, tc::BR, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return VNPCC_ERR_DRIVER;
    auto kern = tc::gemm_wgrad_tf32_kernel<BNW, STAGES>;
    static bool attr_done_dev[64] = {false};      // the attribute is per device
    bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return last_error();
        attr_done = true;
    }
    const int gm = (Cout + tc::BM - 1) / tc::BM, gn = (K + BNW - 1) / BNW;
    long long splits, rows_per_split;
    plan_wgrad_splits(R, (long long)gm * gn, sm_count(), &splits, &rows_per_split);
    // G is accumulated with atomics: clear it first (rows of K floats with pitch ldg)
    cudaMemset2DAsync(G, (size_t)ldg * sizeof(float), 0, (size_t)K * sizeof(float), (size_t)Cout, st);
    count_launch(), kern<<<dim3(gm, gn, (unsigned)splits), tc::NUM_THREADS, L::TOTAL, st>>>(mdy, mx, G, (size_t)ldg, R, Cout, K,
                                                                                          rows_per_split);
    return last_error();
}

}  // extern "C"
