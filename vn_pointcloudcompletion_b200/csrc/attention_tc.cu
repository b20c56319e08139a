// attention_tc.cu -- tensor-core (tcgen05 / TMEM, TF32 operands, fp32 accumulation) forward of the VN multi-head attention core
// (models/transformer.py:89-100), throughput mode.  Same contract and ROW layout as vnpcc_vn_attention_fwd (attention.cu).
//
// One CTA = 128 query tokens of one (sample, head) = the 128 TMEM lanes; keys / values stream through in tiles of 64.
//   S [128 q x 64 keys]  = Q K^T     tcgen05.mma kind::tf32, M128 N64,  K = 3 x 64 (a head's [48, 3] feature, each component's
//                                    48 channels zero-padded to 64 so that it fills two 128-byte swizzle rows)
//   O [128 q x 192]     += P V       M128 N192, K = 64 keys; P is written by the softmax threads (one thread = one query row = one
//                                    TMEM lane, so the row maximum and row sum are thread-local) into shared memory in the
//                                    K-major SWIZZLE_128B operand layout, V is staged transposed ([feature, key]) in the same layout
// Softmax is two-pass (pass 1: row maxima from S alone; pass 2: P = exp(S - m), O += P V with the accumulator resident in TMEM),
// which costs a second Q K^T but needs no accumulator rescaling.
// Operand staging: TMA, 3-D tensor maps over qkv viewed as [token][component][column].  A box of 32 columns x 1 component x T tokens is
// exactly one k-block (T rows x 128 bytes) of the K-major SWIZZLE_128B layout; the second box of a component covers channels 32..63
// of the head, i.e. 16 real channels and 16 that belong to the NEXT head.  Q's copies of those 16 columns are zeroed after its (single)
// load, so whatever K holds there is multiplied by zero; V's extra columns only produce output columns that are never read.  V is the
// MN-major B operand of the second MMA (feature contiguous): boxes with the 128B_ATOM_32B swizzle <-> UMMA SWIZZLE_128B_BASE32B, the
// pairing the weight-gradient GEMM (gemm_tcgen05.cu) uses.  One thread issues TMA and MMA in an order that keeps the tensor pipe fed (see
// attn_fwd_tc_kernel): pass 1 double-buffers K and S, pass 2 queues Q K^T of the next tile right behind P V; the stored P leaves as
// bulk-tensor stores of the swizzled operand tile.  The backward (stored-P route) is three streaming GEMMs (attn_acc_gemm_kernel, separate
// TMA-producer and MMA-issuer threads) around attn_ds_tma_kernel, which turns P into dS in place with P / dS moving by TMA.
// Every mbarrier wait is bounded: a protocol error traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "vnpcc.h"
#include "vnpcc_internal.h"

namespace vnpcc {
namespace atc {

constexpr int BQ = 128;         // queries per CTA (TMEM lanes)
constexpr int BKEY = 64;        // keys per tile
constexpr int DP = 64;          // padded channels per component
constexpr int KD = 3 * DP;      // padded head feature = MMA K of S, MMA N of O
constexpr int NT = 128;
constexpr uint32_t SPIN_LIMIT = 1u << 22;

constexpr int Q_BYTES = (KD / 32) * BQ * 128;        // 98304
constexpr int K_BYTES = (KD / 32) * BKEY * 128;      // 49152
constexpr int VT_BYTES = (BKEY / 32) * KD * 128;     // 49152
constexpr int P_BYTES = (BKEY / 32) * BQ * 128;      // 32768
constexpr int TILE_BYTES = Q_BYTES + K_BYTES + VT_BYTES + P_BYTES;
constexpr int SMEM_BYTES = TILE_BYTES + 128 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
    __trap();
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
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
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): LBO 16 B, SBO 1024 B (8 rows x 128 B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(16 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major SWIZZLE_128B_BASE32B descriptor: slabs of {32 MN elements = 128 B} x rows (reduction index); the swizzle atom is 4 rows, one
// K = 8 instruction spans two atoms (SBO = 512 B); LBO = bytes between consecutive 32-element slabs along M / N
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// instruction descriptor: c_format F32 [4,6), a/b format TF32 [7,10)/[10,13), a_major [15], b_major [16] (0 = K-major, 1 = MN-major),
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int b_mn_major = 0, int a_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row r, k index kappa) inside a K-major SWIZZLE_128B tile of `rows` rows: k-blocks of 32 floats (one 128-byte
// swizzle row each), 8-row groups of 1024 bytes, 16-byte chunk index XOR (row mod 8)
__device__ __forceinline__ uint32_t sw128(int rows, int r, int kappa) {
    return (uint32_t)((kappa >> 5) * (rows * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((((kappa & 31) >> 2) ^ (r & 7)) << 4) + (kappa & 3) * 4);
}

// MODE_FWD: rows = queries (resident Q), streamed K / V:   out = softmax(scale Q K^T) V, lse written
// MODE_DV : rows = keys (resident K), streamed Q / dO:      dV = P^T dO with P^T[key, q] = exp(scale K Q^T - lse[q]) (lse read); the
//           same skeleton with the operands' roles swapped, one pass, no normalisation
constexpr int MODE_FWD = 0, MODE_DV = 1;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// bulk-tensor store of one K-major SWIZZLE_128B k-block ([rows] x 128 bytes in shared memory) to a row-major global matrix
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Issue order (one thread issues TMA and MMA; tcgen05 MMAs of one thread execute in order, a commit covers everything issued before it):
//   pass 1 (MODE_FWD)  K tiles alternate between the K and the (still unused) V buffer, S alternates between two TMEM accumulators:
//                      Q K^T of tile i+1 runs while the lane threads reduce tile i to (max, sum); the tile after that is already in flight.
//   pass 2             P V of tile i and Q K^T of tile i+1 are issued back to back as soon as P(i) is in shared memory; the K tile was
//                      fetched during the softmax, the next V tile is fetched as soon as P V has read the current one (during Q K^T and
//                      the next softmax).  The stored P goes out as bulk-tensor stores straight from the swizzled operand tile.
template <int D, int MODE>
__global__ void __launch_bounds__(NT, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_q,
                                                           const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_p,
                                                           int p_tma, int N, int H, int cx, int cy, int cz, float scale,
                                                           float* __restrict__ out, size_t ldo, float* __restrict__ lse,
                                                           float* __restrict__ p_out) {
    // p_out (MODE_FWD, may be NULL; N % 4 == 0): the normalised attention weights P [B*H*N, N], stored for the GEMM-shaped backward;
    // p_tma != 0 (needs N % 128 == 0): map_p describes p_out as a row-major matrix with boxes {32 columns, 128 rows}, SWIZZLE_128B
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + Q_BYTES;
    uint8_t* Vs = Ks + K_BYTES;
    uint8_t* Ps = Vs + VT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TILE_BYTES);
    uint64_t* qfull = bars, *kfull = bars + 1 /* [2] */, *vfull = bars + 3, *s_done = bars + 4 /* [2] */, *o_done = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int q0 = blockIdx.x * BQ;
    const int tok0 = b * N;                       // first token of the sample in the [B*N]-token tensor maps
    const int colq = cx + h * D, colk = cy + h * D, colv = cz + h * D;      // resident rows / streamed K-major tile / streamed MN-major tile
    const int T = (N + BKEY - 1) / BKEY;
    static_assert(K_BYTES == VT_BYTES, "pass 1 uses the V buffer as the second K buffer");

    if (tid == 0) {
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_v);
        if (p_tma) tma_prefetch_desc(&map_p);
        for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);      // this thread's TMEM lane = its query row
    const uint32_t o_tmem = tmem_base + 2 * BKEY;                          // S accumulators at columns 0 and BKEY, O behind them
    constexpr uint32_t idesc_s = make_idesc(BQ, BKEY, 0);
    constexpr uint32_t idesc_o = make_idesc(BQ, KD, 1);

    auto load_k = [&](int k0, uint8_t* dst, uint64_t* bar) {      // 6 boxes {32 cols, 1 component, 64 tokens} -> 6 k-blocks of 8 KB
        mbar_expect_tx(bar, K_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb) tma_load_3d(&map_k, bar, dst + kb * (BKEY * 128), colk + (kb & 1) * 32, kb >> 1, tok0 + k0);
    };
    auto load_v = [&](int k0) {      // 6 slabs {32 features, 64 keys}
        mbar_expect_tx(vfull, VT_BYTES);
#pragma unroll
        for (int sl = 0; sl < KD / 32; ++sl) tma_load_3d(&map_v, vfull, Vs + sl * (BKEY * 128), colv + (sl & 1) * 32, sl >> 1, tok0 + k0);
    };
    auto issue_s = [&](const uint8_t* kbuf, uint32_t s_tmem, uint64_t* done) {
        const uint32_t qa = smem_u32(Qs), ka = smem_u32(kbuf);
#pragma unroll
        for (int ks = 0; ks < KD / 8; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            if ((kb & 1) && kk * 8 + 32 >= D) continue;      // columns D..63 of a component belong to the next head: those k-steps are skipped
            umma_tf32(s_tmem, make_desc(qa + kb * (BQ * 128) + kk * 32), make_desc(ka + kb * (BKEY * 128) + kk * 32), idesc_s, ks != 0 ? 1u : 0u);
        }
        if (done) umma_commit(done);
    };

    // ---- resident tile: one TMA load (its padding columns are never multiplied: see issue_s)
    if (tid == 0) {
        mbar_expect_tx(qfull, Q_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb) tma_load_3d(&map_q, qfull, Qs + kb * (BQ * 128), colq + (kb & 1) * 32, kb >> 1, tok0 + q0);
    }
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const float sc2 = scale * LOG2E;      // scores are handled in the base-2 domain: exp(x) = 2^(x log2 e) is one MUFU.EX2

    // ---- pass 1: row maxima and row sums of the scaled scores (tile t: K buffer t & 1, S accumulator t & 1, barrier parity (t >> 1) & 1)
    float m = -INFINITY, l1 = 0.f;
    if (MODE == MODE_FWD) {
        if (tid == 0) {
            load_k(0, Ks, &kfull[0]);
            if (T > 1) load_k(BKEY, Vs, &kfull[1]);
            mbar_wait(qfull, 0);
            mbar_wait(&kfull[0], 0);
            tc_fence_after();
            issue_s(Ks, tmem_base, &s_done[0]);
        }
        for (int i = 0; i < T; ++i) {
            const int k0 = i * BKEY, cur = i & 1;
            if (tid == 0 && i + 1 < T) {      // accumulator cur ^ 1 was released by the __syncthreads of iteration i - 1
                mbar_wait(&kfull[cur ^ 1], (uint32_t)(((i + 1) >> 1) & 1));
                tc_fence_after();
                issue_s(cur ? Ks : Vs, tmem_base + (cur ^ 1) * BKEY, &s_done[cur ^ 1]);
            }
            mbar_wait(&s_done[cur], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            if (tid == 0 && i + 2 < T) load_k(k0 + 2 * BKEY, cur ? Vs : Ks, &kfull[cur]);      // K buffer cur is free again
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float s[32];
                tmem_ld32(t_row + cur * BKEY + half * 32, s);
                float hm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    s[j] = (k0 + half * 32 + j < N) ? s[j] * sc2 : -INFINITY;
                    hm = fmaxf(hm, s[j]);
                }
                if (hm > -INFINITY) {      // online (max, sum): only the scalar l is rescaled
                    const float mn = fmaxf(m, hm);
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        acc0 += ex2f(s[j] - mn);
                        acc1 += ex2f(s[j + 1] - mn);
                    }
                    l1 = fmaf(l1, ex2f(m - mn), acc0 + acc1);
                    m = mn;
                }
            }
            tc_fence_before();
            __syncthreads();      // every thread has read accumulator cur before the Q K^T after next overwrites it
        }
    }
    const float lse2_row = m + log2f(l1);      // MODE_FWD: P = 2^(S - lse2) is normalised, O needs no final division

    // ---- pass 2: P = exp(S - lse), O += P V   (K buffer = Ks, accumulator 0; their barriers continue with the parity pass 1 left)
    const uint32_t par0 = MODE == MODE_FWD ? (uint32_t)(((T + 1) >> 1) & 1) : 0u;
    uint32_t kph = par0, sph = par0, vph = 0, oph = 0;
    if (tid == 0) {
        load_k(0, Ks, &kfull[0]);
        load_v(0);
        if (MODE != MODE_FWD) mbar_wait(qfull, 0);
        mbar_wait(&kfull[0], kph);
        tc_fence_after();
        issue_s(Ks, tmem_base, &s_done[0]);
    }
    kph ^= 1;
    for (int i = 0; i < T; ++i) {
        const int k0 = i * BKEY;
        mbar_wait(&s_done[0], sph);      // S(i) complete -- and with it (in-order execution) P V of tile i - 1: P and the K buffer are free
        sph ^= 1;
        tc_fence_after();
        if (tid == 0 && i + 1 < T) load_k(k0 + BKEY, Ks, &kfull[0]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float s[32];
            tmem_ld32(t_row + half * 32, s);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = k0 + half * 32 + j;
                float pv = 0.f;
                if (col < N) pv = ex2f(fmaf(s[j], sc2, -(MODE == MODE_FWD ? lse2_row : LOG2E * __ldg(lse + (size_t)bh * N + col))));
                s[j] = pv;
            }
            if (MODE == MODE_FWD && p_out != nullptr && !p_tma && q0 + tid < N) {
                float4* dst = reinterpret_cast<float4*>(p_out + ((size_t)bh * N + q0 + tid) * N + k0 + half * 32);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    if (k0 + half * 32 + j4 * 4 < N) dst[j4] = make_float4(s[j4 * 4], s[j4 * 4 + 1], s[j4 * 4 + 2], s[j4 * 4 + 3]);
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(Ps + sw128(BQ, tid, half * 32 + j4 * 4)) = make_float4(s[j4 * 4], s[j4 * 4 + 1], s[j4 * 4 + 2], s[j4 * 4 + 3]);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();      // P complete (and S consumed)
        if (tid == 0) {
            mbar_wait(vfull, vph);
            tc_fence_after();
            const uint32_t pa = smem_u32(Ps), va = smem_u32(Vs);
#pragma unroll
            for (int ks = 0; ks < BKEY / 8; ++ks) {
                const int kb = ks >> 2, kk = ks & 3;
                umma_tf32(o_tmem, make_desc(pa + kb * (BQ * 128) + kk * 32), make_desc_mn(va + ks * 1024, BKEY * 128), idesc_o, (i | ks) != 0 ? 1u : 0u);
            }
            umma_commit(o_done);
            if (i + 1 < T) {      // the next Q K^T queues up right behind P V
                mbar_wait(&kfull[0], kph);
                tc_fence_after();
                issue_s(Ks, tmem_base, nullptr);
            }
            if (MODE == MODE_FWD && p_tma) {      // the stored P: two bulk-tensor stores out of the operand tile (rows q0.., columns k0..)
                tma_store_2d(&map_p, Ps, k0, bh * N + q0);
                tma_store_2d(&map_p, Ps + BQ * 128, k0 + 32, bh * N + q0);
                tma_store_commit();
                tma_store_wait_read();      // P may be overwritten once s_done fires: it must have been read by then
            }
            if (i + 1 < T) umma_commit(&s_done[0]);
            mbar_wait(o_done, oph);         // P V has read V: fetch the next tile behind Q K^T and the next softmax
            if (i + 1 < T) load_v(k0 + BKEY);
        }
        kph ^= 1;
        vph ^= 1;
        oph ^= 1;
    }
    if (tid == 0 && MODE == MODE_FWD && p_tma) tma_store_wait_all();
    // all MMAs are complete for thread 0; make that visible to everyone before the accumulator is read
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- epilogue: out = O (P is normalised in MODE_FWD and must not be in MODE_DV), lse = ln 2 (m + log2 l)
    const int n = q0 + tid;
#pragma unroll 1
    for (int i = 0; i < KD / 32; ++i) {
        float o[32];
        tmem_ld32(t_row + 2 * BKEY + i * 32, o);
        if (n < N) {
            const int v = i >> 1, c0 = (i & 1) * 32;
            float* dst = out + ((size_t)(b * (size_t)N + n) * 3 + v) * ldo + (size_t)h * D + c0;
            const int cnt = (c0 + 32 <= D) ? 32 : (D - c0);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < cnt) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
    if (MODE == MODE_FWD && n < N) lse[(size_t)bh * N + n] = lse2_row * LN2;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// -------------------------------------------------------------------------------------------------------------------------------
// backward, dQ and dK:  dS = P o (dP - delta) scale,  dQ = dS K,  dK = dS^T Q       (P = exp(scale Q K^T - lse), dP = dO V^T)
// One kernel, two orientations (BMODE):
//   BMODE_DQ: lanes = 128 queries.  resident X = Q, streamed tiles of 32 keys Y = K;  dP = G H^T with G = dO (rows = the queries), H = V
//   BMODE_DK: lanes = 128 keys.     resident X = K, streamed tiles of 32 queries Y = Q; dP^T = G H^T with G = V (rows = the keys), H = dO
// Per tile:  S = X Y^T (M128 N32) and dP = G H^T (M128 N32; G and H stream through a ring of 32-column k-blocks because a second resident
// 96 KB tile does not fit) -> the threads (one per lane) form dS row-locally and store it as the K-major A operand -> acc[128 x 192] +=
// dS Y with Y re-staged MN-major (M128 N192).  The accumulator stays in TMEM for the whole kernel and is stored once: no atomics.
// -------------------------------------------------------------------------------------------------------------------------------
constexpr int BMODE_DQ = 0, BMODE_DK = 1;
constexpr int BT = 32;                               // streamed tokens per tile
constexpr int NS = 3;                                // ring stages of the dP operands
constexpr int BX_BYTES = (KD / 32) * BQ * 128;       // 98304  resident
constexpr int BY_BYTES = (KD / 32) * BT * 128;       // 24576  streamed, K-major
constexpr int BYM_BYTES = (KD / 32) * BT * 128;      // 24576  streamed, MN-major
constexpr int BDS_BYTES = BQ * 128;                  // 16384  dS tile (one k-block of 32 tokens)
constexpr int BG_BYTES = BQ * 128;                   // 16384  one k-block of G
constexpr int BH_BYTES = BT * 128;                   // 4096   one k-block of H
constexpr int BSTAGE_BYTES = BG_BYTES + BH_BYTES;
constexpr int BTILE_BYTES = BX_BYTES + BY_BYTES + BYM_BYTES + BDS_BYTES + NS * BSTAGE_BYTES;
constexpr int BSMEM_BYTES = BTILE_BYTES + 256 + 1024;

template <int D, int BMODE>
__global__ void __launch_bounds__(NT, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_x,      // qkv, K-major SW128, box 128 tokens  (resident X; G = V in BMODE_DK)
                   const __grid_constant__ CUtensorMap map_y,      // qkv, K-major SW128, box 32 tokens   (Y; H = V in BMODE_DQ)
                   const __grid_constant__ CUtensorMap map_ym,     // qkv, MN-major ATOM_32B, box 32 tokens (Y for the last product)
                   const __grid_constant__ CUtensorMap map_do128,  // dO, K-major SW128, box 128 tokens   (G = dO in BMODE_DQ)
                   const __grid_constant__ CUtensorMap map_do32,   // dO, K-major SW128, box 32 tokens    (H = dO in BMODE_DK)
                   int N, int H, int C, float scale, const float* __restrict__ lse, const float* __restrict__ delta,
                   float* __restrict__ dqkv, size_t lddq, float* __restrict__ ds_out) {
    // ds_out (BMODE_DQ only, may be NULL): dS [B*H*N, N] row-major, written tile by tile so that dK = dS^T Q can run as a plain streaming
    // GEMM (attn_acc_gemm_kernel) instead of recomputing S and dP in the other orientation
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint8_t* Xs = smem;
    uint8_t* Ys = Xs + BX_BYTES;
    uint8_t* Ym = Ys + BY_BYTES;
    uint8_t* dSs = Ym + BYM_BYTES;
    uint8_t* ring = dSs + BDS_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BTILE_BYTES);
    uint64_t* xfull = bars, *yfull = bars + 1, *ymfull = bars + 2, *sdp_done = bars + 3, *acc_done = bars + 4;
    uint64_t* sfull = bars + 5, *sfree = bars + 5 + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * NS);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int r0 = blockIdx.x * BQ;               // first resident token (query or key) of this CTA
    const int tok0 = b * N;
    const int colq = h * D, colk = C + h * D, colv = 2 * C + h * D, coldo = h * D;
    const int colx = BMODE == BMODE_DQ ? colq : colk;      // resident
    const int coly = BMODE == BMODE_DQ ? colk : colq;      // streamed
    const int T = (N + BT - 1) / BT;
    constexpr int KBLK = KD / 32;

    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_y);
        tma_prefetch_desc(&map_ym);
        tma_prefetch_desc(&map_do128);
        tma_prefetch_desc(&map_do32);
        for (int i = 0; i < 5 + 2 * NS; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t s_tmem = tmem_base, dp_tmem = tmem_base + 32, acc_tmem = tmem_base + 64;
    constexpr uint32_t idesc_s = make_idesc(BQ, BT, 0);
    constexpr uint32_t idesc_acc = make_idesc(BQ, KD, 1);

    // producer / consumer state of the ring (thread 0 only)
    int n_loaded = 0, n_used = 0;
    auto ring_load = [&](int tile, int kb) {      // k-block kb of G (128 resident rows) and of H (the tile's 32 tokens)
        const int st = n_loaded % NS;
        if (n_loaded >= NS) mbar_wait(&sfree[st], (uint32_t)((n_loaded / NS - 1) & 1));
        uint8_t* g = ring + st * BSTAGE_BYTES;
        mbar_expect_tx(&sfull[st], BSTAGE_BYTES);
        if (BMODE == BMODE_DQ) {
            tma_load_3d(&map_do128, &sfull[st], g, coldo + (kb & 1) * 32, kb >> 1, tok0 + r0);
            tma_load_3d(&map_y, &sfull[st], g + BG_BYTES, colv + (kb & 1) * 32, kb >> 1, tok0 + tile * BT);
        } else {
            tma_load_3d(&map_x, &sfull[st], g, colv + (kb & 1) * 32, kb >> 1, tok0 + r0);
            tma_load_3d(&map_do32, &sfull[st], g + BG_BYTES, coldo + (kb & 1) * 32, kb >> 1, tok0 + tile * BT);
        }
        ++n_loaded;
    };
    auto load_y = [&](int tile) {
        mbar_expect_tx(yfull, BY_BYTES);
#pragma unroll
        for (int kb = 0; kb < KBLK; ++kb) tma_load_3d(&map_y, yfull, Ys + kb * (BT * 128), coly + (kb & 1) * 32, kb >> 1, tok0 + tile * BT);
    };
    auto load_ym = [&](int tile) {
        mbar_expect_tx(ymfull, BYM_BYTES);
#pragma unroll
        for (int sl = 0; sl < KBLK; ++sl) tma_load_3d(&map_ym, ymfull, Ym + sl * (BT * 128), coly + (sl & 1) * 32, sl >> 1, tok0 + tile * BT);
    };

    if (tid == 0) {
        mbar_expect_tx(xfull, BX_BYTES);
#pragma unroll
        for (int kb = 0; kb < KBLK; ++kb) tma_load_3d(&map_x, xfull, Xs + kb * (BQ * 128), colx + (kb & 1) * 32, kb >> 1, tok0 + r0);
        load_y(0);
        load_ym(0);
        for (int kb = 0; kb < NS; ++kb) ring_load(0, kb);
        mbar_wait(xfull, 0);
    }

    // per-lane scalars (BMODE_DQ: this lane's query)
    const int n_row = r0 + tid;
    float lse_r = 0.f, del_r = 0.f;
    if (BMODE == BMODE_DQ && n_row < N) {
        lse_r = __ldg(lse + (size_t)bh * N + n_row);
        del_r = __ldg(delta + (size_t)bh * N + n_row);
    }

    uint32_t yph = 0, ymph = 0, sdph = 0, accph = 0;
    for (int i = 0; i < T; ++i) {
        const int c0 = i * BT;
        if (tid == 0) {
            // S = X Y^T
            mbar_wait(yfull, yph);
            tc_fence_after();
            const uint32_t xa = smem_u32(Xs), ya = smem_u32(Ys);
#pragma unroll
            for (int ks = 0; ks < KD / 8; ++ks) {
                const int kb = ks >> 2, kk = ks & 3;
                if ((kb & 1) && kk * 8 + 32 >= D) continue;
                umma_tf32(s_tmem, make_desc(xa + kb * (BQ * 128) + kk * 32), make_desc(ya + kb * (BT * 128) + kk * 32), idesc_s, ks != 0 ? 1u : 0u);
            }
            // dP = G H^T through the ring
            for (int kb = 0; kb < KBLK; ++kb) {
                const int st = n_used % NS;
                mbar_wait(&sfull[st], (uint32_t)((n_used / NS) & 1));
                tc_fence_after();
                const uint32_t ga = smem_u32(ring + st * BSTAGE_BYTES), ha = ga + BG_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if ((kb & 1) && kk * 8 + 32 >= D) continue;
                    umma_tf32(dp_tmem, make_desc(ga + kk * 32), make_desc(ha + kk * 32), idesc_s, (kb | kk) != 0 ? 1u : 0u);
                }
                umma_commit(&sfree[st]);
                ++n_used;
                // keep the ring full: the next k-blocks of this tile, then the first ones of the next tile
                const int nxt = kb + NS;
                if (nxt < KBLK)
                    ring_load(i, nxt);
                else if (i + 1 < T)
                    ring_load(i + 1, nxt - KBLK);
            }
            umma_commit(sdp_done);
        }
        yph ^= 1;
        mbar_wait(sdp_done, sdph);
        sdph ^= 1;
        tc_fence_after();
        if (tid == 0 && i + 1 < T) load_y(i + 1);      // the K-major Y buffer is free: S is complete
        {
            float sv[32], dv[32];
            tmem_ld32(t_row, sv);
            tmem_ld32(t_row + 32, dv);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = c0 + j;
                float ds = 0.f;
                if (col < N) {
                    const float L = BMODE == BMODE_DQ ? lse_r : __ldg(lse + (size_t)bh * N + col);
                    const float dl = BMODE == BMODE_DQ ? del_r : __ldg(delta + (size_t)bh * N + col);
                    const float p = expf(sv[j] * scale - L);
                    ds = p * (dv[j] - dl) * scale;
                }
                sv[j] = ds;
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(dSs + sw128(BQ, tid, j4 * 4)) = make_float4(sv[j4 * 4], sv[j4 * 4 + 1], sv[j4 * 4 + 2], sv[j4 * 4 + 3]);
            if (BMODE == BMODE_DQ && ds_out != nullptr && n_row < N) {      // host guarantees N % 32 == 0 on this path
                float4* dst = reinterpret_cast<float4*>(ds_out + ((size_t)bh * N + n_row) * N + c0);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) dst[j4] = make_float4(sv[j4 * 4], sv[j4 * 4 + 1], sv[j4 * 4 + 2], sv[j4 * 4 + 3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();      // dS complete, S / dP consumed
        if (tid == 0) {
            mbar_wait(ymfull, ymph);
            tc_fence_after();
            const uint32_t da = smem_u32(dSs), ma = smem_u32(Ym);
#pragma unroll
            for (int ks = 0; ks < BT / 8; ++ks)
                umma_tf32(acc_tmem, make_desc(da + ks * 32), make_desc_mn(ma + ks * 1024, BT * 128), idesc_acc, (i | ks) != 0 ? 1u : 0u);
            umma_commit(acc_done);
            mbar_wait(acc_done, accph);      // dS and the MN-major Y buffer are free again
            if (i + 1 < T) load_ym(i + 1);
        }
        ymph ^= 1;
        accph ^= 1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- epilogue: the accumulator rows are this lane's dQ (or dK)
    const int part = BMODE == BMODE_DQ ? 0 : C;
#pragma unroll 1
    for (int i = 0; i < KD / 32; ++i) {
        float o[32];
        tmem_ld32(t_row + 64 + i * 32, o);
        if (n_row < N) {
            const int v = i >> 1, cc = (i & 1) * 32;
            float* dst = dqkv + ((size_t)(b * (size_t)N + n_row) * 3 + v) * lddq + part + (size_t)h * D + cc;
            const int cnt = (cc + 32 <= D) ? 32 : (D - cc);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < cnt) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// -------------------------------------------------------------------------------------------------------------------------------
// Streaming GEMMs over a stored [B*H*N, N] score-shaped matrix M (P or dS), accumulator [128 lanes x 192] resident in TMEM, 4-stage TMA
// ring (160 KB in flight), one thread issues TMA and MMA.  N % 32 == 0.
//   A_MN = 1: lanes = 128 COLUMNS of M (keys):  acc[key] = sum_q M[q, key] * Z[q]     dK = dS^T Q,  dV = P^T dO
//             A = the MN-major [32 q x 128 keys] block of M (keys contiguous; 4 TMA boxes, 128B_ATOM_32B)
//   A_MN = 0: lanes = 128 ROWS of M (queries):  acc[q] = sum_k M[q, k] * Z[k]         dQ = dS K
//             A = the K-major [128 q x 32 keys] block of M (one SWIZZLE_128B k-block)
//   B = the MN-major tile of Z (32 tokens x 192 features, 6 TMA boxes out of the row layout): Z = Q, dO or K.
// -------------------------------------------------------------------------------------------------------------------------------
constexpr int GK_NS = 4;
constexpr int GK_A_BYTES = 4 * BT * 128;             // 16384 in both orientations
constexpr int GK_B_BYTES = (KD / 32) * BT * 128;     // 24576: 6 slabs of {32 features} x 32 tokens
constexpr int GK_STAGE_BYTES = GK_A_BYTES + GK_B_BYTES;
constexpr int GK_SMEM_BYTES = GK_NS * GK_STAGE_BYTES + 256 + 1024;

template <int D, int A_MN>
__global__ void __launch_bounds__(NT, 1) attn_acc_gemm_kernel(const __grid_constant__ CUtensorMap map_m,      // M: A_MN ? ATOM_32B box {32,32} : SW128 box {32,128}
                                                             const __grid_constant__ CUtensorMap map_z,      // Z rows, MN-major ATOM_32B, box 32 tokens
                                                             int N, int H, int colz0, float* __restrict__ outp, size_t ldo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GK_NS * GK_STAGE_BYTES);
    uint64_t* full = bars, *empty = bars + GK_NS, *done = bars + 2 * GK_NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GK_NS + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int r0 = blockIdx.x * BQ;               // first lane token (key or query) of this CTA
    const int tok0 = b * N;
    const int colz = colz0 + h * D;
    const int T = N / BT;
    if (tid == 0) {
        tma_prefetch_desc(&map_m);
        tma_prefetch_desc(&map_z);
        for (int i = 0; i < 2 * GK_NS + 1; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    // warp roles: thread 0 = TMA producer, thread 32 = MMA issuer (a single thread doing both has to wait for the MMAs of tile i before it
    // may refill their stage, which leaves the tensor pipe idle for a commit / wake-up round trip per tile)
    if (tid == 0) {
        for (int tile = 0; tile < T; ++tile) {
            const int st = tile % GK_NS;
            if (tile >= GK_NS) mbar_wait(&empty[st], (uint32_t)((tile / GK_NS - 1) & 1));
            uint8_t* a = smem + st * GK_STAGE_BYTES;
            mbar_expect_tx(&full[st], GK_STAGE_BYTES);
            if (A_MN) {
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) tma_load_2d(&map_m, &full[st], a + sl * (BT * 128), r0 + sl * 32, bh * N + tile * BT);
            } else {
                tma_load_2d(&map_m, &full[st], a, tile * BT, bh * N + r0);
            }
#pragma unroll
            for (int sl = 0; sl < KD / 32; ++sl)
                tma_load_3d(&map_z, &full[st], a + GK_A_BYTES + sl * (BT * 128), colz + (sl & 1) * 32, sl >> 1, tok0 + tile * BT);
        }
    } else if (tid == 32) {
        constexpr uint32_t idesc = make_idesc(BQ, KD, 1, A_MN);
        for (int i = 0; i < T; ++i) {
            const int st = i % GK_NS;
            mbar_wait(&full[st], (uint32_t)((i / GK_NS) & 1));
            tc_fence_after();
            const uint32_t aa = smem_u32(smem + st * GK_STAGE_BYTES), ba = aa + GK_A_BYTES;
#pragma unroll
            for (int ks = 0; ks < BT / 8; ++ks)
                umma_tf32(tmem_base, A_MN ? make_desc_mn(aa + ks * 1024, BT * 128) : make_desc(aa + ks * 32), make_desc_mn(ba + ks * 1024, BT * 128),
                          idesc, (i | ks) != 0 ? 1u : 0u);
            umma_commit(&empty[st]);
        }
        umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);      // every thread: the accumulator is complete
    tc_fence_after();
    const int n_row = r0 + tid;
#pragma unroll 1
    for (int i = 0; i < KD / 32; ++i) {
        float o[32];
        tmem_ld32(t_row + i * 32, o);
        if (n_row < N) {
            const int v = i >> 1, cc = (i & 1) * 32;
            float* dst = outp + ((size_t)(b * (size_t)N + n_row) * 3 + v) * ldo + (size_t)h * D + cc;
            const int cnt = (cc + 32 <= D) ? 32 : (D - cc);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < cnt) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// -------------------------------------------------------------------------------------------------------------------------------
// dS = P o (dP - delta) scale, IN PLACE over the stored P:  dP = dO V^T with dO (this CTA's 128 queries) resident and V streaming
// in double-buffered 64-key tiles; the dP accumulator is double-buffered in TMEM so that the MMAs of tile i+1 run while the
// lane threads turn tile i into dS.
// -------------------------------------------------------------------------------------------------------------------------------
constexpr int DSK_SMEM_BYTES = Q_BYTES + 2 * K_BYTES + 128 + 1024;

template <int D>
__global__ void __launch_bounds__(NT, 1) attn_ds_kernel(const __grid_constant__ CUtensorMap map_do,      // dO, K-major SW128, box 128 tokens
                                                       const __grid_constant__ CUtensorMap map_v,       // qkv, K-major SW128, box 64 tokens
                                                       int N, int H, int C, float scale, const float* __restrict__ delta, float* __restrict__ pds) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint8_t* dOs = smem;
    uint8_t* Vs = dOs + Q_BYTES;      // two buffers of K_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + 2 * K_BYTES);
    uint64_t* xfull = bars, *vfull = bars + 1, *dp_done = bars + 3;      // vfull[2], dp_done[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int q0 = blockIdx.x * BQ;
    const int tok0 = b * N;
    const int coldo = h * D, colv = 2 * C + h * D;
    const int T = (N + BKEY - 1) / BKEY;
    if (tid == 0) {
        tma_prefetch_desc(&map_do);
        tma_prefetch_desc(&map_v);
        for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr uint32_t idesc = make_idesc(BQ, BKEY, 0);
    auto load_v = [&](int tile) {
        const int bf = tile & 1;
        mbar_expect_tx(&vfull[bf], K_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb)
            tma_load_3d(&map_v, &vfull[bf], Vs + bf * K_BYTES + kb * (BKEY * 128), colv + (kb & 1) * 32, kb >> 1, tok0 + tile * BKEY);
    };
    auto issue_dp = [&](int tile) {
        const int bf = tile & 1;
        mbar_wait(&vfull[bf], (uint32_t)((tile >> 1) & 1));
        tc_fence_after();
        const uint32_t da = smem_u32(dOs), va = smem_u32(Vs + bf * K_BYTES);
#pragma unroll
        for (int ks = 0; ks < KD / 8; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            if ((kb & 1) && kk * 8 + 32 >= D) continue;
            umma_tf32(tmem_base + bf * BKEY, make_desc(da + kb * (BQ * 128) + kk * 32), make_desc(va + kb * (BKEY * 128) + kk * 32), idesc,
                      ks != 0 ? 1u : 0u);
        }
        umma_commit(&dp_done[bf]);
    };
    if (tid == 0) {
        mbar_expect_tx(xfull, Q_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb) tma_load_3d(&map_do, xfull, dOs + kb * (BQ * 128), coldo + (kb & 1) * 32, kb >> 1, tok0 + q0);
        load_v(0);
        if (T > 1) load_v(1);
        mbar_wait(xfull, 0);
        issue_dp(0);
    }
    const int n_row = q0 + tid;
    const float del = n_row < N ? __ldg(delta + (size_t)bh * N + n_row) : 0.f;
    for (int i = 0; i < T; ++i) {
        const int bf = i & 1;
        if (tid == 0 && i + 1 < T) issue_dp(i + 1);      // its TMEM buffer was released by the __syncthreads of iteration i - 1
        mbar_wait(&dp_done[bf], (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        if (tid == 0 && i + 2 < T) load_v(i + 2);        // V buffer bf is free: the MMAs that read it are complete
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float dp[32];
            tmem_ld32(t_row + bf * BKEY + half * 32, dp);
            if (n_row < N) {
                float4* row = reinterpret_cast<float4*>(pds + ((size_t)bh * N + n_row) * N + i * BKEY + half * 32);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    if (i * BKEY + half * 32 + j4 * 4 >= N) break;
                    float4 p = row[j4];
                    p.x = p.x * (dp[j4 * 4] - del) * scale;
                    p.y = p.y * (dp[j4 * 4 + 1] - del) * scale;
                    p.z = p.z * (dp[j4 * 4 + 2] - del) * scale;
                    p.w = p.w * (dp[j4 * 4 + 3] - del) * scale;
                    row[j4] = p;
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// Same computation with the stored P moving by TMA (N % 128 == 0): the lane threads of attn_ds_kernel read and write their own 4 KB-strided
// rows, 32 different 128-byte lines per warp instruction, which makes the load / store unit the bottleneck (6.8 us per 128 x 64 tile).
// Here P arrives as bulk-tensor loads of {32 columns, 128 rows} boxes into a ring of two 16 KB half-tile buffers (K-major SWIZZLE_128B: a
// thread reads its row as conflict-free 16-byte chunks), dS is written back in place and leaves as bulk-tensor stores; the load of the
// next tile's half is issued as soon as the store has read the buffer.
constexpr int DST_HALF_BYTES = BQ * 128;                                      // one k-block: 128 rows x 32 columns
constexpr int DST_TILE_BYTES = Q_BYTES + 2 * K_BYTES + 2 * DST_HALF_BYTES;
constexpr int DST_SMEM_BYTES = DST_TILE_BYTES + 128 + 1024;

template <int D>
__global__ void __launch_bounds__(NT, 1) attn_ds_tma_kernel(const __grid_constant__ CUtensorMap map_do,     // dO, K-major SW128, box 128 tokens
                                                           const __grid_constant__ CUtensorMap map_v,      // qkv, K-major SW128, box 64 tokens
                                                           const __grid_constant__ CUtensorMap map_p,      // P / dS [B*H*N, N], box {32, 128}, SW128
                                                           int N, int H, int C, float scale, const float* __restrict__ delta) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // 1024-byte aligned; plain offset arithmetic keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
    uint8_t* dOs = smem;
    uint8_t* Vs = dOs + Q_BYTES;                      // two buffers of K_BYTES
    uint8_t* Pb = Vs + 2 * K_BYTES;                   // two half-tile buffers of DST_HALF_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DST_TILE_BYTES);
    uint64_t* xfull = bars, *vfull = bars + 1 /* [2] */, *dp_done = bars + 3 /* [2] */, *pfull = bars + 5 /* [2] */;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int q0 = blockIdx.x * BQ;
    const int tok0 = b * N;
    const int coldo = h * D, colv = 2 * C + h * D;
    const int T = (N + BKEY - 1) / BKEY;
    const int prow = bh * N + q0;                     // first row of this CTA in the P matrix
    if (tid == 0) {
        tma_prefetch_desc(&map_do);
        tma_prefetch_desc(&map_v);
        tma_prefetch_desc(&map_p);
        for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr uint32_t idesc = make_idesc(BQ, BKEY, 0);
    auto load_v = [&](int tile) {
        const int bf = tile & 1;
        mbar_expect_tx(&vfull[bf], K_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb)
            tma_load_3d(&map_v, &vfull[bf], Vs + bf * K_BYTES + kb * (BKEY * 128), colv + (kb & 1) * 32, kb >> 1, tok0 + tile * BKEY);
    };
    auto load_p = [&](int tile, int half) {
        mbar_expect_tx(&pfull[half], DST_HALF_BYTES);
        tma_load_2d(&map_p, &pfull[half], Pb + half * DST_HALF_BYTES, tile * BKEY + half * 32, prow);
    };
    auto issue_dp = [&](int tile) {
        const int bf = tile & 1;
        mbar_wait(&vfull[bf], (uint32_t)((tile >> 1) & 1));
        tc_fence_after();
        const uint32_t da = smem_u32(dOs), va = smem_u32(Vs + bf * K_BYTES);
#pragma unroll
        for (int ks = 0; ks < KD / 8; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            if ((kb & 1) && kk * 8 + 32 >= D) continue;
            umma_tf32(tmem_base + bf * BKEY, make_desc(da + kb * (BQ * 128) + kk * 32), make_desc(va + kb * (BKEY * 128) + kk * 32), idesc,
                      ks != 0 ? 1u : 0u);
        }
        umma_commit(&dp_done[bf]);
    };
    if (tid == 0) {
        mbar_expect_tx(xfull, Q_BYTES);
#pragma unroll
        for (int kb = 0; kb < KD / 32; ++kb) tma_load_3d(&map_do, xfull, dOs + kb * (BQ * 128), coldo + (kb & 1) * 32, kb >> 1, tok0 + q0);
        load_v(0);
        if (T > 1) load_v(1);
        load_p(0, 0);
        load_p(0, 1);
        mbar_wait(xfull, 0);
        issue_dp(0);
    }
    const float del = __ldg(delta + (size_t)bh * N + q0 + tid);      // N % 128 == 0: every lane is a real query
    // this thread's row inside a half-tile buffer: 16-byte chunk c sits at chunk (c ^ (row & 7)) of the row's 128 bytes
    const uint32_t row_off = (uint32_t)((tid >> 3) * 1024 + (tid & 7) * 128);
    for (int i = 0; i < T; ++i) {
        const int bf = i & 1;
        if (tid == 0 && i + 1 < T) issue_dp(i + 1);      // its TMEM buffer was released by the last __syncthreads of iteration i - 1
        mbar_wait(&dp_done[bf], (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        if (tid == 0 && i + 2 < T) load_v(i + 2);        // V buffer bf is free: the MMAs that read it are complete
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float dp[32];
            tmem_ld32(t_row + bf * BKEY + half * 32, dp);
            mbar_wait(&pfull[half], (uint32_t)(i & 1));
            uint8_t* mine = Pb + half * DST_HALF_BYTES + row_off;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                float4* q = reinterpret_cast<float4*>(mine + ((j4 ^ (tid & 7)) << 4));
                float4 pv = *q;
                pv.x = pv.x * (dp[j4 * 4] - del) * scale;
                pv.y = pv.y * (dp[j4 * 4 + 1] - del) * scale;
                pv.z = pv.z * (dp[j4 * 4 + 2] - del) * scale;
                pv.w = pv.w * (dp[j4 * 4 + 3] - del) * scale;
                *q = pv;
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();      // the half tile is complete in shared memory (and, after the second half, accumulator bf is consumed)
            if (tid == 0) {
                tma_store_2d(&map_p, Pb + half * DST_HALF_BYTES, i * BKEY + half * 32, prow);
                tma_store_commit();
                if (i + 1 < T) {
                    tma_store_wait_read();      // the buffer may be refilled once the store has read it
                    load_p(i + 1, half);
                }
            }
        }
    }
    if (tid == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

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
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// qkv [tokens*3, cols] (leading dimension ld) viewed as a 3-D tensor {cols, 3 components, tokens}; box = {32 cols, 1, box_tokens}
static bool make_map3(CUtensorMap* m, const float* base, long long tokens, long long cols, long long ld, int box_tokens, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, 3, (cuuint64_t)tokens};
    cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)ld * 3 * sizeof(float)};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)box_tokens};
    cuuint32_t estr[3] = {1, 1, 1};
    // the L2 promotion size must not exceed a row of the tensor (a 48-column dO has 192-byte rows)
    const CUtensorMapL2promotion promo = cols * 4 >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                          : (cols * 4 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// row-major fp32 matrix [rows, cols] (leading dimension = cols); box = {32 cols, box_rows}
static bool make_map2(CUtensorMap* m, const float* base, long long rows, long long cols, int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
              cols * 4 >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
           CUDA_SUCCESS;
}

}  // namespace atc
}  // namespace vnpcc

using namespace vnpcc;

extern "C" {

// tensor-core twin of vnpcc_vn_attention_fwd.  p_out (may be NULL; needs N % 32 == 0): the normalised attention weights P [B*H*N, N],
// stored for the GEMM-shaped backward.  D == 48 only; returns VNPCC_ERR_UNSUPPORTED otherwise.
int vnpcc_vn_attention_fwd_tf32(const float* qkv, long long ld, int B, int N, int H, int D, float scale, float* out, long long ldo, float* lse,
                                float* p_out, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (D != 48 || H <= 0 || ld % 4 != 0 || ldo % 4 != 0 || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || !(scale > 0.f))
        return VNPCC_ERR_UNSUPPORTED;
    if (p_out != nullptr && (N % 32 != 0 || ((uintptr_t)p_out & 15))) return VNPCC_ERR_BAD_ARG;
    CUtensorMap mk, mq, mv;
    const long long tokens = (long long)B * N, cols = 3LL * H * D;
    if (!atc::make_map3(&mk, qkv, tokens, cols, ld, atc::BKEY, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&mq, qkv, tokens, cols, ld, atc::BQ, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&mv, qkv, tokens, cols, ld, atc::BKEY, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return VNPCC_ERR_DRIVER;
    if (cudaFuncSetAttribute(atc::attn_fwd_tc_kernel<48, atc::MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::SMEM_BYTES) !=
        cudaSuccess)
        return VNPCC_ERR_DRIVER;
    dim3 grid((unsigned)((N + atc::BQ - 1) / atc::BQ), (unsigned)(B * H));
    const int C = H * D;
    // the stored P leaves the kernel as bulk-tensor stores of the swizzled operand tile when whole 128-row tiles fit every (sample, head)
    CUtensorMap mp = mk;
    int p_tma = 0;
    if (p_out != nullptr && N % atc::BQ == 0 && atc::make_map2(&mp, p_out, (long long)B * H * N, N, atc::BQ, CU_TENSOR_MAP_SWIZZLE_128B)) p_tma = 1;
    count_launch(), atc::attn_fwd_tc_kernel<48, atc::MODE_FWD><<<grid, atc::NT, atc::SMEM_BYTES, (cudaStream_t)stream>>>(
                        mk, mq, mv, mp, p_tma, N, H, 0, C, 2 * C, scale, out, (size_t)ldo, lse, p_out);
    return last_error();
}

// tensor-core twin of vnpcc_vn_attention_bwd.  dqkv is fully written with plain stores (no pre-zeroing, no atomics).  D == 48 only.
// Three routes, fastest first:
//   p_buf != NULL (the P the forward stored; N % 32 == 0; DESTROYED: it is turned into dS in place):
//        dV = P^T dO (streaming GEMM) -> dS = P o (dO V^T - delta) scale in place -> dQ = dS K, dK = dS^T Q (streaming GEMMs)
//   ds_workspace (B*H*N*N floats, N % 32 == 0): dV by the forward skeleton with swapped roles, dQ by the recomputing kernel which also
//        stores dS, dK = dS^T Q as a streaming GEMM
//   neither: dV as above, dQ and dK by the two orientations of the recomputing kernel
int vnpcc_vn_attention_bwd_tf32(const float* qkv, long long ld, const float* dout, long long lddo, const float* out, long long ldo,
                                const float* lse, int B, int N, int H, int D, float scale, float* dqkv, long long lddq, float* delta,
                                float* ds_workspace, size_t ds_workspace_bytes, float* p_buf, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (D != 48 || H <= 0 || ld % 4 != 0 || lddo % 4 != 0 || ldo % 4 != 0 || lddq % 4 != 0 || ((uintptr_t)qkv & 15) || ((uintptr_t)dout & 15) ||
        ((uintptr_t)out & 15) || ((uintptr_t)dqkv & 15) || !(scale > 0.f))
        return VNPCC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = H * D;
    const long long tokens = (long long)B * N;
    CUtensorMap q128, q64, q32, qm32, do128, do32, dom64, dom32;
    if (!atc::make_map3(&q128, qkv, tokens, 3LL * C, ld, 128, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&q64, qkv, tokens, 3LL * C, ld, 64, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&q32, qkv, tokens, 3LL * C, ld, 32, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&qm32, qkv, tokens, 3LL * C, ld, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !atc::make_map3(&do128, dout, tokens, C, lddo, 128, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&do32, dout, tokens, C, lddo, 32, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !atc::make_map3(&dom64, dout, tokens, C, lddo, 64, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !atc::make_map3(&dom32, dout, tokens, C, lddo, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
        fprintf(stderr, "[vnpcc] attention_bwd_tf32: tensor map encode failed (tokens %lld C %d ld %lld lddo %lld qkv %p dout %p)\n", tokens, C, ld, lddo,
                (const void*)qkv, (const void*)dout);
        return VNPCC_ERR_DRIVER;
    }
    if (cudaFuncSetAttribute(atc::attn_fwd_tc_kernel<48, atc::MODE_DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_bwd_tc_kernel<48, atc::BMODE_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::BSMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_bwd_tc_kernel<48, atc::BMODE_DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::BSMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_acc_gemm_kernel<48, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::GK_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_acc_gemm_kernel<48, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::GK_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_ds_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::DSK_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(atc::attn_ds_tma_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::DST_SMEM_BYTES) != cudaSuccess) {
        fprintf(stderr, "[vnpcc] attention_bwd_tf32: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return VNPCC_ERR_DRIVER;
    }
    int rc = vnpcc_vn_attention_delta(dout, lddo, out, ldo, B, N, H, D, delta, stream);
    if (rc) return rc;
    dim3 grid((unsigned)((N + atc::BQ - 1) / atc::BQ), (unsigned)(B * H));
    const long long mrows = (long long)B * H * N;

    if (p_buf != nullptr && N % 32 == 0 && !((uintptr_t)p_buf & 15)) {
        CUtensorMap mp_mn, mp_k;
        if (!atc::make_map2(&mp_mn, p_buf, mrows, N, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
            !atc::make_map2(&mp_k, p_buf, mrows, N, 128, CU_TENSOR_MAP_SWIZZLE_128B))
            return VNPCC_ERR_DRIVER;
        // dV = P^T dO
        count_launch(), atc::attn_acc_gemm_kernel<48, 1><<<grid, atc::NT, atc::GK_SMEM_BYTES, st>>>(mp_mn, dom32, N, H, 0, dqkv + 2 * C, (size_t)lddq);
        // P -> dS in place
        if (N % atc::BQ == 0)
            count_launch(), atc::attn_ds_tma_kernel<48><<<grid, atc::NT, atc::DST_SMEM_BYTES, st>>>(do128, q64, mp_k, N, H, C, scale, delta);
        else
            count_launch(), atc::attn_ds_kernel<48><<<grid, atc::NT, atc::DSK_SMEM_BYTES, st>>>(do128, q64, N, H, C, scale, delta, p_buf);
        // dQ = dS K ; dK = dS^T Q
        count_launch(), atc::attn_acc_gemm_kernel<48, 0><<<grid, atc::NT, atc::GK_SMEM_BYTES, st>>>(mp_k, qm32, N, H, C, dqkv, (size_t)lddq);
        count_launch(), atc::attn_acc_gemm_kernel<48, 1><<<grid, atc::NT, atc::GK_SMEM_BYTES, st>>>(mp_mn, qm32, N, H, 0, dqkv + C, (size_t)lddq);
        return last_error();
    }

    // dV: resident K (columns C..), streamed Q (columns 0..) K-major, streamed dO MN-major; writes the v part of dqkv
    count_launch(), atc::attn_fwd_tc_kernel<48, atc::MODE_DV><<<grid, atc::NT, atc::SMEM_BYTES, st>>>(q64, q128, dom64, q64, 0, N, H, C, 0, 0, scale,
                                                                                                 dqkv + 2 * C, (size_t)lddq, const_cast<float*>(lse),
                                                                                                 nullptr);
    // dQ (and, when the caller provides B*H*N*N floats of workspace and N % 32 == 0, the dS matrix for the dK GEMM)
    const size_t ds_need = (size_t)B * H * N * N * sizeof(float);
    CUtensorMap mds;
    const bool use_ds = ds_workspace != nullptr && ds_workspace_bytes >= ds_need && N % 32 == 0 && !((uintptr_t)ds_workspace & 15) &&
                        atc::make_map2(&mds, ds_workspace, mrows, N, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    count_launch(), atc::attn_bwd_tc_kernel<48, atc::BMODE_DQ><<<grid, atc::NT, atc::BSMEM_BYTES, st>>>(q128, q32, qm32, do128, do32, N, H, C, scale, lse,
                                                                                                   delta, dqkv, (size_t)lddq,
                                                                                                   use_ds ? ds_workspace : nullptr);
    if (use_ds)
        count_launch(), atc::attn_acc_gemm_kernel<48, 1><<<grid, atc::NT, atc::GK_SMEM_BYTES, st>>>(mds, qm32, N, H, 0, dqkv + C, (size_t)lddq);
    else
        count_launch(), atc::attn_bwd_tc_kernel<48, atc::BMODE_DK><<<grid, atc::NT, atc::BSMEM_BYTES, st>>>(q128, q32, qm32, do128, do32, N, H, C, scale,
                                                                                                       lse, delta, dqkv, (size_t)lddq, nullptr);
    return last_error();
}

}  // extern "C"
