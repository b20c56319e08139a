// vn_math.cuh -- per-lane device helpers shared by the vectorised Vector-Neuron kernels (vn_stream.cu, vn_fused.cu).
// A thread owns four consecutive channels ("lanes" l = 0..3) of one point; V4x3 holds the 3 components x 4 lanes.
#pragma once
#include <cuda_runtime.h>

namespace vnpcc {

constexpr float VS_EPS = 1e-6f;

struct V4x3 {
    float v[3][4];   // [component][channel lane]
};

__device__ __forceinline__ V4x3 ld43(const float* __restrict__ base, size_t ld) {
    V4x3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + c * ld));
        r.v[c][0] = t.x;
        r.v[c][1] = t.y;
        r.v[c][2] = t.z;
        r.v[c][3] = t.w;
    }
    return r;
}
__device__ __forceinline__ V4x3 ld43_rw(const float* base, size_t ld) {
    V4x3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(base + c * ld);
        r.v[c][0] = t.x;
        r.v[c][1] = t.y;
        r.v[c][2] = t.z;
        r.v[c][3] = t.w;
    }
    return r;
}
__device__ __forceinline__ void st43(float* __restrict__ base, size_t ld, const V4x3& r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(base + c * ld) = make_float4(r.v[c][0], r.v[c][1], r.v[c][2], r.v[c][3]);
}
// (a*b).sum over the 3 components: three rounded products summed left to right, no contraction
__device__ __forceinline__ float dot3l(const V4x3& a, const V4x3& b, int l) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.v[0][l], b.v[0][l]), __fmul_rn(a.v[1][l], b.v[1][l])), __fmul_rn(a.v[2][l], b.v[2][l]));
}

struct ChanParams {
    float mean[4], invstd[4], gamma[4], beta[4];
};
__device__ __forceinline__ ChanParams load_params(const float* stat, const float* gamma, const float* beta, int C, int c0) {
    ChanParams p;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        p.mean[l] = stat ? __ldg(stat + c0 + l) : 0.f;
        p.invstd[l] = stat ? __ldg(stat + C + c0 + l) : 0.f;
        p.gamma[l] = stat ? __ldg(gamma + c0 + l) : 0.f;
        p.beta[l] = stat ? __ldg(beta + c0 + l) : 0.f;
    }
    return p;
}

// process-wide switch (vnpcc_set_fast_math): forward kernels use MUFU reciprocal / rsqrt instead of IEEE division / sqrt.
// Off in parity (fp32) mode, on in throughput (TF32) mode where GEMM operands are already rounded to 10 mantissa bits.
bool fast_math_enabled();

// ONE MUFU instruction each: the plain intrinsics (__fdividef, rsqrtf) wrap the same instruction in range-scaling code for denormal /
// huge operands (4-5 extra instructions per call, which the instruction-bound streaming kernels pay per lane).  Every divisor on these
// paths carries the layers' + 1e-6, and the square roots clamp their argument to the smallest normal number (a denormal squared norm
// is treated like FLT_MIN: both are 13 orders of magnitude below the 1e-6 they are added to).
__device__ __forceinline__ float mufu_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(x, 1.17549435e-38f)));
    return y;
}

template <bool FAST>
__device__ __forceinline__ float vdiv(float a, float b) {
    return FAST ? a * mufu_rcp(b) : a / b;
}
template <bool FAST>
__device__ __forceinline__ float vsqrt(float a) {
    return FAST ? (a > 0.f ? a * mufu_rsqrt(a) : 0.f) : sqrtf(a);
}

template <bool FAST>
__device__ __forceinline__ void bn_apply_lane_t(V4x3& v, int l, const ChanParams& cp, float& n_out, float& nhat_out, float& nb_out) {
    const float r = vsqrt<FAST>(__fadd_rn(__fadd_rn(__fmul_rn(v.v[0][l], v.v[0][l]), __fmul_rn(v.v[1][l], v.v[1][l])),
                                          __fmul_rn(v.v[2][l], v.v[2][l])));
    const float n = r + VS_EPS;
    const float nhat = (n - cp.mean[l]) * cp.invstd[l];
    const float nb = nhat * cp.gamma[l] + cp.beta[l];
    if (FAST) {
        const float t = nb * mufu_rcp(n);
        v.v[0][l] *= t;
        v.v[1][l] *= t;
        v.v[2][l] *= t;
    } else {
        v.v[0][l] = v.v[0][l] / n * nb;
        v.v[1][l] = v.v[1][l] / n * nb;
        v.v[2][l] = v.v[2][l] / n * nb;
    }
    n_out = n;
    nhat_out = nhat;
    nb_out = nb;
}

// leaky projection of lane l of v (post-BN) along dv, in place (op-by-op rounding of the eager expression unless FAST)
template <bool FAST>
__device__ __forceinline__ void leaky_lane_t(V4x3& v, const V4x3& dv, int l, float ns, float k) {
    const float dot = dot3l(v, dv, l);
    float in0 = v.v[0][l], in1 = v.v[1][l], in2 = v.v[2][l];
    if (!(dot >= 0.f)) {
        const float a = vdiv<FAST>(dot, __fadd_rn(dot3l(dv, dv, l), VS_EPS));
        in0 = __fsub_rn(in0, __fmul_rn(a, dv.v[0][l]));
        in1 = __fsub_rn(in1, __fmul_rn(a, dv.v[1][l]));
        in2 = __fsub_rn(in2, __fmul_rn(a, dv.v[2][l]));
    }
    v.v[0][l] = __fadd_rn(__fmul_rn(ns, v.v[0][l]), __fmul_rn(k, in0));
    v.v[1][l] = __fadd_rn(__fmul_rn(ns, v.v[1][l]), __fmul_rn(k, in1));
    v.v[2][l] = __fadd_rn(__fmul_rn(ns, v.v[2][l]), __fmul_rn(k, in2));
}

__device__ __forceinline__ void bn_apply_lane(V4x3& v, int l, const ChanParams& cp, float& n_out, float& nhat_out, float& nb_out) {
    const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v.v[0][l], v.v[0][l]), __fmul_rn(v.v[1][l], v.v[1][l])),
                                    __fmul_rn(v.v[2][l], v.v[2][l])));
    const float n = r + VS_EPS;
    const float nhat = (n - cp.mean[l]) * cp.invstd[l];
    const float nb = nhat * cp.gamma[l] + cp.beta[l];
    v.v[0][l] = v.v[0][l] / n * nb;
    v.v[1][l] = v.v[1][l] / n * nb;
    v.v[2][l] = v.v[2][l] / n * nb;
    n_out = n;
    nhat_out = nhat;
    nb_out = nb;
}


// ---- packed fp32x2 helpers (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per issue slot) ----
// A thread's four channel lanes are handled as two pairs; all per-lane scalars of the backward formulas become pairs.
struct f2 {
    float2 v;
};
__device__ __forceinline__ f2 mk2(float a, float b) { return {make_float2(a, b)}; }
__device__ __forceinline__ f2 bc2(float a) { return {make_float2(a, a)}; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return {__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return {__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ f2 fma2p(f2 a, f2 b, f2 c) { return {__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return {__ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v)}; }   // exact: a + (-1)*b
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 rcp2(f2 a) { return mk2(mufu_rcp(a.v.x), mufu_rcp(a.v.y)); }
__device__ __forceinline__ f2 rsqrt2(f2 a) { return mk2(mufu_rsqrt(a.v.x), mufu_rsqrt(a.v.y)); }
// a where the mask half is true, 0 elsewhere
__device__ __forceinline__ f2 sel0(bool mx, bool my, f2 a) { return mk2(mx ? a.v.x : 0.f, my ? a.v.y : 0.f); }
__device__ __forceinline__ f2 dot3p(const f2 (&a)[3], const f2 (&b)[3]) { return fma2p(a[2], b[2], fma2p(a[1], b[1], a[0] * b[0])); }

// packed throughput-mode forward of one lane pair: out = leaky(BN(p), d) = t p - (k a) d  with  t = nb / n (1 without BN),
// a = <BN(p), d> / (d.d + eps) where that is negative and 0 elsewhere.  p is overwritten with the output.
struct BNPair {
    f2 mean, invstd, gamma, beta;
};
template <bool HAS_BN>
__device__ __forceinline__ void leaky_bn_pair_fwd(f2 (&p)[3], const f2 (&d)[3], const BNPair& bn, float k) {
    f2 t = bc2(1.f);
    if (HAS_BN) {
        const f2 pp = dot3p(p, p);
        f2 rs = rsqrt2(pp);
        rs = mk2(pp.v.x > 0.f ? rs.v.x : 0.f, pp.v.y > 0.f ? rs.v.y : 0.f);
        const f2 n = fma2p(pp, rs, bc2(VS_EPS));
        const f2 nhat = (n - bn.mean) * bn.invstd;
        t = fma2p(nhat, bn.gamma, bn.beta) * rcp2(n);
    }
    const f2 s = t * dot3p(p, d);
    const f2 rq = rcp2(dot3p(d, d) + bc2(VS_EPS));
    const f2 ka = sel0(s.v.x < 0.f, s.v.y < 0.f, bc2(k) * (s * rq));
#pragma unroll
    for (int v = 0; v < 3; ++v) p[v] = t * p[v] - ka * d[v];
}
__device__ __forceinline__ BNPair load_bn_pair(const float* stat, const float* gamma, const float* beta, int C, int c) {
    BNPair b;
    b.mean = stat ? mk2(__ldg(stat + c), __ldg(stat + c + 1)) : bc2(0.f);
    b.invstd = stat ? mk2(__ldg(stat + C + c), __ldg(stat + C + c + 1)) : bc2(0.f);
    b.gamma = stat ? mk2(__ldg(gamma + c), __ldg(gamma + c + 1)) : bc2(0.f);
    b.beta = stat ? mk2(__ldg(beta + c), __ldg(beta + c + 1)) : bc2(0.f);
    return b;
}

// approximate reciprocal / square root (MUFU, ~1 ulp): used by the BACKWARD kernels only -- gradients do not need the
// op-by-op IEEE rounding the forward keeps for parity of masks and selections, and IEEE divisions (about ten per lane)
// made those kernels instruction-bound
__device__ __forceinline__ float frcp(float x) { return mufu_rcp(x); }
__device__ __forceinline__ float fsqrt_fast(float x) { return x > 0.f ? x * mufu_rsqrt(x) : 0.f; }

}  // namespace vnpcc
