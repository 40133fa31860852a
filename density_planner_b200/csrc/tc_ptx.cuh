// tcgen05 / TMEM / mbarrier PTX wrappers, operand-image layout and packed-math helpers shared by the tensor-core rollout
// kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

constexpr int TILE = 128;
constexpr float WSCALE = 16.0f;  // weights are stored x16 so that w_lo = fp16(w - fp16(w)) stays a normal fp16 number
constexpr float INV_WSCALE = 1.0f / 16.0f;
constexpr int N3A = 48, N3B = 32;  // layer-3 MMA N (net b padded 24 -> 32: UMMA M=128 needs N % 16 == 0)

// ---- operand image (identical byte layout in global and shared memory, built once per controller) ------------------
constexpr int W2_BYTES = 128 * 128 * 2;
constexpr int W3A_BYTES = N3A * 128 * 2;
constexpr int W3B_BYTES = N3B * 128 * 2;
constexpr int OFF_W2A_HI = 0;
constexpr int OFF_W2A_LO = OFF_W2A_HI + W2_BYTES;
constexpr int OFF_W2B_HI = OFF_W2A_LO + W2_BYTES;
constexpr int OFF_W2B_LO = OFF_W2B_HI + W2_BYTES;
constexpr int OFF_W3A_HI = OFF_W2B_LO + W2_BYTES;
constexpr int OFF_W3A_LO = OFF_W3A_HI + W3A_BYTES;
constexpr int OFF_W3B_HI = OFF_W3A_LO + W3A_BYTES;
constexpr int OFF_W3B_LO = OFF_W3B_HI + W3B_BYTES;
constexpr int OFF_SMALL = OFF_W3B_LO + W3B_BYTES;  // per-net fp32 block (layout owned by each kernel), 1024 floats per net
constexpr int SMALL_NET_FLOATS = 1280;
constexpr int IMAGE_BYTES = OFF_SMALL + 2 * SMALL_NET_FLOATS * 4;
static_assert(IMAGE_BYTES % 16 == 0, "image copied with 16-byte vectors");

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// true in exactly one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xFFFFFFFF;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16, issued by one thread
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum), "r"(0)
        : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// instruction descriptor, kind::f16: D fp32, A/B fp16, both K-major, M = 128 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// shared-memory matrix descriptor, canonical K-major layout without swizzle (cute::UMMA::SmemDescriptor bit layout):
// core matrix = 8 rows x 16 bytes; LBO = byte distance between the two core matrices along K of one MMA,
// SBO = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
constexpr uint32_t B_LBO = 128;          // K-adjacent core matrices are contiguous
constexpr uint32_t B_SBO = 16 * 128;     // K = 128 -> 16 core matrices per 8-row group
constexpr uint32_t B_KSTEP = 2 * B_LBO;  // one MMA consumes K = 16 = two core matrices
constexpr uint64_t B_KSTEP_DESC = B_KSTEP >> 4;

// ---- math helpers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float TWO_LOG2E = 2.8853900817779268f;
// tanh(x) = 1 - 2 / (1 + e^(2x)); absolute error ~1e-7 (MUFU.EX2 + MUFU.RCP), saturates cleanly at +-1
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(x * TWO_LOG2E);
    return fmaf(-2.0f, rcp_approx(1.0f + e), 1.0f);
}
// two tanh at once from pre-scaled arguments y = 2 log2(e) x, in packed fp32x2 arithmetic (FFMA2 / FADD2):
// t = tanh(x), g = 1 - t^2
__device__ __forceinline__ void tanh_gate2(const float2 y, float2 &t, float2 &g) {
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 a = __fadd2_rn(make_float2(ex2_approx(y.x), ex2_approx(y.y)), one);
    const float2 r = make_float2(rcp_approx(a.x), rcp_approx(a.y));
    t = __ffma2_rn(make_float2(-2.0f, -2.0f), r, one);
    g = __ffma2_rn(make_float2(-t.x, -t.y), t, one);
}
// four tanh at once with ONE reciprocal: e = 2^-|y| in (0, 1] so a = 1 + e in (1, 2] and the product of four a's cannot
// overflow; 1/a_i is recovered from R = 1 / (a0 a1 a2 a3) with products (three MUFU.RCP traded for a few FMULs: the
// MUFU pipe, 4 lanes/clk per SM sub-partition, is what bounds the epilogues).  tanh|x| = 2/(1+e) - 1, sign copied from y.
__device__ __forceinline__ float ex2_negabs(float y) { return ex2_approx(-fabsf(y)); }
__device__ __forceinline__ float2 copysign2(const float2 mag, const float2 sgn) {  // mag >= 0
    return make_float2(__uint_as_float(__float_as_uint(mag.x) | (__float_as_uint(sgn.x) & 0x80000000u)),
                       __uint_as_float(__float_as_uint(mag.y) | (__float_as_uint(sgn.y) & 0x80000000u)));
}
// DP_TIE (A/B knob, rollout_tc.cu): the shared reciprocal R (DP_TIE_ON == 0) or the product it is taken of (== 1) of one
// group of four is handed to the caller, which folds `R * 0` into an operand of the NEXT group's pre-activations.  The value
// is unchanged (R is finite), but the next group's MUFU.EX2 now depend on this group: ptxas can no longer hoist all sixteen
// EX2 of a chunk into one burst that serialises the four warps of a scheduler on the XU pipe.
#ifndef DP_TIE_ON
#define DP_TIE_ON 0
#endif
__device__ __forceinline__ void tanh4(const float2 yA, const float2 yB, float2 &tA, float2 &tB, float &tie) {
    const float2 one = make_float2(1.0f, 1.0f), two = make_float2(2.0f, 2.0f), mone = make_float2(-1.0f, -1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_negabs(yA.x), ex2_negabs(yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_negabs(yB.x), ex2_negabs(yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);                       // (a0 a2, a1 a3)
    const float R = rcp_approx(p.x * p.y);
    tie = DP_TIE_ON ? p.x : R;
    const float2 q = make_float2(R * p.y, R * p.x);            // (1/(a0 a2), 1/(a1 a3))
    const float2 rA = __fmul2_rn(q, aB), rB = __fmul2_rn(q, aA);  // (1/a0, 1/a1), (1/a2, 1/a3)
    tA = copysign2(__ffma2_rn(two, rA, mone), yA);
    tB = copysign2(__ffma2_rn(two, rB, mone), yB);
}
__device__ __forceinline__ void tanh4(const float2 yA, const float2 yB, float2 &tA, float2 &tB) {
    const float2 one = make_float2(1.0f, 1.0f), two = make_float2(2.0f, 2.0f), mone = make_float2(-1.0f, -1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_negabs(yA.x), ex2_negabs(yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_negabs(yB.x), ex2_negabs(yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);                       // (a0 a2, a1 a3)
    const float R = rcp_approx(p.x * p.y);
    const float2 q = make_float2(R * p.y, R * p.x);            // (1/(a0 a2), 1/(a1 a3))
    const float2 rA = __fmul2_rn(q, aB), rB = __fmul2_rn(q, aA);  // (1/a0, 1/a1), (1/a2, 1/a3)
    tA = copysign2(__ffma2_rn(two, rA, mone), yA);
    tB = copysign2(__ffma2_rn(two, rB, mone), yB);
}
// The same without |y| and without the sign copy, for arguments whose size is bounded: 2^-y may now be large, and the shared
// reciprocal only works while the product of the four denominators stays a normal number with a normal reciprocal, i.e.
// every |y| < 31.5 (caller's contract: the host derives the bound from the weights).  tanh x = 2/(1 + 2^-y) - 1.
__device__ __forceinline__ void tanh4_bounded(const float2 yA, const float2 yB, float2 &tA, float2 &tB) {
    const float2 one = make_float2(1.0f, 1.0f), two = make_float2(2.0f, 2.0f), mone = make_float2(-1.0f, -1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_approx(-yA.x), ex2_approx(-yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_approx(-yB.x), ex2_approx(-yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);
    const float R = rcp_approx(p.x * p.y);
    const float2 q = make_float2(R * p.y, R * p.x);
    const float2 rA = __fmul2_rn(q, aB), rB = __fmul2_rn(q, aA);
    tA = __ffma2_rn(two, rA, mone);
    tB = __ffma2_rn(two, rB, mone);
}
// ... and the variant that returns r = 1 / (1 + 2^-y) = (1 + tanh x) / 2 itself: a layer whose consumer is linear can take
// r as its operand (W tanh = 2 W r - W 1: the factor and the row sums go into the consumer's scale and bias), which saves
// the 2 r - 1 step; the gate is 1 - tanh^2 = 4 r (1 - r).
__device__ __forceinline__ void sigm4_bounded(const float2 yA, const float2 yB, float2 &rA, float2 &rB, float &tie) {
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_approx(-yA.x), ex2_approx(-yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_approx(-yB.x), ex2_approx(-yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);
    const float R = rcp_approx(p.x * p.y);
    tie = DP_TIE_ON ? fminf(p.x, 3.0e38f) : R;   // the product may overflow to +inf (R = 0 then): keep the tie finite
    const float2 q = make_float2(R * p.y, R * p.x);
    rA = __fmul2_rn(q, aB);
    rB = __fmul2_rn(q, aA);
}
__device__ __forceinline__ void sigm4_bounded(const float2 yA, const float2 yB, float2 &rA, float2 &rB) {
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_approx(-yA.x), ex2_approx(-yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_approx(-yB.x), ex2_approx(-yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);
    const float R = rcp_approx(p.x * p.y);
    const float2 q = make_float2(R * p.y, R * p.x);
    rA = __fmul2_rn(q, aB);
    rB = __fmul2_rn(q, aA);
}
// The same with TWO reciprocals: the pairs (0,2) and (1,3) share one each (packed products only: 11 instructions for four
// arguments instead of 13, 1.5 MUFU per argument instead of 1.25).  The product of two denominators stays a normal number with a
// normal reciprocal while every y > -63 (caller's contract).
__device__ __forceinline__ void sigm4_pairs(const float2 yA, const float2 yB, float2 &rA, float2 &rB) {
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 aA = __fadd2_rn(make_float2(ex2_approx(-yA.x), ex2_approx(-yA.y)), one);
    const float2 aB = __fadd2_rn(make_float2(ex2_approx(-yB.x), ex2_approx(-yB.y)), one);
    const float2 p = __fmul2_rn(aA, aB);                                     // (a0 a2, a1 a3)
    const float2 R = make_float2(rcp_approx(p.x), rcp_approx(p.y));
    rA = __fmul2_rn(R, aB);                                                  // (1/a0, 1/a1)
    rB = __fmul2_rn(R, aA);                                                  // (1/a2, 1/a3)
}
__device__ __forceinline__ uint32_t h2_bits(const __half2 h) { return *reinterpret_cast<const uint32_t *>(&h); }
__device__ __forceinline__ __half2 bits_h2(const uint32_t u) { return *reinterpret_cast<const __half2 *>(&u); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { return h2_bits(__floats2half2_rn(a, b)); }  // a -> low half
// (a, b) -> fp16 pair hi and the fp16 pair of the residuals lo: a ~= hi.x + lo.x to ~22 bits
__device__ __forceinline__ void split_h2(const float2 v, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(v.x, v.y);
    const float2 f = __half22float2(h);
    const float2 d = __fadd2_rn(v, make_float2(-f.x, -f.y));
    hi = h2_bits(h);
    lo = pack_h2(d.x, d.y);
}
// The same split with the hi part formed by TRUNCATING the fp32 mantissa to fp16's 10 bits (one LOP3 per element on the
// integer pipe instead of an fp16->fp32 conversion on the FMA pipe, which bounds the epilogues): hi is exactly
// representable in fp16, lo = v - hi is exact in fp32 and at most one fp16 ulp of hi, so hi + lo carries >= 21 bits.
__device__ __forceinline__ void split_h2_trunc(const float2 v, uint32_t &hi, uint32_t &lo) {
    const float2 f = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    const float2 d = __fadd2_rn(v, make_float2(-f.x, -f.y));
    hi = pack_h2(f.x, f.y);
    lo = pack_h2(d.x, d.y);
}
// The split with sm_100's mixed-precision add (add.rn.f32.f16, SASS FHADD: fp32 + one half of a packed fp16 register, with
// a free negation): hi = RN_fp16(v) in one F2FP for the pair, lo = v - hi in one FHADD per element -- no mantissa mask and no
// fp16->fp32 conversion: 4 instructions per pair instead of 5, and |lo| <= half an fp16 ulp of hi.
__device__ __forceinline__ void split_h2_mixed(const float2 v, uint32_t &hi, uint32_t &lo) {
    hi = pack_h2(v.x, v.y);
    float d0, d1;
    asm("{\n"
        ".reg .b32 n;\n"
        ".reg .b16 nl, nh;\n"
        "neg.f16x2 n, %2;\n"
        "mov.b32 {nl, nh}, n;\n"
        "add.rn.f32.f16 %0, nl, %3;\n"
        "add.rn.f32.f16 %1, nh, %4;\n"
        "}\n"
        : "=f"(d0), "=f"(d1)
        : "r"(hi), "f"(v.x), "f"(v.y));
    lo = pack_h2(d0, d1);
}
// tanh gate 1 - t^2 of a packed fp16 pair, one HFMA2 (the gates only scale the fp16 tangent operands)
__device__ __forceinline__ uint32_t gate_h2(const uint32_t hi) {
    const __half2 h = bits_h2(hi);
    return h2_bits(__hfma2(__hneg2(h), h, __floats2half2_rn(1.0f, 1.0f)));
}

// the gate used by the kernel: fp32 (FFMA2 + F2FP) by default; -DDP_GATE_FROM_HI=1 forms it from the fp16 hi pair (one
// instruction less per feature pair, but the log-density error grows from 0.22 to 0.36 of its tolerance)
#ifndef DP_GATE_FROM_HI
#define DP_GATE_FROM_HI 0
#endif
__device__ __forceinline__ uint32_t gate_of(const float2 t, const uint32_t hi) {
#if DP_GATE_FROM_HI
    return gate_h2(hi);
#else
    const float2 g = __ffma2_rn(make_float2(-t.x, -t.y), t, make_float2(1.0f, 1.0f));
    return pack_h2(g.x, g.y);
#endif
}

// ---- host: canonical no-swizzle K-major image of W [n_rows][128] (rows >= n_valid are zero) --------------------------------
inline size_t canon_index(int r, int k) { return ((size_t)(r / 8) * 16 + (k / 8)) * 64 + (r % 8) * 8 + (k % 8); }

// sum over k of the row's split image (hi + lo), i.e. what an all-ones operand would accumulate
inline double split_image_rowsum(const float *W, int r, float scale) {
    double s = 0.0;
    for (int k = 0; k < 128; ++k) {
        const float w = W[(size_t)r * 128 + k] * scale;
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        s += (double)__half2float(h) + (double)__half2float(l);
    }
    return s;
}

inline void build_split_image(const float *W, int n_valid, int n_rows, float scale, __half *hi, __half *lo) {
    for (int r = 0; r < n_rows; ++r)
        for (int k = 0; k < 128; ++k) {
            const float w = r < n_valid ? W[(size_t)r * 128 + k] * scale : 0.0f;
            const __half h = __float2half_rn(w);
            const __half l = __float2half_rn(w - __half2float(h));
            hi[canon_index(r, k)] = h;
            lo[canon_index(r, k)] = l;
        }
}

}  // namespace tc
