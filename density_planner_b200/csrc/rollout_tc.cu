// K1 (tensor-core version): fused closed-loop rollout + Liouville density on tcgen05 (sm_100a).
//
// Same math as rollout_ffma.cu (reference: systems/ControlAffineSystem.py:409-463 and the closed form of :69-122), but
// the 128-wide layers run on the 5th-generation tensor cores with split-precision fp16 operands:
//     x = x_hi + x_lo  (x_hi = fp16(x), x_lo = fp16(x - x_hi));   a.w ~= a_hi w_hi + a_hi w_lo + a_lo w_hi
// which carries ~22 mantissa bits -- plain fp16/bf16/tf32 operands fail the parity budget because rounding flips the
// actuator clip mask.  The VALUE rows (which decide the clip mask and the states) use the three
// products; the two TANGENT rows (which only enter the divergence) use fp16 activations with hi+lo weights (two
// products): measured effect on the log-density <= 1.4e-4 relative (budget 1e-3), see DESIGN.md.
//
// Mapping: one persistent CTA per SM, 128 samples per tile, sample s <-> TMEM lane s.  Activations never touch shared
// memory: compute threads write them as fp16 pairs straight into TMEM (tcgen05.st) where tcgen05.mma reads them as the
// A operand; the weights (fp16 hi/lo images, 168 KB) sit in shared memory for the whole launch as the B operand in the
// canonical no-swizzle K-major layout; accumulators are two 128-column fp32 buffers in TMEM.  Each layer-2 epilogue
// converts an accumulator in place of the operand that produced it (that operand is dead once its MMAs completed), so
// the whole step fits the 512 TMEM columns:
//     cols   0..127  A_V   value operand, per 16 features: 8 columns of hi pairs then 8 columns of lo pairs
//     cols 128..191  A_T0  d/dtheta operand (hi only)      cols 192..255  A_T1  d/dv operand (hi only)
//     cols 256..383  D0    accumulator                      cols 384..511  D1    accumulator
// Warps 0-7 are compute warps (warp w owns TMEM lanes 32*(w%4).., feature half w/4), warp 8 issues all MMAs from one
// elected thread and signals completion through tcgen05.commit -> mbarrier; compute warps signal operand readiness /
// accumulator drain through mbarriers after tcgen05.wait + tcgen05.fence.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "dp_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr float WSCALE = 16.0f;             // weights are stored x16 so that w_lo stays a normal fp16 number
constexpr float INV_WSCALE = 1.0f / 16.0f;
constexpr int N3A = 48, N3B = 32;           // layer-3 MMA N (net b padded 24 -> 32: UMMA M=128 needs N % 16 == 0)

constexpr int COL_AV = 0, COL_AT0 = 128, COL_AT1 = 192, COL_D0 = 256, COL_D1 = 384;

// ---- shared-memory image (identical byte layout in global memory, built once per controller) --------------------------
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
constexpr int OFF_SMALL = OFF_W3B_LO + W3B_BYTES;   // fp32: per net W1[128][4] b1[128] b2[128] b3[64]
constexpr int SMALL_NET_FLOATS = 128 * 4 + 128 + 128 + 64 + 64 + 64;
constexpr int SP_W1 = 0, SP_B1 = 512, SP_B2 = 640, SP_B3 = 768;
constexpr int SP_C0H = 832, SP_C1H = 896;  // half2 pairs of W1[:,0] and W1[:,1] (tangent seeds of layer 1)
constexpr int IMAGE_BYTES = OFF_SMALL + 2 * SMALL_NET_FLOATS * 4;
// run-time scratch behind the image
constexpr int OFF_IN = IMAGE_BYTES;                  // float [2][128][8]   (two tiles in flight)
constexpr int OFF_OB = OFF_IN + 2 * TILE * 8 * 4;    // float [48][128]
constexpr int OFF_OUT = OFF_OB + 48 * TILE * 4;      // float [2][6][128]
constexpr int OFF_PART = OFF_OUT + 2 * 6 * TILE * 4;  // float [4 groups][4][128]  partial head sums
constexpr int OFF_BAR = OFF_PART + 4 * 4 * TILE * 4;  // 16 mbarriers + tmem pointer
constexpr int SMEM_BYTES = OFF_BAR + 16 * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(IMAGE_BYTES % 16 == 0, "image copied with 16-byte vectors");

enum { E1 = 0, E2A, E2B, E3A, E3B, E4A, E4B, E5, F1, F2, F3, F4, NUM_BARS };
static_assert(NUM_BARS <= 16, "barrier storage");

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
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
// true in exactly one lane of a fully converged warp (elect.sync); the pattern ptxas recognises for single-thread
// issue of uniform-datapath instructions such as tcgen05.mma / tcgen05.commit
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16, issued by one thread
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
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
constexpr uint32_t B_LBO = 128;           // K-adjacent core matrices are contiguous
constexpr uint32_t B_SBO = 16 * 128;      // K = 128 -> 16 core matrices per 8-row group
constexpr uint32_t B_KSTEP = 2 * B_LBO;   // one MMA consumes K = 16 = two core matrices

// ---- math helpers -------------------------------------------------------------------------------------------------------
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
// tanh(x) = 1 - 2 / (1 + e^(2x)); absolute error ~1e-7 (MUFU.EX2 + MUFU.RCP), saturates cleanly at +-1
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(x * 2.8853900817779268f);
    return fmaf(-2.0f, rcp_approx(1.0f + e), 1.0f);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);  // a -> low half (even k), b -> high half
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void split_h2(float a, float b, uint32_t &hi, uint32_t &lo) {
    __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = *reinterpret_cast<uint32_t *>(&h);
    lo = pack_h2(a - f.x, b - f.y);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }

// ---- compute-warp stages --------------------------------------------------------------------------------------------------
// NG warp groups share the 128 features of a sample: thread (lane quadrant q, group g) owns features [FPT*g, FPT*(g+1)).
template <int NG>
struct Cfg {
    static constexpr int NCW = 4 * NG;        // compute warps
    static constexpr int NCT = NCW * 32;      // compute threads
    static constexpr int NTHREADS = NCT + 32; // + MMA warp
    static constexpr int FPT = 128 / NG;      // features per thread
    static constexpr int CPT = FPT / 16;      // 16-feature chunks per thread
};

// one arrival per warp: every lane has completed (tcgen05.wait) and fenced its own TMEM accesses before the warp syncs
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// layer 1 on CUDA cores: in(4) -> this thread's features, value (hi/lo) + two tangents -> TMEM operands
template <int CPT>
__device__ __forceinline__ void stage_l1(const float *sp, const float4 in, uint32_t tm_lane, int g) {
    const float4 *W1 = reinterpret_cast<const float4 *>(sp + SP_W1);
    const float *b1 = sp + SP_B1;
#pragma unroll 1
    for (int i = 0; i < CPT; ++i) {
        const int c = CPT * g + i;  // 16-feature chunk
        uint32_t av[16], at0[8], at1[8];
#pragma unroll
        for (int pq = 0; pq < 8; ++pq) {
            float hv[2], t0[2], t1[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 16 * c + 2 * pq + e;
                const float4 w = W1[j];
                float pre = w.x * in.x;
                pre = fmaf(w.y, in.y, pre);
                pre = fmaf(w.z, in.z, pre);
                pre = fmaf(w.w, in.w, pre);
                pre += b1[j];
                const float t = tanh_fast(pre);
                const float gg = fmaf(-t, t, 1.0f);
                hv[e] = t;
                t0[e] = gg * w.x;
                t1[e] = gg * w.y;
            }
            split_h2(hv[0], hv[1], av[pq], av[8 + pq]);
            at0[pq] = pack_h2(t0[0], t0[1]);
            at1[pq] = pack_h2(t1[0], t1[1]);
        }
        tc_st16(tm_lane + COL_AV + 16 * c, av);
        tc_st8(tm_lane + COL_AT0 + 8 * c, at0);
        tc_st8(tm_lane + COL_AT1 + 8 * c, at1);
    }
}

// this thread's slice of an accumulator -> registers (the accumulator can be handed back to the MMA warp right away)
template <int CPT>
__device__ __forceinline__ void load_slice(uint32_t tm_lane, int col_d, int g, uint32_t (&d)[16 * CPT]) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) tc_ld16(tm_lane + col_d + 16 * (CPT * g + i), &d[16 * i]);
    tc_wait_ld();
}

// layer-2 epilogue, value rows: h2 = tanh(D/16 + b2) -> A_V (hi/lo); keeps g2/16 packed for the tangent rows
template <int CPT>
__device__ __forceinline__ void stage_epi_value(const float *sp, uint32_t tm_lane, int g, const uint32_t (&d)[16 * CPT],
                                                uint32_t (&gs)[8 * CPT]) {
    const float *b2 = sp + SP_B2;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = CPT * g + i;
        uint32_t av[16];
#pragma unroll
        for (int pq = 0; pq < 8; ++pq) {
            float hv[2], gg[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float pre = fmaf(__uint_as_float(d[16 * i + 2 * pq + e]), INV_WSCALE, b2[16 * c + 2 * pq + e]);
                const float t = tanh_fast(pre);
                hv[e] = t;
                gg[e] = fmaf(-t, t, 1.0f) * INV_WSCALE;
            }
            split_h2(hv[0], hv[1], av[pq], av[8 + pq]);
            gs[8 * i + pq] = pack_h2(gg[0], gg[1]);
        }
        tc_st16(tm_lane + COL_AV + 16 * c, av);
    }
}

// layer-2 epilogue, tangent rows: (g2/16) * D -> A_T (hi only), in packed half arithmetic
template <int CPT>
__device__ __forceinline__ void stage_epi_tangent(uint32_t tm_lane, int g, int col_a, const uint32_t (&d)[16 * CPT],
                                                  const uint32_t (&gs)[8 * CPT]) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = CPT * g + i;
        uint32_t at[8];
#pragma unroll
        for (int pq = 0; pq < 8; ++pq) {
            const __half2 dh = __floats2half2_rn(__uint_as_float(d[16 * i + 2 * pq]), __uint_as_float(d[16 * i + 2 * pq + 1]));
            const __half2 r = __hmul2(dh, *reinterpret_cast<const __half2 *>(&gs[8 * i + pq]));
            at[pq] = *reinterpret_cast<const uint32_t *>(&r);
        }
        tc_st8(tm_lane + col_a + 8 * c, at);
    }
}

// ---- per-sample state (owner thread) -----------------------------------------------------------------------------------------
struct SampleState {
    float x0, x1, x2, x3, x4, rho, margin, ur0, ur1;
};

__device__ __forceinline__ void sample_init(const RolloutParams &p, SampleState &st, long long nl) {
    const int T1 = p.T + 1;
    st.x0 = __ldg(p.xe0 + nl * 5 + 0) + __ldg(p.xref + 0 * T1);
    st.x1 = __ldg(p.xe0 + nl * 5 + 1) + __ldg(p.xref + 1 * T1);
    st.x2 = __ldg(p.xe0 + nl * 5 + 2) + __ldg(p.xref + 2 * T1);
    st.x3 = __ldg(p.xe0 + nl * 5 + 3) + __ldg(p.xref + 3 * T1);
    st.x4 = __ldg(p.xe0 + nl * 5 + 4) + __ldg(p.xref + 4 * T1);
    st.rho = p.rho0 ? __ldg(p.rho0 + nl) : ((p.flags & DP_LOG_DENSITY) ? 0.0f : 1.0f);
    st.margin = 1e30f;
    st.ur0 = st.ur1 = 0.0f;
}

// start of time index t: stage the trajectory slice (if stored) and the controller inputs of this sample
__device__ __forceinline__ void sample_step_start(const RolloutParams &p, SampleState &st, int t, float *sIn_t, float *sOut_t,
                                                  int s, bool valid) {
    const int T = p.T, T1 = p.T + 1;
    const bool logd = p.flags & DP_LOG_DENSITY, dens = p.flags & DP_DENSITY, cut = p.flags & DP_CUT;
    const float xr0 = __ldg(p.xref + 0 * T1 + t), xr1 = __ldg(p.xref + 1 * T1 + t);
    const float xr2 = __ldg(p.xref + 2 * T1 + t), xr3 = __ldg(p.xref + 3 * T1 + t);
    if ((t % p.stride) == 0) {
        if (p.flags & DP_ABS_STATE) {
            sOut_t[0 * TILE + s] = st.x0; sOut_t[1 * TILE + s] = st.x1; sOut_t[2 * TILE + s] = st.x2;
            sOut_t[3 * TILE + s] = st.x3; sOut_t[4 * TILE + s] = st.x4;
        } else {
            sOut_t[0 * TILE + s] = st.x0 - xr0; sOut_t[1 * TILE + s] = st.x1 - xr1; sOut_t[2 * TILE + s] = st.x2 - xr2;
            sOut_t[3 * TILE + s] = st.x3 - xr3; sOut_t[4 * TILE + s] = st.x4 - __ldg(p.xref + 4 * T1 + t);
        }
        float r = st.rho;
        if (cut && dens) {
            int fl = 0;
            if (r > 1e30f) { r = 1e30f; fl |= 1; }
            if (!logd && r < 0.0f) { r = 0.0f; fl |= 1; }
            if (r != r) { r = 1e30f; fl |= 2; }
            if (fl && valid && p.status) {
                if (fl & 1) atomicOr(p.status, 1);
                if (fl & 2) atomicOr(p.status + 1, 1);
            }
        }
        sOut_t[5 * TILE + s] = r;
    }
    if (t < T) {
        // Controller.__call__ (sytem_CAR.py:33-34) and the network input of U_FUNC.forward (model_CAR.py:24)
        const float e0 = st.x0 - xr0, e1 = st.x1 - xr1, e3 = st.x3 - xr3;
        const float e2 = __fadd_rn(__fsub_rn(st.x2, xr2), st.x4);
        float4 *dst = reinterpret_cast<float4 *>(sIn_t + 8 * s);
        dst[0] = make_float4(st.x2, st.x3, st.x2 - e2, st.x3 - e3);
        dst[1] = make_float4(e0, e1, e2, e3);
        st.ur0 = __ldg(p.uref + t);
        st.ur1 = __ldg(p.uref + T + t);
    }
}

// one Euler step of state and (log-)density from the head sums (get_next_x :124, get_next_rho(log) :136-157)
__device__ __forceinline__ void sample_advance(const RolloutParams &p, SampleState &st, float pu0, float pu1, float pd0, float pd1) {
    const float dt = p.dt;
    const float u0r = pu0 + st.ur0, u1r = pu1 + st.ur1;
    st.margin = fminf(st.margin, fminf(fabsf(fabsf(u0r) - 3.0f), fabsf(fabsf(u1r) - 3.0f)));
    if (p.flags & DP_DENSITY) {
        const float m0 = (u0r >= -3.0f && u0r <= 3.0f) ? 1.0f : 0.0f;  // inclusive, as the backward of torch.clip
        const float m1 = (u1r >= -3.0f && u1r <= 3.0f) ? 1.0f : 0.0f;
        const float div = __fadd_rn(__fmul_rn(m0, pd0), __fmul_rn(m1, pd1));
        if (p.flags & DP_LOG_DENSITY)
            st.rho = __fadd_rn(st.rho, logf(__fsub_rn(1.0f, __fmul_rn(div, dt))));
        else
            st.rho = __fadd_rn(st.rho, __fmul_rn(__fmul_rn(-div, st.rho), dt));
    }
    const float u0 = fminf(fmaxf(u0r, -3.0f), 3.0f), u1 = fminf(fmaxf(u1r, -3.0f), 3.0f);
    float sn, cs;
    sincosf(st.x2, &sn, &cs);
    st.x0 = __fadd_rn(st.x0, __fmul_rn(__fmul_rn(st.x3, cs), dt));
    st.x1 = __fadd_rn(st.x1, __fmul_rn(__fmul_rn(st.x3, sn), dt));
    st.x2 = __fadd_rn(st.x2, __fmul_rn(u0, dt));
    st.x3 = __fadd_rn(st.x3, __fmul_rn(u1, dt));
}

// coalesced store of a staged slice (128 consecutive samples per row) by one warp group
__device__ __forceinline__ void store_slice(const RolloutParams &p, const float *sOut_t, long long tile, int t, int j) {
    const long long nn = tile * TILE + j;
    if (nn < p.N) {
        const int nd = (p.flags & DP_XY_ONLY) ? 2 : 5;
        const long long slot = t / p.stride;
        if (p.x_out)
            for (int d = 0; d < nd; ++d) p.x_out[(slot * nd + d) * p.ldn + nn] = sOut_t[d * TILE + j];
        if (p.rho_out && (p.flags & DP_DENSITY)) p.rho_out[slot * p.ldn + nn] = sOut_t[5 * TILE + j];
    }
}

// layer-3 read-out of net b: w2 (2x12), d w2[0,:]/d theta, d w2[1,:]/d v  -> sOb[48][128]
__device__ __forceinline__ void readout_b(uint32_t tm_lane, const float *b3, float *sOb, int s) {
    uint32_t dv[16], dw[16];
    tc_ld16(tm_lane + COL_D1, dv);
    tc_ld16(tm_lane + COL_D1 + 16, dw);
    tc_wait_ld();
#pragma unroll
    for (int k = 0; k < 16; ++k) sOb[k * TILE + s] = fmaf(__uint_as_float(dv[k]), INV_WSCALE, b3[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) sOb[(16 + k) * TILE + s] = fmaf(__uint_as_float(dw[k]), INV_WSCALE, b3[16 + k]);
    tc_ld16(tm_lane + COL_D1 + 64, dv);  // theta tangent, columns 0..15 (row 0 = 0..11)
    tc_ld16(tm_lane + COL_D0, dw);       // v tangent, columns 0..15
    tc_wait_ld();
#pragma unroll
    for (int k = 0; k < 12; ++k) sOb[(24 + k) * TILE + s] = __uint_as_float(dv[k]) * INV_WSCALE;
#pragma unroll
    for (int k = 12; k < 16; ++k) sOb[(36 + k - 12) * TILE + s] = __uint_as_float(dw[k]) * INV_WSCALE;
    tc_ld16(tm_lane + COL_D0 + 16, dw);  // v tangent, columns 16..31 (row 1 continues to 23)
    tc_wait_ld();
#pragma unroll
    for (int k = 0; k < 8; ++k) sOb[(40 + k) * TILE + s] = __uint_as_float(dw[k]) * INV_WSCALE;
}

// layer-3 read-out of net a: w1 (12x4) and its two tangents, four z-rows per pass; head sums
__device__ __forceinline__ void readout_a(uint32_t tm_lane, const float *b3, const float *sOb, int s, const float4 xe, float &pu0,
                                          float &pu1, float &pd0, float &pd1) {
    pu0 = pu1 = pd0 = pd1 = 0.0f;
#pragma unroll 1
    for (int ps = 0; ps < 3; ++ps) {
        uint32_t v[16], a[16], b[16];
        tc_ld16(tm_lane + COL_D1 + 16 * ps, v);
        tc_ld16(tm_lane + COL_D1 + 64 + 16 * ps, a);
        tc_ld16(tm_lane + COL_D0 + 16 * ps, b);
        tc_wait_ld();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = 4 * ps + rr;
            const float w0 = fmaf(__uint_as_float(v[4 * rr + 0]), INV_WSCALE, b3[4 * r + 0]);
            const float w1 = fmaf(__uint_as_float(v[4 * rr + 1]), INV_WSCALE, b3[4 * r + 1]);
            const float w2 = fmaf(__uint_as_float(v[4 * rr + 2]), INV_WSCALE, b3[4 * r + 2]);
            const float w3 = fmaf(__uint_as_float(v[4 * rr + 3]), INV_WSCALE, b3[4 * r + 3]);
            float z = w0 * xe.x;
            z = fmaf(w1, xe.y, z); z = fmaf(w2, xe.z, z); z = fmaf(w3, xe.w, z);
            float dzt = __uint_as_float(a[4 * rr + 0]) * xe.x;
            dzt = fmaf(__uint_as_float(a[4 * rr + 1]), xe.y, dzt);
            dzt = fmaf(__uint_as_float(a[4 * rr + 2]), xe.z, dzt);
            dzt = fmaf(__uint_as_float(a[4 * rr + 3]), xe.w, dzt);
            dzt = fmaf(dzt, INV_WSCALE, w2);  // + d xe2 / d theta * w1[r][2]
            float dzv = __uint_as_float(b[4 * rr + 0]) * xe.x;
            dzv = fmaf(__uint_as_float(b[4 * rr + 1]), xe.y, dzv);
            dzv = fmaf(__uint_as_float(b[4 * rr + 2]), xe.z, dzv);
            dzv = fmaf(__uint_as_float(b[4 * rr + 3]), xe.w, dzv);
            dzv = fmaf(dzv, INV_WSCALE, w3);  // + d xe3 / d v * w1[r][3]
            const float tz = tanhf(z);
            const float gz = fmaf(-tz, tz, 1.0f);
            const float w20 = sOb[r * TILE + s], w21 = sOb[(12 + r) * TILE + s];
            pu0 = fmaf(w20, tz, pu0);
            pu1 = fmaf(w21, tz, pu1);
            pd0 += sOb[(24 + r) * TILE + s] * tz + w20 * (gz * dzt);
            pd1 += sOb[(36 + r) * TILE + s] * tz + w21 * (gz * dzv);
        }
    }
}


// ---- layer-3 read-out distributed over the four warp groups (two-tile kernel) ------------------------------------------------
// net b: each group drains its share of {w2 (24 values), d w2[0,:]/d theta (12), d w2[1,:]/d v (12)} into sOb[48][128]
__device__ __forceinline__ void readout_b_load(uint32_t tm_lane, int g, uint32_t (&d)[32]) {
    if (g == 0) tc_ld16(tm_lane + COL_D1, d);
    else if (g == 1) tc_ld16(tm_lane + COL_D1 + 16, d);
    else if (g == 2) tc_ld16(tm_lane + COL_D1 + 64, d);
    else { tc_ld16(tm_lane + COL_D0, d); tc_ld16(tm_lane + COL_D0 + 16, d + 16); }
    tc_wait_ld();
}
__device__ __forceinline__ void readout_b_store(int g, const uint32_t (&d)[32], const float *b3, float *sOb, int s) {
    if (g == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) sOb[k * TILE + s] = fmaf(__uint_as_float(d[k]), INV_WSCALE, b3[k]);
    } else if (g == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sOb[(16 + k) * TILE + s] = fmaf(__uint_as_float(d[k]), INV_WSCALE, b3[16 + k]);
    } else if (g == 2) {
#pragma unroll
        for (int k = 0; k < 12; ++k) sOb[(24 + k) * TILE + s] = __uint_as_float(d[k]) * INV_WSCALE;
    } else {
#pragma unroll
        for (int k = 12; k < 24; ++k) sOb[(36 + k - 12) * TILE + s] = __uint_as_float(d[k]) * INV_WSCALE;
    }
}
// net a: group g owns z-rows 3g..3g+2 (columns 12g..12g+11 of the value / theta / v accumulators); one row at a time
__device__ __forceinline__ void readout_a_row(uint32_t tm_lane, int r, const float *b3, const float *sOb, const float4 xe, int s,
                                              float &pu0, float &pu1, float &pd0, float &pd1) {
    uint32_t v[4], a[4], b[4];
    tc_ld4(tm_lane + COL_D1 + 4 * r, v);
    tc_ld4(tm_lane + COL_D1 + 64 + 4 * r, a);
    tc_ld4(tm_lane + COL_D0 + 4 * r, b);
    tc_wait_ld();
    const float w0 = fmaf(__uint_as_float(v[0]), INV_WSCALE, b3[4 * r + 0]);
    const float w1 = fmaf(__uint_as_float(v[1]), INV_WSCALE, b3[4 * r + 1]);
    const float w2 = fmaf(__uint_as_float(v[2]), INV_WSCALE, b3[4 * r + 2]);
    const float w3 = fmaf(__uint_as_float(v[3]), INV_WSCALE, b3[4 * r + 3]);
    float z = w0 * xe.x;
    z = fmaf(w1, xe.y, z); z = fmaf(w2, xe.z, z); z = fmaf(w3, xe.w, z);
    float dzt = __uint_as_float(a[0]) * xe.x;
    dzt = fmaf(__uint_as_float(a[1]), xe.y, dzt);
    dzt = fmaf(__uint_as_float(a[2]), xe.z, dzt);
    dzt = fmaf(__uint_as_float(a[3]), xe.w, dzt);
    dzt = fmaf(dzt, INV_WSCALE, w2);
    float dzv = __uint_as_float(b[0]) * xe.x;
    dzv = fmaf(__uint_as_float(b[1]), xe.y, dzv);
    dzv = fmaf(__uint_as_float(b[2]), xe.z, dzv);
    dzv = fmaf(__uint_as_float(b[3]), xe.w, dzv);
    dzv = fmaf(dzv, INV_WSCALE, w3);
    const float tz = tanh_fast(z);
    const float gz = fmaf(-tz, tz, 1.0f);
    const float w20 = sOb[r * TILE + s], w21 = sOb[(12 + r) * TILE + s];
    pu0 = fmaf(w20, tz, pu0);
    pu1 = fmaf(w21, tz, pu1);
    pd0 += sOb[(24 + r) * TILE + s] * tz + w20 * (gz * dzt);
    pd1 += sOb[(36 + r) * TILE + s] * tz + w21 * (gz * dzv);
}
// owner: sum the four groups' partial head sums, Euler step, stage the next time index
__device__ __forceinline__ void finalize_step(const RolloutParams &p, SampleState &st, const float *sPart, int t_next, float *sIn_t,
                                              float *sOut_t, int s, bool valid) {
    float acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        acc[k] = (sPart[(0 * 4 + k) * TILE + s] + sPart[(1 * 4 + k) * TILE + s]) + (sPart[(2 * 4 + k) * TILE + s] + sPart[(3 * 4 + k) * TILE + s]);
    sample_advance(p, st, acc[0], acc[1], acc[2], acc[3]);
    sample_step_start(p, st, t_next, sIn_t, sOut_t, s, valid);
}

// layer 1 for one 16-feature chunk into registers (operands are written to TMEM later, when the MMA warp is done with them):
// value hi/lo pairs (16 registers) + g = 1 - h^2 as half2 pairs (8 registers)
__device__ __forceinline__ void l1_chunk(const float *sp, const float4 in, int c, uint32_t (&av)[16], uint32_t (&gp)[8]) {
    const float4 *W1 = reinterpret_cast<const float4 *>(sp + SP_W1);
    const float *b1 = sp + SP_B1;
#pragma unroll
    for (int pq = 0; pq < 8; ++pq) {
        float hv[2], gg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int j = 16 * c + 2 * pq + e;
            const float4 w = W1[j];
            float pre = fmaf(w.x, in.x, b1[j]);
            pre = fmaf(w.y, in.y, pre);
            pre = fmaf(w.z, in.z, pre);
            pre = fmaf(w.w, in.w, pre);
            const float t = tanh_fast(pre);
            hv[e] = t;
            gg[e] = fmaf(-t, t, 1.0f);
        }
        split_h2(hv[0], hv[1], av[pq], av[8 + pq]);
        gp[pq] = pack_h2(gg[0], gg[1]);
    }
}
// write a prefetched layer-1 chunk to the TMEM operands; the tangent seeds g * W1[:,0], g * W1[:,1] are formed here in
// packed half arithmetic (the tangent operands are fp16 anyway)
__device__ __forceinline__ void l1_store(const float *sp, uint32_t tm_lane, int c, const uint32_t (&av)[16], const uint32_t (&gp)[8]) {
    const uint32_t *c0h = reinterpret_cast<const uint32_t *>(sp + SP_C0H), *c1h = reinterpret_cast<const uint32_t *>(sp + SP_C1H);
    uint32_t at0[8], at1[8];
#pragma unroll
    for (int pq = 0; pq < 8; ++pq) {
        const uint32_t k0 = c0h[8 * c + pq], k1 = c1h[8 * c + pq];
        const __half2 g2 = *reinterpret_cast<const __half2 *>(&gp[pq]);
        const __half2 r0 = __hmul2(g2, *reinterpret_cast<const __half2 *>(&k0));
        const __half2 r1 = __hmul2(g2, *reinterpret_cast<const __half2 *>(&k1));
        at0[pq] = *reinterpret_cast<const uint32_t *>(&r0);
        at1[pq] = *reinterpret_cast<const uint32_t *>(&r1);
    }
    tc_st16(tm_lane + COL_AV + 16 * c, av);
    tc_st8(tm_lane + COL_AT0 + 8 * c, at0);
    tc_st8(tm_lane + COL_AT1 + 8 * c, at1);
}

// per-chunk layer-2 epilogues for the two-tile kernel (16 accumulator registers live at a time)
__device__ __forceinline__ void epi_value_chunk(const float *sp, uint32_t tm_lane, int c, const uint32_t (&d)[16], uint32_t *gs8) {
    const float *b2 = sp + SP_B2;
    uint32_t av[16];
#pragma unroll
    for (int pq = 0; pq < 8; ++pq) {
        float hv[2], gg[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float pre = fmaf(__uint_as_float(d[2 * pq + e]), INV_WSCALE, b2[16 * c + 2 * pq + e]);
            const float t = tanh_fast(pre);
            hv[e] = t;
            gg[e] = fmaf(-t, t, 1.0f) * INV_WSCALE;
        }
        split_h2(hv[0], hv[1], av[pq], av[8 + pq]);
        gs8[pq] = pack_h2(gg[0], gg[1]);
    }
    tc_st16(tm_lane + COL_AV + 16 * c, av);
}
__device__ __forceinline__ void epi_tangent_chunk(uint32_t tm_lane, int col_a, int c, const uint32_t (&d)[16], const uint32_t *gs8) {
    uint32_t at[8];
#pragma unroll
    for (int pq = 0; pq < 8; ++pq) {
        const __half2 dh = __floats2half2_rn(__uint_as_float(d[2 * pq]), __uint_as_float(d[2 * pq + 1]));
        const __half2 r = __hmul2(dh, *reinterpret_cast<const __half2 *>(&gs8[pq]));
        at[pq] = *reinterpret_cast<const uint32_t *>(&r);
    }
    tc_st8(tm_lane + col_a + 8 * c, at);
}

// ---- MMA-warp stages (one elected thread) -----------------------------------------------------------------------------------
// The K loop is fully unrolled and the shared-memory descriptors are built once per matrix: advancing K by one MMA
// (two core matrices = 256 bytes) is an immediate add of 16 to the descriptor's address field, so the issue loop is a
// handful of uniform-datapath instructions per tcgen05.mma instead of a dependent descriptor rebuild.
constexpr uint64_t B_KSTEP_DESC = B_KSTEP >> 4;

// D (+)= A_V(hi,lo) x W(hi,lo): three products per K-chunk
__device__ __forceinline__ void mma_value(uint32_t tm, int col_d, uint64_t desc_hi, uint64_t desc_lo, uint32_t idesc) {
    const uint32_t d = tm + col_d, a0 = tm + COL_AV;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint64_t bh = desc_hi + c * B_KSTEP_DESC, bl = desc_lo + c * B_KSTEP_DESC;
        tc_mma_ts(d, a0 + 16 * c, bh, idesc, c > 0);
        tc_mma_ts(d, a0 + 16 * c, bl, idesc, 1);
        tc_mma_ts(d, a0 + 16 * c + 8, bh, idesc, 1);
    }
}
// D (+)= A_T(hi) x W(hi,lo): two products per K-chunk
__device__ __forceinline__ void mma_tangent(uint32_t tm, int col_d, int col_a, uint64_t desc_hi, uint64_t desc_lo, uint32_t idesc) {
    const uint32_t d = tm + col_d, a0 = tm + col_a;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        tc_mma_ts(d, a0 + 8 * c, desc_hi + c * B_KSTEP_DESC, idesc, c > 0);
        tc_mma_ts(d, a0 + 8 * c, desc_lo + c * B_KSTEP_DESC, idesc, 1);
    }
}

// Barrier protocol per net (count NCW for E*, 1 for F*; each completes one phase per net iteration):
//   E1  layer-1 operands written          F1 value accumulator (D0) complete      F2 theta accumulator (D1) complete
//   E2A D0 drained to registers           E2B A_V rewritten (layer-3 value operand)
//   E3A D1 drained                        E3B A_T0 rewritten                     F3 v accumulator (D0) complete
//   E4A D0 drained                        E4B A_T1 rewritten                     F4 layer-3 accumulators complete
//   E5  layer-3 accumulators read out
template <int NG>
__global__ void __launch_bounds__(Cfg<NG>::NTHREADS, 1) rollout_tc_kernel(RolloutParams p) {
    using C = Cfg<NG>;
    constexpr int NCW = C::NCW, NCT = C::NCT, NTHREADS = C::NTHREADS, CPT = C::CPT;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *sIn = reinterpret_cast<float *>(smem + OFF_IN);
    float *sOb = reinterpret_cast<float *>(smem + OFF_OB);
    float *sOut = reinterpret_cast<float *>(smem + OFF_OUT);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 16 * 8);
    const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (i))

    // weight image -> shared memory (flat copy), barriers, TMEM
    {
        const int4 *src = reinterpret_cast<const int4 *>(p.tc);
        int4 *dst = reinterpret_cast<int4 *>(smem);
        for (int i = tid; i < IMAGE_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int i = E1; i <= E5; ++i) mbar_init(BAR(i), NCW);
        for (int i = F1; i <= F4; ++i) mbar_init(BAR(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NCW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // make the generic-proxy writes of the weight image visible to the async proxy (tcgen05.mma reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr_s;

    const int T = p.T, T1 = p.T + 1;
    const long long num_tiles = (p.N + TILE - 1) / TILE;
    const uint32_t sbase = smem_u32(smem);

    if (warp == NCW) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc128 = make_idesc(128), idesc3a = make_idesc(N3A), idesc3b = make_idesc(N3B);
            uint32_t it = 0;
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int i = 0; i < T; ++i) {
#pragma unroll 1
                    for (int net = 0; net < 2; ++net, ++it) {  // net 0 = b (w2), net 1 = a (w1)
                        const uint32_t par = it & 1u;
                        const uint64_t w2h = make_bdesc(sbase + (net == 0 ? OFF_W2B_HI : OFF_W2A_HI), B_LBO, B_SBO);
                        const uint64_t w2l = make_bdesc(sbase + (net == 0 ? OFF_W2B_LO : OFF_W2A_LO), B_LBO, B_SBO);
                        const uint64_t w3h = make_bdesc(sbase + (net == 0 ? OFF_W3B_HI : OFF_W3A_HI), B_LBO, B_SBO);
                        const uint64_t w3l = make_bdesc(sbase + (net == 0 ? OFF_W3B_LO : OFF_W3A_LO), B_LBO, B_SBO);
                        const uint32_t idesc3 = net == 0 ? idesc3b : idesc3a;
                        if (it > 0) mbar_wait(BAR(E5), par ^ 1u);  // previous layer-3 results read out
                        mbar_wait(BAR(E1), par);
                        tc_fence_after();
                        mma_value(tm, COL_D0, w2h, w2l, idesc128);
                        tc_commit(BAR(F1));
                        mma_tangent(tm, COL_D1, COL_AT0, w2h, w2l, idesc128);
                        tc_commit(BAR(F2));
                        mbar_wait(BAR(E2A), par);
                        tc_fence_after();
                        mma_tangent(tm, COL_D0, COL_AT1, w2h, w2l, idesc128);
                        tc_commit(BAR(F3));
                        mbar_wait(BAR(E2B), par);
                        mbar_wait(BAR(E3A), par);
                        tc_fence_after();
                        mma_value(tm, COL_D1, w3h, w3l, idesc3);
                        mbar_wait(BAR(E3B), par);
                        tc_fence_after();
                        mma_tangent(tm, COL_D1 + 64, COL_AT0, w3h, w3l, idesc3);
                        mbar_wait(BAR(E4A), par);
                        mbar_wait(BAR(E4B), par);
                        tc_fence_after();
                        mma_tangent(tm, COL_D0, COL_AT1, w3h, w3l, idesc3);
                        tc_commit(BAR(F4));
                    }
                }
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q = warp & 3, g = warp >> 2;
        const int s = 32 * q + lane;  // sample within the tile == TMEM lane
        const uint32_t tm_lane = tm + ((uint32_t)(32 * q) << 16);
        const bool owner = g == 0;  // owns the sample's state, the layer-3 read-out and the head
        const float *sp_a = reinterpret_cast<const float *>(smem + OFF_SMALL);
        const float *sp_b = sp_a + SMALL_NET_FLOATS;
        const bool logd = p.flags & DP_LOG_DENSITY, dens = p.flags & DP_DENSITY, cut = p.flags & DP_CUT;
        const bool abs_state = p.flags & DP_ABS_STATE;
        const int ndim_out = (p.flags & DP_XY_ONLY) ? 2 : 5;
        const float dt = p.dt;
        uint32_t it = 0;

        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");  // the previous tile's last slice has left sOut
            const long long n = tile * TILE + s;
            const bool valid = n < p.N;
            const long long nl = valid ? n : p.N - 1;
            float x0 = 0, x1 = 0, x2 = 0, x3 = 0, x4 = 0, rho = 0, margin = 1e30f;
            if (owner) {
                x0 = __ldg(p.xe0 + nl * 5 + 0) + __ldg(p.xref + 0 * T1);
                x1 = __ldg(p.xe0 + nl * 5 + 1) + __ldg(p.xref + 1 * T1);
                x2 = __ldg(p.xe0 + nl * 5 + 2) + __ldg(p.xref + 2 * T1);
                x3 = __ldg(p.xe0 + nl * 5 + 3) + __ldg(p.xref + 3 * T1);
                x4 = __ldg(p.xe0 + nl * 5 + 4) + __ldg(p.xref + 4 * T1);
                rho = p.rho0 ? __ldg(p.rho0 + nl) : (logd ? 0.0f : 1.0f);
            }
            for (int i = 0;; ++i) {
                const bool store_now = (i % p.stride) == 0;
                float ur0 = 0, ur1 = 0;
                if (owner) {
                    const float xr0 = __ldg(p.xref + 0 * T1 + i), xr1 = __ldg(p.xref + 1 * T1 + i);
                    const float xr2 = __ldg(p.xref + 2 * T1 + i), xr3 = __ldg(p.xref + 3 * T1 + i);
                    if (store_now) {
                        if (abs_state) {
                            sOut[0 * TILE + s] = x0; sOut[1 * TILE + s] = x1; sOut[2 * TILE + s] = x2;
                            sOut[3 * TILE + s] = x3; sOut[4 * TILE + s] = x4;
                        } else {
                            sOut[0 * TILE + s] = x0 - xr0; sOut[1 * TILE + s] = x1 - xr1; sOut[2 * TILE + s] = x2 - xr2;
                            sOut[3 * TILE + s] = x3 - xr3; sOut[4 * TILE + s] = x4 - __ldg(p.xref + 4 * T1 + i);
                        }
                        float r = rho;
                        if (cut && dens) {
                            int fl = 0;
                            if (r > 1e30f) { r = 1e30f; fl |= 1; }
                            if (!logd && r < 0.0f) { r = 0.0f; fl |= 1; }
                            if (r != r) { r = 1e30f; fl |= 2; }
                            if (fl && valid && p.status) {
                                if (fl & 1) atomicOr(p.status, 1);
                                if (fl & 2) atomicOr(p.status + 1, 1);
                            }
                        }
                        sOut[5 * TILE + s] = r;
                    }
                    if (i < T) {
                        const float e0 = x0 - xr0, e1 = x1 - xr1, e3 = x3 - xr3;
                        const float e2 = __fadd_rn(__fsub_rn(x2, xr2), x4);
                        float4 *dst = reinterpret_cast<float4 *>(sIn + 8 * s);
                        dst[0] = make_float4(x2, x3, x2 - e2, x3 - e3);
                        dst[1] = make_float4(e0, e1, e2, e3);
                        ur0 = __ldg(p.uref + i);
                        ur1 = __ldg(p.uref + T + i);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");  // compute warps only: sIn / sOut visible
                if (store_now && g == 1) {
                    // coalesced store of the staged slice by warp group 1: 128 consecutive samples per row
                    const int j = tid - 128;  // 0..127
                    const long long nn = tile * TILE + j;
                    if (nn < p.N) {
                        const long long slot = i / p.stride;
                        if (p.x_out)
                            for (int d = 0; d < ndim_out; ++d) p.x_out[(slot * ndim_out + d) * p.ldn + nn] = sOut[d * TILE + j];
                        if (p.rho_out && dens) p.rho_out[slot * p.ldn + nn] = sOut[5 * TILE + j];
                    }
                }
                if (i == T) break;
                const float4 in = *reinterpret_cast<const float4 *>(sIn + 8 * s);

                float pu0 = 0, pu1 = 0, pd0 = 0, pd1 = 0;
#pragma unroll 1
                for (int net = 0; net < 2; ++net, ++it) {
                    const uint32_t par = it & 1u;
                    const float *sp = net == 0 ? sp_b : sp_a;
                    // layer 1 (the previous net's layer-3 MMAs completed before F4 was observed, so A_* are free)
                    stage_l1<CPT>(sp, in, tm_lane, g);
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(BAR(E1), lane);
                    // layer-2 epilogues: drain the accumulator slice first, hand the accumulator back, then compute
                    uint32_t d[16 * CPT], gs[8 * CPT];
                    mbar_wait(BAR(F1), par);
                    tc_fence_after();
                    load_slice<CPT>(tm_lane, COL_D0, g, d);
                    tc_fence_before();
                    warp_arrive(BAR(E2A), lane);
                    stage_epi_value<CPT>(sp, tm_lane, g, d, gs);
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(BAR(E2B), lane);
                    mbar_wait(BAR(F2), par);
                    tc_fence_after();
                    load_slice<CPT>(tm_lane, COL_D1, g, d);
                    tc_fence_before();
                    warp_arrive(BAR(E3A), lane);
                    stage_epi_tangent<CPT>(tm_lane, g, COL_AT0, d, gs);
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(BAR(E3B), lane);
                    mbar_wait(BAR(F3), par);
                    tc_fence_after();
                    load_slice<CPT>(tm_lane, COL_D0, g, d);
                    tc_fence_before();
                    warp_arrive(BAR(E4A), lane);
                    stage_epi_tangent<CPT>(tm_lane, g, COL_AT1, d, gs);
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(BAR(E4B), lane);
                    // layer-3 read-out
                    mbar_wait(BAR(F4), par);
                    tc_fence_after();
                    if (owner) {
                        const float *b3 = sp + SP_B3;
                        if (net == 0) {
                            // net b: w2 (2x12) value, d w2[0,:]/d theta, d w2[1,:]/d v  -> sOb[48][128]
                            uint32_t dv[16], dw[16];
                            tc_ld16(tm_lane + COL_D1, dv);
                            tc_ld16(tm_lane + COL_D1 + 16, dw);
                            tc_wait_ld();
#pragma unroll
                            for (int k = 0; k < 16; ++k) sOb[k * TILE + s] = fmaf(__uint_as_float(dv[k]), INV_WSCALE, b3[k]);
#pragma unroll
                            for (int k = 0; k < 8; ++k) sOb[(16 + k) * TILE + s] = fmaf(__uint_as_float(dw[k]), INV_WSCALE, b3[16 + k]);
                            tc_ld16(tm_lane + COL_D1 + 64, dv);  // theta tangent, columns 0..15 (row 0 = 0..11)
                            tc_ld16(tm_lane + COL_D0, dw);       // v tangent, columns 0..15
                            tc_wait_ld();
#pragma unroll
                            for (int k = 0; k < 12; ++k) sOb[(24 + k) * TILE + s] = __uint_as_float(dv[k]) * INV_WSCALE;
#pragma unroll
                            for (int k = 12; k < 16; ++k) sOb[(36 + k - 12) * TILE + s] = __uint_as_float(dw[k]) * INV_WSCALE;
                            tc_ld16(tm_lane + COL_D0 + 16, dw);  // v tangent, columns 16..31 (row 1 continues to 23)
                            tc_wait_ld();
#pragma unroll
                            for (int k = 0; k < 8; ++k) sOb[(40 + k) * TILE + s] = __uint_as_float(dw[k]) * INV_WSCALE;
                        } else {
                            // net a: w1 (12x4) and its two tangents, four z-rows per pass; head partial sums
                            const float4 xe = *reinterpret_cast<const float4 *>(sIn + 8 * s + 4);
#pragma unroll 1
                            for (int ps = 0; ps < 3; ++ps) {
                                uint32_t v[16], a[16], b[16];
                                tc_ld16(tm_lane + COL_D1 + 16 * ps, v);
                                tc_ld16(tm_lane + COL_D1 + 64 + 16 * ps, a);
                                tc_ld16(tm_lane + COL_D0 + 16 * ps, b);
                                tc_wait_ld();
#pragma unroll
                                for (int rr = 0; rr < 4; ++rr) {
                                    const int r = 4 * ps + rr;
                                    const float w0 = fmaf(__uint_as_float(v[4 * rr + 0]), INV_WSCALE, b3[4 * r + 0]);
                                    const float w1 = fmaf(__uint_as_float(v[4 * rr + 1]), INV_WSCALE, b3[4 * r + 1]);
                                    const float w2 = fmaf(__uint_as_float(v[4 * rr + 2]), INV_WSCALE, b3[4 * r + 2]);
                                    const float w3 = fmaf(__uint_as_float(v[4 * rr + 3]), INV_WSCALE, b3[4 * r + 3]);
                                    float z = w0 * xe.x;
                                    z = fmaf(w1, xe.y, z); z = fmaf(w2, xe.z, z); z = fmaf(w3, xe.w, z);
                                    float dzt = __uint_as_float(a[4 * rr + 0]) * xe.x;
                                    dzt = fmaf(__uint_as_float(a[4 * rr + 1]), xe.y, dzt);
                                    dzt = fmaf(__uint_as_float(a[4 * rr + 2]), xe.z, dzt);
                                    dzt = fmaf(__uint_as_float(a[4 * rr + 3]), xe.w, dzt);
                                    dzt = fmaf(dzt, INV_WSCALE, w2);  // + d xe2 / d theta * w1[r][2]
                                    float dzv = __uint_as_float(b[4 * rr + 0]) * xe.x;
                                    dzv = fmaf(__uint_as_float(b[4 * rr + 1]), xe.y, dzv);
                                    dzv = fmaf(__uint_as_float(b[4 * rr + 2]), xe.z, dzv);
                                    dzv = fmaf(__uint_as_float(b[4 * rr + 3]), xe.w, dzv);
                                    dzv = fmaf(dzv, INV_WSCALE, w3);  // + d xe3 / d v * w1[r][3]
                                    const float tz = tanhf(z);
                                    const float gz = fmaf(-tz, tz, 1.0f);
                                    const float w20 = sOb[r * TILE + s], w21 = sOb[(12 + r) * TILE + s];
                                    pu0 = fmaf(w20, tz, pu0);
                                    pu1 = fmaf(w21, tz, pu1);
                                    pd0 += sOb[(24 + r) * TILE + s] * tz + w20 * (gz * dzt);
                                    pd1 += sOb[(36 + r) * TILE + s] * tz + w21 * (gz * dzv);
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    warp_arrive(BAR(E5), lane);
                }

                if (owner) {
                    const float u0r = pu0 + ur0, u1r = pu1 + ur1;
                    margin = fminf(margin, fminf(fabsf(fabsf(u0r) - 3.0f), fabsf(fabsf(u1r) - 3.0f)));
                    if (dens) {
                        const float m0 = (u0r >= -3.0f && u0r <= 3.0f) ? 1.0f : 0.0f;
                        const float m1 = (u1r >= -3.0f && u1r <= 3.0f) ? 1.0f : 0.0f;
                        const float div = __fadd_rn(__fmul_rn(m0, pd0), __fmul_rn(m1, pd1));
                        if (logd)
                            rho = __fadd_rn(rho, logf(__fsub_rn(1.0f, __fmul_rn(div, dt))));
                        else
                            rho = __fadd_rn(rho, __fmul_rn(__fmul_rn(-div, rho), dt));
                    }
                    const float u0 = fminf(fmaxf(u0r, -3.0f), 3.0f), u1 = fminf(fmaxf(u1r, -3.0f), 3.0f);
                    float sn, cs;
                    sincosf(x2, &sn, &cs);
                    x0 = __fadd_rn(x0, __fmul_rn(__fmul_rn(x3, cs), dt));
                    x1 = __fadd_rn(x1, __fmul_rn(__fmul_rn(x3, sn), dt));
                    x2 = __fadd_rn(x2, __fmul_rn(u0, dt));
                    x3 = __fadd_rn(x3, __fmul_rn(u1, dt));
                }
            }
            if (owner && valid && p.clip_margin) p.clip_margin[n] = margin;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NCW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
#undef BAR
}

// ---- two tiles in flight --------------------------------------------------------------------------------------------------------
// The single-tile kernel above is serialised: CUDA cores idle while the value MMA of a net runs, the tensor pipe idles
// during layer 1 and the read-out.  Here every CTA owns TWO 128-sample tiles X and Y that time-share the TMEM columns:
// slots run in the order (X, net b) (X, net a) (Y, net b) (Y, net a); while the MMA warp and the epilogues work on the
// slot's tile, the compute warps evaluate layer 1 of the NEXT slot into registers (one 16-feature chunk while the value
// MMA runs, one while the layer-3 MMAs run) and write it to TMEM as soon as the next slot begins.  The MMA warp's
// program and the barrier protocol are exactly those of the single-tile kernel (it never needs to know whose tile it is).
constexpr int TRACE_SLOTS = 16, TRACE_EVENTS = 16, TRACE_WHO = 4;
#define TRACE(who, ev)                                                                                   \
    do {                                                                                                 \
        if (p.trace && blockIdx.x == 0 && lane == 0 && it < (uint32_t)TRACE_SLOTS)                       \
            p.trace[((who) * TRACE_SLOTS + it) * TRACE_EVENTS + (ev)] = (unsigned long long)clock64();   \
    } while (0)

__global__ void __launch_bounds__(Cfg<4>::NTHREADS, 1) rollout_tc2_kernel(RolloutParams p) {
    using C = Cfg<4>;
    constexpr int NCW = C::NCW, NCT = C::NCT, NTHREADS = C::NTHREADS, CPT = C::CPT;
    static_assert(CPT == 2, "two chunks per thread: one per overlap window");
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *sIn = reinterpret_cast<float *>(smem + OFF_IN);    // [2][128][8]
    float *sOb = reinterpret_cast<float *>(smem + OFF_OB);    // [48][128]
    float *sOut = reinterpret_cast<float *>(smem + OFF_OUT);  // [2][6][128]
    float *sPart = reinterpret_cast<float *>(smem + OFF_PART);  // [4][4][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 16 * 8);
    const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (i))
    {
        const int4 *src = reinterpret_cast<const int4 *>(p.tc);
        int4 *dst = reinterpret_cast<int4 *>(smem);
        for (int i = tid; i < IMAGE_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int i = E1; i <= E5; ++i) mbar_init(BAR(i), NCW);
        for (int i = F1; i <= F4; ++i) mbar_init(BAR(i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NCW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_ptr_s;

    const int T = p.T;
    const long long num_tiles = (p.N + TILE - 1) / TILE;
    const long long num_pairs = (num_tiles + 1) / 2;
    const uint32_t sbase = smem_u32(smem);

    if (warp == NCW) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (uniform control flow, barrier waits by all lanes); the tcgen05 instructions are
        // issued by the one lane elect.sync picks.
        {
            const uint32_t idesc128 = make_idesc(128), idesc3a = make_idesc(N3A), idesc3b = make_idesc(N3B);
            uint32_t it = 0;
            for (long long pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
                for (int i = 0; i < T; ++i) {
#pragma unroll 1
                    for (int slot = 0; slot < 4; ++slot, ++it) {
                        const int net = slot & 1;  // 0 = b (w2), 1 = a (w1)
                        const uint32_t par = it & 1u;
                        const uint64_t w2h = make_bdesc(sbase + (net == 0 ? OFF_W2B_HI : OFF_W2A_HI), B_LBO, B_SBO);
                        const uint64_t w2l = make_bdesc(sbase + (net == 0 ? OFF_W2B_LO : OFF_W2A_LO), B_LBO, B_SBO);
                        const uint64_t w3h = make_bdesc(sbase + (net == 0 ? OFF_W3B_HI : OFF_W3A_HI), B_LBO, B_SBO);
                        const uint64_t w3l = make_bdesc(sbase + (net == 0 ? OFF_W3B_LO : OFF_W3A_LO), B_LBO, B_SBO);
                        const uint32_t idesc3 = net == 0 ? idesc3b : idesc3a;
                        TRACE(0, 0);
                        if (it > 0) mbar_wait(BAR(E5), par ^ 1u);
                        TRACE(0, 1);
                        mbar_wait(BAR(E1), par);
                        tc_fence_after();
                        TRACE(0, 2);
                        if (elect_one()) {
                            mma_value(tm, COL_D0, w2h, w2l, idesc128);
                            tc_commit(BAR(F1));
                        }
                        __syncwarp();
                        TRACE(0, 3);
                        if (elect_one()) {
                            mma_tangent(tm, COL_D1, COL_AT0, w2h, w2l, idesc128);
                            tc_commit(BAR(F2));
                        }
                        __syncwarp();
                        TRACE(0, 4);
                        mbar_wait(BAR(E2A), par);
                        tc_fence_after();
                        TRACE(0, 5);
                        if (elect_one()) {
                            mma_tangent(tm, COL_D0, COL_AT1, w2h, w2l, idesc128);
                            tc_commit(BAR(F3));
                        }
                        __syncwarp();
                        TRACE(0, 6);
                        mbar_wait(BAR(E2B), par);
                        mbar_wait(BAR(E3A), par);
                        tc_fence_after();
                        TRACE(0, 7);
                        if (elect_one()) mma_value(tm, COL_D1, w3h, w3l, idesc3);
                        __syncwarp();
                        TRACE(0, 8);
                        mbar_wait(BAR(E3B), par);
                        tc_fence_after();
                        TRACE(0, 9);
                        if (elect_one()) mma_tangent(tm, COL_D1 + 64, COL_AT0, w3h, w3l, idesc3);
                        __syncwarp();
                        TRACE(0, 10);
                        mbar_wait(BAR(E4A), par);
                        mbar_wait(BAR(E4B), par);
                        tc_fence_after();
                        TRACE(0, 11);
                        if (elect_one()) {
                            mma_tangent(tm, COL_D0, COL_AT1, w3h, w3l, idesc3);
                            tc_commit(BAR(F4));
                        }
                        __syncwarp();
                        TRACE(0, 12);
                    }
                }
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q = warp & 3, g = warp >> 2;
        const int s = 32 * q + lane;
        const uint32_t tm_lane = tm + ((uint32_t)(32 * q) << 16);
        const bool owner = g == 0;
        const float *sp_a = reinterpret_cast<const float *>(smem + OFF_SMALL);
        const float *sp_b = sp_a + SMALL_NET_FLOATS;
        const int c0 = CPT * g, c1 = CPT * g + 1;  // this thread's two 16-feature chunks
        const int who = warp == 0 ? 1 : warp == 4 ? 2 : warp == 15 ? 3 : -1;
#define CTRACE(ev)                   \
    do {                             \
        if (who >= 0) TRACE(who, ev); \
    } while (0)
        uint32_t it = 0;

        for (long long pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");  // the previous pair's last slices have left sOut
            const long long tileX = 2 * pair, tileY = 2 * pair + 1;
            const long long nX = tileX * TILE + s, nY = tileY * TILE + s;
            const bool validX = nX < p.N, validY = nY < p.N;
            SampleState stX, stY;
            if (owner) {
                sample_init(p, stX, validX ? nX : p.N - 1);
                sample_init(p, stY, validY ? nY : p.N - 1);
                sample_step_start(p, stX, 0, sIn, sOut, s, validX);
                sample_step_start(p, stY, 0, sIn + TILE * 8, sOut + 6 * TILE, s, validY);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            if (g == 1) {
                store_slice(p, sOut, tileX, 0, tid - 128);
                store_slice(p, sOut + 6 * TILE, tileY, 0, tid - 128);
            }
            // prologue: layer 1 of the first slot (X, net b)
            uint32_t pav0[16], pav1[16], pg0[8], pg1[8];
            if (T > 0) {
                const float4 in = *reinterpret_cast<const float4 *>(sIn + 8 * s);
                l1_chunk(sp_b, in, c0, pav0, pg0);
                l1_chunk(sp_b, in, c1, pav1, pg1);
            }
            for (int i = 0; i < T; ++i) {
#pragma unroll 1
                for (int slot = 0; slot < 4; ++slot, ++it) {
                    const int P = slot >> 1, net = slot & 1;
                    const uint32_t par = it & 1u;
                    const float *sp = net == 0 ? sp_b : sp_a;
                    const int nslot = (slot + 1) & 3, nP = nslot >> 1;
                    const float *spn = (nslot & 1) == 0 ? sp_b : sp_a;
                    const bool has_next = !(slot == 3 && i == T - 1);
                    CTRACE(0);
                    // 1. this slot's layer-1 operands (computed during the previous slot) -> TMEM
                    l1_store(sp, tm_lane, c0, pav0, pg0);
                    l1_store(sp, tm_lane, c1, pav1, pg1);
                    tc_wait_st();
                    tc_fence_before();
                    warp_arrive(BAR(E1), lane);
                    CTRACE(1);
                    // 2. publish point for sIn / sOut written at the end of the previous slot; pending slice stores;
                    //    first half of the next slot's layer 1 while the value MMA runs
                    asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
                    // the previous slot's head (if it was a net-a slot): its partial sums are complete now
                    if (owner) {
                        if (slot == 2) finalize_step(p, stX, sPart, i + 1, sIn, sOut, s, validX);
                        if (slot == 0 && i > 0) finalize_step(p, stY, sPart, i, sIn + TILE * 8, sOut + 6 * TILE, s, validY);
                    }
                    // slices staged one slot ago are visible now
                    if (g == 1) {
                        if (slot == 3 && ((i + 1) % p.stride) == 0) store_slice(p, sOut, tileX, i + 1, tid - 128);
                        if (slot == 1 && i > 0 && (i % p.stride) == 0) store_slice(p, sOut + 6 * TILE, tileY, i, tid - 128);
                    }
                    float4 inn = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_next) {
                        inn = *reinterpret_cast<const float4 *>(sIn + nP * TILE * 8 + 8 * s);
                        l1_chunk(spn, inn, c0, pav0, pg0);
                    }
                    CTRACE(2);
                    // 3..5 layer-2 epilogues of this slot's tile
                    {
                        uint32_t d[16], gs[16];
                        mbar_wait(BAR(F1), par);
                        tc_fence_after();
                        CTRACE(3);
                        tc_ld16(tm_lane + COL_D0 + 16 * c0, d);
                        tc_wait_ld();
                        epi_value_chunk(sp, tm_lane, c0, d, gs);
                        tc_ld16(tm_lane + COL_D0 + 16 * c1, d);
                        tc_wait_ld();
                        tc_fence_before();
                        warp_arrive(BAR(E2A), lane);  // D0 drained: the v-tangent MMAs may start
                        CTRACE(4);
                        epi_value_chunk(sp, tm_lane, c1, d, gs + 8);
                        tc_wait_st();
                        tc_fence_before();
                        warp_arrive(BAR(E2B), lane);
                        CTRACE(5);
                        // F3 (v-tangent accumulator) implies F2 (theta accumulator, issued earlier): one wait serves both
                        mbar_wait(BAR(F3), par);
                        tc_fence_after();
                        CTRACE(6);
                        tc_ld16(tm_lane + COL_D1 + 16 * c0, d);
                        tc_wait_ld();
                        epi_tangent_chunk(tm_lane, COL_AT0, c0, d, gs);
                        tc_ld16(tm_lane + COL_D1 + 16 * c1, d);
                        tc_wait_ld();
                        tc_fence_before();
                        warp_arrive(BAR(E3A), lane);
                        epi_tangent_chunk(tm_lane, COL_AT0, c1, d, gs + 8);
                        tc_wait_st();
                        tc_fence_before();
                        warp_arrive(BAR(E3B), lane);
                        CTRACE(7);
                        tc_ld16(tm_lane + COL_D0 + 16 * c0, d);
                        tc_wait_ld();
                        epi_tangent_chunk(tm_lane, COL_AT1, c0, d, gs);
                        tc_ld16(tm_lane + COL_D0 + 16 * c1, d);
                        tc_wait_ld();
                        tc_fence_before();
                        warp_arrive(BAR(E4A), lane);
                        epi_tangent_chunk(tm_lane, COL_AT1, c1, d, gs + 8);
                        tc_wait_st();
                        tc_fence_before();
                        warp_arrive(BAR(E4B), lane);
                        CTRACE(9);
                    }
                    // 5b. second half of the next slot's layer 1 while the layer-3 MMAs run
                    if (has_next) l1_chunk(spn, inn, c1, pav1, pg1);
                    CTRACE(10);
                    // 6. layer-3 read-out (and, after net a, the head + Euler step + staging of the next time index)
                    mbar_wait(BAR(F4), par);
                    tc_fence_after();
                    CTRACE(11);
                    // every group drains its share of the layer-3 accumulators, hands them back (E5), then does the math
                    if (net == 0) {
                        uint32_t d[32];
                        readout_b_load(tm_lane, g, d);
                        tc_fence_before();
                        warp_arrive(BAR(E5), lane);
                        readout_b_store(g, d, sp + SP_B3, sOb, s);
                    } else {
                        const float4 xe = *reinterpret_cast<const float4 *>(sIn + P * TILE * 8 + 8 * s + 4);
                        float pu0 = 0.f, pu1 = 0.f, pd0 = 0.f, pd1 = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) readout_a_row(tm_lane, 3 * g + rr, sp + SP_B3, sOb, xe, s, pu0, pu1, pd0, pd1);
                        tc_fence_before();
                        warp_arrive(BAR(E5), lane);
                        float *pp = sPart + g * 4 * TILE + s;
                        pp[0 * TILE] = pu0; pp[1 * TILE] = pu1; pp[2 * TILE] = pd0; pp[3 * TILE] = pd1;
                    }
                    CTRACE(12);
                }
            }
            // the head of the last slot (tile Y, time T) is still pending
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            if (owner && T > 0) finalize_step(p, stY, sPart, T, sIn + TILE * 8, sOut + 6 * TILE, s, validY);
            asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
            if (g == 1 && T > 0 && (T % p.stride) == 0) store_slice(p, sOut + 6 * TILE, tileY, T, tid - 128);
            if (owner && p.clip_margin) {
                if (validX) p.clip_margin[nX] = stX.margin;
                if (validY) p.clip_margin[nY] = stY.margin;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NCW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
#undef BAR
}

// ---- self-test of the MMA building block: D = split3(A) x split3(W)^T through TMEM --------------------------------------------
// A fp32 [128][128], W fp32 [N][128] (N in {128, 48, 32}); mode 0: value path (3 products), 1: tangent path (2 products)
__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(const float *A, const unsigned char *wimg, int n, int mode,
                                                            float *D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wbytes = n * 128 * 2;
    for (int i = tid; i < 2 * wbytes / 16; i += 160) reinterpret_cast<int4 *>(smem)[i] = reinterpret_cast<const int4 *>(wimg)[i];
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp < 4) {
        const uint32_t tm_lane = tm + ((uint32_t)(32 * warp) << 16);
        const float *row = A + (size_t)tid * 128;
        for (int c = 0; c < 8; ++c) {
            uint32_t av[16], at[8];
            for (int pq = 0; pq < 8; ++pq) {
                split_h2(row[16 * c + 2 * pq], row[16 * c + 2 * pq + 1], av[pq], av[8 + pq]);
                at[pq] = av[pq];
            }
            tc_st16(tm_lane + COL_AV + 16 * c, av);
            tc_st8(tm_lane + COL_AT0 + 8 * c, at);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
        const uint64_t w_hi = make_bdesc(smem_u32(smem), B_LBO, B_SBO), w_lo = make_bdesc(smem_u32(smem) + wbytes, B_LBO, B_SBO);
        if (mode == 0)
            mma_value(tm, COL_D0, w_hi, w_lo, make_idesc(n));
        else
            mma_tangent(tm, COL_D0, COL_AT0, w_hi, w_lo, make_idesc(n));
        tc_commit(smem_u32(&bar));
    }
    if (warp < 4) {
        mbar_wait(smem_u32(&bar), 0);
        tc_fence_after();
        const uint32_t tm_lane = tm + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < n / 16; ++c) {
            uint32_t d[16];
            tc_ld16(tm_lane + COL_D0 + 16 * c, d);
            tc_wait_ld();
            for (int k = 0; k < 16; ++k) D[(size_t)tid * n + 16 * c + k] = __uint_as_float(d[k]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}


// ---- diagnostic: tcgen05.mma issue/execute rate for the operand configurations this kernel could use --------------------------
// variant 0: A in TMEM, B no-swizzle K-major (what the rollout uses)   1: A in TMEM, B SWIZZLE_128B K-major
// variant 2: A and B in shared memory, no swizzle                       3: A and B in shared memory, SWIZZLE_128B
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum), "r"(0)
        : "memory");
}

__global__ void __launch_bounds__(64, 1) tc_mma_bench_kernel(int variant, int n, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 + 32) * 1024 / 16; i += 64) reinterpret_cast<int4 *>(smem)[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp == 1) {
        const bool sw = variant & 1, ss = variant >= 2;
        const uint32_t sb = smem_u32(smem), sa = sb + 64 * 1024;
        // 8 K-steps per operand image; swizzled: one 128-byte row per N index (K = 64), 8-row groups 1024 bytes apart
        const uint64_t bdesc = sw ? (make_bdesc(sb, 16, 1024) | (2ull << 61)) : make_bdesc(sb, B_LBO, B_SBO);
        const uint64_t adesc = sw ? (make_bdesc(sa, 16, 1024) | (2ull << 61)) : make_bdesc(sa, B_LBO, B_SBO);
        const uint64_t kstep = sw ? 2 : B_KSTEP_DESC;
        const uint32_t idesc = make_idesc(n);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (ss)
                        tc_mma_ss(tm + 256, adesc + c * kstep, bdesc + c * kstep, idesc, 1);
                    else
                        tc_mma_ts(tm + 256, tm + 8 * c, bdesc + c * kstep, idesc, 1);
                }
            }
            t1 = clock64();
            tc_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t2 = clock64();
        if (t0) {
            out[0] = t1 - t0;  // issue time
            out[1] = t2 - t0;  // until all MMAs completed
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

// ---- host: weight image ---------------------------------------------------------------------------------------------------------
// canonical no-swizzle K-major image of W [n_rows][128] (rows >= n_valid are zero): halfword index of (r, k)
inline size_t canon_index(int r, int k) { return ((size_t)(r / 8) * 16 + (k / 8)) * 64 + (r % 8) * 8 + (k % 8); }

void build_split_image(const float *W, int n_valid, int n_rows, float scale, __half *hi, __half *lo) {
    for (int r = 0; r < n_rows; ++r)
        for (int k = 0; k < 128; ++k) {
            const float w = r < n_valid ? W[(size_t)r * 128 + k] * scale : 0.0f;
            const __half h = __float2half_rn(w);
            const __half l = __float2half_rn(w - __half2float(h));
            hi[canon_index(r, k)] = h;
            lo[canon_index(r, k)] = l;
        }
}

}  // namespace

int dp_tc_prepare(dp_controller *c, const float *blob) {
    std::vector<unsigned char> img(IMAGE_BYTES, 0);
    const float *p = blob;
    const int nouts[2] = {DP_NOUT_A, DP_NOUT_B}, nrows3[2] = {N3A, N3B};
    const int off_w2h[2] = {OFF_W2A_HI, OFF_W2B_HI}, off_w2l[2] = {OFF_W2A_LO, OFF_W2B_LO};
    const int off_w3h[2] = {OFF_W3A_HI, OFF_W3B_HI}, off_w3l[2] = {OFF_W3A_LO, OFF_W3B_LO};
    for (int m = 0; m < 2; ++m) {
        float *sp = reinterpret_cast<float *>(img.data() + OFF_SMALL) + m * SMALL_NET_FLOATS;
        memcpy(sp + SP_W1, p, sizeof(float) * 512);
        for (int j = 0; j < 64; ++j) {
            __half2 c0 = __floats2half2_rn(p[4 * (2 * j)], p[4 * (2 * j + 1)]);
            __half2 c1 = __floats2half2_rn(p[4 * (2 * j) + 1], p[4 * (2 * j + 1) + 1]);
            memcpy(sp + SP_C0H + j, &c0, 4);
            memcpy(sp + SP_C1H + j, &c1, 4);
        }
        p += 512;
        memcpy(sp + SP_B1, p, sizeof(float) * 128); p += 128;
        build_split_image(p, 128, 128, WSCALE, reinterpret_cast<__half *>(img.data() + off_w2h[m]),
                          reinterpret_cast<__half *>(img.data() + off_w2l[m]));
        p += 128 * 128;
        memcpy(sp + SP_B2, p, sizeof(float) * 128); p += 128;
        build_split_image(p, nouts[m], nrows3[m], WSCALE, reinterpret_cast<__half *>(img.data() + off_w3h[m]),
                          reinterpret_cast<__half *>(img.data() + off_w3l[m]));
        p += nouts[m] * 128;
        memcpy(sp + SP_B3, p, sizeof(float) * nouts[m]); p += nouts[m];
    }
    void *d = nullptr;
    DP_CUDA(cudaMalloc(&d, IMAGE_BYTES));
    DP_CUDA(cudaMemcpy(d, img.data(), IMAGE_BYTES, cudaMemcpyHostToDevice));
    c->d_tc = d;
    c->tc_bytes = IMAGE_BYTES;
    return 0;
}

void dp_tc_release(dp_controller *c) {
    if (c->d_tc) cudaFree(c->d_tc);
    c->d_tc = nullptr;
}

int dp_launch_rollout_tc(const dp_controller *c, const RolloutParams &p, cudaStream_t stream) {
    DP_REQUIRE(c->d_tc != nullptr, "DP_IMPL_TC: tensor-core weight image missing");
    static bool attr_set[64] = {};
    static int mode = -1;  // 0: two tiles in flight (default); 2 / 4: single-tile kernel with that many warp groups
    if (mode < 0) {
        const char *e = getenv("DP_TC_MODE");  // tuning / A-B knob
        mode = e ? atoi(e) : 0;
        if (mode != 2 && mode != 4) mode = 0;
        if (getenv("DP_DEBUG")) fprintf(stderr, "[dp] tcgen05 rollout mode %d (0 = two tiles in flight), %d B smem\n", mode, SMEM_BYTES);
    }
    if (!attr_set[c->device & 63]) {
        DP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        DP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        DP_CUDA(cudaFuncSetAttribute(rollout_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set[c->device & 63] = true;
    }
    const long long num_tiles = (p.N + TILE - 1) / TILE;
    if (mode == 0) {
        const long long num_pairs = (num_tiles + 1) / 2;
        const int grid = (int)(num_pairs < c->num_sms ? num_pairs : c->num_sms);
        const char *trace_path = getenv("DP_TC_TRACE");  // diagnostic: dump an in-kernel clock64 timeline of CTA 0
        if (trace_path) {
            RolloutParams q = p;
            const size_t n = (size_t)TRACE_WHO * TRACE_SLOTS * TRACE_EVENTS;
            DP_CUDA(cudaMalloc(&q.trace, n * 8));
            DP_CUDA(cudaMemsetAsync(q.trace, 0, n * 8, stream));
            rollout_tc2_kernel<<<grid, Cfg<4>::NTHREADS, SMEM_BYTES, stream>>>(q);
            DP_CUDA(cudaStreamSynchronize(stream));
            std::vector<unsigned long long> h(n);
            DP_CUDA(cudaMemcpy(h.data(), q.trace, n * 8, cudaMemcpyDeviceToHost));
            cudaFree(q.trace);
            if (FILE *f = fopen(trace_path, "w")) {
                for (int w = 0; w < TRACE_WHO; ++w)
                    for (int sl = 0; sl < TRACE_SLOTS; ++sl) {
                        fprintf(f, "%d %d", w, sl);
                        for (int e = 0; e < TRACE_EVENTS; ++e) fprintf(f, " %llu", h[((size_t)w * TRACE_SLOTS + sl) * TRACE_EVENTS + e]);
                        fprintf(f, "\n");
                    }
                fclose(f);
            }
            return 0;
        }
        rollout_tc2_kernel<<<grid, Cfg<4>::NTHREADS, SMEM_BYTES, stream>>>(p);
    } else {
        const int grid = (int)(num_tiles < c->num_sms ? num_tiles : c->num_sms);
        if (mode == 2)
            rollout_tc_kernel<2><<<grid, Cfg<2>::NTHREADS, SMEM_BYTES, stream>>>(p);
        else
            rollout_tc_kernel<4><<<grid, Cfg<4>::NTHREADS, SMEM_BYTES, stream>>>(p);
    }
    DP_CUDA(cudaGetLastError());
    return 0;
}

// Diagnostic export: one 128 x n x 128 product through the same TMEM-operand MMA path the rollout uses.
// A host fp32 [128][128], W host fp32 [n_valid][128] (zero-padded to n rows), D host fp32 [128][n] = A * (scale*W)^T.
extern "C" int dp_tc_selftest(int device, const float *A, const float *W, int n_valid, int n, int mode, float scale, float *D) {
    DP_REQUIRE(A && W && D && (n == 128 || n == 48 || n == 32) && n_valid <= n, "dp_tc_selftest: bad argument");
    DeviceGuard guard(device);
    std::vector<__half> img((size_t)2 * n * 128);
    build_split_image(W, n_valid, n, scale, img.data(), img.data() + (size_t)n * 128);
    float *dA, *dD;
    unsigned char *dW;
    DP_CUDA(cudaMalloc(&dA, 128 * 128 * 4));
    DP_CUDA(cudaMalloc(&dD, 128 * n * 4));
    DP_CUDA(cudaMalloc(&dW, img.size() * 2));
    DP_CUDA(cudaMemcpy(dA, A, 128 * 128 * 4, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(dW, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemset(dD, 0, 128 * n * 4));
    const int smem = 2 * n * 128 * 2;
    DP_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_selftest_kernel<<<1, 160, smem>>>(dA, dW, n, mode, dD);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    DP_CUDA(cudaMemcpy(D, dD, 128 * n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA);
    cudaFree(dD);
    cudaFree(dW);
    return 0;
}

// Diagnostic export: cycles per tcgen05.mma (M=128, N=n, K=16, fp16) for `iters` x 4 back-to-back instructions.
// out[0] = issue cycles per MMA, out[1] = cycles per MMA until completion.  variant: see tc_mma_bench_kernel.
extern "C" int dp_tc_mma_bench(int device, int variant, int n, int iters, double *out) {
    DP_REQUIRE(out && variant >= 0 && variant <= 3 && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0, "dp_tc_mma_bench: bad argument");
    DeviceGuard guard(device);
    long long *d;
    DP_CUDA(cudaMalloc(&d, 16));
    DP_CUDA(cudaMemset(d, 0, 16));
    const int smem = 96 * 1024;
    DP_CUDA(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_mma_bench_kernel<<<1, 64, smem>>>(variant, n, iters, d);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    long long h[2];
    DP_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    out[0] = (double)h[0] / (4.0 * iters);
    out[1] = (double)h[1] / (4.0 * iters);
    return 0;
}
