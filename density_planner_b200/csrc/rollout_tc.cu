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

#include <vector>

#include "dp_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr int NCW = 8;                      // compute warps
constexpr int NCT = NCW * 32;               // compute threads
constexpr int NTHREADS = NCT + 32;          // + MMA warp
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
constexpr int SMALL_NET_FLOATS = 128 * 4 + 128 + 128 + 64;
constexpr int SP_W1 = 0, SP_B1 = 512, SP_B2 = 640, SP_B3 = 768;
constexpr int IMAGE_BYTES = OFF_SMALL + 2 * SMALL_NET_FLOATS * 4;
// run-time scratch behind the image
constexpr int OFF_IN = IMAGE_BYTES;                  // float [128][8]
constexpr int OFF_OB = OFF_IN + TILE * 8 * 4;        // float [48][128]
constexpr int OFF_OUT = OFF_OB + 48 * TILE * 4;      // float [6][128]
constexpr int OFF_BAR = OFF_OUT + 6 * TILE * 4;      // 16 mbarriers + tmem pointer
constexpr int SMEM_BYTES = OFF_BAR + 16 * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(IMAGE_BYTES % 16 == 0, "image copied with 16-byte vectors");

enum { E1 = 0, E2, E3, E4, E5, F1, F2, F3, F4, NUM_BARS };

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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
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
// layer 1 on CUDA cores: in(4) -> 64 features of this thread's half, value (hi/lo) + two tangents -> TMEM operands
__device__ __forceinline__ void stage_l1(const float *sp, const float4 in, uint32_t tm_lane, int h) {
    const float4 *W1 = reinterpret_cast<const float4 *>(sp + SP_W1);
    const float *b1 = sp + SP_B1;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * h + i;  // 16-feature chunk
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
                const float g = fmaf(-t, t, 1.0f);
                hv[e] = t;
                t0[e] = g * w.x;
                t1[e] = g * w.y;
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

// layer-2 epilogue, value rows: D0 -> h2 = tanh(D/16 + b2) -> A_V (hi/lo), keeps g2/16 packed for the tangent rows
__device__ __forceinline__ void stage_epi_value(const float *sp, uint32_t tm_lane, int h, uint32_t (&gs)[32]) {
    const float *b2 = sp + SP_B2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * h + i;
        uint32_t d[16], av[16];
        tc_ld16(tm_lane + COL_D0 + 16 * c, d);
        tc_wait_ld();
#pragma unroll
        for (int pq = 0; pq < 8; ++pq) {
            float hv[2], g[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float pre = fmaf(__uint_as_float(d[2 * pq + e]), INV_WSCALE, b2[16 * c + 2 * pq + e]);
                const float t = tanh_fast(pre);
                hv[e] = t;
                g[e] = fmaf(-t, t, 1.0f) * INV_WSCALE;
            }
            split_h2(hv[0], hv[1], av[pq], av[8 + pq]);
            gs[8 * i + pq] = pack_h2(g[0], g[1]);
        }
        tc_st16(tm_lane + COL_AV + 16 * c, av);
    }
}

// layer-2 epilogue, tangent rows: D -> (g2/16) * D -> A_T (hi only)
__device__ __forceinline__ void stage_epi_tangent(uint32_t tm_lane, int h, int col_d, int col_a, const uint32_t (&gs)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * h + i;
        uint32_t d[16], at[8];
        tc_ld16(tm_lane + col_d + 16 * c, d);
        tc_wait_ld();
#pragma unroll
        for (int pq = 0; pq < 8; ++pq) {
            const float2 g = unpack_h2(gs[8 * i + pq]);
            at[pq] = pack_h2(__uint_as_float(d[2 * pq]) * g.x, __uint_as_float(d[2 * pq + 1]) * g.y);
        }
        tc_st8(tm_lane + col_a + 8 * c, at);
    }
}

// ---- MMA-warp stages (one elected thread) -----------------------------------------------------------------------------------
// D (+)= A_V(hi,lo) x W(hi,lo): three products per K-chunk
__device__ __forceinline__ void mma_value(uint32_t tm, int col_d, uint32_t w_hi, uint32_t w_lo, uint32_t idesc) {
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        const uint64_t bh = make_bdesc(w_hi + c * B_KSTEP, B_LBO, B_SBO);
        const uint64_t bl = make_bdesc(w_lo + c * B_KSTEP, B_LBO, B_SBO);
        const uint32_t a_hi = tm + COL_AV + 16 * c, a_lo = a_hi + 8;
        tc_mma_ts(tm + col_d, a_hi, bh, idesc, c > 0);
        tc_mma_ts(tm + col_d, a_hi, bl, idesc, 1);
        tc_mma_ts(tm + col_d, a_lo, bh, idesc, 1);
    }
}
// D (+)= A_T(hi) x W(hi,lo): two products per K-chunk
__device__ __forceinline__ void mma_tangent(uint32_t tm, int col_d, int col_a, uint32_t w_hi, uint32_t w_lo, uint32_t idesc) {
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        const uint64_t bh = make_bdesc(w_hi + c * B_KSTEP, B_LBO, B_SBO);
        const uint64_t bl = make_bdesc(w_lo + c * B_KSTEP, B_LBO, B_SBO);
        const uint32_t a = tm + col_a + 8 * c;
        tc_mma_ts(tm + col_d, a, bh, idesc, c > 0);
        tc_mma_ts(tm + col_d, a, bl, idesc, 1);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) rollout_tc_kernel(RolloutParams p) {
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
        for (int i = E1; i <= E5; ++i) mbar_init(BAR(i), NCT);
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
                        const uint32_t w2h = sbase + (net == 0 ? OFF_W2B_HI : OFF_W2A_HI);
                        const uint32_t w2l = sbase + (net == 0 ? OFF_W2B_LO : OFF_W2A_LO);
                        const uint32_t w3h = sbase + (net == 0 ? OFF_W3B_HI : OFF_W3A_HI);
                        const uint32_t w3l = sbase + (net == 0 ? OFF_W3B_LO : OFF_W3A_LO);
                        const uint32_t idesc3 = net == 0 ? idesc3b : idesc3a;
                        if (it > 0) mbar_wait(BAR(E5), par ^ 1u);  // previous layer-3 results drained
                        mbar_wait(BAR(E1), par);                   // layer-1 operands written
                        tc_fence_after();
                        mma_value(tm, COL_D0, w2h, w2l, idesc128);
                        tc_commit(BAR(F1));
                        mma_tangent(tm, COL_D1, COL_AT0, w2h, w2l, idesc128);
                        tc_commit(BAR(F2));
                        mbar_wait(BAR(E2), par);  // D0 drained by the value epilogue (A_V rewritten)
                        tc_fence_after();
                        mma_tangent(tm, COL_D0, COL_AT1, w2h, w2l, idesc128);
                        tc_commit(BAR(F3));
                        mbar_wait(BAR(E3), par);  // D1 drained, A_T0 rewritten
                        tc_fence_after();
                        mma_value(tm, COL_D1, w3h, w3l, idesc3);
                        mma_tangent(tm, COL_D1 + 64, COL_AT0, w3h, w3l, idesc3);
                        mbar_wait(BAR(E4), par);  // D0 drained, A_T1 rewritten
                        tc_fence_after();
                        mma_tangent(tm, COL_D0, COL_AT1, w3h, w3l, idesc3);
                        tc_commit(BAR(F4));
                    }
                }
            }
        }
    } else {
        // ===================== compute warps =====================
        const int q = warp & 3, h = warp >> 2;
        const int s = 32 * q + lane;  // sample within the tile == TMEM lane
        const uint32_t tm_lane = tm + ((uint32_t)(32 * q) << 16);
        const bool owner = h == 0;  // owns the sample's state, the layer-3 read-out and the head
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
                if (store_now && !owner) {
                    // coalesced store of the staged slice by the four non-owner warps: 128 consecutive samples per row
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
                    stage_l1(sp, in, tm_lane, h);
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(BAR(E1));
                    // layer-2 epilogues
                    uint32_t gs[32];
                    mbar_wait(BAR(F1), par);
                    tc_fence_after();
                    stage_epi_value(sp, tm_lane, h, gs);
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(BAR(E2));
                    mbar_wait(BAR(F2), par);
                    tc_fence_after();
                    stage_epi_tangent(tm_lane, h, COL_D1, COL_AT0, gs);
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(BAR(E3));
                    mbar_wait(BAR(F3), par);
                    tc_fence_after();
                    stage_epi_tangent(tm_lane, h, COL_D0, COL_AT1, gs);
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(BAR(E4));
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
                    mbar_arrive(BAR(E5));
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
        const uint32_t w_hi = smem_u32(smem), w_lo = w_hi + wbytes;
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
        memcpy(sp + SP_W1, p, sizeof(float) * 512); p += 512;
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
    if (!attr_set[c->device & 63]) {
        DP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set[c->device & 63] = true;
    }
    const long long num_tiles = (p.N + TILE - 1) / TILE;
    const int grid = (int)(num_tiles < c->num_sms ? num_tiles : c->num_sms);
    rollout_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(p);
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
