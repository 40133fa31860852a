// K1 (fp32 CUDA-core version): fused closed-loop rollout + Liouville density for the CAR system under the C3M
// controller.  Replaces the per-step PyTorch/autograd loop of ControlAffineSystem.compute_density
// (systems/ControlAffineSystem.py:409-463).
//
// Math per sample and step (closed form of dudx_func/dfdx_func/divf_func, :69-122; Controller.__call__
// systems/sytem_CAR.py:25-37; U_FUNC.forward data/trained_controller/model_CAR.py:19-28):
//   xe = x[:4]-xref[:4]; xe2 += d;  in = [th, v, th-xe2, v-xe3]
//   net n in {a: 48 outputs -> w1 (12x4), b: 24 outputs -> w2 (2x12)}:
//      h1 = tanh(W1 in + b1), h2 = tanh(W2 h1 + b2), o = W3 h2 + b3
//      tangents k in {theta, v}: t1 = (1-h1^2) * W1[:,k];  t2 = (1-h2^2) * (W2 t1);  do = W3 t2
//   z = w1 xe; u_raw = w2 tanh(z) + uref; u = clip(u_raw, +-3)
//   div f = m0 du0/dtheta + m1 du1/dv,  m_i = [-3 <= u_raw_i <= 3]
//   l += log(1 - div*dt)  |  rho += (-div*rho)*dt ;   x += [v cos th, v sin th, u0, u1, 0]*dt
//
// Design: one persistent CTA per SM, 256 threads, tiles of 32 samples.  The two 128x128 and two 128x{48,24} weight
// matrices (164 KB fp32, k-major) stay resident in shared memory for the whole launch.  Per step the three rows
// (value, theta-tangent, v-tangent) of the 32 samples form a 96x128 activation matrix in shared memory; the 128-wide
// layers are register-tiled SGEMMs (6 rows x 8 columns per thread, LDS.128 operands), the state and log-density
// stay in registers of one leader lane per sample over the whole horizon, and stored slices go out as coalesced
// 128-byte rows of the [t][d][n] trajectory.
#include "dp_common.cuh"

namespace {

constexpr int S = 32;          // samples per tile
constexpr int NT = 256;        // threads per CTA
constexpr int ROWS = 3 * S;    // value + 2 tangents
constexpr int LDA = 132;       // activation row stride (floats): 16-byte aligned rows, conflict-free float4 stores

constexpr int SM_W = DP_BIG_FLOATS;                 // resident weights
constexpr int SM_A = ROWS * LDA;                    // activations (O_a aliases this after the last GEMM of a step)
constexpr int SM_OB = S * 3 * DP_NOUT_B;            // outputs of net b
constexpr int SM_IN = S * 8;                        // per-sample {in0..3, xe0..3}
constexpr int SM_OUT = 6 * S;                       // staged trajectory slice
constexpr size_t SMEM_BYTES = sizeof(float) * (SM_W + SM_A + SM_OB + SM_IN + SM_OUT);
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(S * 3 * DP_NOUT_A <= SM_A, "O_a must fit in the activation buffer");

__device__ __forceinline__ float dp_tanh(float x) { return tanhf(x); }

// ---- layer 1: in(4) -> 128, value + 2 tangents, written to A[row][col] ------------------------------------------
__device__ __forceinline__ void layer1(const float *__restrict__ small_net, const float *sIn, float *A, int ty, int tx) {
    const float4 *W1 = reinterpret_cast<const float4 *>(small_net + DP_SM_W1);
    const float *b1 = small_net + DP_SM_B1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int c0 = 4 * tx + 64 * half;
        float4 w[4];
        float b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            w[c] = __ldg(W1 + c0 + c);
            b[c] = __ldg(b1 + c0 + c);
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int s = 2 * ty + sl;
            const float4 in = *reinterpret_cast<const float4 *>(sIn + 8 * s);
            float hv[4], t0[4], t1[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float pre = w[c].x * in.x;
                pre = fmaf(w[c].y, in.y, pre);
                pre = fmaf(w[c].z, in.z, pre);
                pre = fmaf(w[c].w, in.w, pre);
                pre += b[c];
                const float h = dp_tanh(pre);
                const float g = 1.0f - h * h;
                hv[c] = h;
                t0[c] = g * w[c].x;  // d/dtheta enters through in0
                t1[c] = g * w[c].y;  // d/dv     enters through in1
            }
            float *row = A + (6 * ty + 3 * sl) * LDA + c0;
            *reinterpret_cast<float4 *>(row) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4 *>(row + LDA) = make_float4(t0[0], t0[1], t0[2], t0[3]);
            *reinterpret_cast<float4 *>(row + 2 * LDA) = make_float4(t1[0], t1[1], t1[2], t1[3]);
        }
    }
}

// ---- layer 2: A[96][128] x W2t[128][128], thread tile 6 rows x 8 cols ---------------------------------------------
__device__ __forceinline__ void layer2_mainloop(const float *A, const float *Wt, int ty, int tx, float (&acc)[6][8]) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
    const float *a_ptr = A + 6 * ty * LDA;
    const float *w_ptr = Wt + 4 * tx;
#pragma unroll 2
    for (int k4 = 0; k4 < DP_HID / 4; ++k4) {
        float4 a[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) a[r] = *reinterpret_cast<const float4 *>(a_ptr + r * LDA + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w0 = *reinterpret_cast<const float4 *>(w_ptr + (4 * k4 + kk) * DP_HID);
            const float4 w1 = *reinterpret_cast<const float4 *>(w_ptr + (4 * k4 + kk) * DP_HID + 64);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
                acc[r][0] = fmaf(av, w0.x, acc[r][0]);
                acc[r][1] = fmaf(av, w0.y, acc[r][1]);
                acc[r][2] = fmaf(av, w0.z, acc[r][2]);
                acc[r][3] = fmaf(av, w0.w, acc[r][3]);
                acc[r][4] = fmaf(av, w1.x, acc[r][4]);
                acc[r][5] = fmaf(av, w1.y, acc[r][5]);
                acc[r][6] = fmaf(av, w1.z, acc[r][6]);
                acc[r][7] = fmaf(av, w1.w, acc[r][7]);
            }
        }
    }
}

__device__ __forceinline__ void layer2_epilogue(const float *__restrict__ small_net, float *A, int ty, int tx,
                                                float (&acc)[6][8]) {
    const float *b2 = small_net + DP_SM_B2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int c0 = 4 * tx + 64 * half;
        float b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = __ldg(b2 + c0 + c);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            float hv[4], t0[4], t1[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float h = dp_tanh(acc[3 * sl][4 * half + c] + b[c]);
                const float g = 1.0f - h * h;
                hv[c] = h;
                t0[c] = g * acc[3 * sl + 1][4 * half + c];
                t1[c] = g * acc[3 * sl + 2][4 * half + c];
            }
            float *row = A + (6 * ty + 3 * sl) * LDA + c0;
            *reinterpret_cast<float4 *>(row) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4 *>(row + LDA) = make_float4(t0[0], t0[1], t0[2], t0[3]);
            *reinterpret_cast<float4 *>(row + 2 * LDA) = make_float4(t1[0], t1[1], t1[2], t1[3]);
        }
    }
}

// ---- layer 3: A[96][128] x W3t[128][NOUT]; thread (s, cg) owns the 3 rows of sample s and NOUT/8 columns ------------
template <int NOUT>
__device__ __forceinline__ void layer3_mainloop(const float *A, const float *Wt, int s, int cg, float (&acc)[3][NOUT / 8]) {
    constexpr int CPT = NOUT / 8;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = 0.0f;
    const float *a_ptr = A + 3 * s * LDA;
    const float *w_ptr = Wt + CPT * cg;
#pragma unroll 2
    for (int k4 = 0; k4 < DP_HID / 4; ++k4) {
        float4 a[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) a[r] = *reinterpret_cast<const float4 *>(a_ptr + r * LDA + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float w[CPT];
            if constexpr (CPT % 2 == 0) {
#pragma unroll
                for (int c = 0; c < CPT; c += 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(w_ptr + (4 * k4 + kk) * NOUT + c);
                    w[c] = t.x;
                    w[c + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPT; ++c) w[c] = w_ptr[(4 * k4 + kk) * NOUT + c];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
            }
        }
    }
}

template <int NOUT>
__device__ __forceinline__ void layer3_store(const float *__restrict__ small_net, float *O, int s, int cg,
                                             float (&acc)[3][NOUT / 8]) {
    constexpr int CPT = NOUT / 8;
    const float *b3 = small_net + DP_SM_B3;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int col = CPT * cg + c;
        O[(s * 3 + 0) * NOUT + col] = acc[0][c] + __ldg(b3 + col);
        O[(s * 3 + 1) * NOUT + col] = acc[1][c];
        O[(s * 3 + 2) * NOUT + col] = acc[2][c];
    }
}

__global__ void __launch_bounds__(NT, 1) rollout_ffma_kernel(RolloutParams p) {
    extern __shared__ __align__(16) float smem[];
    float *sW = smem;
    float *sA = sW + SM_W;
    float *sOb = sA + SM_A;
    float *sIn = sOb + SM_OB;
    float *sOut = sIn + SM_IN;
    float *sOa = sA;  // alias: valid between the last GEMM of a step and the next step's layer 1

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // GEMM tile coordinates
    const int s = tid >> 3, cg = tid & 7;    // per-sample coordinates (8 lanes per sample)
    const bool leader = cg == 0;

    // resident weights: one pass over 164 KB
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.big);
        float4 *dst = reinterpret_cast<float4 *>(sW);
        for (int i = tid; i < SM_W / 4; i += NT) dst[i] = __ldg(src + i);
    }
    const float *small_a = p.small + DP_SM_NET_A;
    const float *small_b = p.small + DP_SM_NET_B;
    const int T = p.T, T1 = p.T + 1;
    const bool logd = p.flags & DP_LOG_DENSITY, dens = p.flags & DP_DENSITY, cut = p.flags & DP_CUT;
    const bool abs_state = p.flags & DP_ABS_STATE;
    const int ndim_out = (p.flags & DP_XY_ONLY) ? 2 : 5;
    const float dt = p.dt;

    const long long num_tiles = (p.N + S - 1) / S;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long n = tile * S + s;
        const bool valid = n < p.N;
        const long long nl = valid ? n : p.N - 1;
        float x0 = 0, x1 = 0, x2 = 0, x3 = 0, x4 = 0, rho = 0, margin = 1e30f;
        if (leader) {
            x0 = __ldg(p.xe0 + nl * 5 + 0) + __ldg(p.xref + 0 * T1);
            x1 = __ldg(p.xe0 + nl * 5 + 1) + __ldg(p.xref + 1 * T1);
            x2 = __ldg(p.xe0 + nl * 5 + 2) + __ldg(p.xref + 2 * T1);
            x3 = __ldg(p.xe0 + nl * 5 + 3) + __ldg(p.xref + 3 * T1);
            x4 = __ldg(p.xe0 + nl * 5 + 4) + __ldg(p.xref + 4 * T1);
            rho = p.rho0 ? __ldg(p.rho0 + nl) : (logd ? 0.0f : 1.0f);
        }
        __syncthreads();  // previous tile's readers of sIn/sOut are done (and weights are loaded on the first pass)

        for (int i = 0;; ++i) {
            const bool store_now = (i % p.stride) == 0;
            float ur0 = 0, ur1 = 0;
            if (leader) {
                const float xr0 = __ldg(p.xref + 0 * T1 + i), xr1 = __ldg(p.xref + 1 * T1 + i);
                const float xr2 = __ldg(p.xref + 2 * T1 + i), xr3 = __ldg(p.xref + 3 * T1 + i);
                if (store_now) {
                    if (abs_state) {
                        sOut[0 * S + s] = x0; sOut[1 * S + s] = x1; sOut[2 * S + s] = x2; sOut[3 * S + s] = x3;
                        sOut[4 * S + s] = x4;
                    } else {
                        sOut[0 * S + s] = x0 - xr0; sOut[1 * S + s] = x1 - xr1; sOut[2 * S + s] = x2 - xr2;
                        sOut[3 * S + s] = x3 - xr3; sOut[4 * S + s] = x4 - __ldg(p.xref + 4 * T1 + i);
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
                    sOut[5 * S + s] = r;
                }
                if (i < T) {
                    // Controller.__call__ (sytem_CAR.py:33-34) and the network input of U_FUNC.forward (model_CAR.py:24)
                    const float e0 = x0 - xr0, e1 = x1 - xr1, e3 = x3 - xr3;
                    const float e2 = __fadd_rn(__fsub_rn(x2, xr2), x4);
                    float4 *dst = reinterpret_cast<float4 *>(sIn + 8 * s);
                    dst[0] = make_float4(x2, x3, x2 - e2, x3 - e3);
                    dst[1] = make_float4(e0, e1, e2, e3);
                    ur0 = __ldg(p.uref + i);
                    ur1 = __ldg(p.uref + T + i);
                }
            }
            __syncthreads();  // S1: sIn / sOut visible; O_a of the previous step no longer read
            if (store_now && tid < 32 * 6) {
                // coalesced store of the staged slice: 32 consecutive samples per (slot, dim)
                const int d = tid >> 5, j = tid & 31;
                const long long nn = tile * S + j;
                if (nn < p.N) {
                    const long long slot = i / p.stride;
                    if (d < 5) {
                        if (p.x_out && d < ndim_out) p.x_out[(slot * ndim_out + d) * p.ldn + nn] = sOut[d * S + j];
                    } else if (p.rho_out && dens) {
                        p.rho_out[slot * p.ldn + nn] = sOut[5 * S + j];
                    }
                }
            }
            if (i == T) break;

            float acc[6][8];
            // ---- net b (w2): 4 -> 128 -> 128 -> 24 ----
            layer1(small_b, sIn, sA, ty, tx);
            __syncthreads();
            layer2_mainloop(sA, sW + DP_BIG_W2B, ty, tx, acc);
            __syncthreads();
            layer2_epilogue(small_b, sA, ty, tx, acc);
            __syncthreads();
            {
                float acc3[3][DP_NOUT_B / 8];
                layer3_mainloop<DP_NOUT_B>(sA, sW + DP_BIG_W3B, s, cg, acc3);
                layer3_store<DP_NOUT_B>(small_b, sOb, s, cg, acc3);
            }
            __syncthreads();  // all reads of A done before net a overwrites it
            // ---- net a (w1): 4 -> 128 -> 128 -> 48 ----
            layer1(small_a, sIn, sA, ty, tx);
            __syncthreads();
            layer2_mainloop(sA, sW + DP_BIG_W2A, ty, tx, acc);
            __syncthreads();
            layer2_epilogue(small_a, sA, ty, tx, acc);
            __syncthreads();
            {
                float acc3[3][DP_NOUT_A / 8];
                layer3_mainloop<DP_NOUT_A>(sA, sW + DP_BIG_W3A, s, cg, acc3);
                __syncthreads();  // A fully consumed -> reuse as O_a
                layer3_store<DP_NOUT_A>(small_a, sOa, s, cg, acc3);
            }
            __syncthreads();

            // ---- head: u = w2 tanh(w1 xe) + uref and the two Jacobian entries on the diagonal of df/dx ----
            float pu0 = 0, pu1 = 0, pd0 = 0, pd1 = 0;
            {
                const float4 xe = *reinterpret_cast<const float4 *>(sIn + 8 * s + 4);
                const float *oa = sOa + s * 3 * DP_NOUT_A;
                const float *ob = sOb + s * 3 * DP_NOUT_B;
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    const int r = cg + 8 * ii;
                    if (r < 12) {
                        const float4 w1r = *reinterpret_cast<const float4 *>(oa + 4 * r);
                        const float4 dth = *reinterpret_cast<const float4 *>(oa + DP_NOUT_A + 4 * r);
                        const float4 dv = *reinterpret_cast<const float4 *>(oa + 2 * DP_NOUT_A + 4 * r);
                        float z = w1r.x * xe.x;
                        z = fmaf(w1r.y, xe.y, z); z = fmaf(w1r.z, xe.z, z); z = fmaf(w1r.w, xe.w, z);
                        float dzt = dth.x * xe.x;
                        dzt = fmaf(dth.y, xe.y, dzt); dzt = fmaf(dth.z, xe.z, dzt); dzt = fmaf(dth.w, xe.w, dzt);
                        dzt += w1r.z;  // d xe2 / d theta = 1
                        float dzv = dv.x * xe.x;
                        dzv = fmaf(dv.y, xe.y, dzv); dzv = fmaf(dv.z, xe.z, dzv); dzv = fmaf(dv.w, xe.w, dzv);
                        dzv += w1r.w;  // d xe3 / d v = 1
                        const float tz = dp_tanh(z);
                        const float gz = 1.0f - tz * tz;
                        const float w20 = ob[r], w21 = ob[12 + r];
                        pu0 = fmaf(w20, tz, pu0);
                        pu1 = fmaf(w21, tz, pu1);
                        pd0 += ob[DP_NOUT_B + r] * tz + w20 * (gz * dzt);           // d u0 / d theta
                        pd1 += ob[2 * DP_NOUT_B + 12 + r] * tz + w21 * (gz * dzv);  // d u1 / d v
                    }
                }
            }
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                pu0 += __shfl_xor_sync(0xffffffffu, pu0, m);
                pu1 += __shfl_xor_sync(0xffffffffu, pu1, m);
                pd0 += __shfl_xor_sync(0xffffffffu, pd0, m);
                pd1 += __shfl_xor_sync(0xffffffffu, pd1, m);
            }
            if (leader) {
                const float u0r = pu0 + ur0, u1r = pu1 + ur1;
                margin = fminf(margin, fminf(fabsf(fabsf(u0r) - 3.0f), fabsf(fabsf(u1r) - 3.0f)));
                if (dens) {
                    // clip mask inclusive at the bounds, as the backward of torch.clip
                    const float m0 = (u0r >= -3.0f && u0r <= 3.0f) ? 1.0f : 0.0f;
                    const float m1 = (u1r >= -3.0f && u1r <= 3.0f) ? 1.0f : 0.0f;
                    const float div = __fadd_rn(__fmul_rn(m0, pd0), __fmul_rn(m1, pd1));
                    if (logd)
                        rho = __fadd_rn(rho, logf(__fsub_rn(1.0f, __fmul_rn(div, dt))));  // :153-156
                    else
                        rho = __fadd_rn(rho, __fmul_rn(__fmul_rn(-div, rho), dt));  // :142-144
                }
                const float u0 = fminf(fmaxf(u0r, -3.0f), 3.0f), u1 = fminf(fmaxf(u1r, -3.0f), 3.0f);
                float sn, cs;
                sincosf(x2, &sn, &cs);
                // get_next_x (:124-128): x + f*dt with f = [v cos th, v sin th, u0, u1, 0]
                x0 = __fadd_rn(x0, __fmul_rn(__fmul_rn(x3, cs), dt));
                x1 = __fadd_rn(x1, __fmul_rn(__fmul_rn(x3, sn), dt));
                x2 = __fadd_rn(x2, __fmul_rn(u0, dt));
                x3 = __fadd_rn(x3, __fmul_rn(u1, dt));
            }
        }
        if (leader && valid && p.clip_margin) p.clip_margin[n] = margin;
    }
}

}  // namespace

int dp_launch_rollout_ffma(const dp_controller *c, const RolloutParams &p, cudaStream_t stream) {
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        DP_CUDA(cudaFuncSetAttribute(rollout_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set[c->device & 63] = true;
    }
    const long long num_tiles = (p.N + S - 1) / S;
    const int grid = (int)(num_tiles < c->num_sms ? num_tiles : c->num_sms);
    rollout_ffma_kernel<<<grid, NT, SMEM_BYTES, stream>>>(p);
    DP_CUDA(cudaGetLastError());
    return 0;
}
