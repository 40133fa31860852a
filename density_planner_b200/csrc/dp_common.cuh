// Shared declarations for the density_planner_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/density_planner_b200.h"

#define DP_HID 128
#define DP_NOUT_A 48
#define DP_NOUT_B 24

// ---- device weight layout -------------------------------------------------------------------------------
// big   : W2t_a[128][128] | W2t_b[128][128] | W3t_a[128][48] | W3t_b[128][24]      (k-major = transposed)
// small : W1a[128][4] b1a[128] b2a[128] b3a[48] | W1b[128][4] b1b[128] b2b[128] b3b[24]
#define DP_BIG_W2A 0
#define DP_BIG_W2B (DP_HID * DP_HID)
#define DP_BIG_W3A (2 * DP_HID * DP_HID)
#define DP_BIG_W3B (2 * DP_HID * DP_HID + DP_HID * DP_NOUT_A)
#define DP_BIG_FLOATS (2 * DP_HID * DP_HID + DP_HID * (DP_NOUT_A + DP_NOUT_B))
#define DP_SM_W1 0
#define DP_SM_B1 (DP_HID * 4)
#define DP_SM_B2 (DP_HID * 5)
#define DP_SM_B3 (DP_HID * 6)
#define DP_SM_NET_A 0
#define DP_SM_NET_B (DP_HID * 6 + DP_NOUT_A)
#define DP_SMALL_FLOATS (2 * DP_HID * 6 + DP_NOUT_A + DP_NOUT_B)

struct dp_workspace {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct dp_controller {
    int device = 0;
    int num_sms = 0;
    float *d_big = nullptr;
    float *d_small = nullptr;
    // tensor-core operand images (built by rollout_tc.cu at create time; may be null)
    void *d_tc = nullptr;
    size_t tc_bytes = 0;
    bool tc_l2_bounded = false;  // layer-2 pre-activations provably small enough for the sign-free tanh (rollout_tc.cu)
    float tc_l2_bound = 0.0f;    // the bound itself (scaled pre-activation units)
    // grow-only device workspace + stream for the *_host entry points
    dp_workspace ws;
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    std::mutex host_mutex;  // the *_host entry points share the workspace and the streams above: one call at a time
};

struct dp_env {
    int device = 0;
    dp_grid_geom geom;
    int Tg = 0;
    float4 *d_table = nullptr;  // [Tg][gx][gy] {p, des_pos_x, des_pos_y, active} (cost.cu env_pack_kernel)
};

// ---- error plumbing ---------------------------------------------------------------------------------------
void dp_set_error(const char *fmt, ...);
#define DP_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            dp_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
            return 1;                                                                                     \
        }                                                                                                 \
    } while (0)
#define DP_REQUIRE(cond, ...)      \
    do {                           \
        if (!(cond)) {             \
            dp_set_error(__VA_ARGS__); \
            return 2;              \
        }                          \
    } while (0)

struct DeviceGuard {
    int old = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&old);
        if (old != dev) cudaSetDevice(dev);
        else old = -1;
    }
    ~DeviceGuard() {
        if (old >= 0) cudaSetDevice(old);
    }
};

// ---- rollout parameter block (shared by the FFMA and tensor-core kernels) -----------------------------------
struct RolloutParams {
    const float *big;
    const float *small;
    const void *tc;
    const float *xe0;
    const float *xref;
    const float *uref;
    const float *rho0;
    float dt;
    long long N;
    int T;
    unsigned flags;
    int stride;
    int Ts;
    float *x_out;
    float *rho_out;
    long long ldn;
    float *clip_margin;
    int *status;
    unsigned long long *trace;  // optional in-kernel timeline (DP_TC_TRACE), normally null
    // several references in one launch (dp_rollout_batch): xref [R][5][T+1], uref [R][2][T], n_per_ref samples each,
    // optional per-reference horizons T_ref[R] <= T (device); R <= 1: one shared reference
    long long R, n_per_ref;
    const int *T_ref;
    // optional: per stored slot, the maximum of the stored (log-)density over the samples, as an order-preserving
    // unsigned image of the float (dp_decode_ordered); zero-initialised by the caller (pipeline.cu)
    unsigned *lmax_enc;
};

// order-preserving unsigned image of a float (atomicMax on floats of either sign) and its inverse
__host__ __device__ inline unsigned dp_encode_ordered(float f) {
    unsigned b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float dp_decode_ordered(unsigned e) {
    const unsigned b = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

int dp_launch_rollout_ffma(const dp_controller *c, const RolloutParams &p, cudaStream_t stream);
int dp_launch_rollout_tc(const dp_controller *c, const RolloutParams &p, cudaStream_t stream);
int dp_tc_prepare(dp_controller *c, const float *host_blob);  // builds c->d_tc
int dp_bin2d_fused_launch(const dp_bin_spec *spec, const dp_env *env, const float *x, const float *y, int64_t sx_t,
                          const float *l, int64_t sl_t, const float *rho0, const float *lmax, int64_t N, int Ts,
                          long long *mass, int *count, long long *cost_q, int cost_scale_log2, cudaStream_t stream);
void dp_tc_release(dp_controller *c);

// ---- small device helpers ----------------------------------------------------------------------------------
// Correctly rounded a / b without the branchy IEEE division sequence: q = RN(a * RN(1/b)), one exact FMA residual and one
// FMA correction (Markstein).  Checked exhaustively against `a / b` for b in {24, 40, 240, 400} over ALL finite floats a:
// the only mismatches have |a| < 1.9e-37 (subnormal intermediate products); below 1e-30 the IEEE division is used.
__device__ __forceinline__ float dp_div(float a, float b, float rb) {
    const float q = __fmul_rn(a, rb);
    const float r = __fmaf_rn(-q, b, a);
    const float q1 = __fmaf_rn(r, rb, q);
    return (fabsf(a) >= 1e-30f || a == 0.0f) ? q1 : __fdiv_rn(a, b);
}

// grid geometry with the reciprocals dp_div needs (hyperparams.py:124,154-155)
struct DpGeom {
    float x_lo, x_rng, y_lo, y_rng, gxm1, gym1;
    float r_x_rng, r_y_rng, r_gxm1, r_gym1;
    int gx, gy;
};
inline DpGeom dp_make_geom(const dp_grid_geom &g) {
    DpGeom q;
    q.x_lo = g.x_min;
    q.x_rng = g.x_max - g.x_min;
    q.y_lo = g.y_min;
    q.y_rng = g.y_max - g.y_min;
    q.gxm1 = (float)(g.gx - 1);
    q.gym1 = (float)(g.gy - 1);
    q.r_x_rng = (float)(1.0 / (double)q.x_rng);
    q.r_y_rng = (float)(1.0 / (double)q.y_rng);
    q.r_gxm1 = (float)(1.0 / (double)q.gxm1);
    q.r_gym1 = (float)(1.0 / (double)q.gym1);
    q.gx = g.gx;
    q.gy = g.gy;
    return q;
}

// The same without the guard, for the hot kernels: a mismatch needs 0 < |a| < 1.9e-37, i.e. a sample within 1e-37 m of the
// grid's lower edge (the cell index rounds to 0 either way) or a gradient-shifted grid position of that magnitude
// (the position then equals the lower edge either way unless that edge is exactly 0).
__device__ __forceinline__ float dp_div_fast(float a, float b, float rb) {
    const float q = __fmul_rn(a, rb);
    return __fmaf_rn(__fmaf_rn(-q, b, a), rb, q);
}

// pos2gridpos for one coordinate, fp32 op for op as torch evaluates motion_planning/utils.py:82-85
template <bool GUARD = true>
__device__ __forceinline__ int dp_cell(float p, float lo, float range, float r_range, float gm1) {
    const float a = __fsub_rn(p, lo);
    const float q = __fmul_rn(GUARD ? dp_div(a, range, r_range) : dp_div_fast(a, range, r_range), gm1);
    return __float2int_rn(__fadd_rn(q, 0.001f));  // round-half-even like torch.round; values are far from int overflow
}
// The same cell, clamped to [0, hi], as an exactly representable FLOAT and as an int, without the XU-pipe conversions
// (F2I for the cell, I2F for the gradient-shifted grid position of the cost): clamp in float, then the magic-number add
// (t + 1.5 * 2^23 carries rn(t) in its low mantissa bits, ties to even like cvt.rni / torch.round).  Rounding is monotone and
// fixes the integers 0 and hi, so rn(clamp(t, 0, hi)) = clamp(rn(t), 0, hi): identical cells (NaN -> 0 like F2I).
__device__ __forceinline__ void dp_cell_clamped(float p, float lo, float range, float r_range, float gm1, float hi, int &cell,
                                                float &cell_f) {
    const float q = __fmul_rn(dp_div_fast(__fsub_rn(p, lo), range, r_range), gm1);
    const float t = fminf(fmaxf(__fadd_rn(q, 0.001f), 0.0f), hi);
    const float m = __fadd_rn(t, 12582912.0f);
    cell = __float_as_int(m) - 0x4B400000;
    cell_f = __fsub_rn(m, 12582912.0f);
}

// L2 prefetch of the line holding p (a hint: the streaming kernels issue it one grid-stride iteration ahead)
__device__ __forceinline__ void dp_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// gridpos2pos for one coordinate (motion_planning/utils.py:100-116)
template <bool GUARD = true>
__device__ __forceinline__ float dp_pos(float g, float lo, float range, float gm1, float r_gm1) {
    return __fadd_rn(__fmul_rn(GUARD ? dp_div(g, gm1, r_gm1) : dp_div_fast(g, gm1, r_gm1), range), lo);
}

// one (sample, slice) evaluation of the collision cost (MotionPlanner.get_cost_coll, motion_planning/MotionPlanner.py:206-222;
// shared by cost.cu and the fused binning pass of bin.cu).  Returns the fp32 term rho*p*sq (0 if inactive) and optionally the gradients.
struct DpEval {
    float term, gx, gy, grho;
    bool active;
};

// the terms of one evaluation from the gathered table entry e = {p, des_pos_x, des_pos_y, active} of the (clamped) cell: the
// desired position (gridpos + 100 * grad through gridpos2pos, MotionPlanner.py:215-217) is a property of the cell and comes
// precomputed from dp_env_create; what is left per sample is the squared distance and the weighting (:218-222), each op
// rounded like torch
__device__ __forceinline__ DpEval dp_cost_term(const float4 e, float x, float y, float rho, bool has_rho) {
    DpEval r;
    r.active = e.w != 0.0f;  // :213-214
    const float ex = __fsub_rn(e.y, x), ey = __fsub_rn(e.z, y);
    const float sq = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    const float w = has_rho ? __fmul_rn(rho, e.x) : e.x;
    r.term = r.active ? __fmul_rn(w, sq) : 0.0f;
    r.gx = r.active ? __fmul_rn(-2.0f, __fmul_rn(w, ex)) : 0.0f;
    r.gy = r.active ? __fmul_rn(-2.0f, __fmul_rn(w, ey)) : 0.0f;
    r.grho = r.active ? __fmul_rn(e.x, sq) : 0.0f;
    return r;
}

__device__ __forceinline__ DpEval dp_cost_eval(const float4 *__restrict__ table, const DpGeom &q, int t, float x, float y, float rho,
                                               bool has_rho) {
    int ix, iy;
    float fx, fy;
    dp_cell_clamped(x, q.x_lo, q.x_rng, q.r_x_rng, q.gxm1, q.gxm1, ix, fx);  // clamp to [0, G-1] (:208-209)
    dp_cell_clamped(y, q.y_lo, q.y_rng, q.r_y_rng, q.gym1, q.gym1, iy, fy);
    const float4 e = __ldg(table + ((size_t)t * q.gx + ix) * q.gy + iy);  // {p, des_pos_x, des_pos_y, active}
    return dp_cost_term(e, x, y, rho, has_rho);
}
