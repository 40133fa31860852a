// Shared declarations for the density_planner_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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
    // grow-only device workspace + stream for the *_host entry points
    dp_workspace ws;
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
};

struct dp_env {
    int device = 0;
    dp_grid_geom geom;
    int Tg = 0;
    float4 *d_table = nullptr;  // [Tg][gx][gy] {p, gradX, gradY, 0}
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
};

int dp_launch_rollout_ffma(const dp_controller *c, const RolloutParams &p, cudaStream_t stream);
int dp_launch_rollout_tc(const dp_controller *c, const RolloutParams &p, cudaStream_t stream);
int dp_tc_prepare(dp_controller *c, const float *host_blob);  // builds c->d_tc
void dp_tc_release(dp_controller *c);

// ---- small device helpers ----------------------------------------------------------------------------------
// pos2gridpos for one coordinate, fp32 op for op as torch evaluates motion_planning/utils.py:82-85
__device__ __forceinline__ int dp_cell(float p, float lo, float range, float gm1) {
    float q = __fmul_rn(__fdiv_rn(__fsub_rn(p, lo), range), gm1);
    return __float2int_rn(__fadd_rn(q, 0.001f));  // round-half-even like torch.round; values are far from int overflow
}
