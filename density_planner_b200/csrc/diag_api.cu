// Diagnostics library (libdensity_planner_b200_diag.so, include/density_planner_b200_diag.h): NOT part of the product ABI.
//   * dp_diag_rollout_ffma: the all-fp32 CUDA-core rollout (rollout_ffma.cu), kept as the fp32 cross-check of the tcgen05
//     kernel in the parity tests; takes a dp_controller created by the product library (same struct, same sources)
//   * dp_probe_ffma_peak, dp_tc_selftest, dp_tc_mma_bench, dp_tc_tmem_bench (tc_diag.cu): micro-probes
// Built from the same headers with -Ddp_set_error=dp_diag_set_error so the two libraries never share an error slot.
#include <stdarg.h>

#include "dp_common.cuh"
#include "../../include/density_planner_b200_diag.h"

static thread_local char g_diag_err[1024] = "";

void dp_set_error(const char *fmt, ...) {  // = dp_diag_set_error in this library
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_diag_err, sizeof(g_diag_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *dp_diag_last_error(void) { return g_diag_err; }

extern "C" int dp_diag_rollout_ffma(const dp_controller *c, const float *xe0, const float *xref, const float *uref,
                                    const float *rho0, float dt, int64_t N, int T, unsigned flags, int stride, float *x_out,
                                    float *rho_out, int64_t ldn, float *clip_margin, int *status, void *stream) {
    DP_REQUIRE(c && xe0 && xref && uref, "dp_diag_rollout_ffma: null argument");
    DP_REQUIRE(T >= 0 && stride >= 1, "dp_diag_rollout_ffma: bad T=%d / stride=%d", T, stride);
    DP_REQUIRE(ldn >= N || (!x_out && !rho_out), "dp_diag_rollout_ffma: ldn=%lld < N=%lld", (long long)ldn, (long long)N);
    if (N <= 0) return 0;
    DeviceGuard guard(c->device);
    RolloutParams p;
    memset(&p, 0, sizeof(p));
    p.big = c->d_big; p.small = c->d_small;
    p.xe0 = xe0; p.xref = xref; p.uref = uref; p.rho0 = rho0;
    p.dt = dt; p.N = N; p.T = T; p.flags = flags; p.stride = stride; p.Ts = T / stride + 1;
    p.x_out = x_out; p.rho_out = rho_out; p.ldn = ldn; p.clip_margin = clip_margin; p.status = status;
    p.R = 1; p.n_per_ref = N;
    return dp_launch_rollout_ffma(c, p, (cudaStream_t)stream);
}

// ---- FP32 FFMA peak probe (the CUDA-core roofline the fp32 cross-check kernel is bounded by) --------------------------------
namespace {
__global__ void __launch_bounds__(256) ffma_probe_kernel(float *out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}
}  // namespace

extern "C" int dp_probe_ffma_peak(int device, double *tflops) {
    DP_REQUIRE(tflops, "dp_probe_ffma_peak: null argument");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    DP_CUDA(cudaGetDeviceProperties(&prop, device));
    float *d;
    DP_CUDA(cudaMalloc(&d, 4));
    const int iters = 20000, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    DP_CUDA(cudaEventCreate(&e0));
    DP_CUDA(cudaEventCreate(&e1));
    ffma_probe_kernel<<<blocks, 256>>>(d, 1000, 0.999f, 0.001f);
    DP_CUDA(cudaDeviceSynchronize());
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        DP_CUDA(cudaEventRecord(e0));
        ffma_probe_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
        DP_CUDA(cudaEventRecord(e1));
        DP_CUDA(cudaEventSynchronize(e1));
        float ms;
        DP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 128.0 * iters * 256.0 * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return 0;
}
