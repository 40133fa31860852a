/*
 * density_planner_b200_diag.h -- diagnostics library (libdensity_planner_b200_diag.so).  NOT part of the product ABI:
 * the fp32 cross-check rollout used by the parity tests and micro-probes of the hardware paths the rollout kernel uses.
 * Every function returns 0 on success; dp_diag_last_error() returns a thread-local message for the last failure.
 */
#ifndef DENSITY_PLANNER_B200_DIAG_H
#define DENSITY_PLANNER_B200_DIAG_H

#include "density_planner_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *dp_diag_last_error(void);

/* dp_rollout (same arguments, same outputs) evaluated by the all-fp32 CUDA-core kernel (32-sample tiles, weights resident
 * in shared memory, register-tiled SGEMM): the independent fp32 cross-check of the tcgen05 kernel.  `c` is a controller
 * created by the product library. */
int dp_diag_rollout_ffma(const dp_controller *c, const float *xe0, const float *xref, const float *uref, const float *rho0,
                         float dt, int64_t N, int T, unsigned flags, int stride, float *x_out, float *rho_out, int64_t ldn,
                         float *clip_margin, int *status, void *stream);

/* One 128 x n x 128 product through the TMEM-operand tcgen05.mma path of the tensor-core rollout (split-precision fp16
 * operands).  A host fp32 [128][128], W host fp32 [n_valid][128] zero-padded to n rows (n in {128, 48, 32}), mode 0 =
 * value path (3 products), 1 = fp16 activations x (hi + lo) weights (2 products); D host fp32 [128][n] receives
 * A * (scale * W)^T. */
int dp_tc_selftest(int device, const float *A, const float *W, int n_valid, int n, int mode, float scale, float *D);

/* Cycles per tcgen05.mma (M=128, N=n, K=16, fp16 operands, fp32 accumulate) over 4*iters back-to-back instructions on one
 * SM.  variant 0: A in TMEM + B no-swizzle (the rollout's configuration), 1: A in TMEM + B SWIZZLE_128B, 2: A and B in
 * shared memory no swizzle, 3: both SWIZZLE_128B.  out[0] issue, out[1] completion. */
int dp_tc_mma_bench(int device, int variant, int n, int iters, double *out);

/* Cycles per warp-level tcgen05.ld (mode 0) / tcgen05.st (mode 1) / dependent ld+st pair (mode 2) of shape 32x32b.x16
 * (2 KB per warp) while all 16 warps of one 512-thread CTA issue them: the TMEM bandwidth the epilogues see. */
int dp_tc_tmem_bench(int device, int mode, int iters, double *cycles_per_op);

/* FP32 FFMA peak probe (the CUDA-core roofline the fp32 cross-check kernel is bounded by); TFLOP/s in *tflops */
int dp_probe_ffma_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* DENSITY_PLANNER_B200_DIAG_H */
