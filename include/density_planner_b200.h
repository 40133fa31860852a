/*
 * density_planner_b200.h -- C ABI of the B200-native hot path of MIT-REALM/density_planner.
 *
 * The reference is pure Python and has NO plugin / FFI interface for this path: the
 * boundary it exposes is the Python method surface of systems/ControlAffineSystem.py, systems/sytem_CAR.py,
 * motion_planning/utils.py and motion_planning/MotionPlanner.py.  Each entry point below names the reference
 * method(s) (file:line under the reference root) whose arithmetic it replaces; INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.  density_planner_b200/ (Python) mirrors the reference
 * method names on top of this ABI.
 *
 * Conventions
 *   - plain C, no C++/torch types.  Every function returns 0 on success, non-zero on error;
 *     dp_last_error() returns a thread-local message for the last failure.
 *   - "dev" pointers are device memory owned by the caller (e.g. torch allocations); "host" pointers are host
 *     memory (pinned memory makes the *_host calls faster, pageable works).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device-pointer calls are asynchronous
 *     on that stream; *_host calls copy in, run, copy out and synchronise before returning.
 *   - all floating point data is fp32 unless noted.  Trajectories are stored time-major SoA:
 *         x[t][d][n]  (t = stored time slot, d = state dim 0..4, n = sample)   rho[t][n]
 *     which is the transposed view of the reference's (N, 5, T+1) / (N, 1, T+1) tensors.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef DENSITY_PLANNER_B200_H
#define DENSITY_PLANNER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_ABI_VERSION 1

/* ---- flags for dp_rollout* ------------------------------------------------------------------------------- */
#define DP_LOG_DENSITY 1u  /* log_density=True: l += log(1 - div f * dt)  (ControlAffineSystem.py:147-157)      */
#define DP_DENSITY 2u      /* compute_density=True (:409); off => rho_out untouched                             */
#define DP_CUT 4u          /* cutting=True: clamp >1e30 (and <0 if linear), NaN -> 1e30 (:451-462)              */
#define DP_ABS_STATE 8u    /* store x instead of the error state x - xref (the reference returns x - xref)      */
#define DP_XY_ONLY 16u     /* store only state dims 0,1 (x_out is [Ts][2][ldn]) -- input of the collision cost  */
/* There is ONE rollout backend: the tcgen05 tensor-core kernel (split-precision fp16 operands, fp32-level accuracy).
 * The all-fp32 CUDA-core cross-check kernel and the micro-probes live in the separate diagnostics library
 * (include/density_planner_b200_diag.h), not in this ABI. */

/* number of fp32 values in a packed controller blob (see dp_controller_create) */
#define DP_CONTROLLER_BLOB_FLOATS 43592

typedef struct dp_controller dp_controller; /* weights are immutable once created; the device-pointer calls may share a
                                             * handle between threads; the *_host calls serialise on its host workspace */
typedef struct dp_env dp_env;               /* occupancy grid + gradients of one environment, repacked on device */

int dp_abi_version(void);
const char *dp_last_error(void);
/* 1 if the CUDA device `device` is an sm_100 part, 0 if not, <0 on error */
int dp_device_ok(int device);

/*
 * The C3M controller (data/trained_controller/model_CAR.py:19-28, loaded at systems/sytem_CAR.py:113-125).
 * host_blob: fp32, per net in order {model_u_w1 (NOUT=48), model_u_w2 (NOUT=24)}:
 *     W1[128][4] b1[128] W2[128][128] b2[128] W3[NOUT][128] b3[NOUT]       (torch Linear.weight layout [out][in])
 */
int dp_controller_create(const float *host_blob, size_t n_floats, int device, dp_controller **out);
void dp_controller_destroy(dp_controller *c);

/*
 * Fused closed-loop rollout with Liouville density.  Replaces the Python loop of
 * ControlAffineSystem.compute_density (systems/ControlAffineSystem.py:409-463), i.e. per step
 * get_next_rholog/get_next_rho (:136-157) -> divf_func (:116) -> dfdx_func (:93) -> dudx_func (:69) -> jacobian
 * (systems/utils.py:5) and get_next_x (:124) -> f_func (:43) -> Controller.__call__ (systems/sytem_CAR.py:25).
 *
 *   xe0   dev [N][5]      initial error states           xref dev [5][T+1]     uref dev [2][T]
 *   rho0  dev [N] or NULL initial (log-)density (NULL: 0 for log, 1 for linear)
 *   stride               store every `stride`-th step (t = 0, stride, ...), Ts = T/stride + 1 slots
 *   x_out   dev [Ts][5][ldn] (or [Ts][2][ldn] with DP_XY_ONLY) or NULL
 *   rho_out dev [Ts][ldn] or NULL
 *   clip_margin dev [N] or NULL: min over (step, input) of | |u_raw| - 3 |  (distance to the actuator clip; audit)
 *   status  dev int[2] or NULL: [0] |= 1 if a density was clamped, [1] |= 1 if a NaN was replaced (DP_CUT)
 */
int dp_rollout(const dp_controller *c, const float *xe0, const float *xref, const float *uref, const float *rho0,
               float dt, int64_t N, int T, unsigned flags, int stride, float *x_out, float *rho_out, int64_t ldn,
               float *clip_margin, int *status, void *stream);

/*
 * Several references in one launch (the raw-data generator's 20-sample trajectories,
 * data_generation/compute_rawdata.py:26, cannot fill a GPU one at a time): R references with n_per_ref samples each.
 *   xe0 dev [R*n_per_ref][5];  xref dev [R][5][T+1];  uref dev [R][2][T] (rows padded to the common T)
 *   T_ref dev int[R] or NULL: horizon of every reference (<= T; slots beyond it are left untouched)
 *   outputs as dp_rollout with N = R*n_per_ref samples, reference r owning samples [r*n_per_ref, (r+1)*n_per_ref)
 * References of at most 64 samples are packed, 128 / n_per_ref of them per 128-lane tile (the reference row, its horizon and
 * the reference speed are per-lane properties there), so R references of 20 samples fill 120 of 128 lanes; larger references
 * own whole tiles.  Either way the results equal R separate dp_rollout calls bit for bit.
 */
int dp_rollout_batch(const dp_controller *c, const float *xe0, const float *xref, const float *uref, const int *T_ref,
                     const float *rho0, float dt, int64_t R, int64_t n_per_ref, int T, unsigned flags, int stride,
                     float *x_out, float *rho_out, int64_t ldn, float *clip_margin, int *status, void *stream);

/* Same with HOST buffers (x_out/rho_out host, ld = N); status is a host int[2].  Copies are part of the call: the samples
 * are processed in chunks and the device->host copy of chunk k overlaps the rollout of chunk k+1.  Thread safe on a shared
 * controller: concurrent calls serialise on the controller's grow-only device workspace. */
int dp_rollout_host(dp_controller *c, const float *xe0, const float *xref, const float *uref, const float *rho0,
                    float dt, int64_t N, int T, unsigned flags, int stride, float *x_out, float *rho_out,
                    float *clip_margin, int *status);

/* Chunk sizes dp_rollout_host uses for N samples when one wave (two 128-sample tiles per SM) holds `wave` samples: writes
 * up to max_chunks sizes, returns their number (negative: max_chunks too small), 0 for N <= 0.  Host logic only, no device. */
int dp_host_chunk_schedule(int64_t N, int64_t wave, int64_t *sizes, int max_chunks);

/*
 * Single-step API: u_func (:26), f_func (:43), dudx_func (:69), dfdx_func (:93), divf_func (:116) of
 * systems/ControlAffineSystem.py for N states sharing one (xref, uref).
 *   x dev [N][5]; xref5 dev [5]; uref2 dev [2]
 *   outputs (each may be NULL): u_raw [N][2], u [N][2] (clipped), dudx [N][2][5] (clip mask applied, as autograd
 *   of torch.clip gives), f [N][5], div [N]
 */
int dp_step(const dp_controller *c, const float *x, const float *xref5, const float *uref2, int64_t N, float *u_raw,
            float *u, float *dudx, float *f, float *div, void *stream);

/* ---- grid geometry (hyperparams.py:124,154-155) ----------------------------------------------------------- */
typedef struct dp_grid_geom {
    float x_min, x_max, y_min, y_max; /* args.environment_size */
    int gx, gy;                       /* args.grid_size        */
} dp_grid_geom;

/* pos2gridpos (motion_planning/utils.py:70-97): round_half_even(((p - lo)/(hi - lo))*(G-1) + 0.001), fp32 op for op.
 * px, py dev [n] -> ix, iy dev int32 [n] (either pair may be NULL).  clamp_hi < 0: no clamp, else clamp to [0,clamp] */
int dp_pos2gridpos(const dp_grid_geom *g, const float *px, const float *py, int64_t n, int32_t *ix, int32_t *iy,
                   int clamp_hi_x, int clamp_hi_y, void *stream);
/* The same for float64 positions (python floats, lists and numpy float64 arrays are evaluated in double precision by the
 * reference, :79-85): px, py dev double[n] -> ix, iy dev int64[n], no clamp */
int dp_pos2gridpos_f64(const dp_grid_geom *g, const double *px, const double *py, int64_t n, int64_t *ix, int64_t *iy,
                       void *stream);
/* gridpos2pos (motion_planning/utils.py:100-116) on float grid positions */
int dp_gridpos2pos(const dp_grid_geom *g, const float *gx, const float *gy, int64_t n, float *px, float *py,
                   void *stream);

/*
 * Environment grids (motion_planning/simulation_objects.py:29-44,77-94): grid, grid_gradientX, grid_gradientY,
 * each (G0, G1, Tg) fp32 with time innermost (the reference layout).  `on_device` says where the three inputs live.
 * They are repacked once into a time-major float4 table {p, gX, gY, 0}[t][ix][iy] so that one 16-byte gather serves
 * a collision-cost evaluation.
 */
int dp_env_create(const dp_grid_geom *g, int Tg, const float *grid, const float *gradx, const float *grady,
                  int on_device, int device, dp_env **out);
void dp_env_destroy(dp_env *e);

/*
 * Collision cost, forward and analytic backward: MotionPlanner.get_cost_coll (motion_planning/MotionPlanner.py:192-224)
 * and MotionPlannerGrad.get_cost_coll_initialize (motion_planning/MotionPlannerGrad.py:309-336).
 *   x, y   dev, element (n, t) at x[n*sx_n + t*sx_t] (strides in elements; covers both the reference (N,5,Ts) layout
 *          and this library's [t][d][n] layout)
 *   rho    dev (same addressing with sr_n, sr_t) or NULL (=> weights 1, the *_initialize variant)
 *   Ts     number of time slices evaluated (slice t uses env time index t)
 *   cost_slices   dev double[Ts] or NULL: per-slice sum (or per-slice max with get_max)
 *   cost_samples  dev double[N]  or NULL: per-sample sum over slices (get_cost_coll_initialize)
 *   gx, gy, grho  dev or NULL: d(sum cost)/dx, /dy, /drho, same addressing as x / rho (written, not accumulated)
 *   total  dev double[1] or NULL: sum over slices of cost_slices (needs cost_slices workspace inside; deterministic)
 */
int dp_cost_coll(const dp_env *e, const float *x, const float *y, int64_t sx_n, int64_t sx_t, const float *rho,
                 int64_t sr_n, int64_t sr_t, int64_t N, int Ts, int get_max, double *cost_slices, double *cost_samples,
                 float *gx, float *gy, float *grho, int64_t sg_n, int64_t sg_t, double *total, void *stream);

/*
 * Deterministic 2-D binning with integer accumulation (one time slice per call, or Ts slices with strides).
 * mode 0: occupancy-grid cells, pred2grid (motion_planning/utils.py:255-280): cell = clamp(pos2gridpos(p), 0, G)
 *         (clamp to G, not G-1, as the reference does); grid arrays are (gx+1) x (gy+1) per slice.
 * mode 1: uniform bins like np.histogramdd (systems/utils.py:105-129): bins (bx, by) over [lo, hi], right edge
 *         inclusive, samples outside dropped; grid arrays are bx x by per slice.
 *   w      dev weights (same addressing as x with sw_n, sw_t) or NULL (weight 1)
 *   mass   dev int64 [Ts][cells]: sum of llrint(w * 2^mass_scale_log2)  (shared-memory window + atomics: exact, order-free)
 *   count  dev int32 [Ts][cells]
 * Either of mass / count may be NULL.  The caller zeroes the grids (the call accumulates into them).
 */
typedef struct dp_bin_spec {
    int mode;            /* 0 grid cells, 1 uniform bins */
    dp_grid_geom geom;   /* mode 0 */
    double lo_x, hi_x, lo_y, hi_y; /* mode 1 */
    int bx, by;          /* mode 1 */
    int mass_scale_log2; /* fixed-point scale of the mass accumulator */
} dp_bin_spec;

int dp_bin2d(const dp_bin_spec *spec, const float *x, const float *y, int64_t sx_n, int64_t sx_t, const float *w,
             int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count, void *stream);

/*
 * dp_bin2d (mode 0) plus, in the same pass over the samples, the obstacle-grid dot product of the binned density:
 *   occ_dot dev double[Ts]:  sum_s w_s * p[t][cell_s]  =  sum_cells mass[t][cell] * p[t][cell] / 2^mass_scale_log2,
 * i.e. the quantity check_collision (motion_planning/utils.py:231-252) thresholds, for the sum-binned density.  Samples on
 * the clamp cell G (outside the environment grid) contribute nothing.  mass / count may be NULL.  Per-slice sums are
 * reduced in a fixed order (deterministic).
 */
int dp_bin2d_occ(const dp_bin_spec *spec, const dp_env *env, const float *x, const float *y, int64_t sx_n, int64_t sx_t,
                 const float *w, int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count,
                 double *occ_dot, void *stream);

/*
 * Density normaliser of EgoVehicle.predict_density (motion_planning/simulation_objects.py:295-299), per time slot:
 *     rho[t][n] = exp(l[t][n] - max_n l[t][.]) * rho0[n] / sum_n(...)
 * rholog dev [Ts][ldn] -> rho dev [Ts][ldn] (may alias).  Optional lmax / lsum dev float[Ts] receive the per-slot
 * max and (unnormalised) sum so a multi-GPU caller can merge them with an allreduce and call dp_normalize_apply.
 */
int dp_normalize_stats(const float *rholog, const float *rho0, int64_t N, int Ts, int64_t ldn, float *lmax,
                       double *lsum_at_max, void *stream);
int dp_normalize_apply(const float *rholog, const float *rho0, int64_t N, int Ts, int64_t ldn, const float *lmax,
                       const double *lsum, float *rho, void *stream);

/*
 * Reference-trajectory integrator: ControlAffineSystem.compute_xref_traj
 * (systems/ControlAffineSystem.py:341-354, get_next_xref :130-134) for B trajectories, and the gradient of a scalar loss
 * w.r.t. the reference inputs and the start state (what autograd computes through up2ref_traj :370-379).
 *   xref0 dev [B][5] (xref0_batched != 0) or [5] shared by all trajectories;  uref dev [B][2][T];  xref dev [B][5][T+1]
 *   gxref dev [B][5][T+1] upstream gradient  ->  guref dev [B][2][T],  gxref0 dev [B][5] or NULL
 */
int dp_xref_traj(const float *xref0, int xref0_batched, const float *uref, float dt, int B, int T, float *xref,
                 void *stream);
int dp_xref_traj_bwd(const float *xref, const float *gxref, float dt, int B, int T, float *guref, float *gxref0,
                     void *stream);

/*
 * Goal and state-bounds terms of MotionPlanner.get_cost:
 *   get_cost_goal   motion_planning/MotionPlanner.py:121-147   sum_s rho[s,-1] * |x[s,:2,-1] - goal|^2      (or max)
 *   get_cost_bounds motion_planning/MotionPlanner.py:149-190   sum over (s, d, t) with x < lo_d of rho (x - lo_d)^2, the
 *                   same above hi_d (or the max of each part)
 *   x dev: element (n, d, t) at x[n*sx_n + d*sx_d + t*sx_t], d = 0..4;  rho dev (n, t) at rho[n*sr_n + t*sr_t]
 *   lo5, hi5, goal_xy: HOST float[5], [5], [2]
 *   out3 dev double[3] = {goal, below-bound part, above-bound part};  flags2 dev int[2] = {any x < lo, any x > hi} over
 *   the first four state dimensions (the reference enters a part only then, :173,:181; in_bounds = neither flag)
 * dp_cost_state_bwd writes (sum mode) gx(n,d,t) and grho(n,t) = s_goal * d goal + s_lo * d below + s_hi * d above.
 */
int dp_cost_state(const float *x, int64_t sx_n, int64_t sx_d, int64_t sx_t, const float *rho, int64_t sr_n, int64_t sr_t,
                  int64_t N, int Ts, const float *lo5, const float *hi5, const float *goal_xy, int get_max, double *out3,
                  int *flags2, void *stream);
int dp_cost_state_bwd(const float *x, int64_t sx_n, int64_t sx_d, int64_t sx_t, const float *rho, int64_t sr_n, int64_t sr_t,
                      int64_t N, int Ts, const float *lo5, const float *hi5, const float *goal_xy, float s_goal, float s_lo,
                      float s_hi, float *gx, int64_t sg_n, int64_t sg_d, int64_t sg_t, float *grho, int64_t sgr_n,
                      int64_t sgr_t, void *stream);

/*
 * Device-resident pipeline (north-star data flow): sampled initial states -> closed-loop rollout with Liouville log-density
 * -> density weights -> time-sliced occupancy grids + collision cost, with the trajectories never leaving the GPU.
 * Replaces the chain  EgoVehicle.predict_density (LE branch + normaliser, motion_planning/simulation_objects.py:288-299)
 * -> pred2grid per slice (motion_planning/utils.py:255-280) / MotionPlanner.get_cost_coll (motion_planning/MotionPlanner.py:
 * 192-224) / check_collision (motion_planning/utils.py:231-252).
 *
 * The weight of sample n in slice t is the normaliser's NUMERATOR  w = exp(l[t][n] - max_n l[t][.]) * rho0[n]
 * (simulation_objects.py:296-297); the denominator is the sum of the mass grid, applied in dp_pipeline_finalize.  All
 * accumulators are integers, so the result buffer is bit-identical for any execution order and any sharding of the
 * samples over GPUs: a multi-GPU caller needs one MAX allreduce of lmax[Ts] between phases 1 and 2 and one SUM allreduce
 * of the int64 buffer between phases 2 and 3.
 *
 * Integer result buffer (int64 words, dp_pipeline_words), cells = (gx+1)*(gy+1), gc = Ts*cells:
 *     [0, gc)                        int64 mass[Ts][cells]   = sum of llrint(w * 2^spec.mass_scale_log2)
 *     [gc, gc + (gc+1)/2)            int32 count[Ts][cells]  (two counts per word; counts are non-negative and their global
 *                                                            sum is < 2^31, so a 64-bit SUM never carries between them)
 *     [gc + (gc+1)/2, ... + Ts)      int64 cost[Ts]          = sum of llrint(w * p(cell) * |des(cell) - pos|^2 * 2^cost_scale_log2)
 */
typedef struct dp_pipeline dp_pipeline;

/* capacity = most samples one call may carry (device memory: (24 + 12*Ts) bytes per sample); T, stride as in dp_rollout;
 * spec: mode 0, its mass_scale_log2 must come from DECLARED global bounds (total samples over all GPUs, largest rho0) so
 * that every shard uses the same fixed point; the controller and the environment must outlive the pipeline. */
int dp_pipeline_create(const dp_controller *c, const dp_env *env, const dp_bin_spec *spec, int T, int stride, int64_t capacity,
                       int cost_scale_log2, dp_pipeline **out);
void dp_pipeline_destroy(dp_pipeline *p);
int dp_pipeline_words(const dp_pipeline *p, int64_t *words);

/* Phase 1.  xe0 [N][5], rho0 [N] or NULL (= 1), xref [5][T+1], uref [2][T]: host pointers if inputs_on_host (copied in on
 * `stream`), else device pointers.  lmax_dev: dev float[Ts] receives the per-slice max of the log-density over THIS shard
 * (-inf for an empty shard).  Asynchronous on `stream`. */
int dp_pipeline_rollout(dp_pipeline *p, const float *xe0, const float *rho0, const float *xref, const float *uref,
                        int inputs_on_host, float dt, int64_t N, float *lmax_dev, void *stream);
/* Phase 2.  lmax_dev: the GLOBAL per-slice max.  ibuf_dev: dev int64[dp_pipeline_words], zeroed and filled by the call. */
int dp_pipeline_bin(dp_pipeline *p, const float *lmax_dev, int64_t *ibuf_dev, void *stream);
/* Phase 3.  From the (summed) integer buffer: out3_dev dev double[3][Ts] = per slice {sum of weights,
 * collision cost of the normalised density (get_cost_coll's slice term), occupancy dot product of the normalised density
 * sum_s rho_s * p[cell_s] (what check_collision thresholds)}. */
int dp_pipeline_finalize(const dp_pipeline *p, const int64_t *ibuf_dev, double *out3_dev, void *stream);
/* Copies the stored slices of the last dp_pipeline_rollout (N samples) into caller buffers (either may be NULL):
 * xy_dev dev float[Ts][2][N] absolute positions, l_dev dev float[Ts][N] log-density. */
int dp_pipeline_copy_slices(const dp_pipeline *p, float *xy_dev, float *l_dev, void *stream);

/* All three phases for ONE GPU with HOST buffers (copies inside the call, synchronises before returning):
 * ibuf_host int64[dp_pipeline_words] or NULL, out3_host double[3][Ts]. */
int dp_rollout_fused(dp_pipeline *p, const float *xe0, const float *rho0, const float *xref, const float *uref, float dt,
                     int64_t N, int64_t *ibuf_host, double *out3_host);

/*
 * The density-predictor MLP of the planner's density phase: density_training/utils.py:8-40
 * (NeuralNetwork, nn_type MLP, ReLU) as evaluated by get_nn_prediction (:161-180), which EgoVehicle.predict_density calls
 * once per prediction time point (motion_planning/simulation_objects.py:280-287).  All rows (time points x samples) of one
 * call run through the whole layer stack in ONE fused tcgen05 kernel (csrc/dnn.cu); the backward call returns the gradient
 * w.r.t. the input rows (the planner differentiates w.r.t. the input parameters, the weights are frozen).
 *   weights[l] host fp32 [dims[l+1]][dims[l]] (torch Linear.weight layout), biases[l] host fp32 [dims[l+1]],
 *   l = 0 .. n_layers-1; activation 0 = ReLU, 1 = tanh (args.activation, density_training/utils.py:14-19) after every layer
 *   but the last; n_layers <= 8, every width <= 160.
 *   x dev [rows][dims[0]] -> y dev [rows][dims[n_layers]]
 *   masks dev uint32 [rows][dp_dnn_mask_words()] or NULL: what the backward pass needs of the forward pass (ReLU sign bits, or
 *   the tanh activations as fp16 pairs), written by forward, read by backward
 *   gy dev [rows][dims[n_layers]] -> gx dev [rows][dims[0]]
 */
typedef struct dp_dnn dp_dnn;
int dp_dnn_create(const float *const *weights, const float *const *biases, const int *dims, int n_layers, int activation,
                  int device, dp_dnn **out);
void dp_dnn_destroy(dp_dnn *h);
int dp_dnn_mask_words(const dp_dnn *h, int *words_per_row);
int dp_dnn_forward(const dp_dnn *h, const float *x, int64_t rows, float *y, uint32_t *masks, void *stream);
int dp_dnn_backward(const dp_dnn *h, const float *gy, const uint32_t *masks, int64_t rows, float *gx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DENSITY_PLANNER_B200_H */
