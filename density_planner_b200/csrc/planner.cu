// Planner-side terms next to the hot path:
//   * reference-trajectory integrator, forward and adjoint: ControlAffineSystem.compute_xref_traj / up2ref_traj
//     (systems/ControlAffineSystem.py:341-379, get_next_xref :130-134, CAR dynamics systems/sytem_CAR.py:128-176).  The
//     reference builds a 1000-node autograd graph per optimisation step; here the forward is one launch and the
//     gradient w.r.t. the inputs one more (discrete adjoint of the same Euler scheme).
//   * goal and state-bounds cost terms of MotionPlanner.get_cost (motion_planning/MotionPlanner.py:121-190), forward sums
//     (fp64 block partials, fixed-order finalize: deterministic) and analytic gradients.
#include "dp_common.cuh"

namespace {

// ---- integrator ---------------------------------------------------------------------------------------------------
// one thread per trajectory; xref [B][5][T+1], uref [B][2][T] (the reference's layouts, time innermost)
__global__ void xref_fwd_kernel(const float *__restrict__ xref0, int x0_stride, const float *__restrict__ uref, float dt, int B, int T,
                                float *__restrict__ xref) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *x0 = xref0 + (size_t)b * x0_stride;
    float x[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) x[d] = x0[d];
    float *o = xref + (size_t)b * 5 * (T + 1);
    const float *u = uref + (size_t)b * 2 * T;
#pragma unroll
    for (int d = 0; d < 5; ++d) o[(size_t)d * (T + 1)] = x[d];
    for (int k = 0; k < T; ++k) {
        // get_next_xref: x + (a(x) + b(x) u) dt, each op rounded as torch evaluates it
        float sn, cs;
        sincosf(x[2], &sn, &cs);
        const float u0 = u[k], u1 = u[T + k];
        x[0] = __fadd_rn(x[0], __fmul_rn(__fmul_rn(x[3], cs), dt));
        x[1] = __fadd_rn(x[1], __fmul_rn(__fmul_rn(x[3], sn), dt));
        x[2] = __fadd_rn(x[2], __fmul_rn(u0, dt));
        x[3] = __fadd_rn(x[3], __fmul_rn(u1, dt));
#pragma unroll
        for (int d = 0; d < 5; ++d) o[(size_t)d * (T + 1) + k + 1] = x[d];
    }
}

// discrete adjoint: lambda_k = g_k + J_k^T lambda_{k+1},  d loss / d u_k = dt * lambda_{k+1}[2:4]
__global__ void xref_bwd_kernel(const float *__restrict__ xref, const float *__restrict__ gxref, float dt, int B, int T,
                                float *__restrict__ guref, float *__restrict__ gxref0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *xr = xref + (size_t)b * 5 * (T + 1);
    const float *g = gxref + (size_t)b * 5 * (T + 1);
    float *gu = guref + (size_t)b * 2 * T;
    float l[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = T; k >= 1; --k) {
#pragma unroll
        for (int d = 0; d < 5; ++d) l[d] += g[(size_t)d * (T + 1) + k];
        gu[k - 1] = l[2] * dt;
        gu[T + k - 1] = l[3] * dt;
        const float th = xr[(size_t)2 * (T + 1) + k - 1], v = xr[(size_t)3 * (T + 1) + k - 1];
        float sn, cs;
        sincosf(th, &sn, &cs);
        const float lth = l[2] + dt * v * (cs * l[1] - sn * l[0]);
        const float lv = l[3] + dt * (cs * l[0] + sn * l[1]);
        l[2] = lth;
        l[3] = lv;
    }
    if (gxref0)
#pragma unroll
        for (int d = 0; d < 5; ++d) gxref0[(size_t)b * 5 + d] = l[d] + g[(size_t)d * (T + 1)];
}

// ---- goal / bounds costs ------------------------------------------------------------------------------------------------
struct StateArgs {
    const float *x, *rho;
    long long sx_n, sx_d, sx_t, sr_n, sr_t, N;
    int Ts, get_max;
    float lo[5], hi[5], goal[2];
    double *partial;  // [B][4]: goal, below-bound part, above-bound part, (unused)
    int *flags;       // [2]: any violation of the lower / upper bounds in the first four state dimensions
    float *gx, *grho;
    long long sg_n, sg_d, sg_t, sgr_n, sgr_t;
    float s_goal, s_lo, s_hi;
};

constexpr int SB = 256;

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < SB / 32; ++w) tot += sh[w];
    return tot;
}
__device__ __forceinline__ double block_max(double v, double *sh) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < SB / 32; ++w) tot = fmax(tot, sh[w]);
    return tot;
}

// forward: one thread per (sample, time) pair; terms are non-negative, so 0 is the neutral element of both sum and max
__global__ void __launch_bounds__(SB) state_cost_fwd_kernel(StateArgs a) {
    __shared__ double sh[SB / 32];
    double goal = 0.0, below = 0.0, above = 0.0;
    int f_lo = 0, f_hi = 0;
    const long long total = a.N * a.Ts;
    for (long long i = (long long)blockIdx.x * SB + threadIdx.x; i < total; i += (long long)gridDim.x * SB) {
        const long long n = i / a.Ts;
        const int t = (int)(i - n * a.Ts);
        const float r = a.rho ? a.rho[n * a.sr_n + t * a.sr_t] : 1.0f;
        const float *xp = a.x + n * a.sx_n + t * a.sx_t;
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const float v = xp[d * a.sx_d];
            if (v < a.lo[d]) {
                const float e = __fsub_rn(v, a.lo[d]);
                const double term = (double)__fmul_rn(r, __fmul_rn(e, e));
                below = a.get_max ? fmax(below, term) : below + term;
                if (d < 4) f_lo = 1;
            }
            if (v > a.hi[d]) {
                const float e = __fsub_rn(v, a.hi[d]);
                const double term = (double)__fmul_rn(r, __fmul_rn(e, e));
                above = a.get_max ? fmax(above, term) : above + term;
                if (d < 4) f_hi = 1;
            }
        }
        if (t == a.Ts - 1) {
            const float e0 = __fsub_rn(xp[0], a.goal[0]), e1 = __fsub_rn(xp[a.sx_d], a.goal[1]);
            const float sq = __fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1));
            const double term = (double)__fmul_rn(r, sq);
            goal = a.get_max ? fmax(goal, term) : goal + term;
        }
    }
    if (f_lo) atomicOr(a.flags, 1);
    if (f_hi) atomicOr(a.flags + 1, 1);
    const double g = a.get_max ? block_max(goal, sh) : block_sum(goal, sh);
    const double lo = a.get_max ? block_max(below, sh) : block_sum(below, sh);
    const double hi = a.get_max ? block_max(above, sh) : block_sum(above, sh);
    if (threadIdx.x == 0) {
        a.partial[(size_t)blockIdx.x * 4 + 0] = g;
        a.partial[(size_t)blockIdx.x * 4 + 1] = lo;
        a.partial[(size_t)blockIdx.x * 4 + 2] = hi;
    }
}

__global__ void state_cost_finalize_kernel(const double *partial, int B, int get_max, double *out) {
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s = get_max ? fmax(s, partial[(size_t)b * 4 + threadIdx.x]) : s + partial[(size_t)b * 4 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// backward of the summed terms: d/dx and d/drho scaled by the upstream factors s_goal, s_lo, s_hi
__global__ void __launch_bounds__(SB) state_cost_bwd_kernel(StateArgs a) {
    const long long total = a.N * a.Ts;
    for (long long i = (long long)blockIdx.x * SB + threadIdx.x; i < total; i += (long long)gridDim.x * SB) {
        const long long n = i / a.Ts;
        const int t = (int)(i - n * a.Ts);
        const float r = a.rho ? a.rho[n * a.sr_n + t * a.sr_t] : 1.0f;
        const float *xp = a.x + n * a.sx_n + t * a.sx_t;
        float *gp = a.gx + n * a.sg_n + t * a.sg_t;
        float gr = 0.0f;
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            const float v = xp[d * a.sx_d];
            float gxd = 0.0f;
            if (v < a.lo[d]) {
                const float e = v - a.lo[d];
                gxd += a.s_lo * 2.0f * r * e;
                gr += a.s_lo * e * e;
            }
            if (v > a.hi[d]) {
                const float e = v - a.hi[d];
                gxd += a.s_hi * 2.0f * r * e;
                gr += a.s_hi * e * e;
            }
            if (d < 2 && t == a.Ts - 1) {
                const float e = v - a.goal[d];
                gxd += a.s_goal * 2.0f * r * e;
                gr += a.s_goal * e * e;
            }
            gp[d * a.sg_d] = gxd;
        }
        if (a.grho) a.grho[n * a.sgr_n + t * a.sgr_t] = gr;
    }
}

}  // namespace

extern "C" {

int dp_xref_traj(const float *xref0, int xref0_batched, const float *uref, float dt, int B, int T, float *xref, void *stream) {
    DP_REQUIRE(xref0 && uref && xref && B >= 0 && T >= 0, "dp_xref_traj: bad argument");
    if (B == 0) return 0;
    xref_fwd_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(xref0, xref0_batched ? 5 : 0, uref, dt, B, T, xref);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_xref_traj_bwd(const float *xref, const float *gxref, float dt, int B, int T, float *guref, float *gxref0, void *stream) {
    DP_REQUIRE(xref && gxref && guref && B >= 0 && T >= 0, "dp_xref_traj_bwd: bad argument");
    if (B == 0) return 0;
    xref_bwd_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(xref, gxref, dt, B, T, guref, gxref0);
    DP_CUDA(cudaGetLastError());
    return 0;
}

static int fill_state_args(StateArgs &a, const float *x, int64_t sx_n, int64_t sx_d, int64_t sx_t, const float *rho, int64_t sr_n,
                           int64_t sr_t, int64_t N, int Ts, const float *lo5, const float *hi5, const float *goal_xy) {
    DP_REQUIRE(x && lo5 && hi5 && goal_xy && N >= 0 && Ts >= 1, "dp_cost_state: bad argument");
    memset(&a, 0, sizeof(a));
    a.x = x; a.rho = rho;
    a.sx_n = sx_n; a.sx_d = sx_d; a.sx_t = sx_t; a.sr_n = sr_n; a.sr_t = sr_t;
    a.N = N; a.Ts = Ts;
    for (int d = 0; d < 5; ++d) { a.lo[d] = lo5[d]; a.hi[d] = hi5[d]; }
    a.goal[0] = goal_xy[0]; a.goal[1] = goal_xy[1];
    return 0;
}

int dp_cost_state(const float *x, int64_t sx_n, int64_t sx_d, int64_t sx_t, const float *rho, int64_t sr_n, int64_t sr_t,
                  int64_t N, int Ts, const float *lo5, const float *hi5, const float *goal_xy, int get_max, double *out3,
                  int *flags2, void *stream_) {
    StateArgs a;
    if (int rc = fill_state_args(a, x, sx_n, sx_d, sx_t, rho, sr_n, sr_t, N, Ts, lo5, hi5, goal_xy)) return rc;
    DP_REQUIRE(out3 && flags2, "dp_cost_state: null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    DP_CUDA(cudaMemsetAsync(flags2, 0, 2 * sizeof(int), stream));
    if (N == 0) {
        DP_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(double), stream));
        return 0;
    }
    long long Bk = (N * Ts + SB - 1) / SB;
    if (Bk > 148 * 8) Bk = 148 * 8;
    double *partial = nullptr;
    DP_CUDA(cudaMallocAsync(&partial, sizeof(double) * 4 * Bk, stream));
    a.partial = partial;
    a.flags = flags2;
    a.get_max = get_max;
    state_cost_fwd_kernel<<<(unsigned)Bk, SB, 0, stream>>>(a);
    state_cost_finalize_kernel<<<1, 32, 0, stream>>>(partial, (int)Bk, get_max, out3);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaFreeAsync(partial, stream));
    return 0;
}

int dp_cost_state_bwd(const float *x, int64_t sx_n, int64_t sx_d, int64_t sx_t, const float *rho, int64_t sr_n, int64_t sr_t,
                      int64_t N, int Ts, const float *lo5, const float *hi5, const float *goal_xy, float s_goal, float s_lo,
                      float s_hi, float *gx, int64_t sg_n, int64_t sg_d, int64_t sg_t, float *grho, int64_t sgr_n, int64_t sgr_t,
                      void *stream) {
    StateArgs a;
    if (int rc = fill_state_args(a, x, sx_n, sx_d, sx_t, rho, sr_n, sr_t, N, Ts, lo5, hi5, goal_xy)) return rc;
    DP_REQUIRE(gx, "dp_cost_state_bwd: null gradient output");
    if (N == 0) return 0;
    a.gx = gx; a.grho = grho;
    a.sg_n = sg_n; a.sg_d = sg_d; a.sg_t = sg_t; a.sgr_n = sgr_n; a.sgr_t = sgr_t;
    a.s_goal = s_goal; a.s_lo = s_lo; a.s_hi = s_hi;
    long long Bk = (N * Ts + SB - 1) / SB;
    if (Bk > 148 * 8) Bk = 148 * 8;
    state_cost_bwd_kernel<<<(unsigned)Bk, SB, 0, (cudaStream_t)stream>>>(a);
    DP_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
