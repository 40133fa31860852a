// K3: deterministic grid binning with integer fixed-point accumulation, and the density normaliser.
//   pred2grid            motion_planning/utils.py:255-280   (mode 0: occupancy-grid cells, clamp to [0, G])
//   get_density_map      systems/utils.py:105-129           (mode 1: np.histogramdd bins)
//   normaliser           motion_planning/simulation_objects.py:295-299
// Samples of a warp that fall into the same cell are aggregated with match/redux before ONE 64-bit and ONE 32-bit
// atomic per distinct cell reach L2; integer sums make the result independent of execution order and of how the
// samples are sharded over GPUs (the per-GPU grids are summed with an integer allreduce).
#include "dp_common.cuh"

namespace {

struct BinArgs {
    int mode;
    float x_lo, x_rng, y_lo, y_rng, gxm1, gym1;  // mode 0
    int cx, cy;                                  // cells per axis of the output arrays
    double lo_x, hi_x, lo_y, hi_y;               // mode 1
    double scale;
    const float *x, *y, *w;
    long long sx_n, sx_t, sw_n, sw_t, N;
    int Ts;
    long long *mass;
    int *count;
};

// np.histogramdd bin index for `v` over `bins` uniform bins on [lo, hi] (right edge inclusive); -1 if outside.
// Edges are lo + k*step like np.linspace; the comparison is done in fp64 against those edges as numpy's searchsorted does.
__device__ __forceinline__ int hist_bin(double v, double lo, double hi, int bins) {
    if (!(v >= lo) || !(v <= hi)) return -1;
    if (v == hi) return bins - 1;
    const double step = (hi - lo) / bins;
    int k = (int)floor((v - lo) / step);
    k = min(max(k, 0), bins - 1);
    // fix-up against the actual edges
    // edge(k) = k*step + lo with both roundings, as np.linspace evaluates it (no FMA contraction)
    while (k > 0 && v < __dadd_rn(__dmul_rn((double)k, step), lo)) --k;
    while (k < bins - 1 && v >= __dadd_rn(__dmul_rn((double)(k + 1), step), lo)) ++k;
    return k;
}

__global__ void __launch_bounds__(256) bin2d_kernel(BinArgs a) {
    const int t = blockIdx.y;
    const float *xp = a.x + (long long)t * a.sx_t;
    const float *yp = a.y + (long long)t * a.sx_t;
    const float *wp = a.w ? a.w + (long long)t * a.sw_t : nullptr;
    const long long cells = (long long)a.cx * a.cy;
    long long *mass = a.mass ? a.mass + (long long)t * cells : nullptr;
    int *count = a.count ? a.count + (long long)t * cells : nullptr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nmax = ((a.N + 31) / 32) * 32;  // keep warps converged for the match/redux collectives
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < nmax; n += stride) {
        int cell = -1;
        long long q = 0;
        if (n < a.N) {
            const float x = __ldg(xp + n * a.sx_n), y = __ldg(yp + n * a.sx_n);
            int ix, iy;
            if (a.mode == 0) {
                ix = min(max(dp_cell(x, a.x_lo, a.x_rng, a.gxm1), 0), a.cx - 1);  // clamp to [0, G] (:261-262), cx = G+1
                iy = min(max(dp_cell(y, a.y_lo, a.y_rng, a.gym1), 0), a.cy - 1);
            } else {
                ix = hist_bin((double)x, a.lo_x, a.hi_x, a.cx);
                iy = hist_bin((double)y, a.lo_y, a.hi_y, a.cy);
            }
            if (ix >= 0 && iy >= 0) {
                cell = ix * a.cy + iy;
                const double w = wp ? (double)__ldg(wp + n * a.sw_n) : 1.0;
                q = __double2ll_rn(w * a.scale);
            }
        }
        // warp aggregation: lanes with the same cell elect a leader that issues the atomics for the group
        const unsigned grp = __match_any_sync(0xffffffffu, cell);
        const int leader = __ffs(grp) - 1;
        const unsigned lo = (unsigned)(q & 0xFFFFFF), mid = (unsigned)((q >> 24) & 0xFFFFFF);
        const int hi = (int)(q >> 48);
        const unsigned slo = __reduce_add_sync(grp, lo), smid = __reduce_add_sync(grp, mid);
        const int shi = __reduce_add_sync(grp, hi);
        if (cell >= 0 && (int)(threadIdx.x & 31) == leader) {
            const long long sum = (long long)slo + ((long long)smid << 24) + ((long long)shi << 48);
            if (mass) atomicAdd(reinterpret_cast<unsigned long long *>(mass + cell), (unsigned long long)sum);
            if (count) atomicAdd(count + cell, __popc(grp));
        }
    }
}

// ---- normaliser ----------------------------------------------------------------------------------------------------
// one block per time slot: max, then sum of exp(l - max) * rho0 in fp64 with a fixed reduction order
__global__ void __launch_bounds__(1024) normalize_stats_kernel(const float *__restrict__ rholog,
                                                               const float *__restrict__ rho0, long long N, long long ldn,
                                                               float *lmax, double *lsum) {
    __shared__ float shf[32];
    __shared__ double shd[32];
    __shared__ float s_max;
    const int t = blockIdx.x;
    const float *l = rholog + (long long)t * ldn;
    float m = -INFINITY;
    for (long long n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, __ldg(l + n));
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = -INFINITY;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mm = fmaxf(mm, shf[i]);
        s_max = mm;
    }
    __syncthreads();
    const float mx = s_max;
    double s = 0.0;
    for (long long n = threadIdx.x; n < N; n += blockDim.x) {
        const float r0 = rho0 ? __ldg(rho0 + n) : 1.0f;
        s += (double)__fmul_rn(expf(__fsub_rn(__ldg(l + n), mx)), r0);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += shd[i];
        lmax[t] = mx;
        lsum[t] = tot;
    }
}

__global__ void normalize_apply_kernel(const float *__restrict__ rholog, const float *__restrict__ rho0, long long N,
                                       long long ldn, const float *__restrict__ lmax, const double *__restrict__ lsum,
                                       float *rho) {
    const int t = blockIdx.y;
    const float mx = lmax[t];
    const float sm = (float)lsum[t];
    const float *l = rholog + (long long)t * ldn;
    float *o = rho + (long long)t * ldn;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
        const float r0 = rho0 ? __ldg(rho0 + n) : 1.0f;
        o[n] = __fdiv_rn(__fmul_rn(expf(__fsub_rn(l[n], mx)), r0), sm);
    }
}

}  // namespace

int dp_bin2d(const dp_bin_spec *spec, const float *x, const float *y, int64_t sx_n, int64_t sx_t, const float *w,
             int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count, void *stream) {
    DP_REQUIRE(spec && x && y && (mass || count), "dp_bin2d: null argument");
    DP_REQUIRE(spec->mode == 0 || spec->mode == 1, "dp_bin2d: mode must be 0 or 1");
    DP_REQUIRE(spec->mass_scale_log2 >= 0 && spec->mass_scale_log2 <= 62, "dp_bin2d: bad mass_scale_log2");
    if (N <= 0 || Ts <= 0) return 0;
    BinArgs a;
    a.mode = spec->mode;
    if (a.mode == 0) {
        a.x_lo = spec->geom.x_min; a.x_rng = spec->geom.x_max - spec->geom.x_min;
        a.y_lo = spec->geom.y_min; a.y_rng = spec->geom.y_max - spec->geom.y_min;
        a.gxm1 = (float)(spec->geom.gx - 1); a.gym1 = (float)(spec->geom.gy - 1);
        a.cx = spec->geom.gx + 1; a.cy = spec->geom.gy + 1;
        a.lo_x = a.hi_x = a.lo_y = a.hi_y = 0;
    } else {
        DP_REQUIRE(spec->bx > 0 && spec->by > 0 && spec->hi_x > spec->lo_x && spec->hi_y > spec->lo_y, "dp_bin2d: bad bins");
        a.cx = spec->bx; a.cy = spec->by;
        a.lo_x = spec->lo_x; a.hi_x = spec->hi_x; a.lo_y = spec->lo_y; a.hi_y = spec->hi_y;
        a.x_lo = a.x_rng = a.y_lo = a.y_rng = a.gxm1 = a.gym1 = 0;
    }
    a.scale = ldexp(1.0, spec->mass_scale_log2);
    a.x = x; a.y = y; a.w = w;
    a.sx_n = sx_n; a.sx_t = sx_t; a.sw_n = sw_n; a.sw_t = sw_t; a.N = N; a.Ts = Ts;
    a.mass = mass; a.count = count;
    long long B = (N + 255) / 256;
    const long long target = (148LL * 8 + Ts - 1) / Ts;
    if (B > target) B = target;
    if (B < 1) B = 1;
    bin2d_kernel<<<dim3((unsigned)B, (unsigned)Ts), 256, 0, (cudaStream_t)stream>>>(a);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_normalize_stats(const float *rholog, const float *rho0, int64_t N, int Ts, int64_t ldn, float *lmax,
                       double *lsum_at_max, void *stream) {
    DP_REQUIRE(rholog && lmax && lsum_at_max, "dp_normalize_stats: null argument");
    if (N <= 0 || Ts <= 0) return 0;
    normalize_stats_kernel<<<Ts, 1024, 0, (cudaStream_t)stream>>>(rholog, rho0, N, ldn, lmax, lsum_at_max);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_normalize_apply(const float *rholog, const float *rho0, int64_t N, int Ts, int64_t ldn, const float *lmax,
                       const double *lsum, float *rho, void *stream) {
    DP_REQUIRE(rholog && lmax && lsum && rho, "dp_normalize_apply: null argument");
    if (N <= 0 || Ts <= 0) return 0;
    long long B = (N + 255) / 256;
    const long long target = (148LL * 8 + Ts - 1) / Ts;
    if (B > target) B = target;
    if (B < 1) B = 1;
    normalize_apply_kernel<<<dim3((unsigned)B, (unsigned)Ts), 256, 0, (cudaStream_t)stream>>>(rholog, rho0, N, ldn, lmax,
                                                                                             lsum, rho);
    DP_CUDA(cudaGetLastError());
    return 0;
}
