// K3: deterministic grid binning with integer fixed-point accumulation, and the density normaliser.
//   pred2grid            motion_planning/utils.py:255-280   (mode 0: occupancy-grid cells, clamp to [0, G])
//   get_density_map      systems/utils.py:105-129           (mode 1: np.histogramdd bins)
//   check_collision      motion_planning/utils.py:231-252   (occupancy dot product, fused: see occ_dot)
//   normaliser           motion_planning/simulation_objects.py:295-299
// The samples of a slice cluster around the reference trajectory, so every block keeps a WIN x WIN window of cells (int64
// fixed-point mass + int32 count) in shared memory, centred on a probe of its own samples; samples are accumulated with
// shared-memory atomics and the window is flushed to the global grids once at the end (one atomic per touched cell per
// block).  Samples outside the window fall back to warp-aggregated global atomics (match.any + redux).  Integer sums make
// the result independent of execution order and of how the samples are sharded over GPUs (the per-GPU grids are summed
// with an integer allreduce).  Optionally the same pass gathers the environment's occupancy of every sample's cell and
// returns, per slice, sum_s w_s * p[cell_s] = sum_cells mass[cell] * p[cell]: the obstacle-grid dot product of the binned
// density, without a second pass over the grid.
#include <stdlib.h>

#include "dp_common.cuh"

namespace {

constexpr int WIN = 80;                  // window edge in cells
constexpr int BIN_SMEM = WIN * WIN * 12 + 64;
constexpr int MAX_BIN_PARTIALS = 1 << 15;

struct BinArgs {
    int mode;
    DpGeom q;                                    // mode 0
    int cx, cy;                                  // cells per axis of the output arrays
    double lo_x, hi_x, lo_y, hi_y;               // mode 1
    double scale;
    float scale_f;  // the same power of two as a float
    const float *x, *y, *w;
    long long sx_n, sx_t, sw_n, sw_t, N;
    int Ts;
    long long *mass;
    int *count;
    const float4 *table;  // environment table [t][gx][gy] {p, gX, gY, 0} or null
    int egx, egy;         // its cells per axis
    double *partial;      // [Ts][B] partial occupancy dot products or null
    int B;
    int prefetch;         // lean kernel: L2 prefetch distance of the streamed rows in iterations (0 off; DP_BIN_PREFETCH)
    int agg;              // warp aggregation of same-cell samples before the shared atomics: 0 off, 1 on, 2 decided per block
    // fused pipeline pass (FUSED kernels, dp_pipeline_bin): the weight column holds the log-density l and the weight is
    // exp(l - lmax[t]) * rho0[n] (simulation_objects.py:296-297, before the division by the sum); the same pass
    // accumulates the collision cost of the weighted samples (MotionPlanner.py:206-222) in int64 fixed point
    const float *lmax;    // [Ts]
    const float *rho0;    // [N] or null (= 1)
    long long *cost_q;    // [Ts]
    float cost_scale_f;
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

template <int MODE>
__device__ __forceinline__ void sample_cell(const BinArgs &a, float x, float y, int &ix, int &iy) {
    if (MODE == 0) {
        ix = min(max(dp_cell<false>(x, a.q.x_lo, a.q.x_rng, a.q.r_x_rng, a.q.gxm1), 0), a.cx - 1);  // clamp to [0, G] (:261-262), cx = G+1
        iy = min(max(dp_cell<false>(y, a.q.y_lo, a.q.y_rng, a.q.r_y_rng, a.q.gym1), 0), a.cy - 1);
    } else {
        ix = hist_bin((double)x, a.lo_x, a.hi_x, a.cx);
        iy = hist_bin((double)y, a.lo_y, a.hi_y, a.cy);
    }
}

// sum of a 64-bit quantity over the lanes of `grp` (a match.any group): three 24-bit limbs through redux.sync
__device__ __forceinline__ long long group_sum_ll(unsigned grp, long long q) {
    const unsigned lo = (unsigned)(q & 0xFFFFFF), mid = (unsigned)((q >> 24) & 0xFFFFFF);
    const int hi = (int)(q >> 48);
    const unsigned rlo = __reduce_add_sync(grp, lo), rmid = __reduce_add_sync(grp, mid);
    const int rhi = __reduce_add_sync(grp, hi);
    return (long long)rlo + ((long long)rmid << 24) + ((long long)rhi << 48);
}

// The int64 mass of a window cell is kept as two 32-bit words: a native 32-bit shared-memory atomic on the low word returns
// the old value, which tells exactly whether this addition carried into the high word (a 64-bit shared atomicAdd would
// be a compare-and-swap loop that degrades under the contention of a tight sample cloud).  Where many samples of a warp
// fall into the same cell (tight clouds: the planner's few hundred samples, late slices of a contracting closed loop) the
// warp first sums them (match.any + redux) and only the group leaders touch shared memory; whether that pays is decided
// per block from a probe of 32 consecutive samples (agg == 2) -- for a diffuse cloud the match costs more than it saves.
template <int MODE, bool VEC, bool FUSED, int NT>
__global__ void __launch_bounds__(NT, 2) bin2d_kernel(BinArgs a) {
    extern __shared__ __align__(16) unsigned char bsm[];
    unsigned *slo = reinterpret_cast<unsigned *>(bsm);
    unsigned *shi = reinterpret_cast<unsigned *>(bsm + WIN * WIN * 4);
    int *scount = reinterpret_cast<int *>(bsm + WIN * WIN * 8);
    int *sorg = reinterpret_cast<int *>(bsm + WIN * WIN * 12);
    __shared__ double shd[NT / 32];
    __shared__ long long shq[NT / 32];
    const int t = blockIdx.y, tid = threadIdx.x;
    const float *xp = a.x + (long long)t * a.sx_t;
    const float *yp = a.y + (long long)t * a.sx_t;
    const float *wp = a.w ? a.w + (long long)t * a.sw_t : nullptr;
    const long long cells = (long long)a.cx * a.cy;
    long long *mass = a.mass ? a.mass + (long long)t * cells : nullptr;
    int *count = a.count ? a.count + (long long)t * cells : nullptr;
    const float4 *tab = a.table ? a.table + (size_t)t * a.egx * a.egy : nullptr;
    const float mx = FUSED ? __ldg(a.lmax + t) : 0.0f;
    // this block's samples: a contiguous chunk of the slice
    long long chunk = (a.N + gridDim.x - 1) / gridDim.x;
    if (VEC) chunk = (chunk + 3) & ~3LL;
    const long long n_lo = min((long long)blockIdx.x * chunk, a.N), n_hi = min(n_lo + chunk, a.N);
    // window origin: mean cell of 32 probe samples spread over the chunk (any choice gives the same integer grids); the
    // same warp looks at 32 consecutive samples to see how many share a cell
    if (tid < 32) {
        int ix = 0, iy = 0, ok = 0;
        if (n_lo < n_hi) {
            const long long n = n_lo + (n_hi - n_lo) * tid / 32;
            sample_cell<MODE>(a, __ldg(xp + n * a.sx_n), __ldg(yp + n * a.sx_n), ix, iy);
            ok = (ix >= 0 && iy >= 0) ? 1 : 0;
        }
        int sx = ok ? ix : 0, sy = ok ? iy : 0, sk = ok;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, m);
            sy += __shfl_xor_sync(0xffffffffu, sy, m);
            sk += __shfl_xor_sync(0xffffffffu, sk, m);
        }
        int use_agg = a.agg == 1;
        if (a.agg == 2) {
            int c2 = -1 - tid;  // lanes without a sample never match
            if (n_lo + tid < n_hi) {
                int jx, jy;
                sample_cell<MODE>(a, __ldg(xp + (n_lo + tid) * a.sx_n), __ldg(yp + (n_lo + tid) * a.sx_n), jx, jy);
                c2 = jx * a.cy + jy;
            }
            const unsigned grp = __match_any_sync(0xffffffffu, c2);
            const unsigned leaders = __ballot_sync(0xffffffffu, (int)(__ffs(grp) - 1) == tid);
            use_agg = __popc(leaders) <= 20;  // at least 3 of 8 samples ride on another lane's atomic
        }
        if (tid == 0) {
            const int mx_ = sk ? sx / sk : 0, my_ = sk ? sy / sk : 0;
            sorg[0] = min(max(mx_ - WIN / 2, 0), max(a.cx - WIN, 0));
            sorg[1] = min(max(my_ - WIN / 2, 0), max(a.cy - WIN, 0));
            sorg[2] = use_agg;
        }
    }
    for (int i = tid; i < WIN * WIN; i += NT) {
        slo[i] = 0u;
        shi[i] = 0u;
        scount[i] = 0;
    }
    __syncthreads();
    const int ox = sorg[0], oy = sorg[1];
    const bool agg = sorg[2] != 0;
    double dot = 0.0;
    long long cq = 0;
    // one sample: window cell -> shared atomics; outside the window -> warp-aggregated global atomics (all lanes take part
    // in the collectives, `live` tells whether this lane carries a sample)
    auto put = [&](bool live, float x, float y, float w, float r0) {
        int cell = -1, wcell = -1;
        long long q = 0;
        if (live) {
            int ix, iy;
            sample_cell<MODE>(a, x, y, ix, iy);
            if (ix >= 0 && iy >= 0) {
                if (FUSED) w = __fmul_rn(expf(__fsub_rn(w, mx)), r0);  // simulation_objects.py:296-297
                // w * 2^s is exact in fp32 (power-of-two scale), so this equals llrint((double)w * 2^s)
                q = wp ? __float2ll_rn(__fmul_rn(w, a.scale_f)) : (long long)a.scale;
                const int wx = ix - ox, wy = iy - oy;
                if (wx >= 0 && wx < WIN && wy >= 0 && wy < WIN)
                    wcell = wx * WIN + wy;
                else
                    cell = ix * a.cy + iy;
                if (tab) {
                    if (FUSED) {
                        // the cost clamps the cell to [0, G-1] (MotionPlanner.py:208-209), the binning to [0, G]
                        const int jx = min(ix, a.egx - 1), jy = min(iy, a.egy - 1);
                        const float4 e = __ldg(tab + (size_t)jx * a.egy + jy);
                        const DpEval ev = dp_cost_term(e, x, y, w, true);
                        cq += __float2ll_rn(__fmul_rn(ev.term, a.cost_scale_f));
                    } else if (ix < a.egx && iy < a.egy) {
                        dot += (double)q * (double)__ldg(&tab[(size_t)ix * a.egy + iy].x);
                    }
                }
            }
        }
        if (agg) {
            const unsigned grp = __match_any_sync(0xffffffffu, wcell);
            const long long qs = group_sum_ll(grp, q);
            if (wcell >= 0 && (int)(tid & 31) == (int)(__ffs(grp) - 1)) {
                if (mass) {
                    const unsigned ql = (unsigned)qs, qh = (unsigned)((unsigned long long)qs >> 32);
                    const unsigned old = atomicAdd(slo + wcell, ql);
                    const unsigned carry = (old + ql < old) ? 1u : 0u;
                    if (qh + carry) atomicAdd(shi + wcell, qh + carry);
                }
                if (count) atomicAdd(scount + wcell, __popc(grp));
            }
        } else if (wcell >= 0) {
            if (mass) {
                const unsigned ql = (unsigned)q, qh = (unsigned)((unsigned long long)q >> 32);
                const unsigned old = atomicAdd(slo + wcell, ql);
                const unsigned carry = (old + ql < old) ? 1u : 0u;
                if (qh + carry) atomicAdd(shi + wcell, qh + carry);
            }
            if (count) atomicAdd(scount + wcell, 1);
        }
        if (__any_sync(0xffffffffu, cell >= 0)) {
            const unsigned grp = __match_any_sync(0xffffffffu, cell);
            const long long sum = group_sum_ll(grp, q);
            if (cell >= 0 && (int)(tid & 31) == (int)(__ffs(grp) - 1)) {
                if (mass) atomicAdd(reinterpret_cast<unsigned long long *>(mass + cell), (unsigned long long)sum);
                if (count) atomicAdd(count + cell, __popc(grp));
            }
        }
    };
    const long long span = n_hi - n_lo;
    if (VEC) {
        // contiguous, 16-byte aligned rows (chunk boundaries are multiples of four): LDG.128 of x, y, w, four samples per
        // thread and iteration
        const long long g_lo = n_lo >> 2, g_hi = n_hi >> 2;  // full groups of four
        const long long gmax = g_lo + (((g_hi - g_lo) + 31) / 32) * 32;
        for (long long g = g_lo + tid; g < gmax; g += NT) {
            const bool live = g < g_hi;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv, wv = xv, rv = make_float4(1.f, 1.f, 1.f, 1.f);
            if (live) {
                xv = __ldg(reinterpret_cast<const float4 *>(xp) + g);
                yv = __ldg(reinterpret_cast<const float4 *>(yp) + g);
                if (wp) wv = __ldg(reinterpret_cast<const float4 *>(wp) + g);
                if (FUSED && a.rho0) rv = __ldg(reinterpret_cast<const float4 *>(a.rho0) + g);
            }
            put(live, xv.x, yv.x, wv.x, rv.x);
            put(live, xv.y, yv.y, wv.y, rv.y);
            put(live, xv.z, yv.z, wv.z, rv.z);
            put(live, xv.w, yv.w, wv.w, rv.w);
        }
        // the last block of the slice also takes the (at most three) samples behind the last full group
        const long long t_lo = g_hi << 2;
        const bool tail = (n_hi == a.N) && (n_lo < a.N) && tid < 32;
        if (tail && __any_sync(0xffffffffu, t_lo + tid < a.N)) {
            const long long n = t_lo + tid;
            const bool live = n < a.N;
            put(live, live ? __ldg(xp + n) : 0.f, live ? __ldg(yp + n) : 0.f, (live && wp) ? __ldg(wp + n) : 0.f,
                (FUSED && live && a.rho0) ? __ldg(a.rho0 + n) : 1.f);
        }
    } else {
        const long long nmax = ((span + 31) / 32) * 32;  // keep warps converged for the match/redux collectives
        for (long long k = tid; k < nmax; k += NT) {
            const long long n = n_lo + k;
            const bool live = k < span;
            put(live, live ? __ldg(xp + n * a.sx_n) : 0.f, live ? __ldg(yp + n * a.sx_n) : 0.f,
                (live && wp) ? __ldg(wp + n * a.sw_n) : 0.f, (FUSED && live && a.rho0) ? __ldg(a.rho0 + n) : 1.f);
        }
    }
    __syncthreads();
    // flush the window: one atomic per touched cell
    for (int i = tid; i < WIN * WIN; i += NT) {
        const int wx = i / WIN, wy = i - wx * WIN;
        const int ix = ox + wx, iy = oy + wy;
        if (ix < a.cx && iy < a.cy) {
            const long long g = (long long)ix * a.cy + iy;
            const unsigned long long m = ((unsigned long long)shi[i] << 32) | slo[i];
            const int c = scount[i];
            if (mass && m) atomicAdd(reinterpret_cast<unsigned long long *>(mass + g), m);
            if (count && c) atomicAdd(count + g, c);
        }
    }
    if (FUSED) {  // integer sum: exact and independent of the order (one atomic per block)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) cq += __shfl_xor_sync(0xffffffffu, cq, m);
        if ((tid & 31) == 0) shq[tid >> 5] = cq;
        __syncthreads();
        if (tid == 0) {
            long long tot = 0;
            for (int w = 0; w < NT / 32; ++w) tot += shq[w];
            if (tot) atomicAdd(reinterpret_cast<unsigned long long *>(a.cost_q + t), (unsigned long long)tot);
        }
    } else if (a.partial) {  // fixed-order block reduction of the occupancy dot product
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
        if ((tid & 31) == 0) shd[tid >> 5] = dot;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < NT / 32; ++w) tot += shd[w];
            a.partial[(size_t)t * a.B + blockIdx.x] = tot;
        }
    }
}

// ---- lean kernel (mode 0, contiguous 16-byte aligned rows, mass and count both wanted): the default path ---------------------
// ncu of the general kernel above at the config-3 size (profiles/r02_bin2d_before_ncu.txt) showed it ISSUE bound (issue slots
// 91 % active, 101 thread instructions per sample, shared-atomic wavefronts only 21 % of peak, DRAM 39 %), not atomic bound:
// the warp collectives that keep lanes converged for match.any / redux, the run-time knobs and the 64-bit index arithmetic were
// most of the instructions.  This kernel has no warp collective at all in the sample loop:
//   * cell index: the reference's fp32 op sequence up to q + 0.001, then clamp IN FLOAT to [0, G] and round with the magic-number
//     add (t + 1.5 * 2^23 has rn(t) in its low mantissa bits, ties to even like cvt.rni / torch.round).  Rounding is monotone
//     and fixes the integers 0 and G, so rn(clamp(t)) = clamp(rn(t)): same cells as dp_cell + integer clamp, NaN -> 0 as before;
//     two F2I (XU pipe) and the integer clamps become FMNMX + FADD, and the window offset absorbs the magic constant.
//   * a sample outside the block's window goes straight to global atomics from its own lane (no match.any): such samples are
//     rare for the clouds this path sees, and a diffuse cloud is bound by the global atomics either way.
//   * no aggregation before the shared atomics: measured 15x slower with match.any (gpurun r2: 5.96 ms vs 0.41 ms), and even a
//     cloud contracted to a few cells costs ~6 wavefronts per atomic, i.e. about the DRAM time of the pass.
// Results are integer sums, so they equal the general kernel's bit for bit (tests: knob A/B fingerprints, permutation / shard
// invariance).
__device__ __forceinline__ int lean_cell_biased(float p, float lo, float rng, float r_rng, float gm1, float cmax) {
    const float q = __fmul_rn(dp_div_fast(__fsub_rn(p, lo), rng, r_rng), gm1);
    const float t = fminf(fmaxf(__fadd_rn(q, 0.001f), 0.0f), cmax);
    return __float_as_int(__fadd_rn(t, 12582912.0f));  // 0x4B400000 + rn(t)
}
constexpr int LEAN_MAGIC = 0x4B400000;

template <bool FUSED, bool OCC, bool HASW, int NT>
__global__ void __launch_bounds__(NT, 2) bin2d_lean_kernel(BinArgs a) {
    constexpr int VW = FUSED ? 2 : 4;  // samples per thread and iteration (the fused pass carries more state per sample)
    extern __shared__ __align__(16) unsigned char bsm[];
    unsigned *slo = reinterpret_cast<unsigned *>(bsm);
    unsigned *shi = reinterpret_cast<unsigned *>(bsm + WIN * WIN * 4);
    int *scount = reinterpret_cast<int *>(bsm + WIN * WIN * 8);
    int *sorg = reinterpret_cast<int *>(bsm + WIN * WIN * 12);
    __shared__ double shd[NT / 32];
    __shared__ long long shq[NT / 32];
    const int t = blockIdx.y, tid = threadIdx.x;
    const float4 *xp = reinterpret_cast<const float4 *>(a.x + (long long)t * a.sx_t);
    const float4 *yp = reinterpret_cast<const float4 *>(a.y + (long long)t * a.sx_t);
    const float4 *wp = HASW ? reinterpret_cast<const float4 *>(a.w + (long long)t * a.sw_t) : nullptr;
    const float4 *rp = (FUSED && a.rho0) ? reinterpret_cast<const float4 *>(a.rho0) : nullptr;
    const int cells = a.cx * a.cy;
    unsigned long long *mass = reinterpret_cast<unsigned long long *>(a.mass + (long long)t * cells);
    int *count = a.count + (long long)t * cells;
    const float4 *tab = (FUSED || OCC) ? a.table + (size_t)t * a.egx * a.egy : nullptr;
    const float mx = FUSED ? __ldg(a.lmax + t) : 0.0f;
    const float cmx = (float)(a.cx - 1), cmy = (float)(a.cy - 1);
    long long chunk = (a.N + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 3) & ~3LL;
    const long long n_lo = min((long long)blockIdx.x * chunk, a.N), n_hi = min(n_lo + chunk, a.N);
    if (tid < 32) {  // window origin: mean cell of 32 probe samples spread over the chunk (any choice gives the same grids)
        int sx = 0, sy = 0, sk = 0;
        if (n_lo < n_hi) {
            const long long n = n_lo + (n_hi - n_lo) * tid / 32;
            const float *xs = reinterpret_cast<const float *>(xp), *ys = reinterpret_cast<const float *>(yp);
            sx = lean_cell_biased(__ldg(xs + n), a.q.x_lo, a.q.x_rng, a.q.r_x_rng, a.q.gxm1, cmx) - LEAN_MAGIC;
            sy = lean_cell_biased(__ldg(ys + n), a.q.y_lo, a.q.y_rng, a.q.r_y_rng, a.q.gym1, cmy) - LEAN_MAGIC;
            sk = 1;
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, m);
            sy += __shfl_xor_sync(0xffffffffu, sy, m);
            sk += __shfl_xor_sync(0xffffffffu, sk, m);
        }
        if (tid == 0) {
            sorg[0] = min(max((sk ? sx / sk : 0) - WIN / 2, 0), max(a.cx - WIN, 0));
            sorg[1] = min(max((sk ? sy / sk : 0) - WIN / 2, 0), max(a.cy - WIN, 0));
        }
    }
    for (int i = tid; i < WIN * WIN; i += NT) {
        slo[i] = 0u;
        shi[i] = 0u;
        scount[i] = 0;
    }
    __syncthreads();
    const int ox = sorg[0], oy = sorg[1];
    const int oxb = ox + LEAN_MAGIC, oyb = oy + LEAN_MAGIC;
    const long long qconst = (long long)a.scale;
    double dot = 0.0;
    long long cq = 0;
    auto put = [&](float x, float y, float w, float r0) {
        const int wx = lean_cell_biased(x, a.q.x_lo, a.q.x_rng, a.q.r_x_rng, a.q.gxm1, cmx) - oxb;
        const int wy = lean_cell_biased(y, a.q.y_lo, a.q.y_rng, a.q.r_y_rng, a.q.gym1, cmy) - oyb;
        if (FUSED) w = __fmul_rn(expf(__fsub_rn(w, mx)), r0);  // simulation_objects.py:296-297
        // w * 2^s is exact in fp32 (power-of-two scale), so this equals llrint((double)w * 2^s)
        const long long q = HASW ? __float2ll_rn(__fmul_rn(w, a.scale_f)) : qconst;
        if ((unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN) {
            const int wc = wx * WIN + wy;
            const unsigned ql = (unsigned)q, qh = (unsigned)((unsigned long long)q >> 32);
            const unsigned old = atomicAdd(slo + wc, ql);
            const unsigned up = qh + ((old + ql < old) ? 1u : 0u);
            if (up) atomicAdd(shi + wc, up);
            atomicAdd(scount + wc, 1);
        } else {
            const int ix = wx + ox, iy = wy + oy;
            const int cell = ix * a.cy + iy;
            atomicAdd(mass + cell, (unsigned long long)q);
            atomicAdd(count + cell, 1);
            // occupancy dot product: window cells contribute mass * p once at the flush, stray samples here
            if (OCC && ix < a.egx && iy < a.egy) dot += (double)q * (double)__ldg(&tab[ix * a.egy + iy].x);
        }
        if (FUSED) {
            // the cost clamps the cell to [0, G-1] (MotionPlanner.py:208-209), the binning to [0, G]
            const int jx = min(wx + ox, a.egx - 1), jy = min(wy + oy, a.egy - 1);
            const float4 e = __ldg(tab + jx * a.egy + jy);
            const DpEval ev = dp_cost_term(e, x, y, w, true);
            cq += __float2ll_rn(__fmul_rn(ev.term, a.cost_scale_f));
        }
    };
    const long long g_lo = n_lo >> 2, g_hi = n_hi >> 2;  // full groups of four (chunk boundaries are multiples of four)
    if (VW == 4) {
        const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
        const long long pf = (long long)a.prefetch * NT;
        for (long long g = g_lo + tid; g < g_hi; g += NT) {
            if (pf && g + pf < g_hi) {  // a later iteration's rows on their way into L2 while this one computes
                dp_prefetch_l2(xp + g + pf);
                dp_prefetch_l2(yp + g + pf);
                if (HASW) dp_prefetch_l2(wp + g + pf);
            }
            const float4 xv = __ldg(xp + g), yv = __ldg(yp + g);
            const float4 wv = HASW ? __ldg(wp + g) : ones;
            const float4 rv = rp ? __ldg(rp + g) : ones;
            put(xv.x, yv.x, wv.x, rv.x);
            put(xv.y, yv.y, wv.y, rv.y);
            put(xv.z, yv.z, wv.z, rv.z);
            put(xv.w, yv.w, wv.w, rv.w);
        }
    } else {
        const float2 ones = make_float2(1.f, 1.f);
        const float2 *xp2 = reinterpret_cast<const float2 *>(xp), *yp2 = reinterpret_cast<const float2 *>(yp);
        const float2 *wp2 = reinterpret_cast<const float2 *>(wp), *rp2 = reinterpret_cast<const float2 *>(rp);
        for (long long g = 2 * g_lo + tid; g < 2 * g_hi; g += NT) {
            const float2 xv = __ldg(xp2 + g), yv = __ldg(yp2 + g);
            const float2 wv = HASW ? __ldg(wp2 + g) : ones;
            const float2 rv = rp ? __ldg(rp2 + g) : ones;
            put(xv.x, yv.x, wv.x, rv.x);
            put(xv.y, yv.y, wv.y, rv.y);
        }
    }
    // the last block of the slice also takes the (at most three) samples behind the last full group
    if (n_hi == a.N && n_lo < a.N && (g_hi << 2) + tid < a.N) {
        const long long n = (g_hi << 2) + tid;
        put(__ldg(reinterpret_cast<const float *>(xp) + n), __ldg(reinterpret_cast<const float *>(yp) + n),
            HASW ? __ldg(reinterpret_cast<const float *>(wp) + n) : 1.f, rp ? __ldg(a.rho0 + n) : 1.f);
    }
    __syncthreads();
    for (int i = tid; i < WIN * WIN; i += NT) {  // flush the window: one atomic per touched cell
        const int c = scount[i];
        if (c) {
            const int wx = i / WIN, wy = i - wx * WIN;
            const int cell = (ox + wx) * a.cy + (oy + wy);  // the window lies inside the grid (origin clamped) unless the grid is smaller
            if (ox + wx < a.cx && oy + wy < a.cy) {
                const unsigned long long m = ((unsigned long long)shi[i] << 32) | slo[i];
                if (m) atomicAdd(mass + cell, m);
                atomicAdd(count + cell, c);
                // sum_s w_s p[cell_s] over this block's window samples = sum_cells (window mass) * p[cell]
                if (OCC && m && ox + wx < a.egx && oy + wy < a.egy)
                    dot += (double)(long long)m * (double)__ldg(&tab[(ox + wx) * a.egy + (oy + wy)].x);
            }
        }
    }
    if (FUSED) {  // integer sum: exact and independent of the order (one atomic per block)
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) cq += __shfl_xor_sync(0xffffffffu, cq, m);
        if ((tid & 31) == 0) shq[tid >> 5] = cq;
        __syncthreads();
        if (tid == 0) {
            long long tot = 0;
            for (int w = 0; w < NT / 32; ++w) tot += shq[w];
            if (tot) atomicAdd(reinterpret_cast<unsigned long long *>(a.cost_q + t), (unsigned long long)tot);
        }
    } else if (OCC) {  // fixed-order block reduction of the occupancy dot product
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
        if ((tid & 31) == 0) shd[tid >> 5] = dot;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < NT / 32; ++w) tot += shd[w];
            a.partial[(size_t)t * a.B + blockIdx.x] = tot;
        }
    }
}

__global__ void bin_dot_finalize_kernel(const double *partial, int B, int Ts, double inv_scale, double *occ_dot) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Ts; t += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += partial[(size_t)t * B + b];
        occ_dot[t] = s * inv_scale;
    }
}

// ---- normaliser ----------------------------------------------------------------------------------------------------
// per time slot: max, then sum of exp(l - max) * rho0 in fp64; grid (B, Ts) partials, fixed reduction order
constexpr int NB = 256;  // threads per block
__global__ void __launch_bounds__(NB) normalize_max_kernel(const float *__restrict__ rholog, long long N, long long ldn,
                                                           float *pmax, int B) {
    __shared__ float shf[NB / 32];
    const int t = blockIdx.y;
    const float *l = rholog + (long long)t * ldn;
    float m = -INFINITY;
    for (long long n = (long long)blockIdx.x * NB + threadIdx.x; n < N; n += (long long)gridDim.x * NB) m = fmaxf(m, __ldg(l + n));
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = -INFINITY;
        for (int i = 0; i < NB / 32; ++i) mm = fmaxf(mm, shf[i]);
        pmax[(size_t)t * B + blockIdx.x] = mm;
    }
}
__global__ void __launch_bounds__(NB) normalize_sum_kernel(const float *__restrict__ rholog, const float *__restrict__ rho0,
                                                           long long N, long long ldn, const float *pmax, double *psum, int B) {
    __shared__ double shd[NB / 32];
    __shared__ float s_max;
    const int t = blockIdx.y;
    if (threadIdx.x == 0) {
        float mm = -INFINITY;
        for (int b = 0; b < B; ++b) mm = fmaxf(mm, pmax[(size_t)t * B + b]);
        s_max = mm;
    }
    __syncthreads();
    const float mx = s_max;
    const float *l = rholog + (long long)t * ldn;
    double s = 0.0;
    for (long long n = (long long)blockIdx.x * NB + threadIdx.x; n < N; n += (long long)gridDim.x * NB) {
        const float r0 = rho0 ? __ldg(rho0 + n) : 1.0f;
        s += (double)__fmul_rn(expf(__fsub_rn(__ldg(l + n), mx)), r0);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < NB / 32; ++i) tot += shd[i];
        psum[(size_t)t * B + blockIdx.x] = tot;
    }
}
__global__ void normalize_finalize_kernel(const float *pmax, const double *psum, int B, int Ts, float *lmax, double *lsum) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Ts; t += gridDim.x * blockDim.x) {
        float mm = -INFINITY;
        double s = 0.0;
        for (int b = 0; b < B; ++b) {
            mm = fmaxf(mm, pmax[(size_t)t * B + b]);
            s += psum[(size_t)t * B + b];
        }
        lmax[t] = mm;
        lsum[t] = s;
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

// tuning / A-B knobs (read once): DP_BIN_THREADS = 768 (default, 40 registers per thread, two blocks per SM), 640 (48 registers), 1024 (32 registers) or 512 (64 registers),
// DP_BIN_AGG = 0 (never aggregate), 1 (always), 2 (default: decided per block from a probe)
static int bin_waves() {  // DP_BIN_BLOCKS_PER_SM: target number of blocks per SM over all slices (default 8; measured at the config-3 size: 2 -> 0.30 ms, 4 -> 0.25 ms, 8 -> 0.23 ms)
    static int s_w = 0;
    if (!s_w) {
        const char *e = getenv("DP_BIN_BLOCKS_PER_SM");
        s_w = (e && atoi(e) >= 1 && atoi(e) <= 64) ? atoi(e) : 8;
    }
    return s_w;
}
static long long bin_min_samples() {  // DP_BIN_MIN_SAMPLES: fewest samples a block is given (default 8 per thread of a 768-thread block)
    static long long s_m = 0;
    if (!s_m) {
        const char *e = getenv("DP_BIN_MIN_SAMPLES");
        s_m = (e && atoll(e) >= 1024) ? atoll(e) : 6144;
    }
    return s_m;
}
static void bin_knobs(int &threads, int &agg) {
    static int s_threads = 0, s_agg = 2;
    if (!s_threads) {
        const char *e = getenv("DP_BIN_THREADS");
        s_threads = (e && atoi(e) == 512) ? 512 : (e && atoi(e) == 1024) ? 1024 : (e && atoi(e) == 640) ? 640 : 768;
        const char *g = getenv("DP_BIN_AGG");
        if (g) s_agg = atoi(g) < 0 ? 0 : (atoi(g) > 2 ? 2 : atoi(g));
    }
    threads = s_threads;
    agg = s_agg;
}

template <int NT>
static int bin2d_dispatch(const BinArgs &a, bool vec, bool fused, dim3 grid, cudaStream_t stream) {
    static bool attr_set_dev[64] = {};
    int cur_dev = 0;
    DP_CUDA(cudaGetDevice(&cur_dev));
    bool &attr_set = attr_set_dev[cur_dev & 63];
    if (!attr_set) {
        DP_CUDA(cudaFuncSetAttribute(bin2d_kernel<0, false, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_kernel<0, true, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_kernel<0, false, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_kernel<0, true, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_kernel<1, false, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_lean_kernel<true, false, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_lean_kernel<false, true, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_lean_kernel<false, true, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_lean_kernel<false, false, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        DP_CUDA(cudaFuncSetAttribute(bin2d_lean_kernel<false, false, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIN_SMEM));
        attr_set = true;
    }
    // default: the lean kernel (no warp collectives); DP_BIN_AGG=1 or an unusual layout selects the general kernel
    const bool lean = a.mode == 0 && vec && a.agg != 1 && a.mass && a.count && (long long)a.cx * a.cy < (1LL << 30) &&
                      (!a.table || (long long)a.egx * a.egy < (1LL << 30));
    if (lean) {
        if (fused)
            bin2d_lean_kernel<true, false, true, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
        else if (a.partial && a.w)
            bin2d_lean_kernel<false, true, true, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
        else if (a.partial)
            bin2d_lean_kernel<false, true, false, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
        else if (a.w)
            bin2d_lean_kernel<false, false, true, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
        else
            bin2d_lean_kernel<false, false, false, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    } else if (a.mode == 1)
        bin2d_kernel<1, false, false, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    else if (fused && vec)
        bin2d_kernel<0, true, true, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    else if (fused)
        bin2d_kernel<0, false, true, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    else if (vec)
        bin2d_kernel<0, true, false, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    else
        bin2d_kernel<0, false, false, NT><<<grid, NT, BIN_SMEM, stream>>>(a);
    DP_CUDA(cudaGetLastError());
    return 0;
}

// fused == true (dp_pipeline_bin): w holds the log-density, lmax / rho0 / cost_q as in BinArgs
static int bin2d_launch(const dp_bin_spec *spec, const dp_env *env, const float *x, const float *y, int64_t sx_n, int64_t sx_t,
                        const float *w, int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count,
                        double *occ_dot, bool fused, const float *lmax, const float *rho0, long long *cost_q,
                        int cost_scale_log2, cudaStream_t stream) {
    DP_REQUIRE(spec && x && y && (mass || count || occ_dot), "dp_bin2d: null argument");
    DP_REQUIRE(spec->mode == 0 || spec->mode == 1, "dp_bin2d: mode must be 0 or 1");
    DP_REQUIRE(spec->mass_scale_log2 >= 0 && spec->mass_scale_log2 <= 62, "dp_bin2d: bad mass_scale_log2");
    DP_REQUIRE(!occ_dot || (env && spec->mode == 0), "dp_bin2d_occ: the occupancy dot product needs an environment and mode 0");
    DP_REQUIRE(!occ_dot || Ts <= env->Tg, "dp_bin2d_occ: Ts=%d exceeds the %d time slices of the environment", Ts, occ_dot ? env->Tg : 0);
    DP_REQUIRE(!fused || (env && spec->mode == 0 && w && lmax && cost_q && Ts <= env->Tg), "dp_pipeline_bin: bad argument");
    if (Ts <= 0) return 0;
    if (N <= 0) {
        if (occ_dot) DP_CUDA(cudaMemsetAsync(occ_dot, 0, sizeof(double) * Ts, stream));
        return 0;
    }
    int NT, agg;
    bin_knobs(NT, agg);
    // the variant with the occupancy dot product carries more state per thread: at 768 threads (40 registers) ptxas
    // rematerialises constants per sample; 640 threads (48 registers) measured 0.243 against 0.263 ms, while the plain
    // binning prefers 768 (0.221 against 0.229 ms) -- profiles/r02_k3_threads.txt.  DP_BIN_THREADS overrides both.
    if (occ_dot && !getenv("DP_BIN_THREADS")) NT = 640;
    BinArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = spec->mode;
    if (a.mode == 0) {
        a.q = dp_make_geom(spec->geom);
        a.cx = spec->geom.gx + 1; a.cy = spec->geom.gy + 1;
    } else {
        DP_REQUIRE(spec->bx > 0 && spec->by > 0 && spec->hi_x > spec->lo_x && spec->hi_y > spec->lo_y, "dp_bin2d: bad bins");
        a.cx = spec->bx; a.cy = spec->by;
        a.lo_x = spec->lo_x; a.hi_x = spec->hi_x; a.lo_y = spec->lo_y; a.hi_y = spec->hi_y;
    }
    a.scale = ldexp(1.0, spec->mass_scale_log2);
    a.scale_f = (float)a.scale;
    a.x = x; a.y = y; a.w = w;
    a.sx_n = sx_n; a.sx_t = sx_t; a.sw_n = sw_n; a.sw_t = sw_t; a.N = N; a.Ts = Ts;
    a.mass = mass; a.count = count;
    a.agg = agg;
    {
        static int s_pf = -1;
        if (s_pf < 0) {
            const char *pe = getenv("DP_BIN_PREFETCH");
            s_pf = (pe && atoi(pe) >= 0 && atoi(pe) <= 8) ? atoi(pe) : 1;
        }
        a.prefetch = s_pf;
    }
    a.lmax = lmax; a.rho0 = rho0; a.cost_q = cost_q;
    a.cost_scale_f = (float)ldexp(1.0, cost_scale_log2);
    // ~2 resident blocks per SM and two waves; every block amortises one window flush over >= 8 samples per thread
    // every block pays for zeroing and flushing its 77 KB window: at least bin_min_samples() samples per block
    long long B = (N + bin_min_samples() - 1) / bin_min_samples();
    // rounded DOWN: the blocks of all slices then fit the intended number of full waves (two resident blocks per SM); one
    // block more per slice puts an almost empty wave at the end (measured on the cost kernels, profiles/r02_cost_knobs.txt)
    long long target = (148LL * bin_waves()) / Ts;
    if (target < 1) target = 1;
    if (B > target) B = target;
    if (B * Ts > MAX_BIN_PARTIALS) B = MAX_BIN_PARTIALS / Ts;
    if (B < 1) B = 1;
    a.B = (int)B;
    double *partial = nullptr;
    if (occ_dot || fused) {
        a.table = env->d_table;
        a.egx = env->geom.gx;
        a.egy = env->geom.gy;
    }
    if (occ_dot) {
        DP_REQUIRE(Ts <= MAX_BIN_PARTIALS, "dp_bin2d_occ: too many slices");
        DP_CUDA(cudaMallocAsync(&partial, sizeof(double) * B * Ts, stream));
        a.partial = partial;
    }
    auto al16 = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
    const bool vec = a.mode == 0 && sx_n == 1 && (!w || sw_n == 1) && al16(x) && al16(y) && al16(w) && al16(rho0) &&
                     sx_t % 4 == 0 && (!w || sw_t % 4 == 0);
    const dim3 grid((unsigned)B, (unsigned)Ts);
    const int rc = NT == 512   ? bin2d_dispatch<512>(a, vec, fused, grid, stream)
                   : NT == 640 ? bin2d_dispatch<640>(a, vec, fused, grid, stream)
                   : NT == 768 ? bin2d_dispatch<768>(a, vec, fused, grid, stream)
                               : bin2d_dispatch<1024>(a, vec, fused, grid, stream);
    if (rc) return rc;
    if (occ_dot) {
        bin_dot_finalize_kernel<<<(Ts + 127) / 128, 128, 0, stream>>>(partial, a.B, Ts, 1.0 / a.scale, occ_dot);
        DP_CUDA(cudaGetLastError());
        DP_CUDA(cudaFreeAsync(partial, stream));
    }
    return 0;
}

// fused pipeline pass (pipeline.cu): weights exp(l - lmax[t]) * rho0[n] formed in the kernel, occupancy grids and the
// collision cost of the weighted samples (int64 fixed point) from one pass over x, y, l
int dp_bin2d_fused_launch(const dp_bin_spec *spec, const dp_env *env, const float *x, const float *y, int64_t sx_t,
                          const float *l, int64_t sl_t, const float *rho0, const float *lmax, int64_t N, int Ts,
                          long long *mass, int *count, long long *cost_q, int cost_scale_log2, cudaStream_t stream) {
    return bin2d_launch(spec, env, x, y, 1, sx_t, l, 1, sl_t, N, Ts, mass, count, nullptr, true, lmax, rho0, cost_q,
                        cost_scale_log2, stream);
}

int dp_bin2d(const dp_bin_spec *spec, const float *x, const float *y, int64_t sx_n, int64_t sx_t, const float *w,
             int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count, void *stream) {
    DP_REQUIRE(mass || count, "dp_bin2d: null argument");
    return bin2d_launch(spec, nullptr, x, y, sx_n, sx_t, w, sw_n, sw_t, N, Ts, mass, count, nullptr, false, nullptr, nullptr,
                        nullptr, 0, (cudaStream_t)stream);
}

int dp_bin2d_occ(const dp_bin_spec *spec, const dp_env *env, const float *x, const float *y, int64_t sx_n, int64_t sx_t,
                 const float *w, int64_t sw_n, int64_t sw_t, int64_t N, int Ts, long long *mass, int *count, double *occ_dot,
                 void *stream) {
    DP_REQUIRE(env && occ_dot, "dp_bin2d_occ: null argument");
    DeviceGuard guard(env->device);
    return bin2d_launch(spec, env, x, y, sx_n, sx_t, w, sw_n, sw_t, N, Ts, mass, count, occ_dot, false, nullptr, nullptr,
                        nullptr, 0, (cudaStream_t)stream);
}

int dp_normalize_stats(const float *rholog, const float *rho0, int64_t N, int Ts, int64_t ldn, float *lmax,
                       double *lsum_at_max, void *stream_) {
    DP_REQUIRE(rholog && lmax && lsum_at_max, "dp_normalize_stats: null argument");
    if (N <= 0 || Ts <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    long long B = (N + NB * 8 - 1) / (NB * 8);
    const long long target = (148LL * 16 + Ts - 1) / Ts;
    if (B > target) B = target;
    if (B < 1) B = 1;
    float *pmax = nullptr;
    double *psum = nullptr;
    DP_CUDA(cudaMallocAsync(&psum, (sizeof(double) + sizeof(float)) * B * Ts, stream));
    pmax = reinterpret_cast<float *>(psum + B * Ts);
    normalize_max_kernel<<<dim3((unsigned)B, (unsigned)Ts), NB, 0, stream>>>(rholog, N, ldn, pmax, (int)B);
    normalize_sum_kernel<<<dim3((unsigned)B, (unsigned)Ts), NB, 0, stream>>>(rholog, rho0, N, ldn, pmax, psum, (int)B);
    normalize_finalize_kernel<<<(Ts + 127) / 128, 128, 0, stream>>>(pmax, psum, (int)B, Ts, lmax, lsum_at_max);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaFreeAsync(psum, stream));
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
