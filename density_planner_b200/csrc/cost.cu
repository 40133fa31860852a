// K2: collision cost, forward + analytic backward, and the environment table it gathers from.
//   MotionPlanner.get_cost_coll              motion_planning/MotionPlanner.py:192-224
//   MotionPlannerGrad.get_cost_coll_initialize  motion_planning/MotionPlannerGrad.py:309-336
//   pos2gridpos / gridpos2pos                motion_planning/utils.py:70-116
// HBM-bound: 12 B streamed per (sample, slice) forward (+12 B gradients), one 16-byte gather from a time-major
// {p, gradX, gradY} table that stays L2-resident because samples cluster around the reference trajectory.
#include <stdlib.h>

#include "dp_common.cuh"

namespace {

constexpr int CB = 256;  // threads per block
constexpr int MAX_PARTIALS = 1 << 16;

typedef DpGeom Geom;
typedef DpEval Eval;
#define make_geom dp_make_geom
#define eval_one dp_cost_eval

__device__ __forceinline__ double block_reduce_sum(double v, double *sh) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double tot = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    return tot;  // valid in thread 0
}
__device__ __forceinline__ float block_reduce_max(float v, float *sh) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float tot = -INFINITY;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot = fmaxf(tot, sh[w]);
    return tot;
}

struct CostArgs {
    const float4 *table;
    Geom q;
    const float *x, *y, *rho;
    long long sx_n, sx_t, sr_n, sr_t, sg_n, sg_t;
    long long N;
    int Ts, get_max;
    float *gx, *gy, *grho;
    double *partial;      // [Ts][B]   sum mode
    float *partial_max;   // [Ts][B]   max mode (-inf = no active sample)
    double *cost_samples; // [N] or null (generic kernel only)
    int B;
    int prefetch;         // L2 prefetch distance of the streamed rows in grid-stride iterations (0 off; DP_COST_PREFETCH)
};

// (A) main path: samples contiguous (sx_n == 1).  grid (B, Ts): block (b, t) strides over the samples of slice t.
__global__ void __launch_bounds__(CB) cost_nfast_kernel(CostArgs a) {
    __shared__ double shd[CB / 32];
    __shared__ float shf[CB / 32];
    const int t = blockIdx.y;
    const bool has_rho = a.rho != nullptr;
    const float *xp = a.x + (long long)t * a.sx_t;
    const float *yp = a.y + (long long)t * a.sx_t;
    const float *rp = has_rho ? a.rho + (long long)t * a.sr_t : nullptr;
    const bool want_g = a.gx != nullptr;
    double acc = 0.0;
    float mx = -INFINITY;
    const long long stride = (long long)gridDim.x * CB;
    for (long long n0 = (long long)blockIdx.x * CB + threadIdx.x; n0 < a.N; n0 += 4 * stride) {
        float xv[4], yv[4], rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long n = n0 + j * stride;
            const bool ok = n < a.N;
            xv[j] = ok ? __ldg(xp + n * a.sx_n) : 0.0f;
            yv[j] = ok ? __ldg(yp + n * a.sx_n) : 0.0f;
            rv[j] = (ok && has_rho) ? __ldg(rp + n * a.sr_n) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long n = n0 + j * stride;
            if (n < a.N) {
                const Eval e = eval_one(a.table, a.q, t, xv[j], yv[j], rv[j], has_rho);
                acc += (double)e.term;
                if (e.active) mx = fmaxf(mx, e.term);
                if (want_g) {
                    const long long o = n * a.sg_n + (long long)t * a.sg_t;
                    a.gx[o] = e.gx;
                    a.gy[o] = e.gy;
                    if (a.grho) a.grho[o] = e.grho;
                }
            }
        }
    }
    if (a.get_max) {
        const float m = block_reduce_max(mx, shf);
        if (threadIdx.x == 0) a.partial_max[(size_t)t * a.B + blockIdx.x] = m;
    } else {
        const double s = block_reduce_sum(acc, shd);
        if (threadIdx.x == 0) a.partial[(size_t)t * a.B + blockIdx.x] = s;
    }
}

// (V) the same for 16-byte aligned rows: every thread evaluates four consecutive samples per iteration from three LDG.128
// (x, y, rho) and writes the gradients with STG.128 -- a quarter of the memory instructions of (A) for the same bytes.
template <bool WANT_G, int MINB>
__global__ void __launch_bounds__(CB, MINB) cost_vec4_kernel(CostArgs a) {
    __shared__ double shd[CB / 32];
    __shared__ float shf[CB / 32];
    const int t = blockIdx.y;
    const bool has_rho = a.rho != nullptr;
    const float *xp = a.x + (long long)t * a.sx_t;
    const float *yp = a.y + (long long)t * a.sx_t;
    const float *rp = has_rho ? a.rho + (long long)t * a.sr_t : nullptr;
    double acc = 0.0;
    float mx = -INFINITY;
    const long long n4 = a.N >> 2;  // full groups of four
    const long long stride = (long long)gridDim.x * CB;
    const long long pf = a.prefetch * stride;
    for (long long g = (long long)blockIdx.x * CB + threadIdx.x; g < n4; g += stride) {
        if (pf && g + pf < n4) {  // the rows of a later iteration on their way into L2 while this one computes
            dp_prefetch_l2(reinterpret_cast<const float4 *>(xp) + g + pf);
            dp_prefetch_l2(reinterpret_cast<const float4 *>(yp) + g + pf);
            if (has_rho) dp_prefetch_l2(reinterpret_cast<const float4 *>(rp) + g + pf);
        }
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(xp) + g);
        const float4 yv = __ldg(reinterpret_cast<const float4 *>(yp) + g);
        const float4 rv = has_rho ? __ldg(reinterpret_cast<const float4 *>(rp) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
        const Eval e0 = eval_one(a.table, a.q, t, xv.x, yv.x, rv.x, has_rho);
        const Eval e1 = eval_one(a.table, a.q, t, xv.y, yv.y, rv.y, has_rho);
        const Eval e2 = eval_one(a.table, a.q, t, xv.z, yv.z, rv.z, has_rho);
        const Eval e3 = eval_one(a.table, a.q, t, xv.w, yv.w, rv.w, has_rho);
        acc += (double)e0.term;  // sample order within the thread: fixed
        acc += (double)e1.term;
        acc += (double)e2.term;
        acc += (double)e3.term;
        if (e0.active) mx = fmaxf(mx, e0.term);
        if (e1.active) mx = fmaxf(mx, e1.term);
        if (e2.active) mx = fmaxf(mx, e2.term);
        if (e3.active) mx = fmaxf(mx, e3.term);
        if (WANT_G) {
            const long long o = (long long)t * a.sg_t;
            reinterpret_cast<float4 *>(a.gx + o)[g] = make_float4(e0.gx, e1.gx, e2.gx, e3.gx);
            reinterpret_cast<float4 *>(a.gy + o)[g] = make_float4(e0.gy, e1.gy, e2.gy, e3.gy);
            if (a.grho) reinterpret_cast<float4 *>(a.grho + o)[g] = make_float4(e0.grho, e1.grho, e2.grho, e3.grho);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.N & 3)) {  // tail samples
        const long long n = (n4 << 2) + threadIdx.x;
        const Eval e = eval_one(a.table, a.q, t, __ldg(xp + n), __ldg(yp + n), has_rho ? __ldg(rp + n) : 0.0f, has_rho);
        acc += (double)e.term;
        if (e.active) mx = fmaxf(mx, e.term);
        if (WANT_G) {
            const long long o = n + (long long)t * a.sg_t;
            a.gx[o] = e.gx;
            a.gy[o] = e.gy;
            if (a.grho) a.grho[o] = e.grho;
        }
    }
    if (a.get_max) {
        const float m = block_reduce_max(mx, shf);
        if (threadIdx.x == 0) a.partial_max[(size_t)t * a.B + blockIdx.x] = m;
    } else {
        const double s = block_reduce_sum(acc, shd);
        if (threadIdx.x == 0) a.partial[(size_t)t * a.B + blockIdx.x] = s;
    }
}

// (G) generic strides / per-sample sums: one thread per sample loops over the slices; per-slice block reduction.
__global__ void __launch_bounds__(128) cost_generic_kernel(CostArgs a) {
    __shared__ double shd[4];
    __shared__ float shf[4];
    const long long n = (long long)blockIdx.x * 128 + threadIdx.x;
    const bool ok = n < a.N;
    const bool has_rho = a.rho != nullptr;
    double mine = 0.0;
    for (int t = 0; t < a.Ts; ++t) {
        Eval e;
        e.term = 0; e.gx = 0; e.gy = 0; e.grho = 0; e.active = false;
        if (ok) {
            const float x = __ldg(a.x + n * a.sx_n + (long long)t * a.sx_t);
            const float y = __ldg(a.y + n * a.sx_n + (long long)t * a.sx_t);
            const float r = has_rho ? __ldg(a.rho + n * a.sr_n + (long long)t * a.sr_t) : 0.0f;
            e = eval_one(a.table, a.q, t, x, y, r, has_rho);
            mine += (double)e.term;
            if (a.gx) {
                const long long o = n * a.sg_n + (long long)t * a.sg_t;
                a.gx[o] = e.gx;
                a.gy[o] = e.gy;
                if (a.grho) a.grho[o] = e.grho;
            }
        }
        if (a.get_max) {
            const float m = block_reduce_max(e.active ? e.term : -INFINITY, shf);
            if (threadIdx.x == 0) a.partial_max[(size_t)t * a.B + blockIdx.x] = m;
        } else {
            const double s = block_reduce_sum((double)e.term, shd);
            if (threadIdx.x == 0) a.partial[(size_t)t * a.B + blockIdx.x] = s;
        }
    }
    if (ok && a.cost_samples) a.cost_samples[n] = mine;
}

// fixed-order reduction of the per-block partials: slice sums and the total
__global__ void cost_finalize_kernel(const double *partial, const float *partial_max, int B, int Ts, int get_max,
                                     double *cost_slices, double *total) {
    __shared__ double sh[1024];
    double tot = 0.0;
    for (int t = threadIdx.x; t < Ts; t += blockDim.x) {
        double s = 0.0;
        if (get_max) {
            float m = -INFINITY;
            for (int b = 0; b < B; ++b) m = fmaxf(m, partial_max[(size_t)t * B + b]);
            s = (m == -INFINITY) ? 0.0 : (double)m;  // slices without an active sample add nothing (:213)
        } else {
            for (int b = 0; b < B; ++b) s += partial[(size_t)t * B + b];
        }
        if (cost_slices) cost_slices[t] = s;
        tot += s;  // thread-local over t = tid, tid+blockDim, ... (fixed order)
    }
    sh[threadIdx.x] = tot;
    __syncthreads();
    if (threadIdx.x == 0 && total) {
        double s = 0.0;
        for (int i = 0; i < (int)blockDim.x; ++i) s += sh[i];
        *total = s;
    }
}

__global__ void env_pack_kernel(const float *__restrict__ grid, const float *__restrict__ gradx,
                                const float *__restrict__ grady, int G0, int G1, int Tg, Geom q, float4 *__restrict__ table) {
    // input (G0,G1,Tg) time innermost; output [t][ix][iy].  Tile-transposed through shared memory.
    // Everything of the collision cost that depends on the CELL alone is evaluated here, once per cell instead of once per
    // sample: des_gridpos = gridpos + 100 * grad and des_pos = gridpos2pos(des_gridpos) (MotionPlanner.py:215-217), op for op
    // in fp32 as the per-sample code did (so the costs and gradients stay bit-identical), and the activity test (:213-214).
    // Entry = {p, des_pos_x, des_pos_y, active ? 1 : 0}.
    __shared__ float4 tile[32][33];
    const long long cell0 = (long long)blockIdx.x * 32;  // cell = ix*G1 + iy
    const int t0 = blockIdx.y * 32;
    const long long cells = (long long)G0 * G1;
    {
        const int tl = threadIdx.x, cl = threadIdx.y;  // read: t fastest
        for (int c = cl; c < 32; c += blockDim.y) {
            const long long cell = cell0 + c;
            const int t = t0 + tl;
            if (cell < cells && t < Tg) {
                const size_t i = (size_t)cell * Tg + t;
                const float gX = gradx[i], gY = grady[i];
                const int ix = (int)(cell / G1), iy = (int)(cell - (long long)ix * G1);
                const float dgx = __fadd_rn(__int2float_rn(ix), __fmul_rn(100.0f, gX));
                const float dgy = __fadd_rn(__int2float_rn(iy), __fmul_rn(100.0f, gY));
                tile[c][tl] = make_float4(grid[i], dp_pos<false>(dgx, q.x_lo, q.x_rng, q.gxm1, q.r_gxm1),
                                          dp_pos<false>(dgy, q.y_lo, q.y_rng, q.gym1, q.r_gym1),
                                          (gX != 0.0f || gY != 0.0f) ? 1.0f : 0.0f);
            }
        }
    }
    __syncthreads();
    {
        const int cl = threadIdx.x, tl = threadIdx.y;  // write: cell fastest
        for (int tt = tl; tt < 32; tt += blockDim.y) {
            const long long cell = cell0 + cl;
            const int t = t0 + tt;
            if (cell < cells && t < Tg) table[(size_t)t * cells + cell] = tile[cl][tt];
        }
    }
}

}  // namespace

// workspace for the partial sums lives with the environment (one in-flight dp_cost_coll per dp_env)
struct dp_env_ws {
    double *partial;
    float *partial_max;
};
static dp_env_ws *env_ws(const dp_env *e);

struct dp_env_full : dp_env {
    dp_env_ws ws;
};
static dp_env_ws *env_ws(const dp_env *e) { return &static_cast<dp_env_full *>(const_cast<dp_env *>(e))->ws; }

int dp_env_create(const dp_grid_geom *g, int Tg, const float *grid, const float *gradx, const float *grady, int on_device,
                  int device, dp_env **out) {
    DP_REQUIRE(g && grid && gradx && grady && out && Tg > 0 && g->gx > 1 && g->gy > 1, "dp_env_create: bad argument");
    DeviceGuard guard(device);
    dp_env_full *e = new dp_env_full();
    e->device = device;
    e->geom = *g;
    e->Tg = Tg;
    const size_t cells = (size_t)g->gx * g->gy, n = cells * Tg;
    const float *d_in[3] = {grid, gradx, grady};
    float *tmp[3] = {nullptr, nullptr, nullptr};
    if (!on_device) {
        for (int i = 0; i < 3; ++i) {
            DP_CUDA(cudaMalloc(&tmp[i], n * sizeof(float)));
            DP_CUDA(cudaMemcpy(tmp[i], d_in[i], n * sizeof(float), cudaMemcpyHostToDevice));
            d_in[i] = tmp[i];
        }
    }
    DP_CUDA(cudaMalloc(&e->d_table, n * sizeof(float4)));
    DP_CUDA(cudaMalloc(&e->ws.partial, MAX_PARTIALS * sizeof(double)));
    DP_CUDA(cudaMalloc(&e->ws.partial_max, MAX_PARTIALS * sizeof(float)));
    dim3 blk(32, 8), grd((unsigned)((cells + 31) / 32), (unsigned)((Tg + 31) / 32));
    env_pack_kernel<<<grd, blk>>>(d_in[0], d_in[1], d_in[2], g->gx, g->gy, Tg, make_geom(*g), e->d_table);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 3; ++i)
        if (tmp[i]) cudaFree(tmp[i]);
    *out = e;
    return 0;
}

void dp_env_destroy(dp_env *e0) {
    if (!e0) return;
    dp_env_full *e = static_cast<dp_env_full *>(e0);
    DeviceGuard guard(e->device);
    cudaFree(e->d_table);
    cudaFree(e->ws.partial);
    cudaFree(e->ws.partial_max);
    delete e;
}

int dp_cost_coll(const dp_env *e, const float *x, const float *y, int64_t sx_n, int64_t sx_t, const float *rho,
                 int64_t sr_n, int64_t sr_t, int64_t N, int Ts, int get_max, double *cost_slices, double *cost_samples,
                 float *gx, float *gy, float *grho, int64_t sg_n, int64_t sg_t, double *total, void *stream_) {
    DP_REQUIRE(e && x && y, "dp_cost_coll: null argument");
    DP_REQUIRE(Ts >= 0 && Ts <= e->Tg, "dp_cost_coll: Ts=%d exceeds the %d time slices of the environment", Ts, e->Tg);
    DP_REQUIRE((gx == nullptr) == (gy == nullptr), "dp_cost_coll: gx and gy must be given together");
    DP_REQUIRE(!(grho && !gx), "dp_cost_coll: grho needs gx/gy");
    cudaStream_t stream = (cudaStream_t)stream_;
    DeviceGuard guard(e->device);
    if (N <= 0 || Ts == 0) {
        if (cost_slices && Ts > 0) DP_CUDA(cudaMemsetAsync(cost_slices, 0, sizeof(double) * Ts, stream));
        if (total) DP_CUDA(cudaMemsetAsync(total, 0, sizeof(double), stream));
        return 0;
    }
    CostArgs a;
    a.table = e->d_table;
    a.q = make_geom(e->geom);
    a.x = x; a.y = y; a.rho = rho;
    a.sx_n = sx_n; a.sx_t = sx_t; a.sr_n = sr_n; a.sr_t = sr_t; a.sg_n = sg_n; a.sg_t = sg_t;
    a.N = N; a.Ts = Ts; a.get_max = get_max;
    a.gx = gx; a.gy = gy; a.grho = grho;
    a.partial = env_ws(e)->partial;
    a.partial_max = env_ws(e)->partial_max;
    a.cost_samples = cost_samples;
    {
        static int s_pf = -1;  // DP_COST_PREFETCH: L2 prefetch distance in grid-stride iterations (A/B knob, default below)
        if (s_pf < 0) {
            const char *pe = getenv("DP_COST_PREFETCH");
            s_pf = (pe && atoi(pe) >= 0 && atoi(pe) <= 8) ? atoi(pe) : 1;
        }
        a.prefetch = s_pf;
    }
    const bool nfast = (sx_n == 1) && (!rho || sr_n == 1) && (!gx || sg_n == 1) && !cost_samples;
    if (nfast) {
        long long B = (N + CB * 4 - 1) / (CB * 4);
        const long long target = (148LL * 16 + Ts - 1) / Ts;  // ~16 blocks per SM over all slices (two waves of 8)
        if (B > target) B = target;
        if (B * Ts > MAX_PARTIALS) B = MAX_PARTIALS / Ts;
        if (B < 1) B = 1;
        a.B = (int)B;
        DP_REQUIRE((long long)a.B * Ts <= MAX_PARTIALS, "dp_cost_coll: too many slices");
        auto al16 = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
        const bool vec = al16(x) && al16(y) && al16(rho) && al16(gx) && al16(gy) && al16(grho) && (sx_t % 4 == 0) &&
                         (!rho || sr_t % 4 == 0) && (!gx || sg_t % 4 == 0);
        if (vec) {
            // two full waves of resident blocks over all slices (the kernel strides over the samples of its slice)
            static int s_minb = 0, s_waves = 0;  // A/B knobs: resident blocks per SM of the forward kernel, waves over all slices
            if (!s_minb) {
                const char *me = getenv("DP_COST_FWD_MINB");
                s_minb = (me && (atoi(me) == 4 || atoi(me) == 6 || atoi(me) == 8)) ? atoi(me) : 5;
                const char *we = getenv("DP_COST_WAVES");
                s_waves = (we && atoi(we) >= 1 && atoi(we) <= 16) ? atoi(we) : 4;
            }
            // Measured at the config-3 size (profiles/r02_cost_knobs.txt): four resident blocks per SM (63 registers, four
            // evaluations in flight per thread) and just under four FULL waves of blocks over all slices -- 23 blocks per
            // slice = 3.92 waves: 0.274 / 0.404 ms (forward / forward + backward); 24 per slice = 4.09 waves, i.e. a fifth,
            // almost empty wave: 0.280 / 0.427 ms; two waves 0.297 / 0.427 ms; six resident blocks (39 registers) 0.294 ms.
            const int minb = gx ? 4 : s_minb;
            long long Bv = ((long long)s_waves * minb * 148) / Ts;
            const long long Bmax = (N + CB * 4 - 1) / (CB * 4);  // at least one vector iteration per thread
            if (Bv > Bmax) Bv = Bmax;
            if (Bv * Ts > MAX_PARTIALS) Bv = MAX_PARTIALS / Ts;
            if (Bv < 1) Bv = 1;
            a.B = (int)Bv;
            const dim3 grid((unsigned)Bv, (unsigned)Ts);
            if (gx)
                cost_vec4_kernel<true, 4><<<grid, CB, 0, stream>>>(a);
            else if (minb == 4)
                cost_vec4_kernel<false, 4><<<grid, CB, 0, stream>>>(a);
            else if (minb == 8)
                cost_vec4_kernel<false, 8><<<grid, CB, 0, stream>>>(a);
            else if (minb == 5)  // forward default: 48 registers, 0.219 ms against 0.226 ms at four blocks (profiles/r02_cost_knobs.txt)
                cost_vec4_kernel<false, 5><<<grid, CB, 0, stream>>>(a);
            else
                cost_vec4_kernel<false, 6><<<grid, CB, 0, stream>>>(a);
        }
        else
            cost_nfast_kernel<<<dim3((unsigned)B, (unsigned)Ts), CB, 0, stream>>>(a);
    } else {
        const long long B = (N + 127) / 128;
        DP_REQUIRE(B * Ts <= MAX_PARTIALS,
                   "dp_cost_coll: strided/per-sample path supports N*Ts/128 <= %d (got N=%lld, Ts=%d); use the "
                   "sample-contiguous [t][n] layout for large sample sets", MAX_PARTIALS, (long long)N, Ts);
        a.B = (int)B;
        cost_generic_kernel<<<(unsigned)B, 128, 0, stream>>>(a);
    }
    DP_CUDA(cudaGetLastError());
    if (cost_slices || total) {
        cost_finalize_kernel<<<1, 128, 0, stream>>>(a.partial, a.partial_max, a.B, Ts, get_max, cost_slices, total);
        DP_CUDA(cudaGetLastError());
    }
    return 0;
}

// ---- pos2gridpos / gridpos2pos as standalone ops -------------------------------------------------------------------
namespace {
__global__ void pos2gridpos_kernel(Geom q, const float *px, const float *py, long long n, int *ix, int *iy, int chx, int chy) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (px) {
            int v = dp_cell(px[i], q.x_lo, q.x_rng, q.r_x_rng, q.gxm1);
            if (chx >= 0) v = min(max(v, 0), chx);
            ix[i] = v;
        }
        if (py) {
            int v = dp_cell(py[i], q.y_lo, q.y_rng, q.r_y_rng, q.gym1);
            if (chy >= 0) v = min(max(v, 0), chy);
            iy[i] = v;
        }
    }
}
// float64 inputs (python floats / lists / numpy float64 arrays reach the reference's pos2gridpos as float64 and are
// evaluated in double precision there, motion_planning/utils.py:79-85): the same ops in fp64, round half to even
__global__ void pos2gridpos_f64_kernel(double lo_x, double rng_x, double gxm1, double lo_y, double rng_y, double gym1,
                                       const double *px, const double *py, long long n, long long *ix, long long *iy) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (px) ix[i] = __double2ll_rn(__dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(px[i], lo_x), rng_x), gxm1), 0.001));
        if (py) iy[i] = __double2ll_rn(__dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(py[i], lo_y), rng_y), gym1), 0.001));
    }
}
__global__ void gridpos2pos_kernel(Geom q, const float *gx, const float *gy, long long n, float *px, float *py) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (gx) px[i] = dp_pos(gx[i], q.x_lo, q.x_rng, q.gxm1, q.r_gxm1);
        if (gy) py[i] = dp_pos(gy[i], q.y_lo, q.y_rng, q.gym1, q.r_gym1);
    }
}
}  // namespace

int dp_pos2gridpos(const dp_grid_geom *g, const float *px, const float *py, int64_t n, int32_t *ix, int32_t *iy,
                   int clamp_hi_x, int clamp_hi_y, void *stream) {
    DP_REQUIRE(g && (!px || ix) && (!py || iy), "dp_pos2gridpos: bad argument");
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    pos2gridpos_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(make_geom(*g), px, py, n, ix, iy, clamp_hi_x,
                                                                         clamp_hi_y);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_pos2gridpos_f64(const dp_grid_geom *g, const double *px, const double *py, int64_t n, int64_t *ix, int64_t *iy,
                       void *stream) {
    DP_REQUIRE(g && (!px || ix) && (!py || iy), "dp_pos2gridpos_f64: bad argument");
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    // the reference subtracts / divides by the python numbers of args.environment_size (ints or floats -> float64)
    pos2gridpos_f64_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (double)g->x_min, (double)g->x_max - (double)g->x_min, (double)(g->gx - 1), (double)g->y_min,
        (double)g->y_max - (double)g->y_min, (double)(g->gy - 1), px, py, n, reinterpret_cast<long long *>(ix),
        reinterpret_cast<long long *>(iy));
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_gridpos2pos(const dp_grid_geom *g, const float *gx, const float *gy, int64_t n, float *px, float *py, void *stream) {
    DP_REQUIRE(g && (!gx || px) && (!gy || py), "dp_gridpos2pos: bad argument");
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    gridpos2pos_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(make_geom(*g), gx, gy, n, px, py);
    DP_CUDA(cudaGetLastError());
    return 0;
}
