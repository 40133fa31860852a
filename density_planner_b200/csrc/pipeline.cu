// Device-resident pipeline: sampled initial states -> closed-loop rollout with Liouville log-density -> density weights ->
// time-sliced occupancy grids + collision cost.  Nothing but the inputs (24 B per sample) and the results (grids, Ts costs)
// crosses PCIe; the trajectories never leave the GPU.  Reference chain this replaces:
//   EgoVehicle.predict_density, LE branch + normaliser     motion_planning/simulation_objects.py:288-299
//   pred2grid (per time slice)                             motion_planning/utils.py:255-280
//   MotionPlanner.get_cost_coll                            motion_planning/MotionPlanner.py:192-224
//   check_collision (occupancy dot product)                motion_planning/utils.py:231-252
// Phases (a multi-GPU caller puts one collective between them, see density_planner_b200/pipeline.py):
//   1. dp_pipeline_rollout   the tcgen05 rollout stores the (x, y) slices and the log-density l on the device and reduces the
//                            per-slice max of l on the fly (atomicMax of an order-preserving integer image)
//      [allreduce MAX of lmax[Ts]]
//   2. dp_pipeline_bin       ONE pass over the stored slices: w = exp(l - lmax) * rho0 (the normaliser's numerator,
//                            simulation_objects.py:296-297), int64 fixed-point mass + int32 count per cell, and the
//                            collision cost of the weighted samples in int64 fixed point -- all integer sums, so the buffer
//                            is identical for any execution order and any sharding of the samples
//      [allreduce SUM of the int64 buffer]
//   3. dp_pipeline_finalize  per slice: sum of the weights (= sum of the mass grid: the normaliser's denominator), the
//                            normalised collision cost and the occupancy dot product of the normalised density
#include "dp_common.cuh"

struct dp_pipeline {
    const dp_controller *c = nullptr;
    const dp_env *env = nullptr;
    dp_bin_spec spec;
    int T = 0, stride = 1, Ts = 0;
    int64_t cap = 0, ldn = 0, cells = 0, N = 0;
    int cost_scale_log2 = 0;
    bool has_rho0 = false;
    float *d_xe0 = nullptr, *d_rho0 = nullptr, *d_xref = nullptr, *d_uref = nullptr;
    float *d_xy = nullptr;   // [Ts][2][ldn] absolute positions
    float *d_l = nullptr;    // [Ts][ldn] log-density
    unsigned *d_lmax_enc = nullptr;
    float *d_lmax = nullptr;       // [Ts]
    long long *d_ibuf = nullptr;   // dp_rollout_fused only
    double *d_out3 = nullptr;      // [3][Ts]
    cudaStream_t stream = nullptr; // dp_rollout_fused only
};

namespace {

__global__ void lmax_decode_kernel(const unsigned *enc, int Ts, float *lmax) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Ts) lmax[t] = enc[t] ? dp_decode_ordered(enc[t]) : -INFINITY;
}

// one block per slice; fixed order: thread-strided partials, shuffle tree, warp partials summed by thread 0
__global__ void __launch_bounds__(1024) pipeline_finalize_kernel(const long long *mass, const long long *cost_q, const float4 *table,
                                                                 int cx, int cy, int egx, int egy, double inv_mass_scale,
                                                                 double inv_cost_scale, int Ts, double *out3) {
    __shared__ long long shs[32];
    __shared__ double shd[32];
    const int t = blockIdx.x;
    const long long cells = (long long)cx * cy;
    const long long *m = mass + (long long)t * cells;
    const float4 *tab = table + (size_t)t * egx * egy;
    long long s = 0;
    double dot = 0.0;
    for (long long i = threadIdx.x; i < cells; i += blockDim.x) {
        const long long v = m[i];
        if (v) {
            s += v;
            const int ix = (int)(i / cy), iy = (int)(i - (long long)ix * cy);
            if (ix < egx && iy < egy) dot += (double)v * (double)__ldg(&tab[(size_t)ix * egy + iy].x);
        }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, k);
        dot += __shfl_xor_sync(0xffffffffu, dot, k);
    }
    if ((threadIdx.x & 31) == 0) {
        shs[threadIdx.x >> 5] = s;
        shd[threadIdx.x >> 5] = dot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long S = 0;
        double D = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            S += shs[w];
            D += shd[w];
        }
        const double wsum = (double)S * inv_mass_scale;
        out3[t] = wsum;
        out3[Ts + t] = S ? ((double)cost_q[t] * inv_cost_scale) / wsum : 0.0;
        out3[2 * Ts + t] = S ? D / (double)S : 0.0;
    }
}

}  // namespace

extern "C" {

int dp_pipeline_create(const dp_controller *c, const dp_env *env, const dp_bin_spec *spec, int T, int stride, int64_t capacity,
                       int cost_scale_log2, dp_pipeline **out) {
    DP_REQUIRE(c && env && spec && out, "dp_pipeline_create: null argument");
    DP_REQUIRE(spec->mode == 0, "dp_pipeline_create: the pipeline bins occupancy-grid cells (mode 0)");
    DP_REQUIRE(T >= 0 && stride >= 1 && capacity > 0, "dp_pipeline_create: bad T / stride / capacity");
    DP_REQUIRE(c->device == env->device, "dp_pipeline_create: controller and environment live on different devices");
    DP_REQUIRE(spec->geom.gx == env->geom.gx && spec->geom.gy == env->geom.gy, "dp_pipeline_create: grid geometry mismatch");
    DP_REQUIRE(T / stride + 1 <= env->Tg, "dp_pipeline_create: %d stored slices exceed the %d time slices of the environment",
               T / stride + 1, env->Tg);
    DP_REQUIRE(cost_scale_log2 >= 0 && cost_scale_log2 <= 40, "dp_pipeline_create: bad cost_scale_log2");
    DeviceGuard guard(c->device);
    dp_pipeline *p = new dp_pipeline();
    p->c = c; p->env = env; p->spec = *spec;
    p->T = T; p->stride = stride; p->Ts = T / stride + 1;
    p->cap = capacity;
    p->ldn = (capacity + 3) & ~(int64_t)3;
    p->cells = (int64_t)(spec->geom.gx + 1) * (spec->geom.gy + 1);
    p->cost_scale_log2 = cost_scale_log2;
    const size_t n = (size_t)p->ldn, Ts = (size_t)p->Ts;
    int64_t words = 0;
    dp_pipeline_words(p, &words);
    if (cudaMalloc(&p->d_xe0, n * 5 * 4) || cudaMalloc(&p->d_rho0, n * 4) || cudaMalloc(&p->d_xref, 5 * (size_t)(T + 1) * 4) ||
        cudaMalloc(&p->d_uref, 2 * (size_t)(T > 0 ? T : 1) * 4) || cudaMalloc(&p->d_xy, Ts * 2 * n * 4) ||
        cudaMalloc(&p->d_l, Ts * n * 4) || cudaMalloc(&p->d_lmax_enc, Ts * 4) || cudaMalloc(&p->d_lmax, Ts * 4) ||
        cudaMalloc(&p->d_ibuf, (size_t)words * 8) || cudaMalloc(&p->d_out3, 3 * Ts * 8) ||
        cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)) {
        dp_set_error("dp_pipeline_create: device allocation failed (%zu samples x %zu slices): %s", n, Ts,
                     cudaGetErrorString(cudaGetLastError()));
        dp_pipeline_destroy(p);
        return 1;
    }
    *out = p;
    return 0;
}

void dp_pipeline_destroy(dp_pipeline *p) {
    if (!p) return;
    DeviceGuard guard(p->c->device);
    cudaFree(p->d_xe0); cudaFree(p->d_rho0); cudaFree(p->d_xref); cudaFree(p->d_uref); cudaFree(p->d_xy); cudaFree(p->d_l);
    cudaFree(p->d_lmax_enc); cudaFree(p->d_lmax); cudaFree(p->d_ibuf); cudaFree(p->d_out3);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int dp_pipeline_words(const dp_pipeline *p, int64_t *words) {
    DP_REQUIRE(p && words, "dp_pipeline_words: null argument");
    const int64_t gc = (int64_t)p->Ts * p->cells;
    *words = gc + (gc + 1) / 2 + p->Ts;
    return 0;
}

int dp_pipeline_rollout(dp_pipeline *p, const float *xe0, const float *rho0, const float *xref, const float *uref,
                        int inputs_on_host, float dt, int64_t N, float *lmax_dev, void *stream_) {
    DP_REQUIRE(p && lmax_dev && (N <= 0 || (xe0 && xref && uref)), "dp_pipeline_rollout: null argument");  // an empty shard may pass null inputs
    DP_REQUIRE(N >= 0 && N <= p->cap, "dp_pipeline_rollout: N=%lld exceeds the capacity %lld", (long long)N, (long long)p->cap);
    cudaStream_t st = (cudaStream_t)stream_;
    DeviceGuard guard(p->c->device);
    const cudaMemcpyKind kind = inputs_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    p->N = N;
    p->has_rho0 = rho0 != nullptr;
    DP_CUDA(cudaMemsetAsync(p->d_lmax_enc, 0, sizeof(unsigned) * p->Ts, st));
    if (N > 0) {
        DP_CUDA(cudaMemcpyAsync(p->d_xe0, xe0, sizeof(float) * 5 * N, kind, st));
        if (rho0) DP_CUDA(cudaMemcpyAsync(p->d_rho0, rho0, sizeof(float) * N, kind, st));
        DP_CUDA(cudaMemcpyAsync(p->d_xref, xref, sizeof(float) * 5 * (p->T + 1), kind, st));
        if (p->T > 0) DP_CUDA(cudaMemcpyAsync(p->d_uref, uref, sizeof(float) * 2 * p->T, kind, st));
        RolloutParams q;
        memset(&q, 0, sizeof(q));
        q.big = p->c->d_big; q.small = p->c->d_small; q.tc = p->c->d_tc;
        q.xe0 = p->d_xe0; q.xref = p->d_xref; q.uref = p->d_uref; q.rho0 = nullptr;
        q.dt = dt; q.N = N; q.T = p->T; q.stride = p->stride; q.Ts = p->Ts;
        // predict_density calls compute_density with cutting=False, log_density=True (simulation_objects.py:288-291)
        q.flags = DP_LOG_DENSITY | DP_DENSITY | DP_ABS_STATE | DP_XY_ONLY;
        q.x_out = p->d_xy; q.rho_out = p->d_l; q.ldn = p->ldn;
        q.R = 1; q.n_per_ref = N;
        q.lmax_enc = p->d_lmax_enc;
        if (int rc = dp_launch_rollout_tc(p->c, q, st)) return rc;
    }
    lmax_decode_kernel<<<(p->Ts + 127) / 128, 128, 0, st>>>(p->d_lmax_enc, p->Ts, lmax_dev);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_pipeline_bin(dp_pipeline *p, const float *lmax_dev, int64_t *ibuf_dev, void *stream_) {
    DP_REQUIRE(p && lmax_dev && ibuf_dev, "dp_pipeline_bin: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    DeviceGuard guard(p->c->device);
    int64_t words = 0;
    dp_pipeline_words(p, &words);
    DP_CUDA(cudaMemsetAsync(ibuf_dev, 0, (size_t)words * 8, st));
    if (p->N <= 0) return 0;
    const int64_t gc = (int64_t)p->Ts * p->cells;
    long long *mass = reinterpret_cast<long long *>(ibuf_dev);
    int *count = reinterpret_cast<int *>(ibuf_dev + gc);
    long long *cost_q = reinterpret_cast<long long *>(ibuf_dev + gc + (gc + 1) / 2);
    return dp_bin2d_fused_launch(&p->spec, p->env, p->d_xy, p->d_xy + p->ldn, 2 * p->ldn, p->d_l, p->ldn,
                                 p->has_rho0 ? p->d_rho0 : nullptr, lmax_dev, p->N, p->Ts, mass, count, cost_q,
                                 p->cost_scale_log2, st);
}

int dp_pipeline_finalize(const dp_pipeline *p, const int64_t *ibuf_dev, double *out3_dev, void *stream_) {
    DP_REQUIRE(p && ibuf_dev && out3_dev, "dp_pipeline_finalize: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    DeviceGuard guard(p->c->device);
    const int64_t gc = (int64_t)p->Ts * p->cells;
    pipeline_finalize_kernel<<<p->Ts, 1024, 0, st>>>(reinterpret_cast<const long long *>(ibuf_dev),
                                                      reinterpret_cast<const long long *>(ibuf_dev + gc + (gc + 1) / 2),
                                                      p->env->d_table, p->spec.geom.gx + 1, p->spec.geom.gy + 1, p->env->geom.gx,
                                                      p->env->geom.gy, ldexp(1.0, -p->spec.mass_scale_log2),
                                                      ldexp(1.0, -p->cost_scale_log2), p->Ts, out3_dev);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_pipeline_copy_slices(const dp_pipeline *p, float *xy_dev, float *l_dev, void *stream_) {
    DP_REQUIRE(p && (xy_dev || l_dev), "dp_pipeline_copy_slices: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    DeviceGuard guard(p->c->device);
    if (p->N <= 0) return 0;
    const size_t row = sizeof(float) * (size_t)p->N;
    if (xy_dev)
        DP_CUDA(cudaMemcpy2DAsync(xy_dev, row, p->d_xy, sizeof(float) * (size_t)p->ldn, row, (size_t)p->Ts * 2,
                                  cudaMemcpyDeviceToDevice, st));
    if (l_dev)
        DP_CUDA(cudaMemcpy2DAsync(l_dev, row, p->d_l, sizeof(float) * (size_t)p->ldn, row, (size_t)p->Ts, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dp_rollout_fused(dp_pipeline *p, const float *xe0, const float *rho0, const float *xref, const float *uref, float dt,
                     int64_t N, int64_t *ibuf_host, double *out3_host) {
    DP_REQUIRE(p && out3_host, "dp_rollout_fused: null argument");
    DeviceGuard guard(p->c->device);
    cudaStream_t st = p->stream;
    if (int rc = dp_pipeline_rollout(p, xe0, rho0, xref, uref, 1, dt, N, p->d_lmax, st)) return rc;
    if (int rc = dp_pipeline_bin(p, p->d_lmax, reinterpret_cast<int64_t *>(p->d_ibuf), st)) return rc;
    if (int rc = dp_pipeline_finalize(p, reinterpret_cast<int64_t *>(p->d_ibuf), p->d_out3, st)) return rc;
    int64_t words = 0;
    dp_pipeline_words(p, &words);
    if (ibuf_host) DP_CUDA(cudaMemcpyAsync(ibuf_host, p->d_ibuf, (size_t)words * 8, cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaMemcpyAsync(out3_host, p->d_out3, sizeof(double) * 3 * p->Ts, cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
