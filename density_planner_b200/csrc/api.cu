// C ABI glue: error state, controller objects, rollout entry points (device and host buffers).
#include <stdarg.h>

#include <vector>

#include "dp_common.cuh"

static thread_local char g_err[1024] = "";

void dp_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

int dp_abi_version(void) { return DP_ABI_VERSION; }
const char *dp_last_error(void) { return g_err; }

int dp_device_ok(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        dp_set_error("cudaGetDeviceProperties(%d) failed", device);
        return -1;
    }
    return prop.major == 10 ? 1 : 0;
}

int dp_controller_create(const float *blob, size_t n_floats, int device, dp_controller **out) {
    DP_REQUIRE(blob && out, "dp_controller_create: null argument");
    DP_REQUIRE(n_floats == DP_CONTROLLER_BLOB_FLOATS, "dp_controller_create: blob has %zu floats, expected %d", n_floats,
               DP_CONTROLLER_BLOB_FLOATS);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    DP_CUDA(cudaGetDeviceProperties(&prop, device));
    DP_REQUIRE(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback",
               device, prop.major, prop.minor);
    std::vector<float> big(DP_BIG_FLOATS), small(DP_SMALL_FLOATS);
    const float *p = blob;
    const int nouts[2] = {DP_NOUT_A, DP_NOUT_B};
    const int big_w2[2] = {DP_BIG_W2A, DP_BIG_W2B}, big_w3[2] = {DP_BIG_W3A, DP_BIG_W3B};
    const int small_net[2] = {DP_SM_NET_A, DP_SM_NET_B};
    for (int m = 0; m < 2; ++m) {
        const int no = nouts[m];
        float *sm = small.data() + small_net[m];
        memcpy(sm + DP_SM_W1, p, sizeof(float) * DP_HID * 4); p += DP_HID * 4;
        memcpy(sm + DP_SM_B1, p, sizeof(float) * DP_HID); p += DP_HID;
        const float *W2 = p; p += DP_HID * DP_HID;
        memcpy(sm + DP_SM_B2, p, sizeof(float) * DP_HID); p += DP_HID;
        const float *W3 = p; p += no * DP_HID;
        memcpy(sm + DP_SM_B3, p, sizeof(float) * no); p += no;
        for (int j = 0; j < DP_HID; ++j)
            for (int k = 0; k < DP_HID; ++k) big[big_w2[m] + k * DP_HID + j] = W2[j * DP_HID + k];
        for (int j = 0; j < no; ++j)
            for (int k = 0; k < DP_HID; ++k) big[big_w3[m] + k * no + j] = W3[j * DP_HID + k];
    }
    dp_controller *c = new dp_controller();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    DP_CUDA(cudaMalloc(&c->d_big, sizeof(float) * DP_BIG_FLOATS));
    DP_CUDA(cudaMalloc(&c->d_small, sizeof(float) * DP_SMALL_FLOATS));
    DP_CUDA(cudaMemcpy(c->d_big, big.data(), sizeof(float) * DP_BIG_FLOATS, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(c->d_small, small.data(), sizeof(float) * DP_SMALL_FLOATS, cudaMemcpyHostToDevice));
    DP_CUDA(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
    if (dp_tc_prepare(c, blob) != 0) {
        dp_controller_destroy(c);
        return 3;
    }
    *out = c;
    return 0;
}

void dp_controller_destroy(dp_controller *c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    dp_tc_release(c);
    cudaFree(c->d_big);
    cudaFree(c->d_small);
    cudaFree(c->ws.ptr);
    if (c->host_stream) cudaStreamDestroy(c->host_stream);
    if (c->copy_stream) {
        cudaStreamDestroy(c->copy_stream);
        for (int b = 0; b < 2; ++b) {
            cudaEventDestroy(c->ev_done[b]);
            cudaEventDestroy(c->ev_copied[b]);
        }
    }
    delete c;
}

int dp_rollout(const dp_controller *c, const float *xe0, const float *xref, const float *uref, const float *rho0, float dt,
               int64_t N, int T, unsigned flags, int stride, float *x_out, float *rho_out, int64_t ldn, float *clip_margin,
               int *status, void *stream) {
    DP_REQUIRE(c && xe0 && xref && uref, "dp_rollout: null argument");
    DP_REQUIRE(T >= 0 && stride >= 1, "dp_rollout: bad T=%d / stride=%d", T, stride);
    DP_REQUIRE(ldn >= N || (!x_out && !rho_out), "dp_rollout: ldn=%lld < N=%lld", (long long)ldn, (long long)N);
    if (N <= 0) return 0;
    DeviceGuard guard(c->device);
    RolloutParams p;
    p.big = c->d_big; p.small = c->d_small; p.tc = c->d_tc;
    p.xe0 = xe0; p.xref = xref; p.uref = uref; p.rho0 = rho0;
    p.dt = dt; p.N = N; p.T = T; p.flags = flags; p.stride = stride; p.Ts = T / stride + 1;
    p.x_out = x_out; p.rho_out = rho_out; p.ldn = ldn; p.clip_margin = clip_margin; p.status = status; p.trace = nullptr; p.R = 1; p.n_per_ref = N; p.T_ref = nullptr; p.lmax_enc = nullptr;
    return dp_launch_rollout_tc(c, p, (cudaStream_t)stream);  // the tcgen05 kernel: the one rollout backend of this library
}

int dp_rollout_batch(const dp_controller *c, const float *xe0, const float *xref, const float *uref, const int *T_ref,
                     const float *rho0, float dt, int64_t R, int64_t n_per_ref, int T, unsigned flags, int stride, float *x_out,
                     float *rho_out, int64_t ldn, float *clip_margin, int *status, void *stream) {
    DP_REQUIRE(c && xe0 && xref && uref, "dp_rollout_batch: null argument");
    DP_REQUIRE(T >= 0 && stride >= 1 && R >= 0 && n_per_ref >= 0, "dp_rollout_batch: bad sizes");
    const int64_t N = R * n_per_ref;
    DP_REQUIRE(ldn >= N || (!x_out && !rho_out), "dp_rollout_batch: ldn=%lld < N=%lld", (long long)ldn, (long long)N);
    if (N <= 0) return 0;
    DeviceGuard guard(c->device);
    RolloutParams p;
    p.big = c->d_big; p.small = c->d_small; p.tc = c->d_tc;
    p.xe0 = xe0; p.xref = xref; p.uref = uref; p.rho0 = rho0;
    p.dt = dt; p.N = N; p.T = T; p.flags = flags; p.stride = stride; p.Ts = T / stride + 1;
    p.x_out = x_out; p.rho_out = rho_out; p.ldn = ldn; p.clip_margin = clip_margin; p.status = status; p.trace = nullptr;
    p.R = R; p.n_per_ref = n_per_ref; p.T_ref = T_ref; p.lmax_enc = nullptr;
    return dp_launch_rollout_tc(c, p, (cudaStream_t)stream);
}

static int ws_reserve(dp_controller *c, size_t bytes) {
    if (c->ws.bytes >= bytes) return 0;
    if (c->ws.ptr) cudaFree(c->ws.ptr);
    c->ws.ptr = nullptr;
    c->ws.bytes = 0;
    DP_CUDA(cudaMalloc(&c->ws.ptr, bytes));
    c->ws.bytes = bytes;
    return 0;
}

// Chunk sizes of dp_rollout_host (pure host logic, no CUDA call): one wave first (the copy engine starts early), then two
// waves per chunk, and what is left split so that the last chunk -- the only copy nothing overlaps -- is at most one wave.
int dp_host_chunk_schedule(int64_t N, int64_t wave, int64_t *sizes, int max_chunks) {
    if (N <= 0 || wave <= 0 || !sizes) return 0;
    int n = 0;
    int64_t rem = N;
    auto push = [&](int64_t v) {
        if (n < max_chunks) sizes[n] = v;
        ++n;
        rem -= v;
    };
    push(wave < rem ? wave : rem);
    while (rem > 2 * wave) push(2 * wave);
    if (rem > wave) push(wave);
    if (rem > 0) push(rem);
    return n <= max_chunks ? n : -n;
}

int dp_rollout_host(dp_controller *c, const float *xe0, const float *xref, const float *uref, const float *rho0, float dt,
                    int64_t N, int T, unsigned flags, int stride, float *x_out, float *rho_out, float *clip_margin,
                    int *status) {
    DP_REQUIRE(c && xe0 && xref && uref, "dp_rollout_host: null argument");
    DP_REQUIRE(T >= 0 && stride >= 1, "dp_rollout_host: bad T / stride");
    if (N <= 0) return 0;
    std::lock_guard<std::mutex> lock(c->host_mutex);
    DeviceGuard guard(c->device);
    const int Ts = T / stride + 1;
    const int nd = (flags & DP_XY_ONLY) ? 2 : 5;
    // Samples are processed in chunks of two whole waves (one wave = 2 tiles of 128 samples per SM); the device->host copy
    // of chunk k overlaps the rollout of chunk k+1 (two output buffers, a compute stream and a copy stream).  Small chunks
    // keep the exposed tail (the last chunk's copy) short: ~14 chunks for 1M samples.
    const int64_t wave = 2 * 128LL * c->num_sms;
    int64_t chunk = 2 * wave;
    if (chunk > N) chunk = N;
    // the first chunk is a single wave, so that the copy engine starts after one wave instead of two, and what is left at
    // the end is split so that the last chunk -- whose copy nothing overlaps -- is at most one wave
    std::vector<int64_t> sizes((size_t)(N / (2 * wave) + 4));
    const int64_t nchunks = dp_host_chunk_schedule(N, wave, sizes.data(), (int)sizes.size());
    DP_REQUIRE(nchunks > 0, "dp_rollout_host: chunk schedule");
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_xe0 = al(sizeof(float) * N * 5), b_xref = al(sizeof(float) * 5 * (T + 1)),
                 b_uref = al(sizeof(float) * 2 * (T > 0 ? T : 1));
    const size_t b_rho0 = rho0 ? al(sizeof(float) * N) : 0;
    const size_t b_x = x_out ? al(sizeof(float) * (size_t)Ts * nd * chunk) : 0;
    const size_t b_rho = rho_out ? al(sizeof(float) * (size_t)Ts * chunk) : 0;
    const size_t b_m = clip_margin ? al(sizeof(float) * N) : 0;
    const size_t b_st = 256;
    if (ws_reserve(c, b_xe0 + b_xref + b_uref + b_rho0 + 2 * (b_x + b_rho) + b_m + b_st)) return 1;
    char *base = (char *)c->ws.ptr;
    float *d_xe0 = (float *)base; base += b_xe0;
    float *d_xref = (float *)base; base += b_xref;
    float *d_uref = (float *)base; base += b_uref;
    float *d_rho0 = rho0 ? (float *)base : nullptr; base += b_rho0;
    float *d_x[2] = {nullptr, nullptr}, *d_rho[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        d_x[b] = x_out ? (float *)base : nullptr; base += b_x;
        d_rho[b] = rho_out ? (float *)base : nullptr; base += b_rho;
    }
    float *d_m = clip_margin ? (float *)base : nullptr; base += b_m;
    int *d_st = (int *)base;
    if (!c->copy_stream) {
        DP_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            DP_CUDA(cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming));
            DP_CUDA(cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
        }
    }
    cudaStream_t st = c->host_stream, cp = c->copy_stream;
    DP_CUDA(cudaMemcpyAsync(d_xref, xref, sizeof(float) * 5 * (T + 1), cudaMemcpyHostToDevice, st));
    if (T > 0) DP_CUDA(cudaMemcpyAsync(d_uref, uref, sizeof(float) * 2 * T, cudaMemcpyHostToDevice, st));
    DP_CUDA(cudaMemsetAsync(d_st, 0, 2 * sizeof(int), st));
    int64_t off = 0;
    for (int64_t k = 0; k < nchunks; off += sizes[k], ++k) {
        const int64_t n = sizes[k];
        const int b = (int)(k & 1);
        DP_CUDA(cudaMemcpyAsync(d_xe0 + off * 5, xe0 + off * 5, sizeof(float) * n * 5, cudaMemcpyHostToDevice, st));
        if (rho0) DP_CUDA(cudaMemcpyAsync(d_rho0 + off, rho0 + off, sizeof(float) * n, cudaMemcpyHostToDevice, st));
        if (k >= 2) DP_CUDA(cudaStreamWaitEvent(st, c->ev_copied[b], 0));  // buffer b has been drained to the host
        int rc = dp_rollout(c, d_xe0 + off * 5, d_xref, d_uref, rho0 ? d_rho0 + off : nullptr, dt, n, T, flags, stride, d_x[b],
                            d_rho[b], chunk, d_m ? d_m + off : nullptr, d_st, st);
        if (rc) return rc;
        DP_CUDA(cudaEventRecord(c->ev_done[b], st));
        DP_CUDA(cudaStreamWaitEvent(cp, c->ev_done[b], 0));
        if (x_out)
            DP_CUDA(cudaMemcpy2DAsync(x_out + off, sizeof(float) * N, d_x[b], sizeof(float) * chunk, sizeof(float) * n,
                                      (size_t)Ts * nd, cudaMemcpyDeviceToHost, cp));
        if (rho_out)
            DP_CUDA(cudaMemcpy2DAsync(rho_out + off, sizeof(float) * N, d_rho[b], sizeof(float) * chunk, sizeof(float) * n,
                                      (size_t)Ts, cudaMemcpyDeviceToHost, cp));
        DP_CUDA(cudaEventRecord(c->ev_copied[b], cp));
    }
    int h_st[2] = {0, 0};
    if (clip_margin) DP_CUDA(cudaMemcpyAsync(clip_margin, d_m, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaMemcpyAsync(h_st, d_st, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    DP_CUDA(cudaStreamSynchronize(st));
    DP_CUDA(cudaStreamSynchronize(cp));
    if (status) { status[0] = h_st[0]; status[1] = h_st[1]; }
    return 0;
}

}  // extern "C"
