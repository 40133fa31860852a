// K4: the density-predictor MLP of the planner's density phase as ONE fused tcgen05 kernel per direction (sm_100a).
//
// Reference: density_training/utils.py:8-40 (NeuralNetwork, MLP branch: Linear(31,150), 6 x Linear(150,150), Linear(150,6),
// ReLU) evaluated by get_nn_prediction (:161-180), which EgoVehicle.predict_density calls once per prediction time point
// (motion_planning/simulation_objects.py:280-287: 101 calls of a (N, 31) batch).  Here all time points x samples are the rows
// of one launch, and the whole layer stack runs per 128-row tile without the activations ever leaving the SM:
//
//   * layer widths are padded to 160 (a real dense contraction: M = 128 rows, N = 160, K = 160 per layer); the weights of a
//     layer are two fp16 images (hi, lo of w * 16) in the canonical no-swizzle K-major UMMA layout, 100 KB per layer, streamed
//     through a DOUBLE-BUFFERED shared-memory stage by the bulk-copy engine (cp.async.bulk + mbarrier expect_tx): the copy of
//     layer l + 1 is issued together with the MMAs of layer l and lands while the epilogue of layer l runs.
//   * activations live in TMEM: the A operand (fp16 hi / lo pairs, 160 columns) is written with tcgen05.st by the row's own
//     thread, the fp32 accumulator (160 columns) is drained with tcgen05.ld.  Split precision like the rollout kernel:
//     a.w ~= a_hi w_hi + a_hi w_lo + a_lo w_hi (three tcgen05.mma per K step, ~22 mantissa bits) -- the forward pass has to
//     match the reference's fp32 torch.nn.Linear stack to 1e-4 absolute on positions of order 10.
//   * epilogue per layer and thread (thread = row = TMEM lane): scale, bias, ReLU, fp16 hi / lo split (mixed-precision add),
//     operand store; the ReLU sign bits (160 per layer and row) go to global memory for the backward pass.
//   * backward (gradient w.r.t. the INPUT rows only -- the planner differentiates w.r.t. the input parameters `up`, the
//     weights are frozen): the same kernel body over the transposed images in reverse layer order, the epilogue multiplies by
//     the stored ReLU mask instead of adding a bias.
// One CTA per SM (512 TMEM columns, 205 KB of shared memory), 128 threads, persistent over the row tiles.
#include <vector>

#include "dp_common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tc;

constexpr int DW = 160;                        // padded layer width
constexpr int D_IMG = DW * DW * 2;             // one fp16 image
constexpr int D_LAYER = 2 * D_IMG;             // hi + lo
constexpr uint32_t D_LBO = 128;                // K-adjacent core matrices (8 rows x 16 bytes) are contiguous
constexpr uint32_t D_SBO = (DW / 8) * 128;     // next 8-row group
constexpr int D_MAX_LAYERS = 8;
constexpr int D_MASK_WORDS = DW / 32;          // ReLU: sign bits per layer and row
constexpr int D_TANH_WORDS = DW / 2;           // tanh: the activations themselves as fp16 pairs (the gate is 1 - h^2)
constexpr int D_OFF_BIAS = 2 * D_LAYER;        // float [8][160]
constexpr int D_OFF_BAR = D_OFF_BIAS + D_MAX_LAYERS * DW * 4;
constexpr int D_SMEM = D_OFF_BAR + 64;
static_assert(D_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t T_OPER = 0, T_ACC = 256;    // TMEM columns of the operand and the accumulator

struct DnnParams {
    const unsigned char *img;  // layer images in pass order, D_LAYER bytes each
    const float *bias;         // [n_layers][DW], forward only
    const float *in;           // [rows][in_w]
    float *out;                // [rows][out_w]
    uint32_t *masks;           // [rows][n_layers - 1][D_MASK_WORDS or D_TANH_WORDS]: written by forward, read by backward
    long long rows;
    int n_layers, in_w, out_w;
    int nk[D_MAX_LAYERS];      // K steps (of 16) per pass
};

__device__ __forceinline__ float2 f2_of(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

template <bool BWD, bool TANH>
__global__ void __launch_bounds__(128, 1) dnn_kernel(DnnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + D_OFF_BAR);  // [0], [1]: weight buffers, [2]: MMA batch
    uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(smem + D_OFF_BAR + 32);
    float *sBias = reinterpret_cast<float *>(smem + D_OFF_BIAS);
    if (tid == 0) {
        mbar_init(smem_u32(bars), 1);
        mbar_init(smem_u32(bars + 1), 1);
        mbar_init(smem_u32(bars + 2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!BWD)
        for (int i = tid; i < p.n_layers * DW; i += 128) sBias[i] = p.bias[i];
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *tmem_ptr_s;
    const uint32_t lane_off = (uint32_t)(32 * warp) << 16;
    const uint32_t OPER = tm + T_OPER, ACC = tm + T_ACC;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t mma_bar = smem_u32(bars + 2);
    constexpr uint32_t idesc = make_idesc(DW);
    const long long ntiles = (p.rows + 127) / 128;
    const int n_hidden = p.n_layers - 1;

    // weights of global pass number gi (buffer gi & 1): one thread arms the buffer's mbarrier with the byte count and issues
    // the bulk copies; the data arrives through the async proxy, which is also the proxy tcgen05.mma reads it through
    auto load_weights = [&](uint32_t gi, int pass) {
        const uint32_t bar = smem_u32(bars + (gi & 1));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)D_LAYER) : "memory");
        constexpr int PIECE = 25600;  // four pieces of 25,600 bytes
        const unsigned char *src = p.img + (size_t)pass * D_LAYER;
        for (int off = 0; off < D_LAYER; off += PIECE)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             sbase + (gi & 1) * D_LAYER + off),
                         "l"(src + off), "r"(PIECE), "r"(bar)
                         : "memory");
    };
    if (blockIdx.x < ntiles && tid == 0) load_weights(0, 0);

    uint32_t g = 0;    // global pass counter of this CTA: weight buffer g & 1, its barrier has completed (g >> 1) phases before
    uint32_t mph = 0;  // phase of the MMA barrier
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row = tile * 128 + tid;
        const bool valid = row < p.rows;
        // operand of the first pass: the row's inputs, zero padded to whole K steps
        for (int c = 0; c < p.nk[0]; ++c) {
            uint32_t av[16];
#pragma unroll
            for (int pq = 0; pq < 8; ++pq) {
                const int j = 16 * c + 2 * pq;
                float2 v;
                v.x = (valid && j < p.in_w) ? __ldg(p.in + row * p.in_w + j) : 0.0f;
                v.y = (valid && j + 1 < p.in_w) ? __ldg(p.in + row * p.in_w + j + 1) : 0.0f;
                split_h2_mixed(v, av[pq], av[8 + pq]);
            }
            st16(OPER + lane_off + 16 * c, av);
        }
        for (int pass = 0; pass < p.n_layers; ++pass, ++g) {
            // every thread's operand columns are in TMEM and ordered before the barrier
            wait_st();
            fence_before();
            __syncthreads();
            if (warp == 0) {
                mbar_wait(smem_u32(bars + (g & 1)), (g >> 1) & 1u);  // this pass's weights have landed
                fence_after();
                if (elect_one()) {
                    const uint64_t dh = make_bdesc(sbase + (g & 1) * D_LAYER, D_LBO, D_SBO);
                    const uint64_t dl = make_bdesc(sbase + (g & 1) * D_LAYER + D_IMG, D_LBO, D_SBO);
                    for (int kc = 0; kc < p.nk[pass]; ++kc) {
                        const uint64_t bh = dh + kc * B_KSTEP_DESC, bl = dl + kc * B_KSTEP_DESC;
                        mma_ts(ACC, OPER + 16 * kc, bh, idesc, kc > 0 ? 1u : 0u);   // a_hi w_hi
                        mma_ts(ACC, OPER + 16 * kc, bl, idesc, 1u);                 // a_hi w_lo
                        mma_ts(ACC, OPER + 16 * kc + 8, bh, idesc, 1u);             // a_lo w_hi
                    }
                    commit(mma_bar);
                    // next pass's weights into the other buffer: its last reader, the MMA batch of pass g - 1, completed
                    // before that pass's epilogue started
                    if (pass + 1 < p.n_layers)
                        load_weights(g + 1, pass + 1);
                    else if (tile + gridDim.x < ntiles)
                        load_weights(g + 1, 0);
                }
                __syncwarp();
            }
            mbar_wait(mma_bar, mph);
            mph ^= 1u;
            fence_after();
            const bool last = pass + 1 == p.n_layers;
            const float2 inv = make_float2(INV_WSCALE, INV_WSCALE);
            if (!last && !TANH) {
                // hidden layer: forward = scale, bias, ReLU (sign bits kept), backward = scale, ReLU mask of the layer below;
                // either way the result becomes the next pass's operand
                const int hid = BWD ? n_hidden - 1 - pass : pass;
                uint32_t mw[D_MASK_WORDS];
#pragma unroll
                for (int w = 0; w < D_MASK_WORDS; ++w)
                    mw[w] = BWD ? ((valid && p.masks) ? __ldg(p.masks + (row * n_hidden + hid) * D_MASK_WORDS + w) : 0u) : 0u;
#pragma unroll
                for (int c = 0; c < DW / 16; ++c) {
                    uint32_t d[16], av[16];
                    ld16(ACC + lane_off + 16 * c, d);
                    wait_ld();
                    uint32_t bits = BWD ? (mw[c >> 1] >> (16 * (c & 1))) & 0xFFFFu : 0u;
#pragma unroll
                    for (int pq = 0; pq < 8; ++pq) {
                        float2 y;
                        if (!BWD) {
                            const float2 b = *reinterpret_cast<const float2 *>(sBias + pass * DW + 16 * c + 2 * pq);
                            y = __ffma2_rn(f2_of(d[2 * pq], d[2 * pq + 1]), inv, b);
                            bits |= (y.x > 0.0f ? 1u : 0u) << (2 * pq);
                            bits |= (y.y > 0.0f ? 1u : 0u) << (2 * pq + 1);
                            y.x = fmaxf(y.x, 0.0f);
                            y.y = fmaxf(y.y, 0.0f);
                        } else {
                            y = __fmul2_rn(f2_of(d[2 * pq], d[2 * pq + 1]), inv);
                            y.x = ((bits >> (2 * pq)) & 1u) ? y.x : 0.0f;
                            y.y = ((bits >> (2 * pq + 1)) & 1u) ? y.y : 0.0f;
                        }
                        split_h2_mixed(y, av[pq], av[8 + pq]);
                    }
                    if (!BWD) mw[c >> 1] |= bits << (16 * (c & 1));
                    st16(OPER + lane_off + 16 * c, av);
                }
                if (!BWD && valid && p.masks) {
#pragma unroll
                    for (int w = 0; w < D_MASK_WORDS; ++w) p.masks[(row * n_hidden + hid) * D_MASK_WORDS + w] = mw[w];
                }
            } else if (!last) {
                // tanh hidden layer (the reference's other activation, density_training/utils.py:16-17): forward keeps the
                // activations as fp16 pairs -- the hi half of the operand split -- and the backward pass forms the gate 1 - h^2
                // from them (11 bits on the gate: the gradient contract is 2e-3)
                const int hid = BWD ? n_hidden - 1 - pass : pass;
                uint32_t *hrow = p.masks ? p.masks + (row * n_hidden + hid) * D_TANH_WORDS : nullptr;
#pragma unroll
                for (int c = 0; c < DW / 16; ++c) {
                    uint32_t d[16], av[16], hw[8];
                    ld16(ACC + lane_off + 16 * c, d);
                    if (BWD) {
#pragma unroll
                        for (int pq = 0; pq < 8; ++pq) hw[pq] = (valid && hrow) ? __ldg(hrow + 8 * c + pq) : 0u;
                    }
                    wait_ld();
#pragma unroll
                    for (int pq = 0; pq < 8; ++pq) {
                        float2 y;
                        if (!BWD) {
                            const float2 b = *reinterpret_cast<const float2 *>(sBias + pass * DW + 16 * c + 2 * pq);
                            y = __ffma2_rn(f2_of(d[2 * pq], d[2 * pq + 1]), inv, b);
                            y.x = tanh_fast(y.x);
                            y.y = tanh_fast(y.y);
                        } else {
                            const float2 h = __half22float2(bits_h2(hw[pq]));
                            y = __fmul2_rn(f2_of(d[2 * pq], d[2 * pq + 1]), inv);
                            y.x *= fmaf(-h.x, h.x, 1.0f);
                            y.y *= fmaf(-h.y, h.y, 1.0f);
                        }
                        split_h2_mixed(y, av[pq], av[8 + pq]);
                    }
                    if (!BWD && valid && hrow) {
#pragma unroll
                        for (int pq = 0; pq < 8; ++pq) hrow[8 * c + pq] = av[pq];
                    }
                    st16(OPER + lane_off + 16 * c, av);
                }
            } else {
                // output layer: forward adds the bias, backward is the plain product; the first out_w columns go to global
                for (int c = 0; 16 * c < p.out_w; ++c) {
                    uint32_t d[16];
                    ld16(ACC + lane_off + 16 * c, d);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = 16 * c + j;
                        if (valid && col < p.out_w) {
                            float y = __uint_as_float(d[j]) * INV_WSCALE;
                            if (!BWD) y += sBias[pass * DW + col];
                            p.out[row * p.out_w + col] = y;
                        }
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

// canonical no-swizzle K-major image of a padded DW x DW matrix B[n][k]: core matrix = 8 rows x 8 fp16 (128 bytes), the
// DW / 8 core matrices of an 8-row group are contiguous along K
inline size_t dnn_canon(int n, int k) { return ((size_t)(n / 8) * (DW / 8) + (k / 8)) * 64 + (n % 8) * 8 + (k % 8); }

}  // namespace

struct dp_dnn {
    int device = 0, num_sms = 0;
    int n_layers = 0;
    int tanh_act = 0;  // 0: ReLU, 1: tanh
    int dims[D_MAX_LAYERS + 1] = {};
    unsigned char *d_fwd = nullptr;  // [n_layers] images B[n = output feature][k = input feature] of layer l
    unsigned char *d_bwd = nullptr;  // [n_layers] images B[n = input feature][k = output feature] of layer n_layers-1-pass
    float *d_bias = nullptr;         // [n_layers][DW]
};

extern "C" {

int dp_dnn_create(const float *const *weights, const float *const *biases, const int *dims, int n_layers, int activation,
                  int device, dp_dnn **out) {
    DP_REQUIRE(weights && biases && dims && out, "dp_dnn_create: null argument");
    DP_REQUIRE(activation == 0 || activation == 1, "dp_dnn_create: activation is 0 (ReLU) or 1 (tanh)");
    DP_REQUIRE(n_layers >= 1 && n_layers <= D_MAX_LAYERS, "dp_dnn_create: 1..%d layers supported (got %d)", D_MAX_LAYERS, n_layers);
    for (int l = 0; l <= n_layers; ++l)
        DP_REQUIRE(dims[l] >= 1 && dims[l] <= DW, "dp_dnn_create: layer widths up to %d supported (dims[%d] = %d)", DW, l, dims[l]);
    int ok = dp_device_ok(device);
    DP_REQUIRE(ok == 1, "dp_dnn_create: device %d is not an sm_100 GPU", device);
    DeviceGuard guard(device);
    std::vector<unsigned char> fwd((size_t)n_layers * D_LAYER, 0), bwd((size_t)n_layers * D_LAYER, 0);
    std::vector<float> bias((size_t)n_layers * DW, 0.0f);
    for (int l = 0; l < n_layers; ++l) {
        const int nin = dims[l], nout = dims[l + 1];
        __half *fh = reinterpret_cast<__half *>(fwd.data() + (size_t)l * D_LAYER), *fl = fh + DW * DW;
        __half *bh = reinterpret_cast<__half *>(bwd.data() + (size_t)(n_layers - 1 - l) * D_LAYER), *bl = bh + DW * DW;
        for (int o = 0; o < nout; ++o) {
            for (int i = 0; i < nin; ++i) {
                const float w = weights[l][(size_t)o * nin + i] * WSCALE;  // torch Linear.weight layout [out][in]
                const __half h = __float2half_rn(w);
                const __half lo = __float2half_rn(w - __half2float(h));
                fh[dnn_canon(o, i)] = h;
                fl[dnn_canon(o, i)] = lo;
                bh[dnn_canon(i, o)] = h;
                bl[dnn_canon(i, o)] = lo;
            }
            bias[(size_t)l * DW + o] = biases[l][o];
        }
    }
    dp_dnn *h = new dp_dnn();
    h->device = device;
    h->n_layers = n_layers;
    h->tanh_act = activation;
    for (int l = 0; l <= n_layers; ++l) h->dims[l] = dims[l];
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaMalloc(&h->d_fwd, fwd.size()) || cudaMalloc(&h->d_bwd, bwd.size()) ||
        cudaMalloc(&h->d_bias, bias.size() * 4) || cudaMemcpy(h->d_fwd, fwd.data(), fwd.size(), cudaMemcpyHostToDevice) ||
        cudaMemcpy(h->d_bwd, bwd.data(), bwd.size(), cudaMemcpyHostToDevice) ||
        cudaMemcpy(h->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice)) {
        dp_set_error("dp_dnn_create: %s", cudaGetErrorString(cudaGetLastError()));
        dp_dnn_destroy(h);
        return 1;
    }
    h->num_sms = prop.multiProcessorCount;
    DP_CUDA(cudaFuncSetAttribute(dnn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    DP_CUDA(cudaFuncSetAttribute(dnn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    DP_CUDA(cudaFuncSetAttribute(dnn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    DP_CUDA(cudaFuncSetAttribute(dnn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    *out = h;
    return 0;
}

void dp_dnn_destroy(dp_dnn *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaFree(h->d_fwd);
    cudaFree(h->d_bwd);
    cudaFree(h->d_bias);
    delete h;
}

int dp_dnn_mask_words(const dp_dnn *h, int *words_per_row) {
    DP_REQUIRE(h && words_per_row, "dp_dnn_mask_words: null argument");
    *words_per_row = (h->n_layers - 1) * (h->tanh_act ? D_TANH_WORDS : D_MASK_WORDS);
    return 0;
}

int dp_dnn_forward(const dp_dnn *h, const float *x, int64_t rows, float *y, uint32_t *masks, void *stream) {
    DP_REQUIRE(h && x && y, "dp_dnn_forward: null argument");
    if (rows <= 0) return 0;
    DeviceGuard guard(h->device);
    DnnParams p;
    memset(&p, 0, sizeof(p));
    p.img = h->d_fwd; p.bias = h->d_bias; p.in = x; p.out = y; p.masks = masks; p.rows = rows;
    p.n_layers = h->n_layers; p.in_w = h->dims[0]; p.out_w = h->dims[h->n_layers];
    for (int l = 0; l < h->n_layers; ++l) p.nk[l] = (h->dims[l] + 15) / 16;
    const long long ntiles = (rows + 127) / 128;
    const int grid = (int)(ntiles < h->num_sms ? ntiles : h->num_sms);
    if (h->tanh_act)
        dnn_kernel<false, true><<<grid, 128, D_SMEM, (cudaStream_t)stream>>>(p);
    else
        dnn_kernel<false, false><<<grid, 128, D_SMEM, (cudaStream_t)stream>>>(p);
    DP_CUDA(cudaGetLastError());
    return 0;
}

int dp_dnn_backward(const dp_dnn *h, const float *gy, const uint32_t *masks, int64_t rows, float *gx, void *stream) {
    DP_REQUIRE(h && gy && gx && (masks || h->n_layers == 1), "dp_dnn_backward: null argument");
    if (rows <= 0) return 0;
    DeviceGuard guard(h->device);
    DnnParams p;
    memset(&p, 0, sizeof(p));
    p.img = h->d_bwd; p.bias = nullptr; p.in = gy; p.out = gx; p.masks = const_cast<uint32_t *>(masks); p.rows = rows;
    p.n_layers = h->n_layers; p.in_w = h->dims[h->n_layers]; p.out_w = h->dims[0];
    for (int q = 0; q < h->n_layers; ++q) p.nk[q] = (h->dims[h->n_layers - q] + 15) / 16;  // pass q runs layer n_layers-1-q transposed
    const long long ntiles = (rows + 127) / 128;
    const int grid = (int)(ntiles < h->num_sms ? ntiles : h->num_sms);
    if (h->tanh_act)
        dnn_kernel<true, true><<<grid, 128, D_SMEM, (cudaStream_t)stream>>>(p);
    else
        dnn_kernel<true, false><<<grid, 128, D_SMEM, (cudaStream_t)stream>>>(p);
    DP_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
