// Diagnostics of the tcgen05 building blocks used by the rollout kernel (rollout_tc.cu): a self-test of the split-precision
// TMEM-operand MMA path and an issue/execute-rate probe of tcgen05.mma for the operand configurations the kernel could use.
#include <vector>

#include "dp_common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace tc;

constexpr int COL_AV = 0, COL_AT0 = 128, COL_D0 = 256;

// D (+)= A_V(hi,lo) x W(hi,lo): three products per K-chunk (operand chunk c = 8 columns of hi pairs then 8 of lo pairs)
__device__ __forceinline__ void mma_value(uint32_t tm, int col_d, uint64_t desc_hi, uint64_t desc_lo, uint32_t idesc) {
    const uint32_t d = tm + col_d, a0 = tm + COL_AV;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint64_t bh = desc_hi + c * B_KSTEP_DESC, bl = desc_lo + c * B_KSTEP_DESC;
        mma_ts(d, a0 + 16 * c, bh, idesc, c > 0);
        mma_ts(d, a0 + 16 * c, bl, idesc, 1);
        mma_ts(d, a0 + 16 * c + 8, bh, idesc, 1);
    }
}
// D (+)= A_T(hi) x W(hi,lo): two products per K-chunk
__device__ __forceinline__ void mma_tangent(uint32_t tm, int col_d, int col_a, uint64_t desc_hi, uint64_t desc_lo, uint32_t idesc) {
    const uint32_t d = tm + col_d, a0 = tm + col_a;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        mma_ts(d, a0 + 8 * c, desc_hi + c * B_KSTEP_DESC, idesc, c > 0);
        mma_ts(d, a0 + 8 * c, desc_lo + c * B_KSTEP_DESC, idesc, 1);
    }
}

// ---- self-test of the MMA building block: D = split3(A) x split3(W)^T through TMEM --------------------------------------------
// A fp32 [128][128], W fp32 [N][128] (N in {128, 48, 32}); mode 0: value path (3 products), 1: tangent path (2 products)
__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(const float *A, const unsigned char *wimg, int n, int mode,
                                                            float *D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wbytes = n * 128 * 2;
    for (int i = tid; i < 2 * wbytes / 16; i += 160) reinterpret_cast<int4 *>(smem)[i] = reinterpret_cast<const int4 *>(wimg)[i];
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp < 4) {
        const uint32_t tm_lane = tm + ((uint32_t)(32 * warp) << 16);
        const float *row = A + (size_t)tid * 128;
        for (int c = 0; c < 8; ++c) {
            uint32_t av[16], at[8];
            for (int pq = 0; pq < 8; ++pq) {
                split_h2(make_float2(row[16 * c + 2 * pq], row[16 * c + 2 * pq + 1]), av[pq], av[8 + pq]);
                at[pq] = av[pq];
            }
            st16(tm_lane + COL_AV + 16 * c, av);
            st8(tm_lane + COL_AT0 + 8 * c, at);
        }
        wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    if (warp == 4 && lane == 0) {
        const uint64_t w_hi = make_bdesc(smem_u32(smem), B_LBO, B_SBO), w_lo = make_bdesc(smem_u32(smem) + wbytes, B_LBO, B_SBO);
        if (mode == 0)
            mma_value(tm, COL_D0, w_hi, w_lo, make_idesc(n));
        else
            mma_tangent(tm, COL_D0, COL_AT0, w_hi, w_lo, make_idesc(n));
        commit(smem_u32(&bar));
    }
    if (warp < 4) {
        mbar_wait(smem_u32(&bar), 0);
        fence_after();
        const uint32_t tm_lane = tm + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < n / 16; ++c) {
            uint32_t d[16];
            ld16(tm_lane + COL_D0 + 16 * c, d);
            wait_ld();
            for (int k = 0; k < 16; ++k) D[(size_t)tid * n + 16 * c + k] = __uint_as_float(d[k]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 4) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}


// ---- diagnostic: tcgen05.mma issue/execute rate for the operand configurations this kernel could use --------------------------
// variant 0: A in TMEM, B no-swizzle K-major (what the rollout uses)   1: A in TMEM, B SWIZZLE_128B K-major
// variant 2: A and B in shared memory, no swizzle                       3: A and B in shared memory, SWIZZLE_128B
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum), "r"(0)
        : "memory");
}

__global__ void __launch_bounds__(64, 1) tc_mma_bench_kernel(int variant, int n, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 + 32) * 1024 / 16; i += 64) reinterpret_cast<int4 *>(smem)[i] = make_int4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_ptr;
    if (warp == 1) {
        const bool sw = variant & 1, ss = variant >= 2;
        const uint32_t sb = smem_u32(smem), sa = sb + 64 * 1024;
        // 8 K-steps per operand image; swizzled: one 128-byte row per N index (K = 64), 8-row groups 1024 bytes apart
        const uint64_t bdesc = sw ? (make_bdesc(sb, 16, 1024) | (2ull << 61)) : make_bdesc(sb, B_LBO, B_SBO);
        const uint64_t adesc = sw ? (make_bdesc(sa, 16, 1024) | (2ull << 61)) : make_bdesc(sa, B_LBO, B_SBO);
        const uint64_t kstep = sw ? 2 : B_KSTEP_DESC;
        const uint32_t idesc = make_idesc(n);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (ss)
                        tc_mma_ss(tm + 256, adesc + c * kstep, bdesc + c * kstep, idesc, 1);
                    else
                        mma_ts(tm + 256, tm + 8 * c, bdesc + c * kstep, idesc, 1);
                }
            }
            t1 = clock64();
            commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t2 = clock64();
        if (t0) {
            out[0] = t1 - t0;  // issue time
            out[1] = t2 - t0;  // until all MMAs completed
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

// ---- diagnostic: TMEM load / store throughput with all 16 warps of a CTA busy (what the epilogues see) ------------------------
// mode 0: tcgen05.ld 32x32b.x16 back to back (one wait per 4), mode 1: tcgen05.st 32x32b.x16, mode 2: ld + dependent st
__global__ void __launch_bounds__(512, 1) tc_tmem_bench_kernel(int mode, int iters, long long *out) {
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = tmem_ptr + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (warp >> 2);
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = tid + k;
    st16(tm, r);
    wait_st();
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (mode == 0) {
                ld16(tm + 16 * j, r);
            } else if (mode == 1) {
                st16(tm + 16 * j, r);
            } else {
                ld16(tm + 16 * j, r);
                wait_ld();
                r[0] += 1;
                st16(tm + 16 * j + 64, r);
            }
        }
        if (mode == 0) wait_ld();
        if (mode != 0) wait_st();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    if (r[3] == 0x7fffffff) out[1] = r[5];
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512) : "memory");
    }
}

}  // namespace

// Diagnostic: cycles per warp-level tcgen05.ld / st (32x32b.x16 = 2 KB) with 16 warps of one CTA issuing concurrently.
extern "C" int dp_tc_tmem_bench(int device, int mode, int iters, double *cycles_per_op) {
    DP_REQUIRE(cycles_per_op && mode >= 0 && mode <= 2 && iters > 0, "dp_tc_tmem_bench: bad argument");
    DeviceGuard guard(device);
    long long *d;
    DP_CUDA(cudaMalloc(&d, 16));
    DP_CUDA(cudaMemset(d, 0, 16));
    tc_tmem_bench_kernel<<<1, 512>>>(mode, iters, d);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    long long h[2];
    DP_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    *cycles_per_op = (double)h[0] / (4.0 * iters);  // per warp-level instruction, 16 warps in flight
    return 0;
}

// Diagnostic export: one 128 x n x 128 product through the same TMEM-operand MMA path the rollout uses.
// A host fp32 [128][128], W host fp32 [n_valid][128] (zero-padded to n rows), D host fp32 [128][n] = A * (scale*W)^T.
extern "C" int dp_tc_selftest(int device, const float *A, const float *W, int n_valid, int n, int mode, float scale, float *D) {
    DP_REQUIRE(A && W && D && (n == 128 || n == 48 || n == 32) && n_valid <= n, "dp_tc_selftest: bad argument");
    DeviceGuard guard(device);
    std::vector<__half> img((size_t)2 * n * 128);
    build_split_image(W, n_valid, n, scale, img.data(), img.data() + (size_t)n * 128);
    float *dA, *dD;
    unsigned char *dW;
    DP_CUDA(cudaMalloc(&dA, 128 * 128 * 4));
    DP_CUDA(cudaMalloc(&dD, 128 * n * 4));
    DP_CUDA(cudaMalloc(&dW, img.size() * 2));
    DP_CUDA(cudaMemcpy(dA, A, 128 * 128 * 4, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemcpy(dW, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    DP_CUDA(cudaMemset(dD, 0, 128 * n * 4));
    const int smem = 2 * n * 128 * 2;
    DP_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_selftest_kernel<<<1, 160, smem>>>(dA, dW, n, mode, dD);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    DP_CUDA(cudaMemcpy(D, dD, 128 * n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA);
    cudaFree(dD);
    cudaFree(dW);
    return 0;
}

// Diagnostic export: cycles per tcgen05.mma (M=128, N=n, K=16, fp16) for `iters` x 4 back-to-back instructions.
// out[0] = issue cycles per MMA, out[1] = cycles per MMA until completion.  variant: see tc_mma_bench_kernel.
extern "C" int dp_tc_mma_bench(int device, int variant, int n, int iters, double *out) {
    DP_REQUIRE(out && variant >= 0 && variant <= 3 && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0, "dp_tc_mma_bench: bad argument");
    DeviceGuard guard(device);
    long long *d;
    DP_CUDA(cudaMalloc(&d, 16));
    DP_CUDA(cudaMemset(d, 0, 16));
    const int smem = 96 * 1024;
    DP_CUDA(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_mma_bench_kernel<<<1, 64, smem>>>(variant, n, iters, d);
    DP_CUDA(cudaGetLastError());
    DP_CUDA(cudaDeviceSynchronize());
    long long h[2];
    DP_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    out[0] = (double)h[0] / (4.0 * iters);
    out[1] = (double)h[1] / (4.0 * iters);
    return 0;
}
