// Single-step API kernel: u_func (:26), f_func (:43), dudx_func (:69), dfdx_func (:93), divf_func (:116) of
// systems/ControlAffineSystem.py for N states sharing one (xref, uref).  Not the throughput path (that is the fused
// rollout); one warp per sample, full 2x5 Jacobian by forward-mode tangents in the three directions that pass through
// the MLPs (theta, v, d) plus the two that only enter through xe (px, py).
#include "dp_common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int NC = 4;  // value + 3 tangents

// one MLP for the warp's sample: h[c][128] in shared memory per warp
template <int NOUT>
__device__ void mlp_warp(const float *__restrict__ W2t, const float *__restrict__ W3t, const float *__restrict__ small_net,
                         const float in[4], const float din[3][4], float (*h)[DP_HID], float (*h2)[DP_HID],
                         float (*out)[DP_NOUT_A], int lane) {
    const float4 *W1 = reinterpret_cast<const float4 *>(small_net + DP_SM_W1);
    const float *b1 = small_net + DP_SM_B1, *b2 = small_net + DP_SM_B2, *b3 = small_net + DP_SM_B3;
    for (int j = lane; j < DP_HID; j += 32) {
        const float4 w = __ldg(W1 + j);
        float pre = w.x * in[0];
        pre = fmaf(w.y, in[1], pre); pre = fmaf(w.z, in[2], pre); pre = fmaf(w.w, in[3], pre);
        pre += __ldg(b1 + j);
        const float t = tanhf(pre), g = 1.0f - t * t;
        h[0][j] = t;
        for (int d = 0; d < 3; ++d)
            h[1 + d][j] = g * (w.x * din[d][0] + w.y * din[d][1] + w.z * din[d][2] + w.w * din[d][3]);
    }
    __syncwarp();
    for (int j = lane; j < DP_HID; j += 32) {
        float a[NC] = {0, 0, 0, 0};
        for (int k = 0; k < DP_HID; ++k) {
            const float w = __ldg(W2t + k * DP_HID + j);
            for (int c = 0; c < NC; ++c) a[c] = fmaf(w, h[c][k], a[c]);
        }
        const float t = tanhf(a[0] + __ldg(b2 + j)), g = 1.0f - t * t;
        h2[0][j] = t;
        for (int c = 1; c < NC; ++c) h2[c][j] = g * a[c];
    }
    __syncwarp();
    for (int j = lane; j < NOUT; j += 32) {
        float a[NC] = {0, 0, 0, 0};
        for (int k = 0; k < DP_HID; ++k) {
            const float w = __ldg(W3t + k * NOUT + j);
            for (int c = 0; c < NC; ++c) a[c] = fmaf(w, h2[c][k], a[c]);
        }
        out[0][j] = a[0] + __ldg(b3 + j);
        for (int c = 1; c < NC; ++c) out[c][j] = a[c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32) step_kernel(const float *__restrict__ big, const float *__restrict__ small,
                                                          const float *__restrict__ x, const float *__restrict__ xref5,
                                                          const float *__restrict__ uref2, long long N, float *u_raw,
                                                          float *u, float *dudx, float *f, float *div) {
    __shared__ float sh[WARPS][NC][DP_HID];
    __shared__ float sh2[WARPS][NC][DP_HID];
    __shared__ float soa[WARPS][NC][DP_NOUT_A];
    __shared__ float sob[WARPS][NC][DP_NOUT_A];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long n = (long long)blockIdx.x * WARPS + warp; n < N; n += (long long)gridDim.x * WARPS) {
        float xs[5], xr[4];
        for (int k = 0; k < 5; ++k) xs[k] = __ldg(x + n * 5 + k);
        for (int k = 0; k < 4; ++k) xr[k] = __ldg(xref5 + k);
        float xe[4];
        xe[0] = xs[0] - xr[0]; xe[1] = xs[1] - xr[1]; xe[3] = xs[3] - xr[3];
        xe[2] = __fadd_rn(__fsub_rn(xs[2], xr[2]), xs[4]);
        const float in[4] = {xs[2], xs[3], xs[2] - xe[2], xs[3] - xe[3]};
        // tangent directions through the MLP input: theta -> in0, v -> in1, d -> in2 = theta - xe2 (d/dd = -1)
        const float din[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, -1, 0}};
        mlp_warp<DP_NOUT_A>(big + DP_BIG_W2A, big + DP_BIG_W3A, small + DP_SM_NET_A, in, din, sh[warp], sh2[warp], soa[warp], lane);
        mlp_warp<DP_NOUT_B>(big + DP_BIG_W2B, big + DP_BIG_W3B, small + DP_SM_NET_B, in, din, sh[warp], sh2[warp], sob[warp], lane);
        if (lane == 0) {
            const float(*oa)[DP_NOUT_A] = soa[warp];
            const float(*ob)[DP_NOUT_A] = sob[warp];
            float tz[12], gz[12], dz[3][12];
            for (int i = 0; i < 12; ++i) {
                float z = 0;
                for (int j = 0; j < 4; ++j) z = fmaf(oa[0][4 * i + j], xe[j], z);
                tz[i] = tanhf(z);
                gz[i] = 1.0f - tz[i] * tz[i];
                for (int d = 0; d < 3; ++d) {
                    float s = 0;
                    for (int j = 0; j < 4; ++j) s = fmaf(oa[1 + d][4 * i + j], xe[j], s);
                    // d xe / d theta = e2, d xe / d v = e3, d xe / d d = e2
                    dz[d][i] = s + oa[0][4 * i + (d == 1 ? 3 : 2)];
                }
            }
            float ur[2], J[2][5];
            for (int r = 0; r < 2; ++r) {
                float acc = 0;
                for (int i = 0; i < 12; ++i) acc = fmaf(ob[0][12 * r + i], tz[i], acc);
                ur[r] = acc + __ldg(uref2 + r);
                for (int sdir = 0; sdir < 2; ++sdir) {  // px, py: only through xe0 / xe1
                    float a = 0;
                    for (int i = 0; i < 12; ++i) a += ob[0][12 * r + i] * (gz[i] * oa[0][4 * i + sdir]);
                    J[r][sdir] = a;
                }
                for (int d = 0; d < 3; ++d) {
                    float a = 0;
                    for (int i = 0; i < 12; ++i) a += ob[1 + d][12 * r + i] * tz[i] + ob[0][12 * r + i] * (gz[i] * dz[d][i]);
                    J[r][2 + d] = a;
                }
            }
            const float m0 = (ur[0] >= -3.0f && ur[0] <= 3.0f) ? 1.0f : 0.0f;
            const float m1 = (ur[1] >= -3.0f && ur[1] <= 3.0f) ? 1.0f : 0.0f;
            const float uc0 = fminf(fmaxf(ur[0], -3.0f), 3.0f), uc1 = fminf(fmaxf(ur[1], -3.0f), 3.0f);
            if (u_raw) { u_raw[2 * n] = ur[0]; u_raw[2 * n + 1] = ur[1]; }
            if (u) { u[2 * n] = uc0; u[2 * n + 1] = uc1; }
            if (dudx)
                for (int k = 0; k < 5; ++k) {
                    dudx[10 * n + k] = m0 * J[0][k];
                    dudx[10 * n + 5 + k] = m1 * J[1][k];
                }
            if (f) {
                float sn, cs;
                sincosf(xs[2], &sn, &cs);
                f[5 * n + 0] = xs[3] * cs; f[5 * n + 1] = xs[3] * sn; f[5 * n + 2] = uc0; f[5 * n + 3] = uc1;
                f[5 * n + 4] = 0.0f;
            }
            if (div) div[n] = __fadd_rn(__fmul_rn(m0, J[0][2]), __fmul_rn(m1, J[1][3]));
        }
        __syncwarp();
    }
}

}  // namespace

int dp_step(const dp_controller *c, const float *x, const float *xref5, const float *uref2, int64_t N, float *u_raw,
            float *u, float *dudx, float *f, float *div, void *stream) {
    DP_REQUIRE(c && x && xref5 && uref2, "dp_step: null argument");
    if (N <= 0) return 0;
    DeviceGuard g(c->device);
    long long blocks = (N + WARPS - 1) / WARPS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    step_kernel<<<(int)blocks, WARPS * 32, 0, (cudaStream_t)stream>>>(c->d_big, c->d_small, x, xref5, uref2, N, u_raw, u,
                                                                     dudx, f, div);
    DP_CUDA(cudaGetLastError());
    return 0;
}
