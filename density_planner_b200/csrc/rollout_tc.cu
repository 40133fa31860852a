// K1 (tensor-core version) -- placeholder until the tcgen05 kernel lands; selecting DP_IMPL_TC fails loudly.
#include "dp_common.cuh"

int dp_tc_prepare(dp_controller *c, const float *) {
    c->d_tc = nullptr;
    c->tc_bytes = 0;
    return 0;
}
void dp_tc_release(dp_controller *c) {
    if (c->d_tc) cudaFree(c->d_tc);
    c->d_tc = nullptr;
}
int dp_launch_rollout_tc(const dp_controller *, const RolloutParams &, cudaStream_t) {
    dp_set_error("DP_IMPL_TC: tensor-core rollout kernel not built into this library");
    return 4;
}
