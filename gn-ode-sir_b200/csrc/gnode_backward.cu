// Backward reverse sweep (a10) -- placeholder until the kernels land.
#include "gnode_common.cuh"

using namespace gnode;

extern "C" size_t gnode_backward_workspace_bytes(gnode_batch_t b) {
    (void)b;
    return 256;
}

extern "C" int gnode_rollout_backward(gnode_batch_t, const float*, int64_t, const gnode_params_t*, int32_t,
                                      const float*, const float*, const float*, int32_t, float*, void*, size_t,
                                      void*) {
    set_error("gnode_rollout_backward: not implemented yet");
    return GNODE_ERR_UNSUPPORTED;
}
