// placeholder until the head kernels land (replaced in the next milestone)
#include "api.h"
using namespace grl;
extern "C" size_t grl_head_workspace_bytes(int, int, int) { return 0; }
extern "C" int grl_head_forward(grl_handle* h, const grl_head_params*, const float*, int, int, int, float*, float*, float*, float*,
                                float*, void*, size_t, int, void*) { return set_error(h, GRL_EINVAL, "head not built yet"); }
extern "C" int grl_head_backward(grl_handle* h, const grl_head_params*, const float*, int, int, const float*, const float*,
                                 const float*, const float*, const float*, float*, const grl_head_grads*, void*, size_t, void*) {
    return set_error(h, GRL_EINVAL, "head not built yet");
}
extern "C" int grl_head_ws_lookup(int, int, int, const char*, size_t*, size_t*) { return GRL_EINVAL; }
