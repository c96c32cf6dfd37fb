// placeholder: backward lands next
#include "head_common.cuh"
using namespace grl;
extern "C" int grl_head_backward(grl_handle* h, const grl_head_params*, const float*, int, int, const float*, const float*,
                                 const float*, const float*, const float*, float*, const grl_head_grads*, void*, size_t, void*) {
    return set_error(h, GRL_EINVAL, "head backward not built yet");
}
