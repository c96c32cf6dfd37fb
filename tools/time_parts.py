"""GPU: time the stand-alone GCE / TRL operators (forward, backward) at B=32, T=8."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import _lib, head, synth  # noqa: E402

B, T = 32, 8
dev = torch.device("cuda")
sd = {k: v.to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
x = synth.make_head_input(B, T).to(dev)
gu, gc = synth.make_head_grads(B, T)
gu, gc = gu.to(dev), gc.to(dev)
lib = _lib.load_library()
h = _lib.get_handle(dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


xu, xc, cm, ws_g = head.gce_forward_raw(sd, x, B, T, True, save=True)
fu, fc, ws_t = head.trl_forward_raw(sd, xu, xc, B, T, True, save=True)
dxu, dxc, _ = head.trl_backward_raw(sd, B, T, ws_t, gu, gc)
for ov in (0, 3):
    lib.grl_set_overlap(h, ov)
    print("overlap=%d gce_fwd %.3f  trl_fwd %.3f  trl_bwd %.3f  gce_bwd %.3f ms" % (
        ov,
        timeit(lambda: head.gce_forward_raw(sd, x, B, T, True, save=True, ws=ws_g)),
        timeit(lambda: head.trl_forward_raw(sd, xu, xc, B, T, True, save=True, ws=ws_t)),
        timeit(lambda: head.trl_backward_raw(sd, B, T, ws_t, gu, gc)),
        timeit(lambda: head.gce_backward_raw(sd, B, T, ws_g, dxu, dxc, None))), flush=True)
