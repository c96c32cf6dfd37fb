"""GPU: the head step with the BatchNorm-backward partial sums of the memory block's bn1 / bn2 taken in the epilogue of the dgrad
GEMM that produces their input gradient (default) vs separate bn_bwd_reduce passes (grl_set_overlap bit 7), A/B on one box."""
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
ws = None


def step():
    global ws
    out = head.head_forward_raw(sd, x, B, T, True, save=True, ws=ws)
    ws = out[-1]
    return out, head.head_backward_raw(sd, x, B, T, ws, gu, gc)


def timeit(n=20):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
for mask in (3, 131):
    lib.grl_set_overlap(h, mask)
    _, (dx, grads) = step()
    res[mask] = {"dx": dx.clone(), **{k: v.clone() for k, v in grads.items()}}
rel = lambda a, b: float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
worst = max((rel(res[3][k], res[131][k]), k) for k in res[3])
print("fused vs separate reduce: worst relative difference %.2e (%s), dx %.2e" % (worst[0], worst[1], rel(res[3]["dx"], res[131]["dx"])), flush=True)
for rep in range(3):
    for mask, name in ((3, "fused into the dgrad epilogue"), (131, "separate bn_bwd_reduce passes")):
        lib.grl_set_overlap(h, mask)
        print("mask %3d (%s)  fwd+bwd %.3f ms" % (mask, name, timeit()), flush=True)
