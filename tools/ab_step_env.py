"""GPU: time the head step (B=32, T=8, fwd+bwd) in THIS process; run it under different environment switches to A/B a build-time
or load-time choice on one box:  for v in a b; do SWITCH=$v python tools/ab_step_env.py; done"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import head, synth  # noqa: E402

B, T = 32, 8
dev = torch.device("cuda")
sd = {k: v.to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
x = synth.make_head_input(B, T).to(dev)
gu, gc = synth.make_head_grads(B, T)
gu, gc = gu.to(dev), gc.to(dev)
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


(fu, fc, *_), (dx, grads) = step()
print("fwd+bwd %s ms   (checks: %.6f %.6f %.6f)" % (" ".join("%.3f" % timeit() for _ in range(3)), float(fu.double().sum()), float(fc.double().sum()),
                                                   float(dx.double().abs().sum())), flush=True)
