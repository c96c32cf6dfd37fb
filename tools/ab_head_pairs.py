"""GPU: the head step with the CTA-pair split-bf16 GEMM (default) vs the single-CTA kernel (grl_set_overlap bit 4), A/B on one box."""
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
for mask in (3, 19):
    lib.grl_set_overlap(h, mask)
    (fu, fc, *_), (dx, grads) = step()
    res[mask] = (fu.clone(), fc.clone(), dx.clone(), grads["temporal_learning_block.forward_f1.0.weight"].clone())
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("pair vs single-CTA kernel: f_uncorr %.2e  f_corr %.2e  dx %.2e  d f1.w %.2e" %
      tuple(rel(a, b) for a, b in zip(res[3], res[19])), flush=True)
NAMES = {3: "default (pairs where K >= 1024 and >= 1 wave)", 19: "single-CTA kernels only", 35: "fp16 GEMMs on the single-CTA kernel",
         67: "pairs for every 256-wide tile"}
for rep in range(3):
    for mask in (3, 35, 67, 19):
        lib.grl_set_overlap(h, mask)
        print("mask %2d (%s)  fwd+bwd %.3f ms" % (mask, NAMES[mask], timeit()), flush=True)
