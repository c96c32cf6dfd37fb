"""GPU: one train-mode head step at a size that routes the big GEMMs to the CTA-pair kernels (split-bf16 and fp16 x1), for
compute-sanitizer runs (racecheck on the remote mbarrier arrivals / multicast commits of gemm_pair.cuh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import head, synth  # noqa: E402

B, T = 16, 2
sd = {k: v.cuda().contiguous() for k, v in synth.make_head_params(0).items()}
x = synth.make_head_input(B, T).cuda()
gu, gc = synth.make_head_grads(B, T)
out = head.head_forward_raw(sd, x, B, T, True, save=True)
dx, grads = head.head_backward_raw(sd, x, B, T, out[-1], gu.cuda(), gc.cuda())
torch.cuda.synchronize()
print("head step OK", float(out[0].sum()), float(dx.abs().sum()))
