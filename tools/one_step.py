"""GPU: warm up, then run ONE B=32,T=8 head fwd+bwd step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
`--single-stream` serialises the internal side stream so per-launch metrics are attributable."""
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
if "--single-stream" in sys.argv:
    lib.grl_set_overlap(h, 0)
ws = None


def step():
    global ws
    out = head.head_forward_raw(sd, x, B, T, True, save=True, ws=ws)
    ws = out[-1]
    head.head_backward_raw(sd, x, B, T, ws, gu, gc)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step done")
