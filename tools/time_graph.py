"""GPU: eager vs CUDA-graph replay of the B=32,T=8 head step (GraphedHeadStep), plus a parity check between the two."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import head, synth  # noqa: E402

B, T = 32, 8
dev = torch.device("cuda")
sd = {k: v.to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
sd2 = {k: v.clone() for k, v in sd.items()}
x = synth.make_head_input(B, T).to(dev)
gu, gc = synth.make_head_grads(B, T)
gu, gc = gu.to(dev), gc.to(dev)


def timeit(fn, n=20):
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


ws = None


def eager():
    global ws
    out = head.head_forward_raw(sd, x, B, T, True, save=True, ws=ws)
    ws = out[-1]
    return out, head.head_backward_raw(sd, x, B, T, ws, gu, gc)


step = head.GraphedHeadStep(sd2, B, T)
step.x.copy_(x); step.d_f_uncorr.copy_(gu); step.d_f_corr.copy_(gc)
# parity: same parameters and buffers -> same outputs (both sides have run the same number of BN updates? no: compare one fresh pair)
sd_a = {k: v.to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
sd_b = {k: v.clone() for k, v in sd_a.items()}
out_a = head.head_forward_raw(sd_a, x, B, T, True, save=True)
dx_a, g_a = head.head_backward_raw(sd_a, x, B, T, out_a[-1], gu, gc)
st_b = head.GraphedHeadStep(sd_b, B, T)                   # construction runs the step twice (warm-up + capture) on sd_b's buffers
for k in sd_b:
    sd_b[k].copy_(synth.make_head_params(0)[k].to(dev))
st_b.x.copy_(x); st_b.d_f_uncorr.copy_(gu); st_b.d_f_corr.copy_(gc)
st_b()
torch.cuda.synchronize()
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
print("graph vs eager: f_uncorr %.1e f_corr %.1e dx %.1e conv1.w grad %.1e running_mean %.1e nbt %s" % (
    rel(st_b.f_uncorr, out_a[0]), rel(st_b.f_corr, out_a[1]), rel(st_b.dx, dx_a),
    rel(st_b.grads["temporal_learning_block.uncorr_memo_forward.conv1.weight"], g_a["temporal_learning_block.uncorr_memo_forward.conv1.weight"]),
    rel(sd_b["temporal_learning_block.uncorr_memo_forward.bn1.running_mean"], sd_a["temporal_learning_block.uncorr_memo_forward.bn1.running_mean"]),
    int(sd_b["temporal_learning_block.uncorr_memo_forward.bn1.num_batches_tracked"])))
for rep in range(2):
    print("eager %.3f ms/step   graph replay %.3f ms/step" % (timeit(lambda: eager()), timeit(lambda: step())), flush=True)
