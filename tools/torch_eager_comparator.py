"""GPU comparator: the reference head's own torch ops (the oracle's line-by-line transcription: F.conv2d 1x1,
F.linear, F.batch_norm + autograd) moved to the SAME B200 -- cuDNN / cuBLAS fp32 (TF32 off = the reference's precision) and
with TF32 on -- next to the hand-written path, same inputs, CUDA events.  Measurement tooling, not the product path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import head, synth  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

B, T = 32, 8
dev = torch.device("cuda")
params = {k: v.to(dev) for k, v in synth.make_head_params(0).items()}
x0 = synth.make_head_input(B, T).to(dev)
gu, gc = synth.make_head_grads(B, T)
gu, gc = gu.to(dev), gc.to(dev)


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def eager_step():
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in params.items()}
    x = x0.clone().requires_grad_(True)

    def step():
        for v in leaf.values():
            if v.requires_grad:
                v.grad = None
        x.grad = None
        out = ho.ref_forward(leaf, x, B, T, True)
        ((out["f_uncorr"] * gu).sum() + (out["f_corr"] * gc).sum()).backward()
    return step


res = {}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    ms = timeit(eager_step())
    res["torch_eager_%s" % ("tf32" if tf32 else "fp32")] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
sd = {k: v.contiguous() for k, v in params.items()}
ws = None


def ours():
    global ws
    out = head.head_forward_raw(sd, x0, B, T, True, save=True, ws=ws)
    ws = out[-1]
    head.head_backward_raw(sd, x0, B, T, ws, gu, gc)


ms = timeit(ours, n=10, warm=3)
res["grl_b200"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
res["workload"] = "GCE+TRL head fwd+bwd, B=32 T=8, train-mode BN, fp32 in/out, one B200"
print(json.dumps(res))
