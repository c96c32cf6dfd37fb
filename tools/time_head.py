"""GPU: time head forward / backward separately, with the internal side stream on and off."""
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


def fwd(train=True, save=True):
    global ws
    out = head.head_forward_raw(sd, x, B, T, train, save=save, ws=ws)
    ws = out[-1]


def bwd():
    head.head_backward_raw(sd, x, B, T, ws, gu, gc)


def step():
    fwd()
    bwd()


for rep in range(2):
    for ov in (0, 1, 2, 3):
        lib.grl_set_overlap(h, ov)
        fwd()
        print("overlap mask=%d  fwd(train,save) %.3f ms   bwd %.3f ms   fwd+bwd %.3f ms   fwd(eval,nosave) %.3f ms" %
              (ov, timeit(fwd, 20), timeit(bwd, 20), timeit(step, 20), timeit(lambda: fwd(False, False), 20)), flush=True)
