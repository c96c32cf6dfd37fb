"""GPU: one gallery-shard search (10k x NG x 2048, top-100), timed; run under ncu for the launch list."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import evaluator  # noqa: E402

NQ, NG, D, K = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 2048, 100
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
gf = torch.nn.functional.normalize(torch.randn((NG, D), generator=g, device=dev))
qf = torch.nn.functional.normalize(torch.randn((NQ, D), generator=g, device=dev))
evaluator.retrieve_topk(qf, gf, K)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2):
    d, i = evaluator.retrieve_topk(qf, gf, K)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 2
print("NG=%d  %.2f ms per search  %.1f alg TFLOP/s" % (NG, dt * 1e3, 2.0 * NQ * NG * D / dt / 1e12))
