"""GPU: the CTA-pair (cta_group::2) coarse GEMM vs the single-CTA one: correctness against the brute-force search, then timing."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import _lib, evaluator  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load_library()
h = _lib.get_handle(dev)
gen = torch.Generator(device=dev).manual_seed(1)
for mask in (11, 3):        # 11 = single-CTA kernel (debug bit 3), 3 = CTA-pair kernel (default)
    lib.grl_set_overlap(h, mask)
    for (nq, ng, dim, k, metric) in ((1100, 3333, 64, 20, 0), (1500, 20000, 256, 100, 1), (2048, 9472, 2048, 50, 0)):
        q = torch.nn.functional.normalize(torch.randn((nq, dim), generator=gen, device=dev))
        g = torch.nn.functional.normalize(torch.randn((ng, dim), generator=gen, device=dev))
        d, i = evaluator.retrieve_topk(q, g, k, metric=metric)
        dx, ix = evaluator.CudaSearchStages.exact(q[:64].contiguous(), g, k, 0, metric)
        torch.cuda.synchronize()
        print("mask %d  %dx%dx%d k=%d metric=%d  exact match: %s" % (mask, nq, ng, dim, k, metric, bool(torch.equal(i[:64], ix) and torch.equal(d[:64], dx))), flush=True)
NQ, NG, D, K = 10000, 400000, 2048, 100
gf = torch.nn.functional.normalize(torch.randn((NG, D), generator=gen, device=dev))
qf = torch.nn.functional.normalize(torch.randn((NQ, D), generator=gen, device=dev))
for rep in range(2):
    for mask in (3, 11):
        lib.grl_set_overlap(h, mask)
        evaluator.retrieve_topk(qf, gf, K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            evaluator.retrieve_topk(qf, gf, K)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("mask %d  NG=%d  %.2f ms per search  %.1f alg TFLOP/s" % (mask, NG, dt * 1e3, 2.0 * NQ * NG * D / dt / 1e12), flush=True)
