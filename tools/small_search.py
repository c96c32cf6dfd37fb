"""GPU: one small search that exercises the 256 x 256 CTA-pair coarse GEMM, the warp-level list merges, the re-score, the proof,
the second chance and the brute-force leg (for compute-sanitizer runs: tools/sanitize.sh)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import evaluator as ev  # noqa: E402

rng = np.random.default_rng(3)
nq, dim, k = 1100, 64, 20
# 3,072 rows: the first chunk is the whole gallery (exact first-chunk selection); 9,000 rows: a full 8,192-column first chunk (sampled
# pre-filter; 1,500 copies of one row send every query's selection back to the exact path with an over-full boundary bin) + one
# more chunk with candidate lists and a warp-level merge
for ng, ncopies in ((3072, 0), (9000, 0), (9000, 1500)):
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    g = rng.standard_normal((ng, dim)).astype(np.float32)
    g[::13] = g[5]                                      # duplicates: some proofs fail -> second chance / brute force
    if ncopies:
        g[rng.permutation(8192)[:ncopies]] = q.mean(axis=0)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    stats = torch.zeros(8, dtype=torch.int32, device="cuda")
    d, i = ev.sharded_topk(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k, 0, stats=stats)
    dx, ix = ev.CudaSearchStages.exact(torch.from_numpy(q[:64]).cuda(), torch.from_numpy(g).cuda(), k, 0, 0)
    torch.cuda.synchronize()
    assert torch.equal(i[:64], ix) and torch.equal(d[:64], dx)
    print("small search OK (%d rows, %d copies), stats" % (ng, ncopies), stats.cpu().tolist())
