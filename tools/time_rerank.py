"""GPU: MARS-shape re-ranking (1,980 queries + 9,330 gallery rows), timed end to end and per stage (run under ncu for the launch list)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import evaluator, synth  # noqa: E402
from grl_b200.rerank import re_ranking  # noqa: E402

nq, nge, dim = 1980, 7350, 2048
qf, gf, qp, gp, qc, gcam = synth.make_eval_set(nq, nge, dim, seed=0, noise=4.0)
tq, tg = torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda()


def dists():
    return evaluator.pairwise_distance_tensor(tq, tg), evaluator.pairwise_distance_tensor(tq, tq), evaluator.pairwise_distance_tensor(tg, tg)


q_g, q_q, g_g = dists()
final = re_ranking(q_g, q_q, g_g)
torch.cuda.synchronize()
for name, fn in (("3 distance matrices", dists), ("re_ranking", lambda: re_ranking(q_g, q_q, g_g))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-22s %.3f ms" % (name, e0.elapsed_time(e1) / 3))
c0, m0 = evaluator.evaluate(q_g, qp, gp, qc, gcam)
c1, m1 = evaluator.evaluate(final, qp, gp, qc, gcam)
print("mAP %.4f -> %.4f, rank-1 %.4f -> %.4f" % (m0, m1, c0[0], c1[0]))
