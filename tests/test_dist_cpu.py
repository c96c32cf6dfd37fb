"""CPU, world_size 2, gloo: the multi-GPU plumbing of gallery-sharded retrieval (shard bounds, all-gather of the
per-shard candidates, merge) with the per-rank kernels replaced by the oracle (test infrastructure only)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grl_b200 import evaluator


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_local(qf, gf, k, idx_base=0, metric=0):
    from oracle import eval_oracle as eo
    d = eo.cosin_dist(qf.numpy(), gf.numpy()) if metric == 0 else eo.pairwise_distance(qf.numpy(), gf.numpy())
    v, i = eo.topk_stable(d, k, idx_base)
    return torch.from_numpy(v), torch.from_numpy(i)


def _oracle_merge(all_d, all_i):
    from oracle import eval_oracle as eo
    k = all_d.shape[2]
    v, i = eo.merge_topk([a.numpy() for a in all_d], [a.numpy() for a in all_i], k)
    return torch.from_numpy(v), torch.from_numpy(i)


def _worker(rank, world, port, ng, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    q = rng.standard_normal((9, 32)).astype(np.float32)
    g = rng.standard_normal((ng, 32)).astype(np.float32)
    g[::7] = g[3]                                   # exact ties across shard boundaries
    lo, n = evaluator.shard_bounds(ng, world, rank)
    d, i = evaluator.sharded_retrieve(torch.from_numpy(q), torch.from_numpy(g[lo:lo + n]), 10, lo,
                                      local_search=_oracle_local, merge=_oracle_merge)
    if rank == 0:
        np.savez(out, d=d.numpy(), i=i.numpy(), q=q, g=g)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_gallery():
    for ng in (1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [evaluator.shard_bounds(ng, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == ng
            for (a, n), (b, _) in zip(spans, spans[1:]):
                assert a + n == b
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_sharded_retrieve_world2_gloo(tmp_path):
    from oracle import eval_oracle as eo
    out = str(tmp_path / "r.npz")
    ng = 101
    mp.spawn(_worker, args=(2, _free_port(), ng, out), nprocs=2, join=True)
    r = np.load(out)
    v_ref, i_ref = eo.topk_stable(eo.cosin_dist(r["q"], r["g"]), 10)
    assert np.array_equal(r["i"], i_ref) and np.array_equal(r["d"], v_ref)
