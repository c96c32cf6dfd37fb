"""CPU, world_size 2, gloo: the multi-GPU plumbing of gallery-sharded retrieval (shard bounds, all-gather of the
per-shard candidates, merge) with the per-rank kernels replaced by the oracle (test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
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


class NumpyStages(object):
    """The per-rank stages of evaluator.sharded_retrieve restated with the oracle (numpy) so that the collectives run on
    CPU tensors over gloo.  KPRIME is tiny so that the completeness proof fails for some queries and the brute-force leg
    (flags -> exact -> all-gather -> merge) is exercised too."""
    KPRIME = 16

    @classmethod
    def kprime(cls, k):
        return cls.KPRIME

    @staticmethod
    def coarse(qf, gf, kp, idx_base, metric):
        from oracle import eval_oracle as eo
        c = eo.coarse_distance(qf.numpy(), gf.numpy(), metric)
        v, i = eo.topk_stable(c, kp, idx_base)
        pad = kp - v.shape[1]
        if pad > 0:
            v = np.concatenate([v, np.full((v.shape[0], pad), np.inf, np.float32)], 1)
            i = np.concatenate([i, np.full((i.shape[0], pad), -1, np.int64)], 1)
        gmax2 = float((gf.numpy().astype(np.float64) ** 2).sum(1).max())
        dirty = torch.zeros(qf.shape[0], dtype=torch.int32)
        if idx_base == 0:
            dirty[1] = 1                            # pretend one candidate buffer overflowed on the first shard
        return torch.from_numpy(v), torch.from_numpy(i), torch.tensor([gmax2], dtype=torch.float32), dirty

    @staticmethod
    def merge(all_d, all_i):
        from oracle import eval_oracle as eo
        k = all_d.shape[2]
        d = np.where(all_i.numpy() < 0, np.inf, all_d.numpy()).astype(np.float32)
        i = np.where(all_i.numpy() < 0, np.iinfo(np.int64).max, all_i.numpy())
        v, i = eo.merge_topk(list(d), list(i), k)
        i = np.where(i == np.iinfo(np.int64).max, -1, i)
        return torch.from_numpy(v), torch.from_numpy(i)

    @staticmethod
    def rescore(qf, gf, cand_i, idx_base, metric):
        from oracle import eval_oracle as eo
        d = eo.exact_distance_fixed(qf.numpy(), gf.numpy(), metric)
        loc = cand_i.numpy() - idx_base
        own = (loc >= 0) & (loc < gf.shape[0])
        out = np.where(own, np.take_along_axis(d, np.clip(loc, 0, gf.shape[0] - 1), 1), np.float32(0)).astype(np.float32)
        return torch.from_numpy(out)

    @staticmethod
    def finalize(qf, cd, ci, ed, gmax2, dirty, k, metric):
        from oracle import eval_oracle as eo
        d, i, f = eo.finalize_topk(qf.numpy(), cd.numpy(), ci.numpy(), ed.numpy(), float(gmax2[0]), k, metric)
        f = np.maximum(f, dirty.numpy().astype(np.int32))
        return torch.from_numpy(d), torch.from_numpy(i), torch.from_numpy(f)

    @staticmethod
    def exact(qf, gf, k, idx_base, metric):
        from oracle import eval_oracle as eo
        v, i = eo.topk_stable(eo.exact_distance_fixed(qf.numpy(), gf.numpy(), metric), k, idx_base)
        pad = k - v.shape[1]
        if pad > 0:
            v = np.concatenate([v, np.full((v.shape[0], pad), np.inf, np.float32)], 1)
            i = np.concatenate([i, np.full((i.shape[0], pad), -1, np.int64)], 1)
        return torch.from_numpy(v), torch.from_numpy(i)


def _inputs(ng):
    rng = np.random.default_rng(5)
    q = rng.standard_normal((9, 32)).astype(np.float32)
    g = rng.standard_normal((ng, 32)).astype(np.float32)
    g[::7] = g[3]                                   # exact ties across shard boundaries
    return q, g


def _worker(rank, world, port, ng, metric, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, g = _inputs(ng)
    lo, n = evaluator.shard_bounds(ng, world, rank)
    d, i = evaluator.sharded_retrieve(torch.from_numpy(q), torch.from_numpy(g[lo:lo + n]), 10, lo, metric=metric, stages=NumpyStages)
    if rank == 0:
        np.savez(out, d=d.numpy(), i=i.numpy())
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


def test_staged_search_single_process_matches_oracle():
    """world_size 1 through the same staged code path (no process group): flags + brute force included."""
    from oracle import eval_oracle as eo
    q, g = _inputs(101)
    for metric in (0, 1):
        d, i = evaluator.sharded_retrieve(torch.from_numpy(q), torch.from_numpy(g), 10, 0, metric=metric, stages=NumpyStages)
        v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), 10)
        assert np.array_equal(i.numpy(), i_ref) and np.array_equal(d.numpy(), v_ref)


@pytest.mark.parametrize("metric", [0, 1])
def test_sharded_retrieve_world2_gloo(tmp_path, metric):
    from oracle import eval_oracle as eo
    out = str(tmp_path / "r.npz")
    ng = 101
    mp.spawn(_worker, args=(2, _free_port(), ng, metric, out), nprocs=2, join=True)
    r = np.load(out)
    q, g = _inputs(ng)
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), 10)
    assert np.array_equal(r["i"], i_ref) and np.array_equal(r["d"], v_ref)


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grl_b200.replicas import GradientAllReduce
    names, shapes = ["a.weight", "b.bias", "c"], [(3, 5), (7,), (2, 2, 2)]
    sync = GradientAllReduce(names, shapes, "cpu")
    results = []
    for step in range(3):                                 # double-buffered: three steps reuse buffer 0
        v = sync.views()
        for i, k in enumerate(names):
            v[k].copy_(torch.full(shapes[i], float((rank + 1) * (step + 1) * (i + 1))))
        sync.start()
        results.append({k: t.clone() for k, t in sync.finish().items()})
    if rank == 0:
        torch.save(results, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo(tmp_path):
    """replicas.GradientAllReduce: flat double-buffered gradient all-reduce, averaged over the ranks."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    results = torch.load(out)
    for step, r in enumerate(results):
        for i, k in enumerate(["a.weight", "b.bias", "c"]):
            expect = (1 + 2) / 2.0 * (step + 1) * (i + 1)          # mean over ranks of (rank + 1) * (step + 1) * (i + 1)
            assert torch.all(r[k] == expect), (step, k)
