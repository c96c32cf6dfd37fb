"""ORACLE (test infrastructure, NOT product code) — CPU restatement of GRL's matching path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.

Follows, function by function:
  cosin_dist                /root/reference/reid/evaluator/attevaluator.py:44-46
  pairwise_distance_tensor  /root/reference/reid/evaluator/attevaluator.py:33-41
  evaluate                  /root/reference/reid/evaluator/eva_functions.py:134-184
  evaluate_seq              /root/reference/reid/evaluator/attevaluator.py:15-30

`evaluate_literal` is the literal transcription (np.argsort, per-query loop).
`evaluate_rankcount` is the sort-free formulation the CUDA kernel implements
(rank of a positive = number of kept gallery items strictly closer, ties broken
by lower gallery index == a *stable* argsort).  On tie-free inputs both agree
exactly; with ties numpy's default introsort is implementation-defined, which is
why the north-star allows ranking differences at ties (|dd| <= 1e-5).
Pinned against the real reference by tests/golden/eval_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def cosin_dist(qf: np.ndarray, gf: np.ndarray) -> np.ndarray:
    """attevaluator.py:44-46: dist = -qf @ gf.T (float32 in, float32 out)."""
    return -(qf @ gf.T)


def pairwise_distance(qf: np.ndarray, gf: np.ndarray) -> np.ndarray:
    """attevaluator.py:33-41: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 1e-12))."""
    xx = (qf * qf).sum(1, keepdims=True)
    yy = (gf * gf).sum(1, keepdims=True).T
    d = xx + yy - 2.0 * (qf @ gf.T)
    return np.sqrt(np.clip(d, 1e-12, None)).astype(qf.dtype)


def evaluate_literal(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=100, kind=None):
    """eva_functions.py:134-184, line for line (list-comp at :172 vectorised, same values)."""
    num_q, num_g = distmat.shape
    if num_g < max_rank:
        max_rank = num_g
    indices = np.argsort(distmat, axis=1) if kind is None else np.argsort(distmat, axis=1, kind=kind)
    matches = (g_pids[indices] == q_pids[:, np.newaxis]).astype(np.int32)
    all_cmc, all_AP = [], []
    num_valid_q = 0.0
    for q_idx in range(num_q):
        order = indices[q_idx]
        remove = (g_pids[order] == q_pids[q_idx]) & (g_camids[order] == q_camids[q_idx])
        keep = np.invert(remove)
        orig_cmc = matches[q_idx][keep]
        if not np.any(orig_cmc):
            continue
        cmc = orig_cmc.cumsum()
        cmc[cmc > 1] = 1
        all_cmc.append(cmc[:max_rank])
        num_valid_q += 1.0
        num_rel = orig_cmc.sum()
        tmp_cmc = orig_cmc.cumsum() / (np.arange(orig_cmc.shape[0]) + 1.0)
        tmp_cmc = tmp_cmc * orig_cmc
        all_AP.append(tmp_cmc.sum() / num_rel)
    assert num_valid_q > 0, "Error: all query identities do not appear in gallery"
    all_cmc = np.asarray(all_cmc).astype(np.float32)
    all_cmc = all_cmc.sum(0) / num_valid_q
    return all_cmc, float(np.mean(all_AP))


def evaluate_rankcount(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=100):
    """Sort-free CMC/mAP: what grl_cmc_map computes.  Returns (cmc f32[max_rank], mAP, ap[nq], first[nq])."""
    num_q, num_g = distmat.shape
    max_rank = min(max_rank, num_g)
    hits = np.zeros(max_rank, np.int64)
    ap = np.full(num_q, -1.0, np.float64)
    first = np.full(num_q, -1, np.int64)
    gidx = np.arange(num_g)
    for qi in range(num_q):
        d = distmat[qi]
        same = g_pids == q_pids[qi]
        keep = ~(same & (g_camids == q_camids[qi]))
        pos = np.nonzero(same & keep)[0]
        if pos.size == 0:
            continue
        ranks = np.empty(pos.size, np.int64)
        for j, g in enumerate(pos):
            before = keep & ((d < d[g]) | ((d == d[g]) & (gidx < g)))
            ranks[j] = before.sum()
        ranks.sort()
        first[qi] = ranks[0]
        if ranks[0] < max_rank:
            hits[ranks[0]:] += 1
        ap[qi] = np.sum((np.arange(pos.size) + 1.0) / (ranks + 1.0)) / pos.size
    valid = ap >= 0
    nv = int(valid.sum())
    assert nv > 0, "Error: all query identities do not appear in gallery"
    cmc = (hits.astype(np.float32) / float(nv)).astype(np.float32)   # same f32 sum / python-float divide as :179-180
    return cmc, float(np.mean(ap[valid])), ap, first


def evaluate_seq(distmat, query_pids, query_camids, gallery_pids, gallery_camids, cmc_topk=(1, 5, 10, 20)):
    """attevaluator.py:15-30 without the prints: returns (rank1, cmc, mAP)."""
    cmc, mAP = evaluate_literal(distmat, np.array(query_pids), np.array(gallery_pids),
                                np.array(query_camids), np.array(gallery_camids))
    return cmc[0], cmc, mAP


def topk_stable(dist: np.ndarray, k: int, idx_base: int = 0):
    """Row-wise k smallest by (distance, index).  Returns (values f32 [nq,k], indices i64 [nq,k])."""
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(dist, order, 1)
    return vals, order.astype(np.int64) + idx_base


def merge_topk(vals_list, idx_list, k: int):
    """Merge per-shard top-k lists; ties broken by lower global index (shard-count invariant)."""
    v = np.concatenate(vals_list, 1)
    i = np.concatenate(idx_list, 1)
    order = np.lexsort((i, v), axis=1)[:, :k]
    return np.take_along_axis(v, order, 1), np.take_along_axis(i, order, 1)
