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


# --------------------------------------------------------------------------------------------
# Two-stage exact search (grl_b200/csrc/eval.cu): numpy restatements used by the tests.
# The reference has no counterpart beyond `-qf @ gf.T` + argsort (attevaluator.py:44-46, eva_functions.py:139); what is
# pinned here is the ARITHMETIC the kernels promise: distances are fp32 inner products in one fixed summation order.
# --------------------------------------------------------------------------------------------
def _fixed_order_reduce(prod):
    """prod [..., dim] float32 products -> [...] float32: lane l sums the float4 chunks l, l+32, ... component by component,
    then an xor-shuffle tree over the 32 lanes (warp_dot_fixed in eval.cu)."""
    dim = prod.shape[-1]
    assert dim % 4 == 0
    nv = dim // 4
    nt = (nv + 31) // 32
    lanes = np.zeros(prod.shape[:-1] + (32,), np.float32)
    for t in range(nt):
        lo = t * 32
        n = min(32, nv - lo)
        blk = prod[..., lo * 4:(lo + n) * 4].reshape(prod.shape[:-1] + (n, 4))
        for c in range(4):
            lanes[..., :n] = lanes[..., :n] + blk[..., c]
    for off in (16, 8, 4, 2, 1):
        lanes = lanes + lanes[..., np.arange(32) ^ off]
    return lanes[..., 0]


def exact_distance_fixed(qf: np.ndarray, gf: np.ndarray, metric: int = 0, block: int = 2048) -> np.ndarray:
    """[nq, ng] float32 distances exactly as rescore_kernel / exact_rows_kernel compute them."""
    qf = np.ascontiguousarray(qf, np.float32)
    gf = np.ascontiguousarray(gf, np.float32)
    out = np.empty((qf.shape[0], gf.shape[0]), np.float32)
    qq = _fixed_order_reduce(qf * qf)
    for c0 in range(0, gf.shape[0], block):
        g = gf[c0:c0 + block]
        dot = _fixed_order_reduce(qf[:, None, :] * g[None, :, :])
        if metric == 0:
            out[:, c0:c0 + block] = -dot
        else:
            gg = _fixed_order_reduce(g * g)
            s = (qq[:, None] + gg[None, :]) - np.float32(2.0) * dot
            out[:, c0:c0 + block] = np.sqrt(np.maximum(s, np.float32(1e-12)))
    return out


def f16_rows(x: np.ndarray):
    """f16_rows_kernel: per-row power-of-two scale so that max|x| lands in [2^14, 2^15), fp16 rounding.
    Returns (x16 as float32, inv_scale [rows], sqnorm [rows])."""
    x = np.ascontiguousarray(x, np.float32)
    m = np.abs(x).max(axis=1)
    e = np.where((m > 0) & np.isfinite(m), np.floor(np.log2(np.where(m > 0, m, 1.0))), 0.0).astype(np.int64)
    e = np.clip(e, -100, 100)
    s = np.ldexp(np.float32(1.0), (14 - e).astype(np.int32)).astype(np.float32)
    x16 = (x * s[:, None]).astype(np.float16).astype(np.float32)
    return x16, np.ldexp(np.float32(1.0), (e - 14).astype(np.int32)).astype(np.float32), _fixed_order_reduce(x * x)


def coarse_distance(qf, gf, metric: int = 0) -> np.ndarray:
    """Stage-1 distances: fp16 operands, exact products, (here) float64 accumulation rounded to float32; metric 1 is the
    SQUARED L2 distance.  The tensor core's accumulation order differs, which the error bound CE covers."""
    q16, qi, qn = f16_rows(qf)
    g16, gi, gn = f16_rows(gf)
    dot = ((q16.astype(np.float64) @ g16.astype(np.float64).T) * qi[:, None].astype(np.float64) * gi[None, :]).astype(np.float32)
    if metric == 0:
        return -dot
    return np.maximum(qn[:, None] + gn[None, :] - np.float32(2.0) * dot, np.float32(1e-12))


def coarse_error_constant(dim: int) -> float:
    return float(np.float32(2.0 ** -10) + np.float32(2.0 ** -17) + np.float32(2.0) * np.float32(dim) * np.float32(2.0 ** -24))


def finalize_topk(qf, coarse_d, cand_i, exact_d, gmax2, k, metric=0):
    """topk_finalize_kernel: (top_d, top_i, flags).  cand_i < 0 marks empty slots; coarse_d ascending per row."""
    nq, kp = cand_i.shape
    top_d = np.full((nq, k), np.inf, np.float32)
    top_i = np.full((nq, k), -1, np.int64)
    flags = np.zeros(nq, np.int32)
    ce = np.float32(coarse_error_constant(qf.shape[1]))
    for r in range(nq):
        valid = cand_i[r] >= 0
        d, i = exact_d[r][valid], cand_i[r][valid]
        order = np.lexsort((i, d))[:k]
        top_d[r, :len(order)] = d[order]
        top_i[r, :len(order)] = i[order]
        if valid.all() and len(order) == k:
            u = top_d[r, k - 1]
            ck = np.float32(coarse_d[r, kp - 1])
            qq = np.float32((qf[r].astype(np.float64) ** 2).sum())
            emax = ce * np.sqrt(qq) * np.sqrt(np.float32(gmax2)) * np.float32(1.00001)
            if metric == 1:
                lb = ck - np.float32(2.0) * emax - np.float32(1e-5) * (qq + np.float32(gmax2))
                ok = lb > 1e-12 and np.sqrt(lb) * np.float32(0.999999) > u
            else:
                ok = ck - emax - np.float32(1e-6) * abs(ck) > u
            flags[r] = 0 if ok else 1
    return top_d, top_i, flags


def rescore_cut(coarse_kth, qq, gmax2, dim, metric=0):
    """rescore_keys_kernel (grl_b200/csrc/search.cu): candidates ranked >= k by the coarse pass whose coarse distance exceeds
    this value are NOT re-scored -- each of the first k candidates is then strictly closer in exact arithmetic
    (|coarse - exact| <= E_max for every pair), so such a row cannot be among the k nearest.  float32 arithmetic as on the device."""
    ce = np.float32(coarse_error_constant(dim))
    qq, g2, ck = np.float32(qq), np.float32(gmax2), np.float32(coarse_kth)
    emax = ce * np.sqrt(qq) * np.sqrt(g2) * np.float32(1.00001)
    if metric == 1:
        return ck + np.float32(4.0) * emax + np.float32(2e-5) * (qq + g2) + np.float32(4e-12)
    return ck + np.float32(2.0) * emax + np.float32(2e-6) * np.abs(ck)


def two_stage_topk(qf, gf, k, kprime, metric=0, skip=True):
    """The whole two-stage search restated with numpy for one shard: coarse K' list -> (optionally skipping) re-score ->
    finalisation + proof -> brute force for flagged rows.  Returns (top_d, top_i, flags, n_rescored)."""
    qf = np.ascontiguousarray(qf, np.float32)
    gf = np.ascontiguousarray(gf, np.float32)
    c = coarse_distance(qf, gf, metric)
    exact = exact_distance_fixed(qf, gf, metric)
    cd, ci = topk_stable(c, kprime)
    nq, kp = ci.shape
    if kp < kprime:
        cd = np.concatenate([cd, np.full((nq, kprime - kp), np.inf, np.float32)], 1)
        ci = np.concatenate([ci, np.full((nq, kprime - kp), -1, np.int64)], 1)
    ed = np.where(ci >= 0, np.take_along_axis(exact, np.clip(ci, 0, gf.shape[0] - 1), 1), np.float32(0)).astype(np.float32)
    qq = _fixed_order_reduce(qf * qf)
    gmax2 = np.float32(_fixed_order_reduce(gf * gf).max())
    n_rescored = int((ci >= 0).sum())
    if skip:
        for r in range(nq):
            if k - 1 < kprime and ci[r, k - 1] >= 0:
                cut = rescore_cut(cd[r, k - 1], qq[r], gmax2, qf.shape[1], metric)
                drop = (np.arange(kprime) >= k) & (cd[r] > cut) & (ci[r] >= 0)
                ed[r, drop] = np.inf
                n_rescored -= int(drop.sum())
    top_d, top_i, flags = finalize_topk(qf, cd, ci, ed, gmax2, k, metric)
    for r in np.nonzero(flags)[0]:
        v, i = topk_stable(exact[r:r + 1], k)
        top_d[r, :v.shape[1]], top_i[r, :i.shape[1]] = v[0], i[0]
    return top_d, top_i, flags, n_rescored
