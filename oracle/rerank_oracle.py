"""ORACLE (test infrastructure, NOT product code) — CPU restatement of k-reciprocal re-ranking.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

Follows /root/reference/reid/evaluator/rerank.py:37-104 (`re_ranking`), stage by stage, in the decomposition the
CUDA path uses (the reference builds dense N x N `V` matrices with Python loops; the stages below carry the same
values in the same float32 operation order):

  stage 1  original_dist   :41-46   O[i][j] = fl32(D[j][i]^2 / max_r D[r][i]^2)       (squared, column-normalised, transposed)
  stage 2  initial_rank    :48      only columns [0, max(k1+1, k2)) are ever read      (:56-57, :64-65, :80)
  stage 3  V rows          :54-76   k-reciprocal set R(i,k1), expanded by R(c, round(k1/2)) of every member c whose
                                    overlap with R(i,k1) exceeds 2/3; weights exp(-O[i, idx]) / sum
  stage 4  query expansion :78-83   V_qe[i] = mean(V[initial_rank[i,:k2]])   (float32, rows added in rank order)
  stage 5  Jaccard         :86-98   temp_min[j] = sum_c min(V[i,c], V[j,c]) over the non-zero columns c of row i in
                                    ascending order (float32), jaccard = 1 - t/(2-t)
  stage 6  blend           :100-104 final = jaccard*(1-lambda) + O*lambda, columns [nq, N)

Ties: np.argsort's default introsort leaves the order of equal distances implementation-defined; this restatement
(like the CUDA path) breaks ties by the lower index (`kind="stable"`).  On tie-free inputs it is pinned to the real
reference by tests/golden/rerank_*.npz (oracle/make_golden.py), bit for bit.
"""
from __future__ import annotations

import numpy as np


def build_original(q_g, q_q, g_g):
    """rerank.py:41-46."""
    full = np.concatenate([np.concatenate([q_q, q_g], axis=1), np.concatenate([q_g.T, g_g], axis=1)], axis=0)
    sq = np.power(full, 2).astype(np.float32)
    return np.ascontiguousarray(np.transpose(1. * sq / np.max(sq, axis=0)))


def k_reciprocal(rank, i, k):
    """Members f of rank[i, :k+1] with i in rank[f, :k+1], in rank order (rerank.py:56-59)."""
    fwd = rank[i, :k + 1]
    back = rank[fwd, :k + 1]
    return fwd[(back == i).any(axis=1)]


def v_row(O, rank, i, k1):
    """Sorted unique expansion indices and their normalised weights for row i (rerank.py:54-76)."""
    half = int(np.around(k1 / 2.))
    R = k_reciprocal(rank, i, k1)
    parts = [R]
    for c in R:
        Rc = k_reciprocal(rank, c, half)
        if np.isin(Rc, R).sum() > 2. / 3 * len(Rc):
            parts.append(Rc)
    idx = np.unique(np.concatenate(parts))
    w = np.exp(-O[i, idx])
    return idx, (1. * w / np.sum(w)).astype(np.float32)


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, return_stages=False):
    q_g = np.asarray(q_g_dist)
    nq, ng = q_g.shape
    N = nq + ng
    O = build_original(q_g, np.asarray(q_q_dist), np.asarray(g_g_dist))
    kk = min(N, max(k1 + 1, k2))
    rank = np.argsort(O, axis=1, kind="stable")[:, :kk].astype(np.int32)
    V = np.zeros((N, N), np.float32)
    for i in range(N):
        idx, w = v_row(O, rank, i, k1)
        V[i, idx] = w
    if k2 != 1:
        Vq = np.zeros_like(V)
        for i in range(N):
            acc = np.zeros(N, np.float32)
            for r in rank[i, :k2]:                       # np.mean over axis 0 adds the rows in order, then divides
                acc = acc + V[r]
            Vq[i] = acc / np.float32(len(rank[i, :k2]))
        V = Vq
    Vt = np.ascontiguousarray(V.T)
    jac = np.zeros((nq, N), np.float32)
    for i in range(nq):
        t = np.zeros(N, np.float32)
        for c in np.nonzero(V[i])[0]:                    # ascending c; zero entries of column c add min(v, 0) = 0
            t = t + np.minimum(V[i, c], Vt[c])
        jac[i] = 1 - t / (2. - t)
    final = jac * (1 - lambda_value) + O[:nq] * lambda_value
    final = final[:, nq:]
    if return_stages:
        return final, dict(O=O, rank=rank, V=V)
    return final
