"""ORACLE (test infrastructure, NOT product code) — CPU restatement of GRL's eval feature tail.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

Follows, in eval mode (running-statistics BatchNorm):
  corr_bn / uncorr_bn + F.normalize          /root/reference/reid/models/grl_model.py:222-226
  Siamese.self_attention                     /root/reference/reid/models/Siamese.py:79-106
  out_feat = cat(x_uncorr, out_frame, feats_corr.mean(1))   /root/reference/reid/evaluator/attevaluator.py:79-80
Pinned against the real reference modules by tests/golden/tail_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn_eval(p, prefix, x):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"], p[prefix + ".bias"],
                        False, 0.1, EPS)


def ref_tail(p, f_uncorr, f_corr):
    """grl_model.py:222-226."""
    n, t, c = f_corr.shape
    x_corr = F.normalize(_bn_eval(p, "corr_bn", f_corr.reshape(n * t, c)).view(n, t, c), p=2, dim=2)
    x_uncorr = F.normalize(_bn_eval(p, "uncorr_bn", f_uncorr), p=2, dim=1)
    return x_uncorr, x_corr


def ref_self_attention(p, x):
    """Siamese.py:79-106 on x [n, T, 2048]."""
    n, t, c = x.shape
    flat = x.reshape(n * t, c)
    q = _bn_eval(p, "siamese.featQ_bn", F.linear(flat, p["siamese.featQ.weight"], p["siamese.featQ.bias"]))
    q = (q / q.norm(2, 1).unsqueeze(1)).view(n, t, -1)
    k = _bn_eval(p, "siamese.featK_bn", F.linear(flat, p["siamese.featK.weight"], p["siamese.featK.bias"]))
    k = (k / k.norm(2, 1).unsqueeze(1)).view(n, t, -1)
    w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)
    pooled = torch.matmul(w, x).sum(1)
    return pooled / pooled.norm(2, 1).unsqueeze(1)


def ref_descriptor(p, f_uncorr, f_corr):
    """Per-clip 6144-d descriptor (attevaluator.py:77-80) from the head outputs."""
    x_uncorr, x_corr = ref_tail(p, f_uncorr, f_corr)
    out_frame = ref_self_attention(p, x_corr)
    return torch.cat((x_uncorr, out_frame, x_corr.mean(dim=1)), dim=1)
