"""Drop-in mirror of reid/evaluator/rerank.py (k-reciprocal re-ranking, Zhong et al. CVPR 2017) on B200.

Same name, arguments, defaults and return value as `re_ranking` (rerank.py:37-104); the work runs in
csrc/rerank.cu through `grl_rerank` (include/grl_b200.h).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ['re_ranking']


def _dev_f32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """q_g_dist [nq, ng], q_q_dist [nq, nq], g_g_dist [ng, ng] -> final_dist [nq, ng] (float32).

    numpy in -> numpy out (like the reference); CUDA tensors in -> CUDA tensor out (stays on the device)."""
    as_numpy = isinstance(q_g_dist, np.ndarray)
    device = q_g_dist.device if torch.is_tensor(q_g_dist) and q_g_dist.is_cuda else torch.device("cuda", torch.cuda.current_device())
    q_g = _dev_f32(q_g_dist, device)
    q_q = _dev_f32(q_q_dist, device)
    g_g = _dev_f32(g_g_dist, device)
    nq, ng = q_g.shape
    if tuple(q_q.shape) != (nq, nq) or tuple(g_g.shape) != (ng, ng):
        raise RuntimeError("re_ranking: q_q %s / g_g %s do not match q_g %s" % (tuple(q_q.shape), tuple(g_g.shape), (nq, ng)))
    lib = _lib.load_library()
    with torch.cuda.device(device):
        h = _lib.get_handle(device)
        ws_bytes = lib.grl_rerank_workspace_bytes(nq, ng, int(k1), int(k2))
        if ws_bytes == 0:
            raise RuntimeError("re_ranking: unsupported sizes (need nq + ng <= 65536, 1 <= k1 <= 31, 1 <= k2 <= 32)")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        final = torch.empty((nq, ng), dtype=torch.float32, device=device)
        _lib.check(h, lib.grl_rerank(h, q_g.data_ptr(), q_q.data_ptr(), g_g.data_ptr(), nq, ng, int(k1), int(k2), float(lambda_value),
                                     final.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(device)), "grl_rerank")
    return final.cpu().numpy() if as_numpy else final
