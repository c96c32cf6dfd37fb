"""Drop-in mirror of the reference's matching / evaluation path on B200.

ATTEvaluator drives the caller's models exactly like the reference (cnn_model(clips) -> siamese_model.self_attention ->
cat); with grl_b200.head.ResNet50_GRL_Model and grl_b200.siamese.Siamese those calls are the CUDA path end to end.

Same names, argument meaning and return types as
  reid/evaluator/attevaluator.py:15-46   evaluate_seq, pairwise_distance_tensor, cosin_dist
  reid/evaluator/eva_functions.py:134-184 evaluate
  reid/evaluator/attevaluator.py:49-163  ATTEvaluator
but every hot call goes through the C ABI (include/grl_b200.h) into sm_100a kernels.
No CPU fallback: tensors must live on a CUDA device (numpy inputs are copied there).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def _as_cuda_f32(x, device=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not x.is_cuda:
        x = x.to(device if device is not None else "cuda", non_blocking=True)
    return x.contiguous().float()


def _distance(qf, gf, metric):
    qf = _as_cuda_f32(qf)
    gf = _as_cuda_f32(gf, qf.device)
    nq, ng = qf.size(0), gf.size(0)
    q2 = qf.view(nq, -1)
    g2 = gf.view(ng, -1)
    dim = q2.size(1)
    if g2.size(1) != dim:
        raise RuntimeError("size mismatch, qf %s vs gf %s" % (tuple(q2.shape), tuple(g2.shape)))
    if dim % 8:   # TMA rows must be 16-byte multiples: zero-pad the feature axis (does not change any distance)
        pad = 8 - dim % 8
        q2 = torch.nn.functional.pad(q2, (0, pad))
        g2 = torch.nn.functional.pad(g2, (0, pad))
        dim += pad
    lib = _lib.load_library()
    with torch.cuda.device(qf.device):
        h = _lib.get_handle(qf.device)
        dist = torch.empty((nq, ng), dtype=torch.float32, device=qf.device)
        ws_bytes = lib.grl_distance_workspace_bytes(nq, ng, dim)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qf.device)
        _lib.check(h, lib.grl_distance(h, metric, q2.data_ptr(), g2.data_ptr(), nq, ng, dim, dist.data_ptr(),
                                       ws.data_ptr(), ws_bytes, _lib.stream_ptr(qf.device)), "grl_distance")
    return dist


def cosin_dist(qf, gf):
    """attevaluator.py:44-46: `-torch.mm(qf, gf.t())` (split-bf16 tcgen05 GEMM, fp32 accumulate)."""
    return _distance(qf, gf, 0)


def pairwise_distance_tensor(query_x, gallery_x):
    """attevaluator.py:33-41: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, min=1e-12))."""
    return _distance(query_x, gallery_x, 1)


def cmc_map_device(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=100):
    """Runs grl_cmc_map; returns device tensors (cmc_hits int32[max_rank], ap f64[nq], first_hit int32[nq])."""
    dist = _as_cuda_f32(distmat)
    dev = dist.device
    nq, ng = dist.shape
    ids = [torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(device=dev, dtype=torch.int64).contiguous()
           for a in (q_pids, g_pids, q_camids, g_camids)]
    if ids[0].numel() != nq or ids[1].numel() != ng or ids[2].numel() != nq or ids[3].numel() != ng:
        raise RuntimeError("evaluate: id arrays do not match distmat shape %s" % ((nq, ng),))
    max_rank = min(max_rank, ng)
    lib = _lib.load_library()
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        hits = torch.empty(max_rank, dtype=torch.int32, device=dev)
        ap = torch.empty(nq, dtype=torch.float64, device=dev)
        first = torch.empty(nq, dtype=torch.int32, device=dev)
        _lib.check(h, lib.grl_cmc_map(h, dist.data_ptr(), dist.stride(0), ids[0].data_ptr(), ids[1].data_ptr(),
                                      ids[2].data_ptr(), ids[3].data_ptr(), nq, ng, max_rank, hits.data_ptr(),
                                      ap.data_ptr(), first.data_ptr(), _lib.stream_ptr(dev)), "grl_cmc_map")
    return hits, ap, first


def evaluate(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=100):
    """eva_functions.py:134-184.  Returns (all_cmc np.float32[max_rank], mAP float) like the reference.

    `distmat` may be a CUDA tensor (stays on the device) or a numpy array (copied to the device).
    When fewer than max_rank gallery rows exist, max_rank shrinks like :136-138 (the reference then
    crashes on ragged rows if junk removal shortens them further; here the curve is simply padded).
    """
    num_q, num_g = distmat.shape
    if num_g < max_rank:
        max_rank = num_g
        print("Note: number of gallery samples is quite small, got {}".format(num_g))
    hits, ap, first = cmc_map_device(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
    hits = hits.cpu().numpy()
    ap = ap.cpu().numpy()
    valid = ap >= 0
    num_valid_q = float(valid.sum())
    assert num_valid_q > 0, "Error: all query identities do not appear in gallery"
    all_cmc = hits.astype(np.float32) / num_valid_q      # :179-180 (float32 sum, python-float divide)
    mAP = np.mean(ap[valid])                             # :182
    return all_cmc, mAP


def evaluate_seq(distmat, query_pids, query_camids, gallery_pids, gallery_camids, path=None, cmc_topk=[1, 5, 10, 20]):
    """attevaluator.py:15-30 (note the pids/camids argument order); prints like the reference, returns Rank-1."""
    query_ids = np.array(query_pids)
    gallery_ids = np.array(gallery_pids)
    query_cams = np.array(query_camids)
    gallery_cams = np.array(gallery_camids)
    cmc_scores, mAP = evaluate(distmat, query_ids, gallery_ids, query_cams, gallery_cams)
    print('Mean AP: {:4.1%}'.format(mAP))
    for r in cmc_topk:
        print("Rank-{:<3}: {:.1%}".format(r, cmc_scores[r - 1]))
    print("------------------")
    return cmc_scores[0]


def argsort_rows(distmat):
    """np.argsort(distmat, axis=1) (eva_functions.py:139) as a stable (distance, index) sort; ng <= 16384."""
    dist = _as_cuda_f32(distmat)
    nq, ng = dist.shape
    lib = _lib.load_library()
    with torch.cuda.device(dist.device):
        h = _lib.get_handle(dist.device)
        order = torch.empty((nq, ng), dtype=torch.int32, device=dist.device)
        _lib.check(h, lib.grl_argsort_rows(h, dist.data_ptr(), dist.stride(0), nq, ng, order.data_ptr(),
                                           _lib.stream_ptr(dist.device)), "grl_argsort_rows")
    return order


# --------------------------------------------------------------------------------------------
# Gallery-sharded retrieval (BASELINE.json configs[4]).  The reference has no counterpart: it materialises the whole
# matrix and argsorts it on the CPU (attevaluator.py:150, eva_functions.py:139).  Gallery rows are split contiguously
# over the ranks, queries are replicated; every rank searches its shard, one all-gather of [nq, k] candidates over
# NCCL/NVLink follows, and every rank merges them with ties broken by the lower global gallery index, so the result
# does not depend on the number of shards.
# --------------------------------------------------------------------------------------------
def shard_bounds(num_gallery, world_size, rank):
    """Contiguous split of gallery rows: returns (first row, number of rows) of `rank`."""
    base, rem = divmod(num_gallery, world_size)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def _padded_features(qf, gf):
    qf = _as_cuda_f32(qf)
    gf = _as_cuda_f32(gf, qf.device)
    if gf.size(1) != qf.size(1):
        raise RuntimeError("size mismatch, qf %s vs gf %s" % (tuple(qf.shape), tuple(gf.shape)))
    if qf.size(1) % 8:      # TMA rows must be 16-byte multiples: zero-pad the feature axis (does not change any distance)
        pad = 8 - qf.size(1) % 8
        qf = torch.nn.functional.pad(qf, (0, pad))
        gf = torch.nn.functional.pad(gf, (0, pad))
    return qf, gf


class PreparedGallery(object):
    """A static gallery shard converted once for the coarse pass (grl_gallery_prepare: fp16 rows with per-row scales, squared
    norms): pass it to retrieve_topk / sharded_retrieve in place of the feature tensor and every search skips the conversion.
    Keeps the fp32 rows (`.gf`, feature axis zero-padded to a multiple of 8) for the exact re-score."""

    def __init__(self, gf):
        gf = _as_cuda_f32(gf)
        if gf.size(1) % 8:
            gf = torch.nn.functional.pad(gf, (0, 8 - gf.size(1) % 8))
        self.gf = gf.contiguous()
        ng, dim = self.gf.shape
        lib = _lib.load_library()
        with torch.cuda.device(gf.device):
            h = _lib.get_handle(gf.device)
            nbytes = lib.grl_gallery_prepared_bytes(ng, dim)
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=gf.device)
            _lib.check(h, lib.grl_gallery_prepare(h, self.gf.data_ptr(), ng, dim, self.buf.data_ptr(), nbytes, _lib.stream_ptr(gf.device)),
                       "grl_gallery_prepare")

    def size(self, i):
        return self.gf.size(i)


def _pad_queries(qf, dim):
    qf = _as_cuda_f32(qf)
    if qf.size(1) > dim or dim - qf.size(1) >= 8:
        raise RuntimeError("size mismatch, qf %s vs prepared gallery [*, %d]" % (tuple(qf.shape), dim))
    return torch.nn.functional.pad(qf, (0, dim - qf.size(1))) if qf.size(1) < dim else qf


def retrieve_topk(qf, gf, k, idx_base=0, metric=0):
    """k nearest rows of this gallery shard per query (grl_dist_topk: coarse fp16 tensor-core pass + exact fp32 re-score).
    `gf`: features [ng, dim] or a PreparedGallery.  Returns CUDA (dist f32 [nq,k], index i64 [nq,k]), the stable top-k of the
    fixed-order fp32 distances."""
    prepared = gf if isinstance(gf, PreparedGallery) else None
    if prepared is not None:
        gf = prepared.gf
        qf = _pad_queries(qf, gf.size(1))
    else:
        qf, gf = _padded_features(qf, gf)
    nq, ng, dim = qf.size(0), gf.size(0), qf.size(1)
    lib = _lib.load_library()
    with torch.cuda.device(qf.device):
        h = _lib.get_handle(qf.device)
        top_d = torch.empty((nq, k), dtype=torch.float32, device=qf.device)
        top_i = torch.empty((nq, k), dtype=torch.int64, device=qf.device)
        ws_bytes = lib.grl_dist_topk_workspace_bytes(nq, ng, dim)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qf.device)
        if prepared is None:
            _lib.check(h, lib.grl_dist_topk(h, metric, qf.data_ptr(), gf.data_ptr(), nq, ng, dim, k, idx_base, top_d.data_ptr(),
                                            top_i.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(qf.device)), "grl_dist_topk")
        else:
            _lib.check(h, lib.grl_dist_topk_prepared(h, metric, qf.data_ptr(), gf.data_ptr(), prepared.buf.data_ptr(), nq, ng, dim, k, idx_base,
                                                     top_d.data_ptr(), top_i.data_ptr(), ws.data_ptr(), ws_bytes,
                                                     _lib.stream_ptr(qf.device)), "grl_dist_topk_prepared")
    return top_d, top_i


def merge_topk(all_d, all_i):
    """[nshards, nq, k] candidate lists -> [nq, k] (grl_topk_merge; (distance, global index) order)."""
    nshards, nq, k = all_d.shape
    all_d = all_d.contiguous()
    all_i = all_i.contiguous()
    lib = _lib.load_library()
    with torch.cuda.device(all_d.device):
        h = _lib.get_handle(all_d.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=all_d.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=all_d.device)
        _lib.check(h, lib.grl_topk_merge(h, all_d.data_ptr(), all_i.data_ptr(), nshards, nq, k, out_d.data_ptr(), out_i.data_ptr(),
                                         _lib.stream_ptr(all_d.device)), "grl_topk_merge")
    return out_d, out_i


class CudaSearchStages(object):
    """The per-rank stages of the two-stage exact search (include/grl_b200.h: grl_coarse_topk, grl_rescore, grl_topk_finalize,
    grl_exact_topk, grl_topk_merge).  sharded_retrieve takes the stages as an object only so that its collectives can be
    exercised on CPU with gloo in tests (which plug in a numpy restatement)."""

    @staticmethod
    def kprime(k):
        return int(_lib.load_library().grl_topk_kprime(k))

    @staticmethod
    def coarse(qf, gf, kp, idx_base, metric, prepared=None):
        nq, ng, dim = qf.size(0), gf.size(0), qf.size(1)
        lib = _lib.load_library()
        with torch.cuda.device(qf.device):
            h = _lib.get_handle(qf.device)
            cd = torch.empty((nq, kp), dtype=torch.float32, device=qf.device)
            ci = torch.empty((nq, kp), dtype=torch.int64, device=qf.device)
            gmax2 = torch.zeros(1, dtype=torch.float32, device=qf.device)
            dirty = torch.zeros(nq, dtype=torch.int32, device=qf.device)
            ws_bytes = lib.grl_coarse_topk_workspace_bytes(nq, ng, dim)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qf.device)
            if prepared is None:
                _lib.check(h, lib.grl_coarse_topk(h, metric, qf.data_ptr(), gf.data_ptr(), nq, ng, dim, kp, idx_base, cd.data_ptr(),
                                                  ci.data_ptr(), gmax2.data_ptr(), dirty.data_ptr(), ws.data_ptr(), ws_bytes,
                                                  _lib.stream_ptr(qf.device)), "grl_coarse_topk")
            else:
                _lib.check(h, lib.grl_coarse_topk_prepared(h, metric, qf.data_ptr(), prepared.buf.data_ptr(), nq, ng, dim, kp, idx_base,
                                                           cd.data_ptr(), ci.data_ptr(), gmax2.data_ptr(), dirty.data_ptr(), ws.data_ptr(),
                                                           ws_bytes, _lib.stream_ptr(qf.device)), "grl_coarse_topk_prepared")
        return cd, ci, gmax2, dirty

    merge = staticmethod(merge_topk)

    @staticmethod
    def rescore(qf, gf, cand_i, idx_base, metric):
        nq, ng, dim = qf.size(0), gf.size(0), qf.size(1)
        lib = _lib.load_library()
        with torch.cuda.device(qf.device):
            h = _lib.get_handle(qf.device)
            ed = torch.empty(cand_i.shape, dtype=torch.float32, device=qf.device)
            _lib.check(h, lib.grl_rescore(h, metric, qf.data_ptr(), gf.data_ptr(), nq, ng, dim, idx_base, cand_i.data_ptr(),
                                          cand_i.size(1), ed.data_ptr(), _lib.stream_ptr(qf.device)), "grl_rescore")
        return ed

    @staticmethod
    def finalize(qf, cd, ci, ed, gmax2, dirty, k, metric):
        nq, dim, kp = qf.size(0), qf.size(1), ci.size(1)
        lib = _lib.load_library()
        with torch.cuda.device(qf.device):
            h = _lib.get_handle(qf.device)
            top_d = torch.empty((nq, k), dtype=torch.float32, device=qf.device)
            top_i = torch.empty((nq, k), dtype=torch.int64, device=qf.device)
            flags = torch.empty(nq, dtype=torch.int32, device=qf.device)
            nflag = torch.zeros(1, dtype=torch.int32, device=qf.device)
            _lib.check(h, lib.grl_topk_finalize(h, metric, qf.data_ptr(), nq, dim, cd.data_ptr(), ci.data_ptr(), ed.data_ptr(), kp,
                                                gmax2.data_ptr(), dirty.data_ptr(), k, top_d.data_ptr(), top_i.data_ptr(),
                                                flags.data_ptr(), nflag.data_ptr(), _lib.stream_ptr(qf.device)), "grl_topk_finalize")
        return top_d, top_i, flags

    @staticmethod
    def exact(qf, gf, k, idx_base, metric):
        nq, ng, dim = qf.size(0), gf.size(0), qf.size(1)
        lib = _lib.load_library()
        with torch.cuda.device(qf.device):
            h = _lib.get_handle(qf.device)
            top_d = torch.empty((nq, k), dtype=torch.float32, device=qf.device)
            top_i = torch.empty((nq, k), dtype=torch.int64, device=qf.device)
            ws_bytes = lib.grl_exact_topk_workspace_bytes(nq, ng, dim)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qf.device)
            _lib.check(h, lib.grl_exact_topk(h, metric, qf.data_ptr(), gf.data_ptr(), nq, ng, dim, k, idx_base, top_d.data_ptr(),
                                             top_i.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(qf.device)), "grl_exact_topk")
        return top_d, top_i


def sharded_retrieve(qf, gf_local, k, idx_base, group=None, metric=0, stages=None):
    """One search over a gallery sharded across the ranks of `group` (torch.distributed; NCCL over NVLink on B200).

    Every rank ranks its shard by the coarse tensor-core distance and keeps K' candidates per query; one all-gather + merge
    gives the global coarse K'; every rank re-scores (fixed-order fp32) the candidates whose gallery rows it owns and an
    all-reduce (sum of disjoint contributions) assembles them; finalisation sorts by exact distance and proves completeness
    per query.  Queries without a proof (rare: near-duplicate galleries) are searched by brute force per shard and merged.
    The result is identical on every rank and for every shard count."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    prepared = gf_local if isinstance(gf_local, PreparedGallery) else None
    if stages is None:
        if world == 1:
            return retrieve_topk(qf, gf_local, k, idx_base=idx_base, metric=metric)
        stages = CudaSearchStages
        if prepared is not None:
            gf_local = prepared.gf
            qf = _pad_queries(qf, gf_local.size(1))
        else:
            qf, gf_local = _padded_features(qf, gf_local)
    kp = stages.kprime(k)
    if prepared is not None:
        cd, ci, gmax2, dirty = stages.coarse(qf, gf_local, kp, idx_base, metric, prepared=prepared)
    else:
        cd, ci, gmax2, dirty = stages.coarse(qf, gf_local, kp, idx_base, metric)
    if world > 1:
        all_d = [torch.empty_like(cd) for _ in range(world)]
        all_i = [torch.empty_like(ci) for _ in range(world)]
        dist.all_gather(all_d, cd.contiguous(), group=group)
        dist.all_gather(all_i, ci.contiguous(), group=group)
        dist.all_reduce(gmax2, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(dirty, op=dist.ReduceOp.MAX, group=group)   # a row whose candidate buffer overflowed on any shard
        cd, ci = stages.merge(torch.stack(all_d), torch.stack(all_i))
    ed = stages.rescore(qf, gf_local, ci, idx_base, metric)
    if world > 1:
        dist.all_reduce(ed, op=dist.ReduceOp.SUM, group=group)       # every candidate is owned by exactly one rank
    top_d, top_i, flags = stages.finalize(qf, cd, ci, ed, gmax2, dirty, k, metric)
    rows = torch.nonzero(flags).flatten()                            # identical on every rank (same inputs to finalize)
    if rows.numel():
        d_x, i_x = stages.exact(qf[rows].contiguous(), gf_local, k, idx_base, metric)
        if world > 1:
            all_d = [torch.empty_like(d_x) for _ in range(world)]
            all_i = [torch.empty_like(i_x) for _ in range(world)]
            dist.all_gather(all_d, d_x.contiguous(), group=group)
            dist.all_gather(all_i, i_x.contiguous(), group=group)
            d_x, i_x = stages.merge(torch.stack(all_d), torch.stack(all_i))
        top_d[rows] = d_x
        top_i[rows] = i_x
    return top_d, top_i


class ATTEvaluator(object):
    """attevaluator.py:49-163 with the hot calls replaced; loaders/models are the caller's (PyTorch/cuDNN backbone)."""

    def __init__(self, cnn_model, Siamese_model, only_eval):
        super(ATTEvaluator, self).__init__()
        self.cnn_model = cnn_model
        self.siamese_model = Siamese_model
        self.only_eval = only_eval

    @torch.no_grad()
    def extract_feature(self, data_loader):
        self.cnn_model.eval()
        self.siamese_model.eval()
        qf, q_pids, q_camids = [], [], []
        for i, inputs in enumerate(data_loader):
            imgs, pids, camids = inputs
            if self.only_eval:                                   # 'dense' mode, attevaluator.py:68-98
                b, n, s, c, h, w = imgs.size()
                imgs = imgs.view(b * n, s, c, h, w).cuda()
                feats = []
                for y in range(int(math.ceil(b * n * 1.0 / 8))):  # chunks of 8 clips (:72-77)
                    clips = imgs[y * 8:(y + 1) * 8]
                    x_uncorr, feats_corr = self.cnn_model(clips)
                    out_frame = self.siamese_model.self_attention(feats_corr)
                    feats.append(torch.cat((x_uncorr, out_frame, feats_corr.mean(dim=1)), dim=1))
                feats = torch.cat(feats, 0).mean(dim=0)           # mean over clips (:83-84 / :94-95)
                qf.append(feats.unsqueeze(0))
            else:                                                 # one clip per tracklet (:100-117)
                imgs = imgs.cuda()
                x_uncorr, feats_corr = self.cnn_model(imgs)
                out_frame = self.siamese_model.self_attention(feats_corr)
                qf.append(torch.cat((x_uncorr, out_frame, feats_corr.mean(dim=1)), dim=1))
            q_pids.extend(pids)
            q_camids.extend(camids)
        qf = torch.cat(qf, 0)
        return qf, np.asarray(q_pids), np.asarray(q_camids)

    def evaluate(self, query, gallery, query_loader, gallery_loader, path, visual, rerank):
        if visual:
            raise NotImplementedError("rank visualisation is debug tooling outside the hot path (SURVEY.md §2)")
        qf, q_pids, q_camids = self.extract_feature(query_loader)
        print('Done, obtained {}-by-{} matrix'.format(qf.size(0), qf.size(1)))
        gf, g_pids, g_camids = self.extract_feature(gallery_loader)
        gf = torch.cat((qf, gf), 0)                               # :143-145 queries join the gallery
        g_pids = np.append(q_pids, g_pids)
        g_camids = np.append(q_camids, g_camids)
        print('Done, obtained {}-by-{} matrix'.format(gf.size(0), gf.size(1)))
        print("Computing distance matrix")
        distmat = cosin_dist(qf, gf)                              # stays on the device (no 74 MB D2H, :150)
        if rerank:                                                # :151-155, all three matrices stay on the device
            print('Applying person re-ranking ...')
            from .rerank import re_ranking
            distmat_qq = pairwise_distance_tensor(qf, qf)
            distmat_gg = pairwise_distance_tensor(gf, gf)
            distmat = re_ranking(distmat, distmat_qq, distmat_gg)
        return evaluate_seq(distmat, q_pids, q_camids, g_pids, g_camids, path)
