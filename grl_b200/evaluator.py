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
# over the ranks; every rank runs the coarse tensor-core pass over its shard, the candidate lists are exchanged by query
# slice over NCCL/NVLink, and every step that follows (merge, exact re-score, completeness proof) is divided over the
# ranks as well; ties are broken by the lower global gallery index, so the result does not depend on the number of shards.
# The protocol lives in the library (grl_sharded_topk); the host only creates the communicator.
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
        ws = _search_workspace(qf.device, ws_bytes)      # persistent: no allocation per search
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


# ---- the search communicator: one NCCL communicator per handle, created by the library (grl_comm_init) from a unique id that
#      travels over the caller's torch.distributed group (any backend: it is 128 bytes of host data)
def comm_info(device=None):
    """(world, rank) of the handle's search communicator; (1, 0) without one."""
    import ctypes as C
    lib = _lib.load_library()
    w, r = C.c_int(), C.c_int()
    lib.grl_comm_info(_lib.get_handle(device), C.byref(w), C.byref(r))
    return w.value, r.value


def init_search_comm(group=None, device=None):
    """Create the library's NCCL communicator over the ranks of `group` (default: the world group).  Rank 0 draws the unique id
    (grl_comm_unique_id), torch.distributed broadcasts it, every rank calls grl_comm_init.  Idempotent per device."""
    import ctypes as C
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return 1, 0
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if comm_info(dev) == (world, rank):
        return world, rank
    lib = _lib.load_library()
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        buf = C.create_string_buffer(_lib.GRL_COMM_ID_BYTES)
        if rank == 0:
            _lib.check(h, lib.grl_comm_unique_id(h, buf, _lib.GRL_COMM_ID_BYTES), "grl_comm_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = C.create_string_buffer(box[0], _lib.GRL_COMM_ID_BYTES)
        _lib.check(h, lib.grl_comm_init(h, ident, _lib.GRL_COMM_ID_BYTES, world, rank), "grl_comm_init")
    return world, rank


def destroy_search_comm(device=None):
    _lib.load_library().grl_comm_destroy(_lib.get_handle(device))


_SEARCH_WS = {}      # device index -> persistent workspace (grown on demand, reused by every search: no allocation per call)


def _search_workspace(dev, nbytes):
    ws = _SEARCH_WS.get(dev.index)
    if ws is None or ws.numel() < nbytes:
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        shift = (-raw.data_ptr()) % 1024
        ws = raw[shift:shift + nbytes]
        _SEARCH_WS[dev.index] = ws
    return ws


def query_slice(num_queries, world_size, rank):
    """Rows of the query block rank `rank` contributes to a sharded search: (first row, number of rows)."""
    qs = -(-num_queries // world_size)
    lo = min(num_queries, rank * qs)
    return lo, max(0, min(qs, num_queries - rank * qs))


def search_stage_ms(device=None):
    """Per-stage device times of the last profiled grl_sharded_topk call (grl_search_profile): dict name -> ms."""
    import ctypes as C
    lib = _lib.load_library()
    h = _lib.get_handle(device)
    ms = (C.c_double * _lib.GRL_SEARCH_STAGES)()
    _lib.check(h, lib.grl_search_stage_ms(h, ms, _lib.GRL_SEARCH_STAGES), "grl_search_stage_ms")
    names = ("query_allgather", "convert", "coarse_pass", "exchange_merge", "rescore", "reduce_scatter", "finalize",
             "result_allgather_unpack", "brute_force")
    return {n: float(v) for n, v in zip(names, ms)}


def sharded_topk(qf, gf_local, k, idx_base, nq=None, metric=0, max_flagged=-1, out=None, stats=None):  # noqa: C901
    """grl_sharded_topk: one search over a gallery sharded across the ranks of the library's search communicator
    (init_search_comm), the whole protocol -- coarse tensor-core pass, NCCL exchanges, owned re-scores, completeness proof,
    brute-force leg -- behind one C call on the current stream.

    qf        all query rows (nq=None), or -- when `nq` is given -- this rank's query_slice of the nq queries (EVERY rank must
              then pass its slice: the mode decides which collectives run)
    gf_local  this rank's gallery rows [ng_local, dim] or a PreparedGallery of them; idx_base = global index of its first row
    out       optional (top_d f32 [nq, k], top_i i64 [nq, k]) CUDA tensors to fill;  stats: optional int32[8] CUDA tensor
    Returns (top_d, top_i): the exact stable top-k of the fixed-order fp32 distances, identical on every rank."""
    prepared = gf_local if isinstance(gf_local, PreparedGallery) else None
    if prepared is not None:
        gf = prepared.gf
        qf = _pad_queries(qf, gf.size(1))
    else:
        qf, gf = _padded_features(qf, gf_local)
    dev = qf.device
    q_rows, dim, ng = qf.size(0), qf.size(1), gf.size(0)
    is_slice = nq is not None
    nq = q_rows if nq is None else int(nq)
    if is_slice:
        world, rank = comm_info(dev)
        if q_rows != query_slice(nq, world, rank)[1]:
            raise RuntimeError("sharded_topk: rank %d of %d must pass %d query rows of %d (query_slice), got %d" %
                               (rank, world, query_slice(nq, world, rank)[1], nq, q_rows))
    lib = _lib.load_library()
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev))
        top_d, top_i = out
        nbytes = lib.grl_sharded_topk_workspace_bytes(h, nq, ng, dim, k, 1 if prepared is not None else 0)
        if nbytes == 0:
            raise RuntimeError("grl_sharded_topk: bad sizes nq=%d ng=%d dim=%d k=%d" % (nq, ng, dim, k))
        ws = _search_workspace(dev, nbytes)
        _lib.check(h, lib.grl_sharded_topk(h, metric, qf.data_ptr() if q_rows else None, 1 if is_slice else 0, gf.data_ptr(),
                                           None if prepared is None else prepared.buf.data_ptr(),
                                           nq, ng, dim, k, idx_base, max_flagged, top_d.data_ptr(), top_i.data_ptr(), _lib.ptr(stats),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)), "grl_sharded_topk")
    return top_d, top_i


class CudaSearchStages(object):
    """The per-rank stage entry points of the two-stage exact search (include/grl_b200.h: grl_coarse_topk, grl_rescore,
    grl_topk_finalize, grl_exact_topk, grl_topk_merge) for hosts that run their own collectives between them
    (staged_retrieve below).  The product path is grl_sharded_topk, which runs the whole protocol in the library."""

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
            ws = _search_workspace(qf.device, ws_bytes)
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


def staged_retrieve(qf, gf_local, k, idx_base, group=None, metric=0, stages=None):
    """The sharded search written against torch.distributed, stage by stage -- the protocol of grl_sharded_topk restated on the
    host: rank r owns query slice r (query_slice) for all list work.

        coarse K' lists of the local shard for every query
        -> all-reduce(max) of the overflow marks and of max |g|^2
        -> exchange by query slice (an all-to-all; spelled as all-gather + slicing here so that gloo can run it) and merge
           -> the global coarse K' of the own slice -> all-gather of the merged slices
        -> every rank re-scores the candidates whose gallery rows it owns (0 elsewhere)
        -> sum over ranks (a reduce-scatter; spelled as all-reduce + slicing here) -> exact distances of the own slice
        -> finalisation + completeness proof of the own slice -> all-gather of the results and flags
        -> flagged queries: brute force per shard + all-gather + merge

    `stages` supplies the per-rank computations (CudaSearchStages, or a numpy restatement in the CPU tests, which is how the
    host-side protocol is exercised over gloo).  The result is identical on every rank and for every shard count."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    prepared = gf_local if isinstance(gf_local, PreparedGallery) else None
    if stages is None:
        stages = CudaSearchStages
        if prepared is not None:
            gf_local = prepared.gf
            qf = _pad_queries(qf, gf_local.size(1))
        else:
            qf, gf_local = _padded_features(qf, gf_local)
    nq = qf.size(0)
    kp = stages.kprime(k)
    if prepared is not None:
        cd, ci, gmax2, dirty = stages.coarse(qf, gf_local, kp, idx_base, metric, prepared=prepared)
    else:
        cd, ci, gmax2, dirty = stages.coarse(qf, gf_local, kp, idx_base, metric)
    lo, n = query_slice(nq, world, rank)
    spans = [query_slice(nq, world, r) for r in range(world)]

    def gather_rows(mine, width, dtype):
        """all-gather of per-slice rows [n, width] (ragged last slices padded to the slice size) -> [nq, width]"""
        qs = spans[0][1]
        pad = torch.zeros((qs, width), dtype=dtype, device=mine.device)
        pad[:mine.size(0)] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat([p[:spans[r][1]] for r, p in enumerate(parts)], 0)

    if world > 1:
        dist.all_reduce(gmax2, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(dirty, op=dist.ReduceOp.MAX, group=group)   # a row whose candidate buffer overflowed on any shard
        all_d = [torch.empty_like(cd) for _ in range(world)]
        all_i = [torch.empty_like(ci) for _ in range(world)]
        dist.all_gather(all_d, cd.contiguous(), group=group)
        dist.all_gather(all_i, ci.contiguous(), group=group)
        md, mi = stages.merge(torch.stack([d[lo:lo + n] for d in all_d]), torch.stack([i[lo:lo + n] for i in all_i])) if n else \
            (cd[:0], ci[:0])
        cd, ci = gather_rows(md, kp, cd.dtype), gather_rows(mi, kp, ci.dtype)
    ed = stages.rescore(qf, gf_local, ci, idx_base, metric)
    if world > 1:
        dist.all_reduce(ed, op=dist.ReduceOp.SUM, group=group)       # every candidate is owned by exactly one rank
    sl = slice(lo, lo + n)
    if n:
        top_d, top_i, flags = stages.finalize(qf[sl], cd[sl], ci[sl], ed[sl], gmax2, dirty[sl], k, metric)
    else:
        top_d, top_i, flags = cd[:0, :k], ci[:0, :k], dirty[:0]
    if world > 1:
        top_d, top_i = gather_rows(top_d, k, top_d.dtype), gather_rows(top_i, k, top_i.dtype)
        flags = gather_rows(flags.view(-1, 1), 1, flags.dtype).view(-1)
    rows = torch.nonzero(flags).flatten()                            # identical on every rank
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


def sharded_retrieve(qf, gf_local, k, idx_base, group=None, metric=0, stages=None, nq=None, out=None, stats=None):
    """One search over a gallery sharded across the ranks of `group` (torch.distributed; NCCL over NVLink on B200): the k
    nearest gallery rows of every query, identical on every rank and for every shard count.

    On the GPU this is ONE library call, grl_sharded_topk (the search communicator over `group` is created on first use);
    `qf` may hold all queries or, with `nq` given, only this rank's query_slice.  With `stages` given the host-side
    restatement of the protocol (staged_retrieve) runs instead -- that is how the CPU tests exercise it over gloo."""
    if stages is not None:
        return staged_retrieve(qf, gf_local, k, idx_base, group=group, metric=metric, stages=stages)
    init_search_comm(group)
    return sharded_topk(qf, gf_local, k, idx_base, nq=nq, metric=metric, out=out, stats=stats)


def evaluate_sharded(distmat_local, q_pids, g_pids_local, q_camids, g_camids_local, idx_base, max_rank=100, group=None):
    """eva_functions.evaluate (eva_functions.py:134-184) over a gallery sharded across the ranks of `group`:
    `distmat_local` [nq, ng_local] holds this rank's gallery columns (global rows idx_base ...), g_pids_local / g_camids_local
    their identities; query ids are replicated.  grl_cmc_map_sharded: all-gather of each query's positives, per-shard rank
    counts, all-reduce(sum).  Returns (all_cmc np.float32[max_rank], mAP) -- on every rank, bit-identical to evaluate() on the
    concatenated matrix.  max_rank is taken as given (the caller knows the global gallery size)."""
    init_search_comm(group)
    dist_l = _as_cuda_f32(distmat_local)
    dev = dist_l.device
    nq, ng = dist_l.shape
    ids = [torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(device=dev, dtype=torch.int64).contiguous()
           for a in (q_pids, g_pids_local, q_camids, g_camids_local)]
    if ids[0].numel() != nq or ids[1].numel() != ng or ids[2].numel() != nq or ids[3].numel() != ng:
        raise RuntimeError("evaluate_sharded: id arrays do not match distmat shape %s" % ((nq, ng),))
    lib = _lib.load_library()
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        hits = torch.empty(max_rank, dtype=torch.int32, device=dev)
        ap = torch.empty(nq, dtype=torch.float64, device=dev)
        first = torch.empty(nq, dtype=torch.int32, device=dev)
        seen = torch.zeros(1, dtype=torch.int32, device=dev)
        max_pos = 64
        while True:
            nbytes = lib.grl_cmc_map_sharded_workspace_bytes(h, nq, max_pos)
            ws = _search_workspace(dev, nbytes)
            _lib.check(h, lib.grl_cmc_map_sharded(h, dist_l.data_ptr(), dist_l.stride(0), ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr(),
                                                  ids[3].data_ptr(), nq, ng, idx_base, max_rank, max_pos, hits.data_ptr(), ap.data_ptr(),
                                                  first.data_ptr(), seen.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                       "grl_cmc_map_sharded")
            need = int(seen.item())                 # identical on every rank (all-reduced): every rank takes the same branch
            if need <= max_pos:
                break
            max_pos = 1 << (need - 1).bit_length()
    hits = hits.cpu().numpy()
    ap = ap.cpu().numpy()
    valid = ap >= 0
    num_valid_q = float(valid.sum())
    assert num_valid_q > 0, "Error: all query identities do not appear in gallery"
    return hits.astype(np.float32) / num_valid_q, np.mean(ap[valid])


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
