"""ctypes binding of libgrl_b200.so (the C ABI in include/grl_b200.h).

Fails loudly when the CUDA library is missing or the device is not sm_100: there is no CPU or
PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgrl_b200.so")

GRL_OK, GRL_EINVAL, GRL_ECUDA, GRL_EARCH, GRL_ENOMEM, GRL_ENCCL = 0, -1, -2, -3, -4, -5
_CODES = {GRL_EINVAL: "GRL_EINVAL", GRL_ECUDA: "GRL_ECUDA", GRL_EARCH: "GRL_EARCH", GRL_ENOMEM: "GRL_ENOMEM", GRL_ENCCL: "GRL_ENCCL"}
GRL_COMM_ID_BYTES = 128
GRL_SEARCH_STAGES = 9

c_float_p = C.c_void_p   # device pointers travel as integers


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
                ("a_mn_major", C.c_int), ("b_mn_major", C.c_int),
                ("lda", C.c_longlong), ("ldb", C.c_longlong), ("ldc", C.c_longlong),
                ("a_bstride", C.c_longlong), ("b_bstride", C.c_longlong), ("c_bstride", C.c_longlong),
                ("alpha", C.c_float), ("row_scale", C.c_void_p), ("col_bias", C.c_void_p),
                ("relu", C.c_int), ("accumulate", C.c_int), ("bn", C.c_int),
                ("col_sum", C.c_void_p), ("col_sq", C.c_void_p), ("planes_hi", C.c_void_p), ("planes_lo", C.c_void_p)]


class BnParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p)]


class HeadParams(C.Structure):
    _fields_ = [("glo_fc_w", C.c_void_p), ("glo_fc_b", C.c_void_p), ("glo_bn", BnParams),
                ("atte0_w", C.c_void_p), ("atte_bn1", BnParams), ("atte2_w", C.c_void_p), ("atte_bn3", BnParams),
                ("atte5_w", C.c_void_p), ("atte_bn6", BnParams),
                ("f1_w", C.c_void_p * 2), ("f1_b", C.c_void_p * 2), ("f2_w", C.c_void_p * 2), ("f2_b", C.c_void_p * 2),
                ("se1_w", C.c_void_p * 2), ("se2_w", C.c_void_p * 2),
                ("memo_conv1_w", C.c_void_p * 2), ("memo_bn1", BnParams * 2),
                ("memo_conv2_w", C.c_void_p * 2), ("memo_bn2", BnParams * 2),
                ("memo_conv3_w", C.c_void_p * 2), ("memo_bn3", BnParams * 2)]


class HeadGrads(C.Structure):
    _fields_ = [("glo_fc_w", C.c_void_p), ("glo_fc_b", C.c_void_p), ("glo_bn_w", C.c_void_p), ("glo_bn_b", C.c_void_p),
                ("atte0_w", C.c_void_p), ("atte_bn1_w", C.c_void_p), ("atte_bn1_b", C.c_void_p),
                ("atte2_w", C.c_void_p), ("atte_bn3_w", C.c_void_p), ("atte_bn3_b", C.c_void_p),
                ("atte5_w", C.c_void_p), ("atte_bn6_w", C.c_void_p), ("atte_bn6_b", C.c_void_p),
                ("f1_w", C.c_void_p * 2), ("f1_b", C.c_void_p * 2), ("f2_w", C.c_void_p * 2), ("f2_b", C.c_void_p * 2),
                ("se1_w", C.c_void_p * 2), ("se2_w", C.c_void_p * 2),
                ("memo_conv1_w", C.c_void_p * 2), ("memo_bn1_w", C.c_void_p * 2), ("memo_bn1_b", C.c_void_p * 2),
                ("memo_conv2_w", C.c_void_p * 2), ("memo_bn2_w", C.c_void_p * 2), ("memo_bn2_b", C.c_void_p * 2),
                ("memo_conv3_w", C.c_void_p * 2), ("memo_bn3_w", C.c_void_p * 2), ("memo_bn3_b", C.c_void_p * 2)]


class TailParams(C.Structure):
    _fields_ = [("corr_bn", BnParams), ("uncorr_bn", BnParams),
                ("featQ_w", C.c_void_p), ("featQ_b", C.c_void_p), ("featQ_bn", BnParams),
                ("featK_w", C.c_void_p), ("featK_b", C.c_void_p), ("featK_bn", BnParams)]


class SiameseParams(C.Structure):
    _fields_ = [("featQ_w", C.c_void_p), ("featQ_b", C.c_void_p), ("featQ_bn", BnParams),
                ("featK_w", C.c_void_p), ("featK_b", C.c_void_p), ("featK_bn", BnParams),
                ("cls_bn", BnParams), ("cls_w", C.c_void_p), ("cls_b", C.c_void_p)]


class SiameseGrads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("featQ_w", "featQ_b", "featQ_bn_w", "featQ_bn_b", "featK_w", "featK_b", "featK_bn_w",
                                          "featK_bn_b", "cls_bn_w", "cls_bn_b", "cls_w", "cls_b")]


_SIGNATURES = {
    "grl_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "grl_destroy": (None, [C.c_void_p]),
    "grl_last_error": (C.c_char_p, [C.c_void_p]),
    "grl_version": (C.c_char_p, []),
    "grl_num_sms": (C.c_int, [C.c_void_p]),
    "grl_launch_count": (C.c_longlong, [C.c_void_p]),
    "grl_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "grl_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "grl_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "grl_gemm_workspace_bytes": (C.c_size_t, [C.POINTER(GemmDesc)]),
    "grl_gemm_bf16x3": (C.c_int, [C.c_void_p, C.POINTER(GemmDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "grl_distance_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "grl_distance": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_cmc_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_cmc_map_sharded_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "grl_cmc_map_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "grl_argsort_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "grl_topk_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "grl_topk_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_topk_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "grl_gallery_prepared_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "grl_gallery_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_coarse_topk_prepared": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_dist_topk_prepared": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_topk_kprime": (C.c_int, [C.c_int]),
    "grl_coarse_topk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "grl_coarse_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_rescore": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                              C.c_int, C.c_void_p, C.c_void_p]),
    "grl_topk_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_exact_topk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "grl_exact_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_dist_topk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "grl_dist_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_comm_unique_id": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "grl_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "grl_comm_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "grl_comm_destroy": (C.c_int, [C.c_void_p]),
    "grl_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "grl_sharded_topk_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "grl_sharded_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_search_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "grl_merge_key_lists": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "grl_search_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "grl_rerank_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "grl_rerank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_head_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "grl_head_forward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "grl_head_backward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(HeadGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_basic_block_workspace_bytes": (C.c_size_t, [C.c_int]),
    "grl_basic_block_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BnParams), C.c_void_p, C.POINTER(BnParams), C.c_void_p,
                                          C.POINTER(BnParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "grl_gce_forward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "grl_gce_backward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(HeadGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_trl_forward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "grl_trl_backward": (C.c_int, [C.c_void_p, C.POINTER(HeadParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.POINTER(HeadGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_eval_descriptor_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "grl_eval_descriptor": (C.c_int, [C.c_void_p, C.POINTER(TailParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_longlong, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_oim_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_oim_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_triplet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_triplet_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_siamese_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "grl_siamese_forward": (C.c_int, [C.c_void_p, C.POINTER(SiameseParams), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_siamese_backward": (C.c_int, [C.c_void_p, C.POINTER(SiameseParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.POINTER(SiameseGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "grl_pair_loss_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_pair_loss_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grl_head_ws_lookup": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
}

_lib = None
_lock = threading.Lock()
_handles = {}


def load_library() -> C.CDLL:
    """dlopen the in-tree library and attach every signature declared in include/grl_b200.h."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "grl_b200: %s is missing. Build it with `python -m grl_b200.build` (or "
                    "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError == header/library mismatch: fail loudly
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES.keys())


class GrlError(RuntimeError):
    pass


def check(handle, code: int, what: str):
    if code != GRL_OK:
        msg = load_library().grl_last_error(handle)
        raise GrlError("%s failed with %s: %s" % (what, _CODES.get(code, code), msg.decode() if msg else ""))


def get_handle(device=None) -> C.c_void_p:
    """One grl_handle per CUDA device (created lazily; DataParallel threads share it safely for launches)."""
    if not torch.cuda.is_available():
        raise RuntimeError("grl_b200 needs a CUDA device (sm_100a); no CPU path exists.")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    with _lock:
        h = _handles.get(idx)
    if h is None:
        lib = load_library()
        torch.cuda.init()
        out = C.c_void_p()
        code = lib.grl_create(idx, C.byref(out))
        if code != GRL_OK:
            raise GrlError("grl_create(%d) failed with %s: %s" % (idx, _CODES.get(code, code), lib.grl_last_error(None).decode()))
        with _lock:
            _handles[idx] = out
        h = out
    return h


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def launch_count(device=None) -> int:
    return int(load_library().grl_launch_count(get_handle(device)))
