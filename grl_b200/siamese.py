"""Host-side mirror of the reference's Siamese module (temporal self-attention used by the evaluator) and the fused
eval descriptor, on top of the C ABI.

  Siamese(input_num, output_num, class_num)                  reid/models/Siamese.py:43-76   (same attributes => same state_dict)
  Siamese.self_attention(input [n, T, 2048]) -> [n, 2048]    reid/models/Siamese.py:79-106
  eval_descriptor(model, siamese, f_uncorr, f_corr)          grl_model.py:222-226 + Siamese.py:79-106 + attevaluator.py:79-80

Evaluation only (eval-mode BatchNorm, no autograd): training-time use of the verification head (Siamese.forward) is a
loss-side neighbour outside the hot path.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


def _bn(mod):
    r = _lib.BnParams()
    r.weight, r.bias = mod.weight.data_ptr(), mod.bias.data_ptr()
    r.running_mean, r.running_var = mod.running_mean.data_ptr(), mod.running_var.data_ptr()
    return r


def _run(params, f_uncorr, f_corr, apply_tail_bn):
    if not f_corr.is_cuda:
        raise RuntimeError("grl_b200 eval descriptor needs CUDA tensors (no CPU path exists)")
    f_corr = f_corr.contiguous().float()
    f_uncorr = f_uncorr.contiguous().float()
    n, T, c = f_corr.shape
    if c != 2048 or tuple(f_uncorr.shape) != (n, 2048):
        raise RuntimeError("eval descriptor expects f_uncorr [n, 2048] and f_corr [n, T, 2048]")
    lib = _lib.load_library()
    dev = f_corr.device
    with torch.cuda.device(dev):
        h = _lib.get_handle(dev)
        out = torch.empty((n, 6144), device=dev)
        nbytes = lib.grl_eval_descriptor_workspace_bytes(n, T)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        rc = lib.grl_eval_descriptor(h, C.byref(params), f_uncorr.data_ptr(), f_corr.data_ptr(), n, T, 1 if apply_tail_bn else 0,
                                     out.data_ptr(), 6144, ws.data_ptr(), nbytes, _lib.stream_ptr(dev))
        _lib.check(h, rc, "grl_eval_descriptor")
    return out


class Siamese(nn.Module):
    """reid/models/Siamese.py:43-76 (parameter container with the reference's attribute names)."""

    def __init__(self, input_num, output_num, class_num):
        super(Siamese, self).__init__()
        if input_num != 2048 or output_num != 512:
            raise RuntimeError("grl_b200 Siamese is built for input_num=2048, output_num=512 (mars_train.py:77)")
        self.input_num, self.output_num, self.class_num, self.feat_num = input_num, output_num, class_num, input_num
        self.featQ = nn.Linear(input_num, output_num)
        self.featQ_bn = nn.BatchNorm1d(output_num)
        self.featK = nn.Linear(input_num, output_num)
        self.featK_bn = nn.BatchNorm1d(output_num)
        self.featV = nn.Linear(input_num, output_num)          # present in the reference's state_dict, unused by self_attention
        self.featV_bn = nn.BatchNorm1d(output_num)
        self.softmax = nn.Softmax(dim=-1)
        self.classifierBN = nn.BatchNorm1d(self.feat_num)
        self.classifierlinear = nn.Linear(self.feat_num, class_num)

    def _params(self, model=None):
        for t in (self.featQ.weight, self.featQ.bias, self.featK.weight, self.featK.bias):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("grl_b200 Siamese: parameters must be contiguous float32 CUDA tensors")
        p = _lib.TailParams()
        p.featQ_w, p.featQ_b, p.featQ_bn = self.featQ.weight.data_ptr(), self.featQ.bias.data_ptr(), _bn(self.featQ_bn)
        p.featK_w, p.featK_b, p.featK_bn = self.featK.weight.data_ptr(), self.featK.bias.data_ptr(), _bn(self.featK_bn)
        if model is not None:
            p.corr_bn, p.uncorr_bn = _bn(model.corr_bn), _bn(model.uncorr_bn)
        return p

    @torch.no_grad()
    def self_attention(self, input):
        """Siamese.py:79-106 in eval mode: temporal self-attention pooling of x_corr [n, T, 2048] -> [n, 2048]."""
        if self.training:
            raise RuntimeError("grl_b200 Siamese.self_attention is the evaluation path (call .eval()); training-mode "
                               "BatchNorm over Q/K is part of the loss-side graph, outside this library")
        n = input.size(0)
        dummy = input.new_zeros((n, 2048))
        return _run(self._params(), dummy, input, apply_tail_bn=False)[:, 2048:4096].contiguous()


@torch.no_grad()
def eval_descriptor(model, siamese, f_uncorr, f_corr):
    """Head outputs (model.head(...)[0:2], before corr_bn / uncorr_bn) -> [n, 6144] per-clip descriptors, one fused call."""
    if model.training or siamese.training:
        raise RuntimeError("eval_descriptor is the evaluation path: call .eval() on both models")
    return _run(siamese._params(model), f_uncorr, f_corr, apply_tail_bn=True)
