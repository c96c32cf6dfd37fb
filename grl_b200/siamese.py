"""Host-side mirror of the reference's Siamese module (temporal self-attention used by the evaluator) and the fused
eval descriptor, on top of the C ABI.

  Siamese(input_num, output_num, class_num)                  reid/models/Siamese.py:43-76   (same attributes => same state_dict)
  Siamese.self_attention(input [n, T, 2048]) -> [n, 2048]    reid/models/Siamese.py:79-106
  eval_descriptor(model, siamese, f_uncorr, f_corr)          grl_model.py:222-226 + Siamese.py:79-106 + attevaluator.py:79-80

  Siamese.forward(x) -> (cls_encode [n, n, 2], siamese_out [2n, 2048])   reid/models/Siamese.py:108-142, with autograd
  PairLoss.forward(score, tar_probe, tar_gallery) -> (loss, prec)        reid/loss/pairloss.py:19-48

`self_attention` / `eval_descriptor` are the evaluation path (eval-mode BatchNorm, no autograd); `forward` is the
verification head as the trainer calls it (reid/train/trainer.py:137, train- or eval-mode BatchNorm, backward in train
mode).  No CPU fallback.
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


_SIA_PARAM_NAMES = ("featQ.weight", "featQ.bias", "featQ_bn.weight", "featQ_bn.bias", "featK.weight", "featK.bias", "featK_bn.weight",
                    "featK_bn.bias", "classifierBN.weight", "classifierBN.bias", "classifierlinear.weight", "classifierlinear.bias")
_SIA_GRAD_FIELDS = ("featQ_w", "featQ_b", "featQ_bn_w", "featQ_bn_b", "featK_w", "featK_b", "featK_bn_w", "featK_bn_b",
                    "cls_bn_w", "cls_bn_b", "cls_w", "cls_b")


def _siamese_params(mod):
    p = _lib.SiameseParams()
    p.featQ_w, p.featQ_b, p.featQ_bn = mod.featQ.weight.data_ptr(), mod.featQ.bias.data_ptr(), _bn(mod.featQ_bn)
    p.featK_w, p.featK_b, p.featK_bn = mod.featK.weight.data_ptr(), mod.featK.bias.data_ptr(), _bn(mod.featK_bn)
    p.cls_bn = _bn(mod.classifierBN)
    p.cls_w, p.cls_b = mod.classifierlinear.weight.data_ptr(), mod.classifierlinear.bias.data_ptr()
    return p


class _SiameseFunction(torch.autograd.Function):
    """(x [2n, T, 2048], *params) -> (cls_encode [n, n, 2], siamese_out [2n, 2048]) through grl_siamese_forward/backward."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        n2, T, D = x.shape
        n = n2 // 2
        xr = torch.cat((x[0::2], x[1::2]), 0).contiguous().float()       # x.view(n, 2, T, -1): probe = [:, 0], gallery = [:, 1] (:117-123)
        lib = _lib.load_library()
        dev = x.device
        with torch.cuda.device(dev):
            h = _lib.get_handle(dev)
            nbytes = lib.grl_siamese_workspace_bytes(n2, T)
            if nbytes == 0:
                raise RuntimeError("grl_b200 Siamese.forward: need an even batch and seq_len <= 32")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            cls = torch.empty((n, n, 2), device=dev)
            out = torch.empty((n2, D), device=dev)
            sp = _siamese_params(mod)
            _lib.check(h, lib.grl_siamese_forward(h, C.byref(sp), xr.data_ptr(), n2, T, 1 if mod.training else 0, cls.data_ptr(),
                                                  out.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr(dev)), "grl_siamese_forward")
        ctx.mod, ctx.ws, ctx.train = mod, ws, mod.training
        ctx.save_for_backward(xr, out, *params)
        return cls, out

    @staticmethod
    def backward(ctx, d_cls, d_out):
        if not ctx.train:
            raise RuntimeError("grl_b200 Siamese: backward with eval-mode BatchNorm is not supported")
        xr, out, *params = ctx.saved_tensors
        mod = ctx.mod
        n2, T, D = xr.shape
        n = n2 // 2
        lib = _lib.load_library()
        dev = xr.device
        with torch.cuda.device(dev):
            h = _lib.get_handle(dev)
            d_cls = torch.zeros((n, n, 2), device=dev) if d_cls is None else d_cls.contiguous().float()
            d_out = None if d_out is None else d_out.contiguous().float()
            dxr = torch.empty_like(xr)
            grads = [torch.empty_like(t) for t in params]
            g = _lib.SiameseGrads()
            for f, t in zip(_SIA_GRAD_FIELDS, grads):
                setattr(g, f, t.data_ptr())
            sp = _siamese_params(mod)
            _lib.check(h, lib.grl_siamese_backward(h, C.byref(sp), xr.data_ptr(), out.data_ptr(), n2, T, d_cls.data_ptr(), _lib.ptr(d_out),
                                                   dxr.data_ptr(), C.byref(g), ctx.ws.data_ptr(), ctx.ws.numel(), _lib.stream_ptr(dev)),
                       "grl_siamese_backward")
        dx = torch.empty_like(dxr)
        dx[0::2] = dxr[:n]
        dx[1::2] = dxr[n:]
        return (None, dx) + tuple(grads)


def _siamese_forward(self, x):
    """Siamese.py:108-142.  x [2n, T, 2048] (samples 2i and 2i+1 form probe / gallery) -> (cls_encode, siamese_out)."""
    if not x.is_cuda:
        raise RuntimeError("grl_b200 Siamese needs CUDA tensors (no CPU path exists)")
    if x.size(0) % 2 != 0:
        raise RuntimeError("the batch size should be even number!")
    sd = dict(self.named_parameters())
    params = [sd[k] for k in _SIA_PARAM_NAMES]
    cls, out = _SiameseFunction.apply(self, x, *params)
    if self.training:
        with torch.no_grad():                                  # two self_attention calls and one classifierBN call (:124-125, :138)
            self.featQ_bn.num_batches_tracked += 2
            self.featK_bn.num_batches_tracked += 2
            self.classifierBN.num_batches_tracked += 1
    return cls, out


Siamese.forward = _siamese_forward


class _PairLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, tar_probe, tar_gallery):
        s = score.contiguous().float()
        n = s.size(0)
        tp, tg = tar_probe.contiguous().long(), tar_gallery.contiguous().long()
        lib = _lib.load_library()
        with torch.cuda.device(s.device):
            h = _lib.get_handle(s.device)
            loss = torch.empty((), device=s.device)
            prec = torch.empty((), device=s.device)
            _lib.check(h, lib.grl_pair_loss_forward(h, s.data_ptr(), tp.data_ptr(), tg.data_ptr(), n, loss.data_ptr(), prec.data_ptr(),
                                                    _lib.stream_ptr(s.device)), "grl_pair_loss_forward")
        ctx.save_for_backward(s, tp, tg)
        ctx.mark_non_differentiable(prec)
        return loss, prec

    @staticmethod
    def backward(ctx, d_loss, _d_prec):
        s, tp, tg = ctx.saved_tensors
        n = s.size(0)
        lib = _lib.load_library()
        with torch.cuda.device(s.device):
            h = _lib.get_handle(s.device)
            g = d_loss.contiguous().float().reshape(1)
            ds = torch.empty_like(s)
            _lib.check(h, lib.grl_pair_loss_backward(h, s.data_ptr(), tp.data_ptr(), tg.data_ptr(), n, g.data_ptr(), ds.data_ptr(),
                                                     _lib.stream_ptr(s.device)), "grl_pair_loss_backward")
        return ds, None, None


class PairLoss(nn.Module):
    """reid/loss/pairloss.py:11-48: BCE of the pair scores [n, n] (already softmax-ed by the trainer, trainer.py:142-147) against
    identity equality; returns (loss, prec) like the reference."""

    def forward(self, score, tar_probe, tar_gallery):
        if score.dim() != 2 or score.size(0) != score.size(1):
            raise RuntimeError("PairLoss expects a square [n, n] score matrix")
        return _PairLossFunction.apply(score, tar_probe, tar_gallery)


@torch.no_grad()
def eval_descriptor(model, siamese, f_uncorr, f_corr):
    """Head outputs (model.head(...)[0:2], before corr_bn / uncorr_bn) -> [n, 6144] per-clip descriptors, one fused call."""
    if model.training or siamese.training:
        raise RuntimeError("eval_descriptor is the evaluation path: call .eval() on both models")
    return _run(siamese._params(model), f_uncorr, f_corr, apply_tail_bn=True)
