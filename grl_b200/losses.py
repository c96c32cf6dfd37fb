"""Drop-in mirrors of the head's loss neighbours on B200.

  OIMLoss      reid/loss/oim.py:33-58      same constructor, buffers (`lut`) and `forward(inputs, targets) -> (loss, logits)`
  TripletLoss  reid/loss/triplet.py:5-90   ('soft' | float margin, batch_hard=True) `forward(feat, id) -> b_loss [B]`

The reference's `OIM` is a legacy (non-static) autograd.Function that current PyTorch refuses to run; the Function below
keeps its contract: backward returns the input gradient computed with the table BEFORE the update, then applies the
momentum update sample by sample in batch order.  Everything runs through the C ABI (csrc/losses.cu); no CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


class _OIMLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, lut, scalar, momentum):
        x = inputs.contiguous().float()
        t = targets.contiguous().long()
        B, D = x.shape
        C = lut.size(0)
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            h = _lib.get_handle(x.device)
            logits = torch.empty((B, C), dtype=torch.float32, device=x.device)
            probs = torch.empty_like(logits)
            row_loss = torch.empty(B, dtype=torch.float32, device=x.device)
            loss = torch.empty((), dtype=torch.float32, device=x.device)
            _lib.check(h, lib.grl_oim_forward(h, x.data_ptr(), t.data_ptr(), lut.data_ptr(), B, C, D, float(scalar), logits.data_ptr(),
                                              probs.data_ptr(), row_loss.data_ptr(), loss.data_ptr(), _lib.stream_ptr(x.device)),
                       "grl_oim_forward")
        ctx.save_for_backward(x, t, probs)
        ctx.lut, ctx.scalar, ctx.momentum = lut, float(scalar), float(momentum)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, d_loss, _d_logits):
        x, t, probs = ctx.saved_tensors
        B, D = x.shape
        lut = ctx.lut
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            h = _lib.get_handle(x.device)
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            g = d_loss.contiguous().float().reshape(1)
            _lib.check(h, lib.grl_oim_backward(h, x.data_ptr(), t.data_ptr(), lut.data_ptr(), probs.data_ptr(), B, lut.size(0), D,
                                               ctx.scalar, ctx.momentum, g.data_ptr(), _lib.ptr(dx), _lib.stream_ptr(x.device)),
                       "grl_oim_backward")
        return dx, None, None, None, None


class OIMLoss(nn.Module):
    def __init__(self, num_features, num_classes, scalar=1.0, momentum=0.5, weight=None, size_average=True):
        super(OIMLoss, self).__init__()
        if weight is not None:
            raise NotImplementedError("class weights are never used by the reference (mars_train.py:85-86)")
        self.num_features = num_features
        self.num_classes = num_classes
        self.momentum = momentum
        self.scalar = scalar
        self.weight = weight
        self.register_buffer('lut', torch.zeros(num_classes, num_features))
        self.size_average = size_average

    #: F.cross_entropy / `self.lut[y]` in the reference raise on a label outside [0, num_classes); the kernels index the table with
    #: it, so the range is checked first (one tiny device reduction + a read-back; set False to skip it in a tuned loop)
    check_targets = True

    def forward(self, inputs, targets):
        if self.check_targets and targets.numel():
            lo, hi = int(targets.min()), int(targets.max())
            if lo < 0 or hi >= self.num_classes:
                raise IndexError("OIMLoss: target %d is out of bounds for %d classes" % (lo if lo < 0 else hi, self.num_classes))
        loss, logits = _OIMLossFunction.apply(inputs, targets, self.lut, self.scalar, self.momentum)
        return loss, logits


class _TripletFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, ids, soft, margin):
        f = feat.contiguous().float()
        t = ids.contiguous().long()
        B, D = f.shape
        lib = _lib.load_library()
        with torch.cuda.device(f.device):
            h = _lib.get_handle(f.device)
            dev = f.device
            loss = torch.empty(B, dtype=torch.float32, device=dev)
            z = torch.empty(B, dtype=torch.float32, device=dev)
            pi = torch.empty(B, dtype=torch.int32, device=dev)
            ni = torch.empty(B, dtype=torch.int32, device=dev)
            pd = torch.empty(B, dtype=torch.float32, device=dev)
            nd = torch.empty(B, dtype=torch.float32, device=dev)
            _lib.check(h, lib.grl_triplet_forward(h, f.data_ptr(), t.data_ptr(), B, D, int(soft), float(margin), loss.data_ptr(), z.data_ptr(),
                                                  pi.data_ptr(), ni.data_ptr(), pd.data_ptr(), nd.data_ptr(), _lib.stream_ptr(dev)),
                       "grl_triplet_forward")
        ctx.save_for_backward(f, z, pi, ni, pd, nd)
        ctx.soft, ctx.margin = int(soft), float(margin)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        f, z, pi, ni, pd, nd = ctx.saved_tensors
        B, D = f.shape
        lib = _lib.load_library()
        with torch.cuda.device(f.device):
            h = _lib.get_handle(f.device)
            g = d_loss.contiguous().float()
            df = torch.empty_like(f)
            _lib.check(h, lib.grl_triplet_backward(h, f.data_ptr(), B, D, ctx.soft, ctx.margin, z.data_ptr(), pi.data_ptr(), ni.data_ptr(),
                                                   pd.data_ptr(), nd.data_ptr(), g.data_ptr(), df.data_ptr(), _lib.stream_ptr(f.device)),
                       "grl_triplet_backward")
        return df, None, None, None


class TripletLoss(nn.Module):
    def __init__(self, margin=0, batch_hard=False, dim=2048):
        super(TripletLoss, self).__init__()
        self.batch_hard = batch_hard
        if isinstance(margin, float) or margin == 'soft':
            self.margin = margin
        else:
            raise NotImplementedError('The margin {} is not recognized in TripletLoss()'.format(margin))

    def forward(self, feat, id=None, pos_mask=None, neg_mask=None, mode='id', dis_func='eu', n_dis=0):
        if mode != 'id' or dis_func != 'eu' or n_dis != 0 or not self.batch_hard:
            raise NotImplementedError("only the configuration the trainer uses: TripletLoss('soft', True)(feat, id) (trainer.py:12,139)")
        if id is None:
            raise RuntimeError('foward is in id mode, please input id!')
        soft = self.margin == 'soft'
        return _TripletFunction.apply(feat, id, soft, 0.0 if soft else self.margin)
