"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the head's loss neighbours.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

  oim_loss      /root/reference/reid/loss/oim.py:8-58   OIM.forward (:12-16), OIM.backward (:18-27), OIMLoss.forward (:50-58)
  triplet_loss  /root/reference/reid/loss/triplet.py:16-90 with the trainer's configuration TripletLoss('soft', True)(feat, id)
                (reid/train/trainer.py:12,139): mode 'id', dis_func 'eu', n_dis 0, batch_hard

Pinning: `triplet_loss` is checked against the REAL reference class (values and autograd gradients, tests/golden/loss_*.npz,
oracle/make_golden.py).  The reference's `OIM` is a legacy autograd.Function with a non-static forward, which PyTorch >= 1.5
refuses to execute, so its backward contract (input gradient from the table BEFORE the update, then the sequential momentum
update) is restated here line by line and is NOT pinned by an execution of the reference: "parity unpinned" for OIM.backward;
its forward (`inputs.mm(lut.t())`, scaling, F.cross_entropy) is plain torch and is what the golden file holds.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def oim_loss(inputs, targets, lut, scalar=1.0, momentum=0.5, d_loss=1.0):
    """Returns (loss, scaled logits, d inputs, updated lut); `lut` itself is not modified."""
    x = inputs.detach().clone().requires_grad_(True)
    logits = x.mm(lut.t())                               # OIM.forward, oim.py:15
    logits = logits * scalar                             # OIMLoss.forward: inputs *= self.scalar, :54
    loss = F.cross_entropy(logits, targets)              # :56
    (loss * d_loss).backward()                           # == grad_outputs.mm(self.lut) through the scaling, :22
    new_lut = lut.clone()
    for xi, y in zip(inputs.detach(), targets):          # :24-26, in batch order
        new_lut[y] = momentum * new_lut[y] + (1. - momentum) * xi
        new_lut[y] /= new_lut[y].norm()
    return loss.detach(), logits.detach(), x.grad.detach(), new_lut


def triplet_loss(feat, ids, margin='soft', d_loss=None):
    """triplet.py:16-90 (batch_hard, mode 'id', dis_func 'eu', n_dis 0).  Returns (b_loss [B], d feat)."""
    f = feat.detach().clone().requires_grad_(True)
    diff = f.unsqueeze(1) - f.unsqueeze(0)
    dist = ((diff ** 2).sum(2) + 1e-12).sqrt()           # cdist, :88-90
    same = ids.unsqueeze(1) == ids.unsqueeze(0)
    eye = torch.eye(f.size(0), dtype=torch.bool)
    pos = (same & ~eye).to(dist.dtype)                   # same_id_mask ^ identity_mask, :32
    max_positive = (dist * pos).max(1)[0]                # :52-53
    min_negative = (dist + 1e5 * same.to(dist.dtype)).min(1)[0]   # :55-57
    z = max_positive - min_negative
    b_loss = torch.log(1 + torch.exp(z)) if margin == 'soft' else torch.clamp(z + margin, min=0)
    g = torch.ones_like(b_loss) if d_loss is None else d_loss
    (b_loss * g).sum().backward()
    return b_loss.detach(), f.grad.detach()
