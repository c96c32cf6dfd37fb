"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the head's loss neighbours.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

  oim_loss      /root/reference/reid/loss/oim.py:8-58   OIM.forward (:12-16), OIM.backward (:18-27), OIMLoss.forward (:50-58)
  triplet_loss  /root/reference/reid/loss/triplet.py:16-90 with the trainer's configuration TripletLoss('soft', True)(feat, id)
                (reid/train/trainer.py:12,139): mode 'id', dis_func 'eu', n_dis 0, batch_hard

Pinning: `triplet_loss` is checked against the REAL reference class (values and autograd gradients, tests/golden/loss_*.npz,
oracle/make_golden.py).  The reference's `OIM` is a legacy autograd.Function with a non-static forward, which PyTorch >= 1.5
refuses to *apply*; its `backward` (input gradient from the table BEFORE the update, then the sequential momentum update,
oim.py:18-27) is nevertheless a plain method, so oracle/make_golden.py executes the REAL `OIM.backward` unbound on a stub that
carries the attributes it reads (saved_tensors, needs_input_grad, lut, momentum) and the golden files hold its `grad_inputs`
and the table rows it rewrote: `oim_loss` below is pinned to them (tests/test_oracle_loss.py).  The forward
(`inputs.mm(lut.t())`, scaling, F.cross_entropy) is plain torch and is pinned the same way.
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


# --------------------------------------------------------------------------------------------
# Verification head in training: Siamese.forward (reid/models/Siamese.py:79-142) and PairLoss (reid/loss/pairloss.py:19-48).
# Functional restatement over a {state_dict key: tensor} dict (BN running buffers updated in place like nn.BatchNorm1d);
# pinned to the REAL reference modules by tests/golden/siamese_*.npz.
# --------------------------------------------------------------------------------------------
def _sia_bn(p, prefix, x, training):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"], p[prefix + ".bias"],
                        training, 0.1, 1e-5)


def _sia_self_attention(p, inp, training):                 # Siamese.py:79-106
    batch, length = inp.size(0), inp.size(1)
    flat = inp.reshape(batch * length, -1)
    q = _sia_bn(p, "featQ_bn", F.linear(flat, p["featQ.weight"], p["featQ.bias"]), training)
    q = (q / q.norm(2, 1).unsqueeze(1)).view(batch, length, -1)
    k = _sia_bn(p, "featK_bn", F.linear(flat, p["featK.weight"], p["featK.bias"]), training)
    k = (k / k.norm(2, 1).unsqueeze(1)).view(batch, length, -1)
    w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)
    pool = torch.matmul(w, inp).sum(1)
    return pool / pool.norm(2, 1).unsqueeze(1)


def siamese_forward(p, x, training=True):
    """Siamese.py:108-142: returns (cls_encode [n, n, 2], siamese_out [2n, D])."""
    n2, T = x.size(0), x.size(1)
    xv = x.view(n2 // 2, 2, T, -1)
    pp = _sia_self_attention(p, xv[:, 0].contiguous(), training)
    pg = _sia_self_attention(p, xv[:, 1].contiguous(), training)
    out = torch.cat((pp, pg))
    diff = (pp.unsqueeze(1) - pg.unsqueeze(0)) ** 2
    y = _sia_bn(p, "classifierBN", diff.view(pp.size(0) * pg.size(0), -1), training)
    z = F.linear(y, p["classifierlinear.weight"], p["classifierlinear.bias"])
    return z.view(pp.size(0), pg.size(0), -1), out


def pair_loss(score, tar_probe, tar_gallery):
    """pairloss.py:19-48: (loss, prec); label[i][j] = (tar_probe[j] == tar_gallery[i])."""
    n = score.size(0)
    mask = tar_probe.unsqueeze(0).expand(n, n).eq(tar_gallery.unsqueeze(1).expand(n, n)).reshape(-1)
    s = score.reshape(-1)
    loss = F.binary_cross_entropy(s, mask.to(s.dtype))
    prec = ((s.detach() > 1 - s.detach()) == mask).to(s.dtype).mean()
    return loss, prec
