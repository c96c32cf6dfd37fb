"""GPU parity: OIMLoss / TripletLoss mirrors (grl_oim_*, grl_triplet_* through the C ABI) vs the reference's golden outputs
and the fp64 oracle.  Tolerance: 1e-5 relative (fp32 arithmetic, north-star bar for head-adjacent tensors is 1e-3)."""
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("name", ["loss_b32", "loss_b12"])
def test_triplet_matches_reference_golden(golden_dir, name):
    from grl_b200.losses import TripletLoss
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, D, C = int(g["B"]), int(g["D"]), int(g["C"])
    feat, ids, _, _ = synth.make_loss_inputs(B, D, C, int(g["seed"]), int(g["n_ids"]))
    gw = torch.linspace(0.5, 1.5, B).cuda()
    for margin, tag in (('soft', 'soft'), (0.3, 'm03')):
        f = feat.cuda().requires_grad_(True)
        b_loss = TripletLoss(margin, True)(f, ids.cuda())
        assert b_loss.shape == (B,) and b_loss.requires_grad
        (b_loss * gw).sum().backward()
        assert np.abs(b_loss.detach().cpu().numpy() - g["tri_%s_loss" % tag]).max() < 2e-6
        assert rel(f.grad.cpu().numpy(), g["tri_%s_dfeat" % tag]) < 1e-5


@pytest.mark.parametrize("name", ["loss_b32", "loss_b12"])
def test_oim_matches_golden_and_oracle(golden_dir, name):
    from grl_b200.losses import OIMLoss
    from oracle import loss_oracle as lo
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, D, C = int(g["B"]), int(g["D"]), int(g["C"])
    feat, _, lut, targets = synth.make_loss_inputs(B, D, C, int(g["seed"]), int(g["n_ids"]))
    crit = OIMLoss(D, C, scalar=30.0, momentum=0.5).cuda()
    assert list(crit.state_dict().keys()) == ["lut"]
    crit.lut.copy_(lut)
    x = feat.cuda().requires_grad_(True)
    loss, logits = crit(x, targets.cuda())
    assert logits.shape == (B, C) and not logits.requires_grad
    assert abs(float(loss.detach()) - float(g["oim_loss"])) < 2e-5 * abs(float(g["oim_loss"]))
    assert rel(logits.cpu().numpy(), g["oim_logits"]) < 1e-5
    before = crit.lut.clone()
    (loss * 0.7).backward()                               # a non-unit upstream gradient
    assert rel(x.grad.cpu().numpy(), 0.7 * g["oim_dx"]) < 1e-5
    _, _, _, new_lut = lo.oim_loss(feat.double(), targets, lut.double(), 30.0, 0.5)
    assert rel(crit.lut.cpu().numpy(), new_lut.numpy()) < 1e-6
    # ... and against the REAL reference's OIM.backward (executed unbound on a stub, oracle/make_golden.py)
    assert rel(x.grad.cpu().numpy(), 0.7 * g["oim_bwd_dx"]) < 1e-5
    rows = g["oim_new_lut_rows"]
    assert rel(crit.lut.cpu().numpy()[rows], g["oim_new_lut_vals"]) < 1e-6
    rest = np.setdiff1d(np.arange(C), rows)
    assert np.array_equal(crit.lut.cpu().numpy()[rest], lut.numpy()[rest])
    assert not torch.equal(before, crit.lut)              # the update happened in backward, not in forward
    # a second step composes on the updated table, like the reference's persistent buffer
    x2 = feat.cuda().requires_grad_(True)
    loss2, _ = crit(x2, targets.cuda())
    loss2.backward()
    l2, _, dx2, lut2 = lo.oim_loss(feat.double(), targets, new_lut, 30.0, 0.5)
    # after one update the table rows ARE the batch's blends: the loss collapses to ~1e-8, so absolute tolerances here
    assert abs(float(loss2) - float(l2)) < 1e-6 and np.abs(x2.grad.cpu().numpy() - dx2.numpy()).max() < 1e-6
    assert rel(crit.lut.cpu().numpy(), lut2.numpy()) < 1e-6


def test_oim_frame_level_shape():
    """trainer.py:118-127: frame-level loss, B*T = 256 rows of 2048-d against 625 identities, every id repeated T times."""
    from grl_b200.losses import OIMLoss
    from oracle import loss_oracle as lo
    feat, ids, lut, targets = synth.make_loss_inputs(32, 2048, 625, 3, 8)
    T = 8
    g = torch.Generator().manual_seed(5)
    frames = torch.nn.functional.normalize(feat.unsqueeze(1) + 0.1 * torch.randn((32, T, 2048), generator=g), dim=2).reshape(32 * T, 2048)
    tx = targets.unsqueeze(1).expand(32, T).reshape(-1)
    crit = OIMLoss(2048, 625, scalar=30.0, momentum=0.5).cuda()
    crit.lut.copy_(lut)
    x = frames.cuda().requires_grad_(True)
    loss, _ = crit(x, tx.cuda())
    loss.backward()
    l, _, dx, new_lut = lo.oim_loss(frames.double(), tx, lut.double(), 30.0, 0.5)
    assert abs(float(loss) - float(l)) < 2e-5 * abs(float(l))
    assert rel(x.grad.cpu().numpy(), dx.numpy()) < 1e-5 and rel(crit.lut.cpu().numpy(), new_lut.numpy()) < 1e-6


def test_losses_reject_unsupported_configurations():
    from grl_b200.losses import OIMLoss, TripletLoss
    with pytest.raises(NotImplementedError):
        TripletLoss('hard', True)
    with pytest.raises(NotImplementedError):
        TripletLoss('soft', False)(torch.zeros((4, 8), device="cuda"), torch.zeros(4, dtype=torch.long, device="cuda"))
    with pytest.raises(NotImplementedError):
        OIMLoss(8, 4, weight=torch.ones(4))


def test_oim_rejects_out_of_range_targets():
    """A label outside [0, C) raises (the reference's F.cross_entropy / lut[y] do), instead of indexing the table with it."""
    from grl_b200.losses import OIMLoss
    crit = OIMLoss(64, 10, scalar=30.0).cuda()
    x = torch.randn(4, 64, device="cuda")
    for bad in (torch.tensor([0, 1, 10, 2]), torch.tensor([0, -1, 3, 2])):
        with pytest.raises(IndexError):
            crit(x, bad.cuda())
