"""CPU: pin the loss oracle to the real reference (tests/golden/loss_*.npz, oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth
from oracle import loss_oracle as lo


@pytest.mark.parametrize("name", ["loss_b32", "loss_b12"])
def test_loss_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, D, C = int(g["B"]), int(g["D"]), int(g["C"])
    feat, ids, lut, targets = synth.make_loss_inputs(B, D, C, int(g["seed"]), int(g["n_ids"]))
    gw = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
    for margin, tag in (('soft', 'soft'), (0.3, 'm03')):
        b_loss, dfeat = lo.triplet_loss(feat.double(), ids, margin, gw)
        assert np.allclose(b_loss.numpy(), g["tri_%s_loss" % tag], rtol=0, atol=1e-12)
        assert np.allclose(dfeat.numpy(), g["tri_%s_dfeat" % tag], rtol=0, atol=1e-12)
    loss, logits, dx, new_lut = lo.oim_loss(feat.double(), targets, lut.double(), 30.0, 0.5)
    assert abs(float(loss) - float(g["oim_loss"])) < 1e-12
    assert np.allclose(logits.numpy(), g["oim_logits"], atol=1e-12) and np.allclose(dx.numpy(), g["oim_dx"], atol=1e-12)
    # the table update: rows of identities in the batch are unit-norm blends, all others untouched
    touched = np.unique(targets.numpy())
    rest = np.setdiff1d(np.arange(C), touched)
    assert np.array_equal(new_lut.numpy()[rest], lut.double().numpy()[rest])
    assert np.allclose(np.linalg.norm(new_lut.numpy()[touched], axis=1), 1.0, atol=1e-12)
    # OIM.backward as executed by the REAL reference method (unbound call on a stub, oracle/make_golden.py): the input
    # gradient and every table row it rewrote
    assert np.allclose(dx.numpy(), g["oim_bwd_dx"], rtol=0, atol=1e-12)
    assert np.array_equal(g["oim_new_lut_rows"], touched)
    assert np.allclose(new_lut.numpy()[touched], g["oim_new_lut_vals"], rtol=0, atol=1e-12)
