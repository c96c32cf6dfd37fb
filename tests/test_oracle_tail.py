"""CPU: the eval-feature-tail oracle is pinned to outputs of the REAL reference modules (tests/golden/tail_*.npz,
made by oracle/make_golden.py: ResNet50_GRL_Model.corr_bn/uncorr_bn + Siamese.self_attention in eval mode)."""
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth
from oracle import tail_oracle as to


@pytest.mark.parametrize("name", ["tail_n5t8", "tail_n3t16"])
def test_tail_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, T = int(g["n"]), int(g["T"])
    p = synth.make_tail_params(10, dtype=torch.float64)
    fu, fc = synth.make_tail_input(n, T, dtype=torch.float64)
    out = to.ref_descriptor(p, fu, fc)
    assert out.shape == (n, 6144)
    assert np.abs(out.numpy() - g["out_feat"]).max() < 1e-13
    xu, xc = to.ref_tail(p, fu, fc)
    assert np.abs(to.ref_self_attention(p, xc).numpy() - g["out_frame"]).max() < 1e-13
    assert np.abs(out.mean(0).numpy() - g["tracklet"]).max() < 1e-13     # attevaluator.py:83-84 mean over clips
    # the three parts are unit-norm-ish, the concat is not
    assert abs(float(out[:, :2048].norm(dim=1).mean()) - 1) < 1e-9 and abs(float(out[:, 2048:4096].norm(dim=1).mean()) - 1) < 1e-9
