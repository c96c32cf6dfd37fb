"""CPU: pin the verification-head oracle (Siamese.forward in training, PairLoss) to the real reference's outputs."""
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth
from helpers_sample import grad_sample
from oracle import loss_oracle as lo


@pytest.mark.parametrize("name", ["siamese_n32t8", "siamese_n6t3"])
def test_siamese_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n2, T = int(g["n2"]), int(g["T"])
    params, x, d_cls, d_out, tar = synth.make_siamese_inputs(n2, T, int(g["seed"]))
    p = {k: (v.double().clone().requires_grad_(v.is_floating_point() and "running" not in k) if v.is_floating_point() else v.clone())
         for k, v in params.items()}
    xin = x.double().clone().requires_grad_(True)
    cls, out = lo.siamese_forward(p, xin, True)
    ((cls * d_cls.double()).sum() + (out * d_out.double()).sum()).backward()
    assert np.allclose(cls.detach().numpy(), g["cls"], atol=1e-10)
    assert np.allclose(grad_sample(out, 512), g["out_sample"], atol=1e-12) and abs(float(out.norm()) - float(g["out_norm"])) < 1e-10
    assert np.allclose(grad_sample(xin.grad, 512), g["dx_sample"], atol=1e-10)
    for k, nrm, smp in zip(g["grad_names"], g["grad_norms"], g["grad_samples"]):
        gr = p[str(k)].grad
        assert abs(float(gr.norm()) - float(nrm)) <= 1e-9 * max(1.0, float(nrm)), k
        s = grad_sample(gr, 16)
        assert np.allclose(s, smp[:s.size], atol=1e-10), k
    off = 0
    for k in g["buf_names"]:
        k = str(k)
        n = max(1, int(params[k].numel()))
        ref = g["buf_values"][off:off + n]
        off += n
        if "num_batches" not in k and not k.startswith("featV"):
            assert np.allclose(p[k].detach().numpy().reshape(-1), ref, atol=1e-12), k
    n = n2 // 2
    score = torch.softmax(torch.from_numpy(g["cls"]).float().view(-1, 2), dim=-1).view(n, n, 2)[:, :, 1].clone().requires_grad_(True)
    tv = tar.view(n, -1)
    loss, prec = lo.pair_loss(score, tv[:, 0], tv[:, 1])
    (loss * 1.7).backward()
    assert abs(float(loss) - float(g["pair_loss"])) < 1e-6 and abs(float(prec) - float(g["pair_prec"])) < 1e-6
    assert np.allclose(score.grad.numpy(), g["pair_dscore"], atol=1e-6)
