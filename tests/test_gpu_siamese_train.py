"""GPU parity: Siamese.forward in training + PairLoss (grl_siamese_*, grl_pair_loss_* through the C ABI) vs the REAL reference's
golden outputs and the fp64 oracle.  Bar: 1e-3 relative like the head (fp32 arithmetic; measured ~1e-5)."""
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth
from helpers_sample import grad_sample

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def build(n2, T, seed):
    from grl_b200.siamese import Siamese
    params, x, d_cls, d_out, tar = synth.make_siamese_inputs(n2, T, seed)
    sia = Siamese(2048, 512, 2)
    sd = sia.state_dict()
    for k, v in params.items():
        assert k in sd and sd[k].shape == v.shape, k
        sd[k] = v.clone()
    sia.load_state_dict(sd)
    return sia.cuda().train(), params, x, d_cls, d_out, tar


@pytest.mark.parametrize("name", ["siamese_n32t8", "siamese_n6t3"])
def test_siamese_forward_backward_matches_reference_golden(golden_dir, name):
    from oracle import loss_oracle as lo
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n2, T = int(g["n2"]), int(g["T"])
    sia, params, x, d_cls, d_out, tar = build(n2, T, int(g["seed"]))
    xin = x.cuda().requires_grad_(True)
    cls, out = sia(xin)
    assert cls.shape == (n2 // 2, n2 // 2, 2) and out.shape == (n2, 2048) and cls.requires_grad and out.requires_grad
    ((cls * d_cls.cuda()).sum() + (out * d_out.cuda()).sum()).backward()
    assert rel(cls.detach().cpu().numpy(), g["cls"]) < 1e-4
    assert rel(grad_sample(out, 512), g["out_sample"]) < 1e-5 and abs(float(out.norm()) - float(g["out_norm"])) < 1e-4 * float(g["out_norm"])
    assert rel(grad_sample(xin.grad, 512), g["dx_sample"]) < 1e-3 and abs(float(xin.grad.norm()) - float(g["dx_norm"])) < 1e-3 * float(g["dx_norm"])
    sd = dict(sia.named_parameters())
    for k, nrm, smp in zip(g["grad_names"], g["grad_norms"], g["grad_samples"]):
        k = str(k)
        gr = sd[k].grad
        assert gr is not None, k
        if float(nrm) < 1e-9:                       # Linear biases in front of a train-mode BN: exactly zero gradient
            assert float(gr.norm()) < 1e-5, k
            continue
        assert abs(float(gr.norm()) - float(nrm)) < 1e-3 * float(nrm), (k, float(gr.norm()), float(nrm))
        s = grad_sample(gr, 16)
        assert np.abs(s - smp[:s.size]).max() < 1e-3 * max(np.abs(smp).max(), 1e-12) + 1e-7, k
    assert sd["featV.weight"].grad is None          # featV takes no part in forward (Siamese.py:99)
    # BN running buffers and counters as the real modules leave them
    off = 0
    bufs = dict(sia.named_buffers())
    for k in g["buf_names"]:
        k = str(k)
        n = max(1, int(params[k].numel()))
        ref = g["buf_values"][off:off + n]
        off += n
        if k.startswith("featV"):
            continue
        got = bufs[k].detach().cpu().double().numpy().reshape(-1)
        if "num_batches" in k:
            assert int(got[0]) == int(ref[0]), k
        else:
            assert rel(got, ref) < 1e-5, (k, rel(got, ref))
    # the full tensors against the fp64 oracle (the fixture only holds samples)
    p64 = {k: v.double() if v.is_floating_point() else v.clone() for k, v in params.items()}
    x64 = x.double().clone().requires_grad_(True)
    cls_o, out_o = lo.siamese_forward(p64, x64, True)
    ((cls_o * d_cls.double()).sum() + (out_o * d_out.double()).sum()).backward()
    assert rel(out.detach().cpu().numpy(), out_o.detach().numpy()) < 1e-5
    assert rel(xin.grad.cpu().numpy(), x64.grad.numpy()) < 1e-3


def test_siamese_eval_mode_forward_and_pairloss(golden_dir):
    from grl_b200.siamese import PairLoss
    from oracle import loss_oracle as lo
    g = np.load(os.path.join(golden_dir, "siamese_n32t8.npz"))
    n2, T = 32, 8
    sia, params, x, _, _, tar = build(n2, T, int(g["seed"]))
    sia.eval()
    with torch.no_grad():
        cls, out = sia(x.cuda())
    p64 = {k: v.double() if v.is_floating_point() else v.clone() for k, v in params.items()}
    cls_o, out_o = lo.siamese_forward(p64, x.double(), False)
    assert rel(cls.cpu().numpy(), cls_o.numpy()) < 1e-4 and rel(out.cpu().numpy(), out_o.numpy()) < 1e-5
    # PairLoss on the reference's own scores
    n = n2 // 2
    score = torch.softmax(torch.from_numpy(g["cls"]).float().view(-1, 2), dim=-1).view(n, n, 2)[:, :, 1].cuda().requires_grad_(True)
    tv = tar.view(n, -1).cuda()
    loss, prec = PairLoss()(score, tv[:, 0], tv[:, 1])
    (loss * 1.7).backward()
    assert abs(float(loss.detach()) - float(g["pair_loss"])) < 1e-5 and abs(float(prec) - float(g["pair_prec"])) < 1e-6
    assert np.abs(score.grad.cpu().numpy() - g["pair_dscore"]).max() < 1e-5


def test_siamese_rejects_odd_batch():
    sia, _, x, _, _, _ = build(6, 3, 1)
    with pytest.raises(RuntimeError):
        sia(x.cuda()[:5])
