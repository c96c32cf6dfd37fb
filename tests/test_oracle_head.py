"""CPU: pin the head oracle to the real reference's outputs (tests/golden, made by
oracle/make_golden.py) and prove the decomposed plan (what the kernels implement)
equals the reference-style restatement incl. every gradient and BN buffer."""
import copy
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth
from oracle import head_oracle as ho


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).detach()
    b = torch.as_tensor(b, dtype=torch.float64).detach()
    return float((a - b).norm() / (b.norm() + 1e-300))


def sample(t, n):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy()


def run_ref(B, T, training, dtype=torch.float64, grads=True):
    p = synth.make_head_params(0, dtype=dtype)
    leaf = {k: (v.clone().requires_grad_(grads) if v.is_floating_point() and "running" not in k else v.clone())
            for k, v in p.items()}
    x = synth.make_head_input(B, T, dtype=dtype).requires_grad_(grads)
    out = ho.ref_forward(leaf, x, B, T, training)
    if grads:
        gu, gc = synth.make_head_grads(B, T, dtype=dtype)
        ((out["f_uncorr"] * gu).sum() + (out["f_corr"] * gc).sum()).backward()
    return leaf, x, out


@pytest.mark.parametrize("name", ["head_train_b2t3", "head_train_b4t2", "head_eval_b3t4"])
def test_ref_restatement_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, T, training = int(g["B"]), int(g["T"]), bool(g["training"])
    has_grads = "dx_sample" in g.files
    leaf, x, out = run_ref(B, T, training, grads=has_grads)
    for k in ("f_uncorr", "f_corr", "corr_map"):
        assert rel(out[k], g[k]) < 1e-12, k
    assert np.allclose(sample(out["x_corr"], 256), g["x_corr_sample"], rtol=1e-12, atol=1e-14)
    assert np.allclose(sample(out["x_uncorr"], 256), g["x_uncorr_sample"], rtol=1e-12, atol=1e-14)
    if has_grads:
        assert abs(float(x.grad.norm()) - float(g["dx_norm"])) < 1e-10 * float(g["dx_norm"])
        assert np.allclose(sample(x.grad, 512), g["dx_sample"], rtol=1e-9, atol=1e-13)
        for name_, norm in zip(g["grad_names"], g["grad_norms"]):
            got = float(leaf[str(name_)].grad.norm())
            assert abs(got - norm) <= 1e-9 * max(norm, 1e-9) + 1e-14, name_
    if training:
        vals = np.concatenate([leaf[str(k)].double().reshape(-1).numpy() for k in g["buf_names"]])
        assert np.allclose(vals, g["buf_values"], rtol=1e-12, atol=1e-14)


def test_tail_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "head_train_b2t3.npz"))
    p = {}
    for n in ("corr_bn", "uncorr_bn"):          # grl_model.py:203-209: weight 1, bias 0, default buffers
        p[n + ".weight"] = torch.ones(2048, dtype=torch.float64)
        p[n + ".bias"] = torch.zeros(2048, dtype=torch.float64)
        p[n + ".running_mean"] = torch.zeros(2048, dtype=torch.float64)
        p[n + ".running_var"] = torch.ones(2048, dtype=torch.float64)
        p[n + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    xu, xc = ho.ref_tail(p, torch.from_numpy(g["f_uncorr"]), torch.from_numpy(g["f_corr"]), True)
    assert rel(xu, g["tail_x_uncorr"]) < 1e-12 and rel(xc, g["tail_x_corr"]) < 1e-12


@pytest.mark.parametrize("training", [True, False])
def test_plan_equals_ref_fp64(training):
    B, T = 2, 3
    leaf, x, out = run_ref(B, T, training)
    p2 = synth.make_head_params(0, dtype=torch.float64)
    gu, gc = synth.make_head_grads(B, T, dtype=torch.float64)
    o2 = ho.plan_head(p2, x.detach(), B, T, training, grads=(gu, gc))
    for k in ("f_uncorr", "f_corr", "corr_map", "x_uncorr", "x_corr"):
        assert rel(o2[k], out[k]) < 1e-12, k
    assert rel(o2["dx"], x.grad) < 1e-11
    for k, v in leaf.items():
        if v.requires_grad:
            g2 = o2["grads"][k].reshape(v.shape)
            # biases feeding a train-mode BN have an exactly-zero gradient: compare absolutely
            assert float((g2 - v.grad).norm()) <= 1e-10 * float(v.grad.norm()) + 1e-12, k
    for k in p2:
        if "running" in k or "num_batches" in k:
            assert rel(p2[k], leaf[k]) < 1e-12, k


def test_fp32_floor_vs_fp64():
    """Documents the precision floor the GPU parity gate is calibrated against."""
    B, T = 2, 3
    _, _, o64 = run_ref(B, T, True, torch.float64, grads=False)
    _, _, o32 = run_ref(B, T, True, torch.float32, grads=False)
    for k in ("f_uncorr", "f_corr", "corr_map"):
        assert rel(o32[k], o64[k]) < 2e-5, k
