"""GPU parity: GCE + TRL head (forward, backward, BN buffers, module API) through the C ABI vs the oracle.

Truth = the fp64 oracle (oracle/head_oracle.py, pinned to the real reference by tests/golden).  Tolerances:
  * forward outputs: 1e-3 relative (north-star); the kernels actually sit near 1e-5, asserted at 1e-4;
  * BN running buffers: 1e-5;
  * backward, conditioned on the CUDA forward's own saved activations: 1e-3 (actual ~1e-5).  This is the sharp
    test of the backward kernels;
  * backward end to end: the head's gradient is a discontinuous function of its inputs (ReLU masks, tiny-batch BN),
    so even the reference's own fp32 run differs from its fp64 run by 1e-3..6e-3 (a forward
    perturbation of 1e-6 relative already moves dx by 2e-3..5e-3, DESIGN.md "Numerics"); whether a given run flips a
    mask is a coin toss, so the reference's floor itself jumps between 1e-7 and 5e-3 from tensor to tensor.  Gates:
      - gradients no kink feeds (f1 / f2 / channel attention weights): 1e-3, the north-star figure;
      - gradients downstream of the memory-update ReLUs and the GCE gate (everything else, incl. dx):
        err(ours, fp64) <= max(KINK_TOL, E2E_FLOOR_MULT * err(reference fp32, fp64)).
    The two scalar gradients of
    corr_atte.6 (BatchNorm2d(1): sums of >= 1024 signed terms that cancel to ~1e-3 of their absolute mass) get an
    absolute allowance instead; their kernels are pinned by the conditioned test above.
"""
import copy
import os

import numpy as np
import pytest
import torch

from grl_b200 import synth

pytestmark = pytest.mark.gpu

E2E_FLOOR_MULT = 10.0     # the tensor-core forward sits at ~1e-5 vs the CPU fp32 reference's ~1e-6: ~3x more mask flips
KINK_TOL = 1e-2           # a handful of flipped ReLU masks among ~1e6 activations
SCALAR_BN_TOL = 1e-1      # corr_atte.6.weight / .bias, see module docstring
# parameters whose gradient is analytically zero: a per-channel constant added in front of a train-mode BatchNorm
# (glo_fc.0.bias -> glo_fc.1;  corr_atte.1.bias -> conv -> corr_atte.3).  Compared against the scale of their layer.
ZERO_GRADS = {"backbone.glo_fc.0.bias": "backbone.glo_fc.1.bias", "backbone.corr_atte.1.bias": "backbone.corr_atte.1.weight"}


def _mods():
    from grl_b200 import _lib, head
    from oracle import head_oracle as ho
    return _lib, head, ho


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def device_params(seed=0):
    return {k: v.cuda().contiguous() for k, v in synth.make_head_params(seed).items()}


@pytest.mark.parametrize("name", ["head_train_b2t3", "head_train_b4t2", "head_eval_b3t4"])
def test_forward_matches_oracle_and_reference_golden(golden_dir, name):
    _, head, ho = _mods()
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, T, training = int(g["B"]), int(g["T"]), bool(g["training"])
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    f_uncorr, f_corr, corr_map, xu, xc, _ = head.head_forward_raw(sd, x, B, T, training, save=training, want_maps=True)
    # against outputs of the REAL reference modules (fp64), committed as fixtures
    for k, v in (("f_uncorr", f_uncorr), ("f_corr", f_corr), ("corr_map", corr_map)):
        assert rel(v, g[k]) < 1e-4, (k, rel(v, g[k]))
    # against the oracle plan, including the stand-alone gated maps and the BN buffers
    p64 = synth.make_head_params(0, dtype=torch.float64)
    o = ho.plan_head(p64, synth.make_head_input(B, T, dtype=torch.float64), B, T, training)
    assert rel(xu, o["x_uncorr"]) < 1e-4 and rel(xc, o["x_corr"]) < 1e-4
    assert corr_map.shape == (B * T, 1, 16, 8) and f_corr.shape == (B, T, 2048) and f_uncorr.shape == (B, 2048)
    for k in p64:
        if "running" in k:
            assert rel(sd[k], p64[k]) < 1e-5, k
    if training:      # running buffers as left behind by the REAL reference (fixture), entry by entry
        off = 0
        for k in g["buf_names"]:
            k = str(k)
            n = max(1, int(sd[k].numel()))
            ref = g["buf_values"][off:off + n]
            off += n
            if "num_batches" in k:
                continue                      # advanced by the nn.Module wrapper (test_module_api_*)
            assert rel(sd[k], ref) < 2e-5, (k, rel(sd[k], ref))


@pytest.mark.parametrize("B,T", [(2, 3), (4, 2), (6, 4)])
def test_backward_matches_oracle_on_saved_activations(B, T):
    """Backward kernels in isolation: fp64 oracle backward evaluated on the CUDA forward's saved state."""
    _, head, ho = _mods()
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    gu, gc = synth.make_head_grads(B, T)
    *_, ws = head.head_forward_raw(sd, x, B, T, True, save=True)
    sv = head.saved_state(ws, B, T)
    gctx, tctx = ho.ctx_from_saved(sv, B, T)
    p64 = synth.make_head_params(0, dtype=torch.float64)
    dx_ref, G = ho.plan_backward_from_ctx(p64, gctx, tctx, gu.double(), gc.double())
    dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu.cuda(), gc.cuda())
    assert rel(dx, dx_ref) < 1e-3, rel(dx, dx_ref)
    worst = {}
    for k in head.head_param_names():
        ref = G[k].reshape(grads[k].shape)
        if k in ZERO_GRADS:
            assert float(grads[k].double().norm()) < 1e-3 * float(G[ZERO_GRADS[k]].norm()), k
            continue
        if B == 2 and k.startswith("backbone.glo_fc"):
            continue      # BatchNorm1d over 2 samples: d(u) is identically ~0 (pure cancellation); covered at B >= 4
        worst[k] = rel(grads[k], ref)
    bad = {k: v for k, v in worst.items() if v >= 1e-3}
    assert not bad, bad


@pytest.mark.parametrize("B,T", [(2, 3), (4, 2)])
def test_backward_end_to_end_vs_fp64_oracle(B, T):
    _, head, ho = _mods()
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    gu, gc = synth.make_head_grads(B, T)
    *_, ws = head.head_forward_raw(sd, x, B, T, True, save=True)
    dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu.cuda(), gc.cuda())

    def ref_run(dtype):
        p = {k: (v.to(dtype).requires_grad_("running" not in k) if v.is_floating_point() else v.clone())
             for k, v in synth.make_head_params(0).items()}
        xx = synth.make_head_input(B, T).to(dtype).requires_grad_(True)
        out = ho.ref_forward(p, xx, B, T, True)
        ((out["f_uncorr"] * gu.to(dtype)).sum() + (out["f_corr"] * gc.to(dtype)).sum()).backward()
        return xx.grad, {k: v.grad for k, v in p.items() if v.requires_grad}

    dx64, g64 = ref_run(torch.float64)
    dx32, g32 = ref_run(torch.float32)
    report = {}

    def gate(name, ours, r64, r32):
        if name in ZERO_GRADS:
            return
        floor = rel(r32, r64)
        err = rel(ours, r64)
        report[name] = (err, floor)
        if name.startswith("backbone.corr_atte.6"):
            tol = SCALAR_BN_TOL
        elif "_f1.0" in name or "_f2.0" in name or "channel_atte" in name:
            tol = 1e-3
        else:
            tol = max(KINK_TOL, E2E_FLOOR_MULT * floor)
        assert err <= tol, (name, err, floor)

    gate("dx", dx, dx64, dx32)
    for k in head.head_param_names():
        if B == 2 and k.startswith("backbone.glo_fc"):
            continue
        gate(k, grads[k], g64[k].reshape(grads[k].shape), g32[k].reshape(grads[k].shape))
    print("end-to-end gradient error (ours vs fp64 | reference fp32 vs fp64):")
    for k, (e, f) in report.items():
        print("  %-72s %.2e | %.2e" % (k, e, f))


def test_module_api_autograd_and_state_dict(golden_dir):
    """Drop-in surface: same state_dict keys as the reference, autograd through model.head, BN bookkeeping."""
    _, head, ho = _mods()
    g = np.load(os.path.join(golden_dir, "head_train_b2t3.npz"))
    B, T = 2, 3
    model = head.ResNet50_GRL_Model(base=torch.nn.Identity()).cuda()
    keys = set(model.state_dict().keys())
    for k in list(g["buf_names"]) + list(g["grad_names"]):
        assert str(k) in keys, k                       # names recorded from the real reference model
    assert "corr_bn.weight" in keys and "uncorr_bn.running_var" in keys
    sd = model.state_dict()
    for k, v in synth.make_head_params(0).items():
        sd[k] = v
    model.load_state_dict(sd)
    model.train()
    x = synth.make_head_input(B, T).cuda().requires_grad_(True)
    f_uncorr, f_corr, corr_map, _, _ = model.head(x, B, T)
    assert f_uncorr.requires_grad and f_corr.requires_grad
    for k, v in (("f_uncorr", f_uncorr), ("f_corr", f_corr), ("corr_map", corr_map)):
        assert rel(v, g[k]) < 1e-4, k
    gu, gc = synth.make_head_grads(B, T)
    torch.autograd.backward([f_uncorr, f_corr], [gu.cuda(), gc.cuda()])
    assert x.grad is not None and abs(float(x.grad.double().norm()) / float(g["dx_norm"]) - 1) < 2e-2
    named = dict(model.named_parameters())
    for k in head.head_param_names():
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all(), k
    msd = model.state_dict()
    assert int(msd["backbone.corr_atte.1.num_batches_tracked"]) == 1
    assert int(msd["temporal_learning_block.uncorr_memo_forward.bn2.num_batches_tracked"]) == T   # one BN call per step
    # eval(): running statistics, no autograd state kept, clips independent
    model.eval()
    with torch.no_grad():
        a = model.head(x.detach(), B, T)[1]
        b = model.head(x.detach()[T:], 1, T)[1]
    assert rel(b, a[1:]) < 1e-5
    with pytest.raises(RuntimeError):
        model.head(x, B, T)                            # eval-mode backward is not a reference use case: refuse loudly


def test_full_size_properties_b32_t8():
    """BASELINE config 2 (B=32, T=8): size-independent properties instead of a CPU oracle run."""
    _, head, ho = _mods()
    B, T = 32, 8
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    gu, gc = synth.make_head_grads(B, T)
    fu, fc, cm, _, _, ws = head.head_forward_raw(sd, x, B, T, True, save=True)
    assert torch.isfinite(fu).all() and torch.isfinite(fc).all() and (cm > 0).all() and (cm < 1).all()
    dx1, g1 = head.head_backward_raw(sd, x, B, T, ws, gu.cuda(), gc.cuda())
    dx2, g2 = head.head_backward_raw(sd, x, B, T, ws, (2 * gu).cuda(), (2 * gc).cuda())
    assert torch.isfinite(dx1).all()
    assert rel(dx2, 2 * dx1) < 1e-4                    # backward is linear in the upstream gradient
    for k in ("temporal_learning_block.forward_f1.0.weight", "backbone.corr_atte.0.weight",
              "temporal_learning_block.uncorr_memo_backward.conv2.weight"):
        assert rel(g2[k], 2 * g1[k]) < 1e-4, k
    # the backward conditioned on its own saved activations, spot-checked on the TRL side at full size:
    # d f2 bias = column sums of dF2 (a checksum of the whole attention path)
    sv = head.saved_state(ws, B, T)
    a_sum = sv["a"][:, 0].permute(1, 0, 2) + sv["a"][:, 1].flip(0).permute(1, 0, 2)      # [B,T,C]: fwd step t, bwd step T-1-t
    dgc_ref = 2 * gc.cuda() * (2 + a_sum)                # the workspace holds the last backward (upstream grads x2)
    from grl_b200.head import ws_view
    dgc = ws_view(ws, B, T, True, "dgc", torch.float32, (B, T, 2048))
    assert rel(dgc, dgc_ref) < 1e-5
    # eval mode: chunk invariance (attevaluator.py:72-77 feeds 8 clips at a time)
    with torch.no_grad():
        full = head.head_forward_raw(sd, x, B, T, False, save=False)[1]
        part = head.head_forward_raw(sd, x[8 * T:16 * T], 8, T, False, save=False)[1]
    assert rel(part, full[8:16]) < 1e-5


def test_full_size_b32_t8_vs_real_reference(golden_dir):
    """The benchmark configuration itself (BASELINE configs[1]: B=32, T=8, train-mode BN) against ONE fp64 run of the REAL
    reference (tests/golden/head_train_b32t8.npz, oracle/make_golden.py --full-size): forward outputs at 1e-4, BN running
    buffers at 2e-5, every parameter gradient at SURVEY.md section 7.2's rule  err <= max(1e-3, 2 * err(reference fp32, fp64))
    -- the reference's own fp32 rounding floor at this size travels in the fixture.  Gradients are stored as strided samples
    (4,096 values per parameter tensor, 262,144 of dx) + whole-tensor norms; both are gated.
    dx is the one documented exception (gate 1e-2): it is the only gradient that is a per-pixel function of the ReLU masks of
    all 16 memory updates, so a mask that flips under a 1e-5 forward perturbation changes whole pixels of it, not a sum over
    thousands of pixels; the reference's own fp32 run already sits at 1.2e-3 here (fixture: floor_dx) with a forward that is
    ~50x closer to fp64 than the split-bf16 tensor-core forward.  The backward kernels themselves are pinned at 1e-3 (measured
    2e-5) by test_backward_matches_oracle_on_saved_activations.  The per-tensor table is printed and written to gpurun_out/
    (committed under profiles/)."""
    from helpers_sample import grad_sample
    _, head, ho = _mods()
    g = np.load(os.path.join(golden_dir, "head_train_b32t8.npz"))
    B, T = int(g["B"]), int(g["T"])
    assert (B, T) == (32, 8)
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    gu, gc = synth.make_head_grads(B, T)
    fu, fc, cm, _, _, ws = head.head_forward_raw(sd, x, B, T, True, save=True)
    for k, v in (("f_uncorr", fu), ("f_corr", fc), ("corr_map", cm)):
        assert rel(v, g[k]) < 1e-4, (k, rel(v, g[k]))
    off = 0
    for k in g["buf_names"]:
        k = str(k)
        n = max(1, int(sd[k].numel()))
        ref = g["buf_values"][off:off + n]
        off += n
        if "num_batches" not in k:
            assert rel(sd[k], ref) < 2e-5, (k, rel(sd[k], ref))
    dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu.cuda(), gc.cuda())
    lines, bad = [], {}

    def gate(name, ours, ref_sample, ref_norm, floor, tol):
        e_s = rel(grad_sample(ours, ref_sample.size), ref_sample)
        e_n = abs(float(ours.double().norm()) / ref_norm - 1.0)
        lines.append("%-72s sample %.2e  norm %.2e | reference fp32 floor %.2e | gate %.2e" % (name, e_s, e_n, floor, tol))
        # the whole-tensor statistic is gated at the rule itself; a strided sample of a tensor whose error sits in the few
        # pixels behind flipped ReLU masks scatters around the whole-tensor figure, so the sample gets 2x
        if e_s > 2.0 * tol or e_n > tol:
            bad[name] = (e_s, e_n, tol)

    # dx: whole-tensor norm, a strided sample, and dx POOLED over the pixels of every (frame, channel) -- its error is spiky
    # (flipped masks change whole pixels), which a strided sample can miss entirely: the pooled form sees every entry
    gate("dx (sample; reference floor on the same sample %.2e)" % float(g["floor_dx_sample"]), dx, g["dx_sample"], float(g["dx_norm"]),
         float(g["floor_dx"]), 1e-2)
    pool = dx.double().reshape(B * T, 2048, 128).sum(2)
    e_pool = rel(pool, g["dx_pool"])
    lines.append("%-72s pooled %.2e                | reference fp32 floor %.2e | gate %.2e" % ("dx (summed over the 16x8 pixels)", e_pool,
                                                                                              float(g["floor_dx_pool"]), 1e-2))
    if e_pool > 1e-2:
        bad["dx_pool"] = (e_pool, float(g["floor_dx_pool"]))
    for k, nrm, smp, fl, numel in zip(g["grad_names"], g["grad_norms"], g["grad_samples"], g["grad_floor"], g["grad_numel"]):
        k = str(k)
        if k in ZERO_GRADS:
            assert float(grads[k].double().norm()) < 1e-3 * float(grads[ZERO_GRADS[k]].double().norm()), k
            continue
        gate(k, grads[k], smp[:min(smp.size, int(numel))], float(nrm), float(fl), max(1e-3, 2.0 * float(fl)))
    table = "full-size (B=32, T=8) gradient error, ours vs the REAL reference in fp64:\n  " + "\n  ".join(lines)
    print(table)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "full_size_gradient_table.txt"), "w") as f:
            f.write(table + "\n")
    assert not bad, bad


def test_config1_b2_t8_forward_vs_oracle():
    """BASELINE configs[0]: head forward, B=2, T=8, train-mode BN (the reference's CPU-runnable case)."""
    _, head, ho = _mods()
    B, T = 2, 8
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    fu, fc, cm, _, _, _ = head.head_forward_raw(sd, x, B, T, True, save=False)
    p64 = synth.make_head_params(0, dtype=torch.float64)
    o = ho.ref_forward(p64, synth.make_head_input(B, T, dtype=torch.float64), B, T, True)
    for k, v in (("f_uncorr", fu), ("f_corr", fc), ("corr_map", cm)):
        assert rel(v, o[k]) < 1e-4, (k, rel(v, o[k]))
    for k in p64:
        if "running" in k:
            assert rel(sd[k], p64[k]) < 2e-5, k


def test_config4_long_tracklet_t16_dense_eval():
    """BASELINE configs[3]: T=16, eval-mode BN, clips of one tracklet fed in chunks of 8 and averaged
    (attevaluator.py:72-95).  Forward vs the fp64 oracle, and chunk invariance of the aggregation."""
    _, head, ho = _mods()
    T, n_clips = 16, 11
    sd = device_params()
    x = synth.make_head_input(n_clips, T, seed=9).cuda()
    feats = []
    with torch.no_grad():
        for y in range(0, n_clips, 8):                         # chunks of 8 clips like the reference
            nb = min(8, n_clips - y)
            fu, fc, *_ = head.head_forward_raw(sd, x[y * T:(y + nb) * T], nb, T, False, save=False)
            feats.append(torch.cat((fu, fc.mean(dim=1)), dim=1))
        dense = torch.cat(feats, 0).mean(dim=0)
        fu_all, fc_all, *_ = head.head_forward_raw(sd, x, n_clips, T, False, save=False)
        whole = torch.cat((fu_all, fc_all.mean(dim=1)), dim=1).mean(dim=0)
    assert rel(dense, whole) < 1e-5                            # eval-mode BN: clips independent, chunking is free
    p64 = synth.make_head_params(0, dtype=torch.float64)
    o = ho.ref_forward(p64, synth.make_head_input(3, T, seed=9, dtype=torch.float64)[: 3 * T], 3, T, False)
    x3 = synth.make_head_input(3, T, seed=9).cuda()
    fu3, fc3, cm3, _, _, _ = head.head_forward_raw(sd, x3, 3, T, False, save=False)
    for k, v in (("f_uncorr", fu3), ("f_corr", fc3), ("corr_map", cm3)):
        assert rel(v, o[k]) < 1e-4, (k, rel(v, o[k]))


def test_standalone_backbone_and_trl_modules_match_fused_head():
    """Backbone.forward (GCE) and TRLBlock.forward called separately, with autograd between them, reproduce the
    fused head: same outputs, same dx, same parameter gradients (SURVEY.md section 8 rows a1 + a2 on their own)."""
    _, head, ho = _mods()
    B, T = 4, 3
    gu, gc = synth.make_head_grads(B, T)

    def build():
        m = head.ResNet50_GRL_Model(base=torch.nn.Identity()).cuda()
        sd = m.state_dict()
        for k, v in synth.make_head_params(0).items():
            sd[k] = v
        m.load_state_dict(sd)
        return m.train()

    fused, split = build(), build()
    x1 = synth.make_head_input(B, T).cuda().requires_grad_(True)
    f_uncorr, f_corr, corr_map, _, _ = fused.head(x1, B, T)
    torch.autograd.backward([f_uncorr, f_corr], [gu.cuda(), gc.cuda()])

    x2 = synth.make_head_input(B, T).cuda().requires_grad_(True)
    x_uncorr, x_corr, cmap2 = split.backbone(x2, B, T)                   # reference signature: (x, b, t)
    assert x_uncorr.shape == (B * T, 2048, 16, 8) and cmap2.shape == (B * T, 1, 16, 8)
    fu2, fc2 = split.temporal_learning_block(x_uncorr.view(B, T, 2048, 16, 8), x_corr.view(B, T, 2048, 16, 8))
    torch.autograd.backward([fu2, fc2], [gu.cuda(), gc.cuda()])
    assert rel(cmap2, corr_map) < 1e-6 and rel(fu2, f_uncorr) < 1e-5 and rel(fc2, f_corr) < 1e-5
    assert rel(x_uncorr + x_corr, x2) < 1e-5                             # x*(1-m) + x*m
    assert rel(x2.grad, x1.grad) < 1e-4
    pf, ps = dict(fused.named_parameters()), dict(split.named_parameters())
    for k in head.head_param_names():
        if k in ZERO_GRADS:
            continue
        assert rel(ps[k].grad, pf[k].grad) < 1e-4, (k, rel(ps[k].grad, pf[k].grad))
    sf, ss = fused.state_dict(), split.state_dict()
    for k in sf:
        if "running" in k or "num_batches" in k:
            assert rel(ss[k], sf[k]) < 1e-6, k
    # a gradient flowing only into corr_map (stand-alone GCE use)
    x3 = synth.make_head_input(B, T).cuda().requires_grad_(True)
    _, _, cm3 = split.backbone(x3, B, T)
    cm3.sum().backward()
    assert torch.isfinite(x3.grad).all() and float(x3.grad.abs().sum()) > 0


def test_fused_head_with_map_outputs_backward_vs_fp64_reference():
    """want_maps=True: gradients arriving on the stand-alone x_uncorr / x_corr / corr_map outputs as well as on
    f_uncorr / f_corr (grl_head_backward's optional d_x_uncorr / d_x_corr / d_corr_map) vs autograd of the fp64 oracle."""
    _, head, ho = _mods()
    B, T = 4, 2
    gu, gc = synth.make_head_grads(B, T)
    rng = np.random.default_rng(4)
    r_u = torch.from_numpy(rng.standard_normal((B * T, 2048, 16, 8)).astype(np.float32)) * 1e-2
    r_c = torch.from_numpy(rng.standard_normal((B * T, 2048, 16, 8)).astype(np.float32)) * 1e-2
    r_m = torch.from_numpy(rng.standard_normal((B * T, 1, 16, 8)).astype(np.float32))
    model = head.ResNet50_GRL_Model(base=torch.nn.Identity()).cuda()
    sd = model.state_dict()
    for k, v in synth.make_head_params(0).items():
        sd[k] = v
    model.load_state_dict(sd)
    model.train()
    x = synth.make_head_input(B, T).cuda().requires_grad_(True)
    fu, fc, cm, xu, xc = model.head(x, B, T, want_maps=True)
    loss = (fu * gu.cuda()).sum() + (fc * gc.cuda()).sum() + (xu * r_u.cuda()).sum() + (xc * r_c.cuda()).sum() + (cm * r_m.cuda()).sum()
    loss.backward()
    p64 = {k: (v.double().requires_grad_("running" not in k) if v.is_floating_point() else v.clone())
           for k, v in synth.make_head_params(0).items()}
    x64 = synth.make_head_input(B, T, dtype=torch.float64).requires_grad_(True)
    o = ho.ref_forward(p64, x64, B, T, True)
    (((o["f_uncorr"] * gu.double()).sum() + (o["f_corr"] * gc.double()).sum() + (o["x_uncorr"] * r_u.double()).sum()
      + (o["x_corr"] * r_c.double()).sum() + (o["corr_map"] * r_m.double()).sum())).backward()
    assert rel(x.grad, x64.grad) < KINK_TOL, rel(x.grad, x64.grad)
    named = dict(model.named_parameters())
    for k in ("backbone.corr_atte.0.weight", "backbone.corr_atte.2.weight", "temporal_learning_block.forward_f1.0.weight"):
        assert rel(named[k].grad, p64[k].grad.reshape(named[k].shape)) < KINK_TOL, k


def test_graphed_head_step_replays_bit_identically():
    """GraphedHeadStep (CUDA graph of grl_head_forward + grl_head_backward, side-stream fork/join included) == the eager
    calls, bit for bit, including the BN running buffers and num_batches_tracked after several replays."""
    from grl_b200 import head
    B, T = 2, 3
    x = synth.make_head_input(B, T).cuda()
    gu, gc = synth.make_head_grads(B, T)
    gu, gc = gu.cuda(), gc.cuda()
    fresh = lambda: {k: v.cuda().contiguous() for k, v in synth.make_head_params(0).items()}
    sd_e, sd_g = fresh(), fresh()
    step = head.GraphedHeadStep(sd_g, B, T)               # construction runs a warm-up step on zero maps + the capture ...
    for k, v in fresh().items():                          # ... and must leave parameters AND BN buffers exactly as loaded
        assert torch.equal(sd_g[k], v), k
    step.x.copy_(x); step.d_f_uncorr.copy_(gu); step.d_f_corr.copy_(gc)
    ws = None
    for it in range(3):
        fu, fc, cm, _, _, ws = head.head_forward_raw(sd_e, x, B, T, True, save=True, ws=ws)
        dx, grads = head.head_backward_raw(sd_e, x, B, T, ws, gu, gc)
        gfu, gfc, gdx, ggr = step()
        torch.cuda.synchronize()
        assert torch.equal(gfu, fu) and torch.equal(gfc, fc) and torch.equal(gdx, dx) and torch.equal(step.corr_map, cm), it
        for k in grads:
            assert torch.equal(ggr[k], grads[k]), (it, k)
    for k in sd_e:
        if "running" in k:
            assert torch.equal(sd_g[k], sd_e[k]), k
    k = "temporal_learning_block.uncorr_memo_forward.bn1.num_batches_tracked"
    assert int(sd_g[k]) == 3 * T                          # the eager raw calls leave the counter to the module wrapper


@pytest.mark.parametrize("B,T,training", [(1, 1, False), (3, 5, True), (5, 8, False), (2, 1, True), (7, 2, True)])
def test_forward_odd_shapes_vs_reference_transcription(B, T, training):
    """Ragged sizes: a single clip / single frame (T = 1: both temporal directions degenerate to one step), odd batch sizes
    (row counts that are not multiples of the 256-row tile pair), train and eval BN.  B = 1 in train mode is excluded:
    BatchNorm1d over one sample raises in the reference (glo_fc.1)."""
    _, head, ho = _mods()
    sd = device_params()
    x = synth.make_head_input(B, T).cuda()
    fu, fc, cm, xu, xc, _ = head.head_forward_raw(sd, x, B, T, training, save=False, want_maps=True)
    p64 = synth.make_head_params(0, dtype=torch.float64)
    o = ho.ref_forward(p64, synth.make_head_input(B, T, dtype=torch.float64), B, T, training)
    for k, v in (("f_uncorr", fu), ("f_corr", fc), ("corr_map", cm), ("x_uncorr", xu), ("x_corr", xc)):
        assert rel(v, o[k]) < 1e-4, (k, rel(v, o[k]))
    if training:
        for k in p64:
            if "running" in k:
                assert rel(sd[k], p64[k]) < 2e-5, k


def test_head_rejects_bad_inputs():
    _, head, _ = _mods()
    sd = device_params()
    with pytest.raises(RuntimeError):
        head.head_forward_raw(sd, torch.zeros((6, 2048, 16, 8)), 2, 3, False, save=False)           # CPU tensor
    with pytest.raises(RuntimeError):
        head.head_forward_raw(sd, torch.zeros((6, 2048, 8, 16), device="cuda"), 2, 3, False, save=False)   # wrong map size
    with pytest.raises(RuntimeError):
        head.head_forward_raw(sd, torch.zeros((5, 2048, 16, 8), device="cuda"), 2, 3, False, save=False)   # b*t mismatch


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process (the reference's nn.DataParallel layout)")
def test_two_devices_in_one_process():
    """nn.DataParallel (mars_train.py:80) drives several GPUs from ONE process: every device needs its own handle, its own
    per-function shared-memory opt-in and its own stream; results must be identical on both."""
    _, head, _ = _mods()
    from grl_b200 import evaluator
    B, T = 2, 3
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        sd = {k: v.to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
        x = synth.make_head_input(B, T).to(dev)
        gu, gc = synth.make_head_grads(B, T)
        fu, fc, cm, _, _, ws = head.head_forward_raw(sd, x, B, T, True, save=True)
        dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu.to(dev), gc.to(dev))
        q = torch.nn.functional.normalize(torch.randn((1100, 64), generator=torch.Generator().manual_seed(1))).to(dev)
        g = torch.nn.functional.normalize(torch.randn((3000, 64), generator=torch.Generator().manual_seed(2))).to(dev)
        td, ti = evaluator.retrieve_topk(q, g, 10)
        torch.cuda.synchronize(dev)
        outs.append((fu.cpu(), fc.cpu(), dx.cpu(), td.cpu(), ti.cpu()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("training", [True, False])
def test_basic_block_forward_callable(training):
    """BasicBlock.forward(x1, x2) (grl_model.py:67-85) on its own: one memory update through grl_basic_block_forward, against the
    fp64 transcription of the reference module; train mode updates the running buffers like nn.BatchNorm2d."""
    _, head, ho = _mods()
    n = 5
    blk = head.BasicBlock(2048, 512).cuda()
    p = synth.make_head_params(0)
    pre = "temporal_learning_block.uncorr_memo_forward."
    sd = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
    blk.load_state_dict(sd)
    blk.train(training)
    g = torch.Generator().manual_seed(3)
    x1 = torch.relu(torch.randn((n, 2048, 16, 8), generator=g))
    x2 = torch.relu(torch.randn((n, 2048, 16, 8), generator=g))
    with torch.no_grad():
        out = blk(x1.cuda(), x2.cuda())
    p64 = {k: v.double().clone() for k, v in p.items()}
    ref = ho._ref_basic_block(p64, pre[:-1], x1.double(), x2.double(), training)
    assert out.shape == x1.shape and rel(out, ref) < 1e-4, rel(out, ref)
    if training:
        for k in ("bn1", "bn2", "bn3"):
            assert rel(getattr(blk, k).running_var, p64[pre + k + ".running_var"]) < 2e-5, k
            assert int(getattr(blk, k).num_batches_tracked) == 1
    with pytest.raises(RuntimeError):
        blk(x1.cuda().requires_grad_(True), x2.cuda())      # forward-only: gradients belong to TRLBlock's fused backward
