"""ORACLE tooling — generate tests/golden/*.npz by running the REAL reference.

Run in the authoring container only (needs /root/reference, which does not exist on
the GPU box):      python oracle/make_golden.py

Recipe: put /root/reference on sys.path; stub
`torch.utils.model_zoo.load_url` (no network) and let ResNet.load_state_dict
ignore the resulting None; build `ResNet50_GRL_Model`; replace `backbone.base`
with Identity so seeded synthetic layer4 maps feed GCE directly; load the seeded
head parameters of grl_b200.synth into the reference's own state_dict.  The
reference modules then run *unmodified* (Backbone.forward, TRLBlock.forward, the
corr_bn/uncorr_bn tail, cosin_dist, pairwise_distance_tensor, evaluate).

Fixtures hold only outputs (+ small gradient samples); inputs/params regenerate
bit-identically from their PCG64 seeds.
"""
from __future__ import annotations

import io
import os
import sys
import contextlib
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    import torch.utils.model_zoo as model_zoo
    model_zoo.load_url = lambda *a, **k: None                      # stub 1: no network
    with contextlib.redirect_stdout(io.StringIO()):
        from reid.models import resnets1
        _orig = resnets1.ResNet.load_state_dict
        resnets1.ResNet.load_state_dict = lambda self, sd, *a, **k: None if sd is None else _orig(self, sd, *a, **k)
        from reid.models.grl_model import ResNet50_GRL_Model
        from reid.evaluator import attevaluator, eva_functions
    return ResNet50_GRL_Model, attevaluator, eva_functions


def build_reference_model(Model, params, dtype):
    with contextlib.redirect_stdout(io.StringIO()):
        model = Model()
    model.backbone.base = torch.nn.Identity()                      # stub 2: synthetic layer4 maps
    sd = model.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.clone()
    model.load_state_dict(sd)
    return model.to(dtype)


def grad_sample(t: torch.Tensor, n=64):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy().copy()


def head_fixture(Model, name, B, T, training, dtype=torch.float64, with_grads=True):
    from grl_b200 import synth
    params = synth.make_head_params(0, dtype=torch.float32)
    model = build_reference_model(Model, params, dtype)
    model.train(training)
    x = synth.make_head_input(B, T).to(dtype).requires_grad_(with_grads)
    gu, gc = synth.make_head_grads(B, T)
    gu, gc = gu.to(dtype), gc.to(dtype)
    # Backbone.forward with base=Identity == GCE on layer4 maps (basebranch.py:52-68)
    x_uncorr, x_corr, corr_map = model.backbone(x, B, T)
    xu5 = x_uncorr.view(B, T, 2048, 16, 8)
    xc5 = x_corr.view(B, T, 2048, 16, 8)
    f_uncorr, f_corr = model.temporal_learning_block(xu5, xc5)      # grl_model.py:131-180
    out = dict(B=B, T=T, training=int(training),
               f_uncorr=f_uncorr.detach().double().numpy(), f_corr=f_corr.detach().double().numpy(),
               corr_map=corr_map.detach().double().numpy(),
               x_corr_sample=grad_sample(x_corr, 256), x_uncorr_sample=grad_sample(x_uncorr, 256))
    # tail (grl_model.py:222-226)
    xc_t = model.corr_bn(f_corr.view(B * T, 2048)).view(B, T, 2048)
    xc_t = torch.nn.functional.normalize(xc_t, p=2, dim=2)
    xu_t = torch.nn.functional.normalize(model.uncorr_bn(f_uncorr.view(B, 2048)).view(B, 2048), p=2, dim=1)
    out["tail_x_corr"] = xc_t.detach().double().numpy()
    out["tail_x_uncorr"] = xu_t.detach().double().numpy()
    if with_grads:
        loss = (f_uncorr * gu).sum() + (f_corr * gc).sum()
        loss.backward()
        out["dx_sample"] = grad_sample(x.grad, 512)
        out["dx_norm"] = float(x.grad.norm())
        names, norms, samples = [], [], []
        for k, v in model.named_parameters():
            if k in params and v.grad is not None:
                names.append(k)
                norms.append(float(v.grad.norm()))
                samples.append(grad_sample(v.grad, 16) if v.grad.numel() >= 16 else
                               np.pad(grad_sample(v.grad, 16), (0, 16 - v.grad.numel())))
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms)
        out["grad_samples"] = np.stack(samples)
    if training:
        bufs = {k: v for k, v in model.state_dict().items() if k in params and ("running" in k or "num_batches" in k)}
        out["buf_names"] = np.array(list(bufs.keys()))
        out["buf_values"] = np.concatenate([v.double().reshape(-1).numpy() for v in bufs.values()])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("f_uncorr", "f_corr")})


def head_fixture_full_size(Model, name="head_train_b32t8", B=32, T=8):
    """The benchmark configuration itself (BASELINE configs[1], B=32, T=8, train-mode BN) through the REAL reference in
    float64 (truth) and in float32 (the reference's own rounding floor per tensor: rel |fp32 - fp64|).  Outputs are stored
    as float32 (gate 1e-4), gradients as norms + strided samples + the fp32 floor.  Takes a few minutes and ~25 GB of RAM."""
    from grl_b200 import synth
    params = synth.make_head_params(0, dtype=torch.float32)
    gu, gc = synth.make_head_grads(B, T)
    res = {}
    for dtype in (torch.float64, torch.float32):
        model = build_reference_model(Model, params, dtype).train()
        x = synth.make_head_input(B, T).to(dtype).requires_grad_(True)
        x_uncorr, x_corr, corr_map = model.backbone(x, B, T)
        f_uncorr, f_corr = model.temporal_learning_block(x_uncorr.view(B, T, 2048, 16, 8), x_corr.view(B, T, 2048, 16, 8))
        ((f_uncorr * gu.to(dtype)).sum() + (f_corr * gc.to(dtype)).sum()).backward()
        r = dict(f_uncorr=f_uncorr.detach().double(), f_corr=f_corr.detach().double(), corr_map=corr_map.detach().double(),
                 dx=x.grad.double())
        r["grads"] = {k: v.grad.double().clone() for k, v in model.named_parameters() if k in params and v.grad is not None}
        r["bufs"] = {k: v.double().clone() for k, v in model.state_dict().items()
                     if k in params and ("running" in k or "num_batches" in k)}
        res[dtype] = r
        del model, x, x_uncorr, x_corr, f_uncorr, f_corr
    t, f = res[torch.float64], res[torch.float32]
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-300))
    names = list(t["grads"].keys())
    # strided samples, stored as float32 (the gates are >= 1e-3).  dx's error is spiky (a flipped ReLU mask changes whole
    # pixels), so a strided sample can miss it entirely: dx is also stored POOLED over the 16x8 pixels of every (frame,
    # channel) -- a linear functional that sees every entry -- with the reference's own fp32 floor for the same functional.
    NS_DX, NS_G = 65536, 4096
    pool = lambda v: v.reshape(B * T, 2048, 128).sum(2)
    smp = lambda v, n: grad_sample(v, n).astype(np.float32)
    pad = lambda a, n: np.pad(a, (0, n - a.size))
    out = dict(B=B, T=T, f_uncorr=t["f_uncorr"].float().numpy(), f_corr=t["f_corr"].float().numpy(),
               corr_map=t["corr_map"].float().numpy(),
               floor_f_uncorr=rel(f["f_uncorr"], t["f_uncorr"]), floor_f_corr=rel(f["f_corr"], t["f_corr"]),
               floor_corr_map=rel(f["corr_map"], t["corr_map"]),
               dx_sample=smp(t["dx"], NS_DX), floor_dx_sample=rel(torch.from_numpy(grad_sample(f["dx"], NS_DX)), torch.from_numpy(grad_sample(t["dx"], NS_DX))),
               dx_pool=pool(t["dx"]).float().numpy(), floor_dx_pool=rel(pool(f["dx"]), pool(t["dx"])),
               dx_norm=float(t["dx"].norm()), floor_dx=rel(f["dx"], t["dx"]),
               grad_names=np.array(names), grad_norms=np.array([float(t["grads"][k].norm()) for k in names]),
               grad_numel=np.array([t["grads"][k].numel() for k in names]),
               grad_samples=np.stack([pad(smp(t["grads"][k], NS_G), NS_G) for k in names]),
               grad_floor=np.array([rel(f["grads"][k], t["grads"][k]) for k in names]),
               buf_names=np.array(list(t["bufs"].keys())),
               buf_values=np.concatenate([v.reshape(-1).numpy() for v in t["bufs"].values()]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, "floors: f_uncorr %.1e f_corr %.1e dx %.1e max grad %.1e" %
          (out["floor_f_uncorr"], out["floor_f_corr"], out["floor_dx"], out["grad_floor"].max()))


def eval_fixture(att, eva, name, nq, ng_extra, dim, seed, noise, max_rank=100, quantize=None):
    from grl_b200 import synth
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(nq, ng_extra, dim, seed=seed, num_ids=25, noise=noise,
                                                 missing_query_frac=0.05)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d_cos = att.cosin_dist(torch.from_numpy(qf), torch.from_numpy(gf)).numpy()
        d_l2 = att.pairwise_distance_tensor(torch.from_numpy(qf), torch.from_numpy(gf)).numpy()
    if quantize:
        d_cos = (np.round(d_cos * quantize) / quantize).astype(np.float32)    # force exact ties
    with contextlib.redirect_stdout(io.StringIO()):
        cmc, mAP = eva.evaluate(d_cos, qp, gp, qc, gc, max_rank=max_rank)
        cmc_l2, mAP_l2 = eva.evaluate(d_l2, qp, gp, qc, gc, max_rank=max_rank)
        rank1 = att.evaluate_seq(d_cos, qp, qc, gp, gc, path=None)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), nq=nq, ng_extra=ng_extra, dim=dim, seed=seed, noise=noise,
                        max_rank=max_rank, quantize=quantize or 0, d_cos=d_cos, d_l2=d_l2,
                        cmc=cmc, mAP=mAP, cmc_l2=cmc_l2, mAP_l2=mAP_l2, rank1=rank1)
    print("wrote", name, "mAP %.4f rank1 %.4f" % (mAP, cmc[0]))


def eval_props_fixture(att, eva, name, nq, ng_extra, dim, seed, noise, topk=50):
    """The reference's OTHER two evaluators on the same distance matrix (they are never called by the
    live path, which makes them an independent cross-check of `evaluate`): `cmc(..., first_match_break=True)`
    (eva_functions.py:18-78) counts the rank of the first true match exactly like `evaluate`'s CMC, and `mean_ap`
    (:81-115, sklearn's average_precision_score) is the same mAP when no two distances tie."""
    from grl_b200 import synth
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(nq, ng_extra, dim, seed=seed, num_ids=25, noise=noise, missing_query_frac=0.05)
    gf = (gf + np.float32(1e-3) * np.random.default_rng(seed).standard_normal(gf.shape).astype(np.float32))   # break exact ties
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d_cos = att.cosin_dist(torch.from_numpy(qf), torch.from_numpy(gf)).numpy()
    assert all(len(np.unique(r)) == len(r) for r in d_cos), "fixture needs a tie-free matrix"
    with contextlib.redirect_stdout(io.StringIO()):
        cmc_ev, map_ev = eva.evaluate(d_cos, qp, gp, qc, gc, max_rank=topk)
    cmc_fmb = eva.cmc(d_cos, qp, gp, qc, gc, topk=topk, first_match_break=True)
    map_sk = eva.mean_ap(d_cos, qp, gp, qc, gc)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), nq=nq, ng_extra=ng_extra, dim=dim, seed=seed, noise=noise, topk=topk,
                        cmc_evaluate=cmc_ev, mAP_evaluate=map_ev, cmc_first_match_break=cmc_fmb, mAP_sklearn=map_sk)
    print("wrote", name, "mAP evaluate %.6f sklearn %.6f | max |cmc - cmc_fmb| %.1e" % (map_ev, map_sk, np.abs(cmc_ev - cmc_fmb).max()))


def tail_fixture(Model, name, n, T):
    """Eval feature tail on the REAL reference modules: model.corr_bn / uncorr_bn (grl_model.py:222-226) and
    reid.models.Siamese.Siamese(2048, 512, 2).self_attention (mars_train.py:77), all in eval mode, float64."""
    from grl_b200 import synth
    from reid.models.Siamese import Siamese
    tp = synth.make_tail_params(10)
    with contextlib.redirect_stdout(io.StringIO()):
        model = Model()
    sia = Siamese(2048, 512, 2)
    msd, ssd = model.state_dict(), sia.state_dict()
    for k, v in tp.items():
        if k.startswith("siamese."):
            assert k[8:] in ssd and ssd[k[8:]].shape == v.shape, k
            ssd[k[8:]] = v.clone()
        else:
            assert k in msd and msd[k].shape == v.shape, k
            msd[k] = v.clone()
    model.load_state_dict(msd)
    sia.load_state_dict(ssd)
    model = model.double().eval()
    sia = sia.double().eval()
    fu, fc = synth.make_tail_input(n, T, dtype=torch.float64)
    with torch.no_grad():
        x_corr = torch.nn.functional.normalize(model.corr_bn(fc.view(n * T, 2048)).view(n, T, 2048), p=2, dim=2)
        x_uncorr = torch.nn.functional.normalize(model.uncorr_bn(fu), p=2, dim=1)
        out_frame = sia.self_attention(x_corr)
        out_feat = torch.cat((x_uncorr, out_frame, x_corr.mean(dim=1)), dim=1)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), n=n, T=T, out_feat=out_feat.numpy(), out_frame=out_frame.numpy(),
                        tracklet=out_feat.mean(dim=0).numpy())
    print("wrote", name, tuple(out_feat.shape))


def rerank_fixture(att, name, nq, ng_extra, dim, seed, k1, k2, lam, noise=0.6):
    """k-reciprocal re-ranking by the REAL reference (reid/evaluator/rerank.py:37-104) on the three distance matrices
    ATTEvaluator.evaluate hands it (attevaluator.py:150-155: cosine q-g, L2 q-q and g-g).  The matrices are stored so
    the restatement and the CUDA path see bit-identical inputs."""
    from grl_b200 import synth
    from reid.evaluator.rerank import re_ranking
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(nq, ng_extra, dim, seed=seed, num_ids=12, noise=noise)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tq, tg = torch.from_numpy(qf), torch.from_numpy(gf)
        q_g = att.cosin_dist(tq, tg).numpy()
        q_q = att.pairwise_distance_tensor(tq, tq).numpy()
        g_g = att.pairwise_distance_tensor(tg, tg).numpy()
    final = re_ranking(q_g, q_q, g_g, k1=k1, k2=k2, lambda_value=lam)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), nq=nq, ng_extra=ng_extra, dim=dim, seed=seed, k1=k1, k2=k2, lam=lam,
                        noise=noise, q_g=q_g, q_q=q_q, g_g=g_g, final=final)
    print("wrote", name, final.shape, final.dtype)


def loss_fixture(name, B, D, C, seed, n_ids):
    """The REAL TripletLoss('soft', True) (reid/loss/triplet.py, as built at reid/train/trainer.py:12) on seeded features:
    values and autograd gradients in float64; plus OIMLoss's forward arithmetic (oim.py:15,54,56) with torch's own ops."""
    from reid.loss.triplet import TripletLoss
    from grl_b200 import synth
    feat, ids, lut, targets = synth.make_loss_inputs(B, D, C, seed, n_ids)
    out = dict(B=B, D=D, C=C, seed=seed, n_ids=n_ids)
    for margin, tag in (('soft', 'soft'), (0.3, 'm03')):
        f = feat.double().clone().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            b_loss = TripletLoss(margin, True)(f, ids)
        gw = torch.linspace(0.5, 1.5, B, dtype=torch.float64)
        (b_loss * gw).sum().backward()
        out["tri_%s_loss" % tag] = b_loss.detach().numpy()
        out["tri_%s_dfeat" % tag] = f.grad.numpy()
    x = feat.double().clone().requires_grad_(True)
    logits = x.mm(lut.double().t()) * 30.0
    loss = torch.nn.functional.cross_entropy(logits, targets)
    loss.backward()
    out.update(oim_loss=float(loss), oim_logits=logits.detach().numpy(), oim_dx=x.grad.numpy())
    # OIM.backward (oim.py:18-27) by the REAL reference code: the class is a legacy (non-static) autograd.Function that modern
    # PyTorch refuses to *apply*, but its `backward` is a plain method -- call it unbound on a stub carrying exactly the
    # attributes it reads (saved_tensors, needs_input_grad, lut, momentum).  grad_outputs = d loss / d (inputs.mm(lut.t())),
    # i.e. the cross-entropy gradient through OIMLoss.forward's `inputs *= self.scalar` (:54).
    from types import SimpleNamespace
    from reid.loss.oim import OIM
    raw = feat.double().mm(lut.double().t()).requires_grad_(True)
    torch.nn.functional.cross_entropy(raw * 30.0, targets).backward()
    stub = SimpleNamespace(saved_tensors=(feat.double(), targets), needs_input_grad=(True, False), lut=lut.double().clone(),
                           momentum=0.5)
    grad_inputs, none = OIM.backward(stub, raw.grad)
    assert none is None
    touched = np.unique(targets.numpy())                 # only these rows may change; the fixture stores them alone
    rest = np.setdiff1d(np.arange(C), touched)
    assert np.array_equal(stub.lut.numpy()[rest], lut.double().numpy()[rest])
    out.update(oim_bwd_dx=grad_inputs.numpy(), oim_new_lut_rows=touched, oim_new_lut_vals=stub.lut.numpy()[touched])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, "triplet soft mean %.4f oim loss %.4f" % (out["tri_soft_loss"].mean(), float(loss)))


def siamese_fixture(name, n2, T, seed):
    """The REAL Siamese(2048, 512, 2).forward in train mode (reid/models/Siamese.py:108-142, as built at mars_train.py:77) and the
    REAL PairLoss (reid/loss/pairloss.py), float64 / float32: outputs, gradients for seeded upstream grads, BN buffers."""
    from reid.models.Siamese import Siamese
    from reid.loss.pairloss import PairLoss
    from grl_b200 import synth
    params, x, d_cls, d_out, tar = synth.make_siamese_inputs(n2, T, seed)
    sia = Siamese(2048, 512, 2)
    sd = sia.state_dict()
    for k, v in params.items():
        assert k in sd and sd[k].shape == v.shape, k
        sd[k] = v.clone()
    sia.load_state_dict(sd)
    sia = sia.double().train()
    xin = x.double().clone().requires_grad_(True)
    cls, out = sia(xin)
    ((cls * d_cls.double()).sum() + (out * d_out.double()).sum()).backward()
    # fixtures stay small: full cls_encode, strided samples + norms of the big tensors (same sampling as the head fixtures)
    res = dict(n2=n2, T=T, seed=seed, cls=cls.detach().numpy(), out_sample=grad_sample(out, 512), out_norm=float(out.norm()),
               dx_sample=grad_sample(xin.grad, 512), dx_norm=float(xin.grad.norm()))
    names, norms, samples = [], [], []
    for k, v in sia.named_parameters():
        if v.grad is not None:
            names.append(k)
            norms.append(float(v.grad.norm()))
            smp = grad_sample(v.grad, 16)
            samples.append(np.pad(smp, (0, 16 - smp.size)))
    res["grad_names"] = np.array(names)
    res["grad_norms"] = np.array(norms)
    res["grad_samples"] = np.stack(samples)
    bufs = {k: v for k, v in sia.state_dict().items() if "running" in k or "num_batches" in k}
    res["buf_names"] = np.array(list(bufs.keys()))
    res["buf_values"] = np.concatenate([v.double().reshape(-1).numpy() for v in bufs.values()])
    # PairLoss on the softmax-ed scores exactly as the trainer feeds it (trainer.py:142-147), float32 like its labels
    n = n2 // 2
    score = torch.softmax(cls.detach().float().view(-1, 2), dim=-1).view(n, n, 2)[:, :, 1].clone().requires_grad_(True)
    tv = tar.view(n, -1)
    loss, prec = PairLoss()(score, tv[:, 0], tv[:, 1])
    (loss * 1.7).backward()
    res.update(pair_loss=float(loss), pair_prec=float(prec), pair_dscore=score.grad.numpy())
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **res)
    print("wrote", name, tuple(cls.shape), "pair loss %.4f prec %.3f" % (float(loss), float(prec)))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    Model, att, eva = import_reference()
    if "--loss" in sys.argv:               # only the loss fixtures
        loss_fixture("loss_b32", 32, 2048, 625, seed=21, n_ids=8)
        loss_fixture("loss_b12", 12, 256, 40, seed=22, n_ids=5)
        return
    if "--eval-props" in sys.argv:         # only the evaluator cross-check fixture
        eval_props_fixture(att, eva, "eval_props", 80, 400, 64, seed=6, noise=1.5)
        return
    if "--full-size" in sys.argv:          # only the benchmark-size head fixture (minutes, ~25 GB of RAM)
        head_fixture_full_size(Model)
        return
    head_fixture(Model, "head_train_b2t3", 2, 3, True)
    head_fixture(Model, "head_train_b4t2", 4, 2, True)
    head_fixture(Model, "head_eval_b3t4", 3, 4, False, with_grads=False)
    tail_fixture(Model, "tail_n5t8", 5, 8)
    tail_fixture(Model, "tail_n3t16", 3, 16)
    eval_fixture(att, eva, "eval_small", 60, 240, 64, seed=3, noise=1.5)
    # NOTE: the reference itself raises (ragged all_cmc, eva_functions.py:164,180) when junk removal leaves
    # fewer than max_rank gallery rows, so "num_g < max_rank" (:136-138) cannot be pinned; use max_rank=10.
    eval_fixture(att, eva, "eval_rank10", 20, 100, 32, seed=4, noise=1.0, max_rank=10)
    eval_fixture(att, eva, "eval_ties", 40, 160, 16, seed=5, noise=1.0, quantize=8)  # exact ties
    eval_props_fixture(att, eva, "eval_props", 80, 400, 64, seed=6, noise=1.5)
    siamese_fixture("siamese_n32t8", 32, 8, seed=31)
    siamese_fixture("siamese_n6t3", 6, 3, seed=32)
    loss_fixture("loss_b32", 32, 2048, 625, seed=21, n_ids=8)
    loss_fixture("loss_b12", 12, 256, 40, seed=22, n_ids=5)
    rerank_fixture(att, "rerank_k20", 48, 160, 64, seed=11, k1=20, k2=6, lam=0.3)
    rerank_fixture(att, "rerank_k6", 30, 100, 32, seed=12, k1=6, k2=3, lam=0.5)
    rerank_fixture(att, "rerank_k5_noqe", 25, 90, 32, seed=13, k1=5, k2=1, lam=0.3)      # k2 == 1 skips :78-83


if __name__ == "__main__":
    main()
