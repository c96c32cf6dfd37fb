"""ORACLE (test infrastructure, NOT product code) — CPU restatement of GRL's GCE + TRL head.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path (grl_b200/) never does.

Two restatements of the same math live here, both dtype-generic (run them in
float64 for "truth", float32 for the reference's own precision floor):

1. `ref_forward`  — a line-by-line functional transcription of the reference
   modules using the same torch ops (F.conv2d 1x1, F.linear, F.batch_norm, mean,
   sigmoid).  autograd over it *is* the reference backward.
     GCE : /root/reference/reid/models/basebranch.py:56-68  (params :38-50)
     TRL : /root/reference/reid/models/grl_model.py:131-180 (BasicBlock :67-85)
     tail: /root/reference/reid/models/grl_model.py:222-226

2. `plan_*`       — the decomposed, pixel-major ([P=B*T*S, C]) formulation that the
   sm_100a kernels implement, with an explicit (hand-derived) backward.  It uses the
   identities F1-F4 of SURVEY.md §7.1.  tests/test_oracle_head.py proves
   plan == ref (forward, every gradient, BN running buffers) in float64, and
   tests/golden/ pins `ref_forward` to outputs of the real reference modules
   (oracle/make_golden.py imports /root/reference to produce them).

PyTorch BatchNorm semantics relied upon (torch/nn/modules/batchnorm.py): biased
variance normalises, unbiased variance goes into running_var, eps=1e-5,
momentum=0.1, num_batches_tracked += 1 per call in train mode.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5
MOM = 0.1
S = 128  # 16 x 8, hard-coded in the reference (basebranch.py:59)
DIRS = (("forward", "foreward"), ("backward", "backward"))
TP = "temporal_learning_block."


# --------------------------------------------------------------------------------------
# 1. reference-style functional forward
# --------------------------------------------------------------------------------------
def _bn(p, prefix, x, training):
    """nn.BatchNorm{1,2}d.forward with default momentum/eps; updates buffers in place."""
    if training:
        p[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"],
                        p[prefix + ".weight"], p[prefix + ".bias"], training, MOM, EPS)


def ref_gce(p, x, b, t, training):
    """basebranch.py:56-68 after `self.base`.  x: [b*t, 2048, 16, 8]."""
    x_4 = x.view(b, t, x.size(1), x.size(2), x.size(3))
    x_glo = x_4.mean(dim=-1).mean(dim=-1).mean(dim=1)                                 # :58
    glo = F.linear(x_glo, p["backbone.glo_fc.0.weight"], p["backbone.glo_fc.0.bias"])
    glo = F.relu(_bn(p, "backbone.glo_fc.1", glo, training))                          # :38-40
    glo = glo.view(b, 1, 1024, 1, 1).expand(b, t, 1024, 16, 8).contiguous().view(b * t, 1024, 16, 8)  # :59
    x_cat = torch.cat((x, glo), dim=1)                                                # :61
    y = F.conv2d(x_cat, p["backbone.corr_atte.0.weight"])
    y = _bn(p, "backbone.corr_atte.1", y, training)
    y = F.conv2d(y, p["backbone.corr_atte.2.weight"])
    y = F.relu(_bn(p, "backbone.corr_atte.3", y, training))
    y = F.conv2d(y, p["backbone.corr_atte.5.weight"])
    y = _bn(p, "backbone.corr_atte.6", y, training)                                   # :42-50,62
    corr_map = torch.sigmoid(y).view(b * t, 1, 16, 8).contiguous()                    # :63
    x_corr = x * corr_map                                                             # :65
    x_uncorr = x * (1 - corr_map)                                                     # :66
    return x_uncorr, x_corr, corr_map


def _ref_basic_block(p, prefix, x1, x2, training):
    """grl_model.py:67-85."""
    x = x1 + x2
    residual = x
    out = F.conv2d(x, p[prefix + ".conv1.weight"])
    out = F.relu(_bn(p, prefix + ".bn1", out, training))
    out = F.conv2d(out, p[prefix + ".conv2.weight"])
    out = F.relu(_bn(p, prefix + ".bn2", out, training))
    out = F.conv2d(out, p[prefix + ".conv3.weight"])
    out = _bn(p, prefix + ".bn3", out, training)
    out = out + residual
    return F.relu(out)


def ref_trl(p, x_uncorr, x_corr, training):
    """grl_model.py:131-180.  x_*: [b, t, c, h, w]."""
    b, t, c, h, w = x_corr.size()
    f_step = {"forward": [], "backward": []}
    memo = {"forward": x_uncorr.mean(dim=1), "backward": x_uncorr.mean(dim=1)}        # :137-138
    for i in range(t):
        for direction, atte in DIRS:
            tau = i if direction == "forward" else t - 1 - i
            xc = x_corr[:, tau]
            xu = x_uncorr[:, tau]
            f1 = F.relu(F.conv2d(memo[direction], p[TP + direction + "_f1.0.weight"], p[TP + direction + "_f1.0.bias"]))
            f2 = F.relu(F.conv2d(xc, p[TP + direction + "_f2.0.weight"], p[TP + direction + "_f2.0.bias"]))
            q = (f1 - f2).pow(2).mean(dim=-1).mean(dim=-1)                             # :149
            a = F.relu(F.linear(q, p[TP + "channel_atte_" + atte + "_corr.0.weight"]))
            a = torch.sigmoid(F.linear(a, p[TP + "channel_atte_" + atte + "_corr.2.weight"]))
            x_temp = xc * a.view(b, c, 1, 1).expand(b, c, h, w) + xc                   # :150
            f_step[direction].append(x_temp.mean(dim=-1).mean(dim=-1))                 # :151
            memo[direction] = _ref_basic_block(p, TP + "uncorr_memo_" + direction, memo[direction], xu, training)
    f_fwd = torch.stack(f_step["forward"], dim=1)
    f_bwd = torch.stack([f_step["backward"][t - 1 - i] for i in range(t)], dim=1)     # :170-174
    f_corr = f_fwd + f_bwd                                                            # :176
    f_uncorr = memo["forward"].mean(dim=-1).mean(dim=-1) + memo["backward"].mean(dim=-1).mean(dim=-1)  # :178
    return f_uncorr, f_corr


def ref_forward(p, x, b, t, training=True):
    """GCE then TRL on synthetic layer4 maps x [b*t,2048,16,8].  Returns a dict."""
    x_uncorr, x_corr, corr_map = ref_gce(p, x, b, t, training)
    xc5 = x_corr.view(b, t, *x_corr.shape[1:])
    xu5 = x_uncorr.view(b, t, *x_uncorr.shape[1:])
    f_uncorr, f_corr = ref_trl(p, xu5, xc5, training)
    return dict(x_uncorr=x_uncorr, x_corr=x_corr, corr_map=corr_map, f_uncorr=f_uncorr, f_corr=f_corr)


def ref_tail(p, f_uncorr, f_corr, training):
    """grl_model.py:222-226 (needs corr_bn.* / uncorr_bn.* in p)."""
    b, t, c = f_corr.shape
    xc = F.normalize(_bn(p, "corr_bn", f_corr.reshape(b * t, c), training).view(b, t, c), p=2, dim=2)
    xu = F.normalize(_bn(p, "uncorr_bn", f_uncorr, training), p=2, dim=1)
    return xu, xc


# --------------------------------------------------------------------------------------
# 2. decomposed pixel-major plan (what the CUDA kernels implement) + explicit backward
# --------------------------------------------------------------------------------------
def to_pm(x):
    """NCHW [N,C,H,W] -> pixel-major [N*H*W, C]."""
    n, c = x.shape[:2]
    return x.reshape(n, c, -1).permute(0, 2, 1).reshape(-1, c)


def from_pm(xp, n):
    """pixel-major [N*S, C] -> NCHW [N,C,16,8]."""
    c = xp.shape[1]
    return xp.reshape(n, S, c).permute(0, 2, 1).reshape(n, c, 16, 8).contiguous()


def _bn_coeffs(p, prefix, Y, training, update):
    """Per-channel affine (a, c) with y_bn = a*Y + c, plus what backward needs."""
    gamma, beta = p[prefix + ".weight"], p[prefix + ".bias"]
    if training:
        n = Y.shape[0]
        mu = Y.mean(0)
        var = Y.var(0, unbiased=False)
        if update:
            with torch.no_grad():
                p[prefix + ".running_mean"].mul_(1 - MOM).add_(MOM * mu.detach())
                p[prefix + ".running_var"].mul_(1 - MOM).add_(MOM * var.detach() * n / max(n - 1, 1))
                p[prefix + ".num_batches_tracked"] += 1
    else:
        mu, var = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    rstd = torch.rsqrt(var + EPS)
    a = gamma * rstd
    c = beta - a * mu
    return a, c, mu, rstd


def _bn_bwd(dOut, Y, mu, rstd, gamma, training):
    xh = (Y - mu) * rstd
    dgamma = (dOut * xh).sum(0)
    dbeta = dOut.sum(0)
    if training:
        n = Y.shape[0]
        dY = gamma * rstd * (dOut - dbeta / n - xh * dgamma / n)
    else:
        dY = gamma * rstd * dOut
    return dY, dgamma, dbeta


def plan_gce_forward(p, x, b, t, training, update=True):
    ctx = {}
    X = to_pm(x)                                                     # [P,C]
    P, C = X.shape
    g = X.view(b, t * S, C).mean(1)                                  # K1 (F: any summation order)
    Wg, bg = p["backbone.glo_fc.0.weight"], p["backbone.glo_fc.0.bias"]
    u = g @ Wg.t() + bg                                              # K2
    ag, cg, mug, rstdg = _bn_coeffs(p, "backbone.glo_fc.1", u, training, update)
    glo = torch.relu(ag * u + cg)
    W1 = p["backbone.corr_atte.0.weight"].view(1024, 3072)
    W1a, W1b = W1[:, :2048], W1[:, 2048:]
    bias1 = glo @ W1b.t()                                            # F1: glo as per-clip bias
    Y1 = X @ W1a.t() + bias1.repeat_interleave(t * S, 0)             # K3
    a1, c1, mu1, rstd1 = _bn_coeffs(p, "backbone.corr_atte.1", Y1, training, update)
    W2 = p["backbone.corr_atte.2.weight"].view(256, 1024)
    Y2 = Y1 @ (W2 * a1).t() + W2 @ c1                                # K4 folded into K5's weights
    a2, c2, mu2, rstd2 = _bn_coeffs(p, "backbone.corr_atte.3", Y2, training, update)
    Z2 = torch.relu(a2 * Y2 + c2)
    w3 = p["backbone.corr_atte.5.weight"].view(256)
    y3 = (Z2 @ w3).view(P, 1)                                        # K6
    a3, c3, mu3, rstd3 = _bn_coeffs(p, "backbone.corr_atte.6", y3, training, update)
    m = torch.sigmoid(a3 * y3 + c3).view(P)
    Xc = X * m[:, None]                                              # K7 (planes only; F3)
    Xu = X * (1 - m)[:, None]
    ctx.update(X=X, g=g, u=u, mug=mug, rstdg=rstdg, glo=glo, Y1=Y1, a1=a1, c1=c1, mu1=mu1, rstd1=rstd1,
               Y2=Y2, mu2=mu2, rstd2=rstd2, Z2=Z2, y3=y3, mu3=mu3, rstd3=rstd3, m=m, b=b, t=t,
               training=training)
    return Xu, Xc, m, ctx


def plan_gce_backward(p, ctx, dXu, dXc, dm_extra=None):
    """dXu,dXc: [P,C] grads wrt x_uncorr / x_corr (pixel-major); dm_extra: [P] grad wrt corr_map."""
    c = ctx
    b, t, tr = c["b"], c["t"], c["training"]
    X, m = c["X"], c["m"]
    P, C = X.shape
    G = {}
    dX = dXc * m[:, None] + dXu * (1 - m)[:, None]
    dm = ((dXc - dXu) * X).sum(1)
    if dm_extra is not None:
        dm = dm + dm_extra
    dz3 = (dm * m * (1 - m)).view(P, 1)
    dy3, G["backbone.corr_atte.6.weight"], G["backbone.corr_atte.6.bias"] = _bn_bwd(
        dz3, c["y3"], c["mu3"], c["rstd3"], p["backbone.corr_atte.6.weight"], tr)
    w3 = p["backbone.corr_atte.5.weight"].view(256)
    G["backbone.corr_atte.5.weight"] = (dy3.t() @ c["Z2"]).view(1, 256, 1, 1)
    dZ2 = dy3 * w3[None, :]
    dA2 = dZ2 * (c["Z2"] > 0)
    dY2, G["backbone.corr_atte.3.weight"], G["backbone.corr_atte.3.bias"] = _bn_bwd(
        dA2, c["Y2"], c["mu2"], c["rstd2"], p["backbone.corr_atte.3.weight"], tr)
    W2 = p["backbone.corr_atte.2.weight"].view(256, 1024)
    # Y2 = Z1 @ W2^T with Z1 = a1*Y1 + c1 (BN folded):  dW2 = (dY2^T Y1) * a1 + colsum(dY2) (x) c1
    G["backbone.corr_atte.2.weight"] = ((dY2.t() @ c["Y1"]) * c["a1"][None, :]
                                        + dY2.sum(0)[:, None] * c["c1"][None, :]).view(256, 1024, 1, 1)
    dZ1 = dY2 @ W2
    dY1, G["backbone.corr_atte.1.weight"], G["backbone.corr_atte.1.bias"] = _bn_bwd(
        dZ1, c["Y1"], c["mu1"], c["rstd1"], p["backbone.corr_atte.1.weight"], tr)
    W1 = p["backbone.corr_atte.0.weight"].view(1024, 3072)
    W1a, W1b = W1[:, :2048], W1[:, 2048:]
    dW1a = dY1.t() @ X
    dX = dX + dY1 @ W1a
    dbias1 = dY1.view(b, t * S, 1024).sum(1)
    dW1b = dbias1.t() @ c["glo"]
    G["backbone.corr_atte.0.weight"] = torch.cat([dW1a, dW1b], 1).view(1024, 3072, 1, 1)
    dglo = dbias1 @ W1b
    dua = dglo * (c["glo"] > 0)
    du, G["backbone.glo_fc.1.weight"], G["backbone.glo_fc.1.bias"] = _bn_bwd(
        dua, c["u"], c["mug"], c["rstdg"], p["backbone.glo_fc.1.weight"], tr)
    G["backbone.glo_fc.0.weight"] = du.t() @ c["g"]
    G["backbone.glo_fc.0.bias"] = du.sum(0)
    dg = du @ p["backbone.glo_fc.0.weight"]
    dX = dX + (dg / (t * S)).repeat_interleave(t * S, 0)
    return dX, G


# The six contractions of the f1 / f2 attention convs (forward, input gradient, weight gradient) go through this hook so that
# tools/exp_f1f2_precision.py can emulate reduced-precision tensor-core variants of exactly those GEMMs ("Numerics").
MM_F12 = [torch.matmul]


def _mm12(a, b, role="fwd"):
    """role: "fwd" (F = X W^T), "dgrad" (dX = dF W), "wgrad" (dW = dF^T X); a hook may take (a, b) or (a, b, role)."""
    f = MM_F12[0]
    try:
        return f(a, b, role)
    except TypeError:
        return f(a, b)


# ... and the contractions of the memory block (conv1/2/3 of BasicBlock) through this one
MM_BLK = [torch.matmul]


def _mmblk(a, b, role="fwd"):
    f = MM_BLK[0]
    try:
        return f(a, b, role)
    except TypeError:
        return f(a, b)


def _rows(b, t, tau):
    """Row indices (in [P]) of frame tau of every clip, ordered (b, s)."""
    base = (torch.arange(b)[:, None] * t + tau) * S + torch.arange(S)[None, :]
    return base.reshape(-1)


def plan_trl_forward(p, Xu, Xc, b, t, training, update=True):
    P, C = Xc.shape
    ctx = dict(b=b, t=t, training=training, Xu=Xu, Xc=Xc, steps=[[], []])
    Gc = Xc.view(b, t, S, C).mean(2)                                 # F4: GAP(x_corr) once
    M0 = Xu.view(b, t, S, C).mean(1).reshape(b * S, C)               # K8 once for both directions
    out = [torch.zeros(b, t, C, dtype=Xc.dtype), torch.zeros(b, t, C, dtype=Xc.dtype)]
    Mfin = []
    F2 = []
    for d, (direction, atte) in enumerate(DIRS):
        Wf2 = p[TP + direction + "_f2.0.weight"].view(C, C)
        F2.append(torch.relu(_mm12(Xc, Wf2.t()) + p[TP + direction + "_f2.0.bias"]))   # F2: not recurrent
    for d, (direction, atte) in enumerate(DIRS):
        mp = TP + "uncorr_memo_" + direction
        Wf1 = p[TP + direction + "_f1.0.weight"].view(C, C)
        bf1 = p[TP + direction + "_f1.0.bias"]
        L1 = p[TP + "channel_atte_" + atte + "_corr.0.weight"]
        L2 = p[TP + "channel_atte_" + atte + "_corr.2.weight"]
        Wc1 = p[mp + ".conv1.weight"].view(512, C)
        Wc2 = p[mp + ".conv2.weight"].view(512, 512)
        Wc3 = p[mp + ".conv3.weight"].view(C, 512)
        M = M0
        for i in range(t):
            tau = i if d == 0 else t - 1 - i
            r = _rows(b, t, tau)
            F1 = torch.relu(_mm12(M, Wf1.t()) + bf1)                       # K10
            E = F1 - F2[d][r]
            q = (E * E).view(b, S, C).mean(1)                        # K11
            h = torch.relu(q @ L1.t())                               # K12
            a = torch.sigmoid(h @ L2.t())
            out[d][:, tau] = (1 + a) * Gc[:, tau]                    # K13 via F4
            Z = M + Xu[r]                                            # K14
            H1 = _mmblk(Z, Wc1.t())
            a1, c1, mu1, rs1 = _bn_coeffs(p, mp + ".bn1", H1, training, update)
            H1p = torch.relu(a1 * H1 + c1)
            H2 = _mmblk(H1p, Wc2.t())
            a2, c2, mu2, rs2 = _bn_coeffs(p, mp + ".bn2", H2, training, update)
            H2p = torch.relu(a2 * H2 + c2)
            H3 = _mmblk(H2p, Wc3.t())
            a3, c3, mu3, rs3 = _bn_coeffs(p, mp + ".bn3", H3, training, update)
            Mn = torch.relu(a3 * H3 + c3 + Z)
            ctx["steps"][d].append(dict(tau=tau, M=M, F1=F1, E=E, q=q, h=h, a=a, Z=Z, H1=H1, mu1=mu1, rs1=rs1,
                                        H1p=H1p, H2=H2, mu2=mu2, rs2=rs2, H2p=H2p, H3=H3, mu3=mu3, rs3=rs3,
                                        Mn=Mn))
            M = Mn
        Mfin.append(M)
    f_corr = out[0] + out[1]
    f_uncorr = Mfin[0].view(b, S, C).mean(1) + Mfin[1].view(b, S, C).mean(1)   # K15
    ctx.update(Gc=Gc, F2=F2)
    return f_uncorr, f_corr, ctx


def plan_trl_backward(p, ctx, d_f_uncorr, d_f_corr):
    b, t, tr = ctx["b"], ctx["t"], ctx["training"]
    Xu, Xc, Gc, F2 = ctx["Xu"], ctx["Xc"], ctx["Gc"], ctx["F2"]
    P, C = Xc.shape
    G = {}
    a_sum = torch.zeros(b, t, C, dtype=Xc.dtype)
    for d in range(2):
        for st in ctx["steps"][d]:
            a_sum[:, st["tau"]] += st["a"]
    dGc = d_f_corr * (2 + a_sum)
    dXc = (dGc / S)[:, :, None, :].expand(b, t, S, C).reshape(P, C).clone()
    dXu = torch.zeros(P, C, dtype=Xc.dtype)
    for d, (direction, atte) in enumerate(DIRS):
        mp = TP + "uncorr_memo_" + direction
        Wf1 = p[TP + direction + "_f1.0.weight"].view(C, C)
        Wf2 = p[TP + direction + "_f2.0.weight"].view(C, C)
        L1 = p[TP + "channel_atte_" + atte + "_corr.0.weight"]
        L2 = p[TP + "channel_atte_" + atte + "_corr.2.weight"]
        Wc1 = p[mp + ".conv1.weight"].view(512, C)
        Wc2 = p[mp + ".conv2.weight"].view(512, 512)
        Wc3 = p[mp + ".conv3.weight"].view(C, 512)
        acc = {k: 0 for k in ("Wc1", "Wc2", "Wc3", "g1", "b1", "g2", "b2", "g3", "b3", "Wf1", "bf1", "L1", "L2")}
        dF2 = torch.zeros(P, C, dtype=Xc.dtype)
        dM = (d_f_uncorr / S)[:, None, :].expand(b, S, C).reshape(b * S, C)
        for st in reversed(ctx["steps"][d]):
            r = _rows(b, t, st["tau"])
            dPre = dM * (st["Mn"] > 0)
            dH3, dg, db = _bn_bwd(dPre, st["H3"], st["mu3"], st["rs3"], p[mp + ".bn3.weight"], tr)
            acc["g3"] = acc["g3"] + dg; acc["b3"] = acc["b3"] + db
            acc["Wc3"] = acc["Wc3"] + _mmblk(dH3.t(), st["H2p"], "wgrad")
            dA2 = _mmblk(dH3, Wc3, "dgrad") * (st["H2p"] > 0)
            dH2, dg, db = _bn_bwd(dA2, st["H2"], st["mu2"], st["rs2"], p[mp + ".bn2.weight"], tr)
            acc["g2"] = acc["g2"] + dg; acc["b2"] = acc["b2"] + db
            acc["Wc2"] = acc["Wc2"] + _mmblk(dH2.t(), st["H1p"], "wgrad")
            dA1 = _mmblk(dH2, Wc2, "dgrad") * (st["H1p"] > 0)
            dH1, dg, db = _bn_bwd(dA1, st["H1"], st["mu1"], st["rs1"], p[mp + ".bn1.weight"], tr)
            acc["g1"] = acc["g1"] + dg; acc["b1"] = acc["b1"] + db
            acc["Wc1"] = acc["Wc1"] + _mmblk(dH1.t(), st["Z"], "wgrad")
            dZ = _mmblk(dH1, Wc1, "dgrad") + dPre
            dXu[r] += dZ
            # reciprocal-attention path
            da = d_f_corr[:, st["tau"]] * Gc[:, st["tau"]]
            ds = da * st["a"] * (1 - st["a"])
            acc["L2"] = acc["L2"] + ds.t() @ st["h"]
            dh = (ds @ L2) * (st["h"] > 0)
            acc["L1"] = acc["L1"] + dh.t() @ st["q"]
            dq = dh @ L1
            dE = (2.0 / S) * dq[:, None, :].expand(b, S, C).reshape(b * S, C) * st["E"]
            dF1 = dE * (st["F1"] > 0)
            dF2[r] = -dE * (F2[d][r] > 0)
            acc["Wf1"] = acc["Wf1"] + _mm12(dF1.t(), st["M"], "wgrad")
            acc["bf1"] = acc["bf1"] + dF1.sum(0)
            dM = dZ + _mm12(dF1, Wf1, "dgrad")
        dXu += (dM / t).view(b, 1, S, C).expand(b, t, S, C).reshape(P, C)
        G[TP + direction + "_f2.0.weight"] = _mm12(dF2.t(), Xc, "wgrad").view(C, C, 1, 1)
        G[TP + direction + "_f2.0.bias"] = dF2.sum(0)
        dXc += _mm12(dF2, Wf2, "dgrad")
        G[TP + direction + "_f1.0.weight"] = acc["Wf1"].view(C, C, 1, 1)
        G[TP + direction + "_f1.0.bias"] = acc["bf1"]
        G[TP + "channel_atte_" + atte + "_corr.0.weight"] = acc["L1"]
        G[TP + "channel_atte_" + atte + "_corr.2.weight"] = acc["L2"]
        G[mp + ".conv1.weight"] = acc["Wc1"].view(512, C, 1, 1)
        G[mp + ".conv2.weight"] = acc["Wc2"].view(512, 512, 1, 1)
        G[mp + ".conv3.weight"] = acc["Wc3"].view(C, 512, 1, 1)
        for k in (1, 2, 3):
            G[mp + ".bn%d.weight" % k] = acc["g%d" % k]
            G[mp + ".bn%d.bias" % k] = acc["b%d" % k]
    return dXu, dXc, G


def ctx_from_saved(sv, b, t):
    """Build the (gce_ctx, trl_ctx) of plan_*_backward from a forward's SAVED state (dict of tensors named as in
    plan_*_forward, e.g. what grl_b200.head.saved_state() reads out of the CUDA workspace), converted to float64.

    This lets a test run the fp64 oracle backward on exactly the activation pattern (ReLU masks, BN statistics) the
    CUDA forward produced: the backward kernels are then checked in isolation, free of the ReLU-kink sensitivity that
    makes end-to-end gradients of this head differ by ~1e-3 even between the reference's own fp32 and fp64 runs."""
    D = {k: v.detach().double().cpu() for k, v in sv.items()}
    P = b * t * S
    a2, c2 = D["bn2_stat"][0], D["bn2_stat"][1]
    g = dict(X=D["X"], g=D["g"], u=D["u"], mug=D["glo_stat"][2], rstdg=D["glo_stat"][3], glo=D["glo"], Y1=D["Y1"],
             a1=D["bn1_stat"][0], c1=D["bn1_stat"][1], mu1=D["bn1_stat"][2], rstd1=D["bn1_stat"][3], Y2=D["Y2"],
             mu2=D["bn2_stat"][2], rstd2=D["bn2_stat"][3], Z2=torch.relu(a2 * D["Y2"] + c2), y3=D["y3"].view(P, 1),
             mu3=D["bn3_stat"][2], rstd3=D["bn3_stat"][3], m=D["m"], b=b, t=t, training=True)
    F2 = [D["F2"][:, :2048], D["F2"][:, 2048:]]
    steps = [[], []]
    for d in range(2):
        for i in range(t):
            tau = i if d == 0 else t - 1 - i
            r = _rows(b, t, tau)
            steps[d].append(dict(tau=tau, M=D["mem"][i, d], F1=D["F1"][i, d], E=D["F1"][i, d] - F2[d][r], q=D["q"][i, d],
                                 h=D["h"][i, d], a=D["a"][i, d], Z=D["Z"][i, d], H1=D["H1"][i, d], mu1=D["sbn1"][i, d, 2],
                                 rs1=D["sbn1"][i, d, 3], H1p=D["H1p"][i, d], H2=D["H2"][i, d], mu2=D["sbn2"][i, d, 2],
                                 rs2=D["sbn2"][i, d, 3], H2p=D["H2p"][i, d], H3=D["H3"][i, d], mu3=D["sbn3"][i, d, 2],
                                 rs3=D["sbn3"][i, d, 3], Mn=D["mem"][i + 1, d]))
    tc = dict(b=b, t=t, training=True, Xu=D["Xu"], Xc=D["Xc"], steps=steps, Gc=D["Gc"].view(b, t, -1), F2=F2)
    return g, tc


def plan_backward_from_ctx(p, gctx, tctx, d_f_uncorr, d_f_corr):
    """fp64 oracle backward on a given forward context.  Returns (dx NCHW, {name: grad})."""
    n = gctx["b"] * gctx["t"]
    with torch.no_grad():
        dXu, dXc, G = plan_trl_backward(p, tctx, d_f_uncorr, d_f_corr)
        dX, G2 = plan_gce_backward(p, gctx, dXu, dXc)
    G.update(G2)
    return from_pm(dX, n), G


def plan_head(p, x, b, t, training=True, grads=None, update=True):
    """Fused head: forward, and backward when grads=(d_f_uncorr, d_f_corr) is given."""
    n = b * t
    with torch.no_grad():
        Xu, Xc, m, gctx = plan_gce_forward(p, x, b, t, training, update)
        f_uncorr, f_corr, tctx = plan_trl_forward(p, Xu, Xc, b, t, training, update)
        out = dict(f_uncorr=f_uncorr, f_corr=f_corr, corr_map=m.view(n, 1, 16, 8),
                   x_uncorr=from_pm(Xu, n), x_corr=from_pm(Xc, n))
        if grads is not None:
            dXu, dXc, G = plan_trl_backward(p, tctx, grads[0], grads[1])
            dX, G2 = plan_gce_backward(p, gctx, dXu, dXc)
            G.update(G2)
            out["dx"] = from_pm(dX, n)
            out["dxu_pm"], out["dxc_pm"] = dXu, dXc
            out["grads"] = G
    return out
