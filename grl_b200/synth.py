"""Seeded synthetic inputs for the GRL hot path.

Everything here is *data generation*, not algorithm: it is shared by the tests,
the golden-vector script, smoke() and bench.py so that
the reference, the oracle and the CUDA path all see identical tensors.

numpy's PCG64 `default_rng` is used (stream stability is guaranteed by numpy),
never torch's global RNG, so fixtures regenerate bit-identically anywhere.

Parameter names/shapes follow the reference `state_dict` keys:
  reid/models/basebranch.py:38-50   (GCE: glo_fc, corr_atte)
  reid/models/grl_model.py:88-128   (TRL: *_f1, *_f2, channel_atte_*, uncorr_memo_*)
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import numpy as np
import torch

C_FEAT = 2048      # layer4 channels           (basebranch.py:43)
C_GLO = 1024       # glo_fc width              (basebranch.py:38)
C_MID = 256        # corr_atte middle width    (basebranch.py:45)
C_MEMO = 512       # BasicBlock planes         (grl_model.py:92)
C_SE = 128         # 2048 // 16                (grl_model.py:104)
H_MAP, W_MAP = 16, 8   # hard-coded spatial size (basebranch.py:59)
S_MAP = H_MAP * W_MAP


def head_param_shapes() -> "OrderedDict[str, tuple]":
    """state_dict keys (minus `backbone.base.*`) of ResNet50_GRL_Model's head."""
    d: "OrderedDict[str, tuple]" = OrderedDict()

    def bn(prefix, n):
        d[prefix + ".weight"] = (n,)
        d[prefix + ".bias"] = (n,)
        d[prefix + ".running_mean"] = (n,)
        d[prefix + ".running_var"] = (n,)
        d[prefix + ".num_batches_tracked"] = ()

    d["backbone.glo_fc.0.weight"] = (C_GLO, C_FEAT)
    d["backbone.glo_fc.0.bias"] = (C_GLO,)
    bn("backbone.glo_fc.1", C_GLO)
    d["backbone.corr_atte.0.weight"] = (C_GLO, C_FEAT + C_GLO, 1, 1)
    bn("backbone.corr_atte.1", C_GLO)
    d["backbone.corr_atte.2.weight"] = (C_MID, C_GLO, 1, 1)
    bn("backbone.corr_atte.3", C_MID)
    d["backbone.corr_atte.5.weight"] = (1, C_MID, 1, 1)
    bn("backbone.corr_atte.6", 1)
    t = "temporal_learning_block."
    for direction, atte in (("forward", "foreward"), ("backward", "backward")):
        m = t + "uncorr_memo_" + direction
        d[m + ".conv1.weight"] = (C_MEMO, C_FEAT, 1, 1)
        bn(m + ".bn1", C_MEMO)
        d[m + ".conv2.weight"] = (C_MEMO, C_MEMO, 1, 1)
        bn(m + ".bn2", C_MEMO)
        d[m + ".conv3.weight"] = (C_FEAT, C_MEMO, 1, 1)
        bn(m + ".bn3", C_FEAT)
        for f in ("_f1", "_f2"):
            d[t + direction + f + ".0.weight"] = (C_FEAT, C_FEAT, 1, 1)
            d[t + direction + f + ".0.bias"] = (C_FEAT,)
        d[t + "channel_atte_" + atte + "_corr.0.weight"] = (C_SE, C_FEAT)
        d[t + "channel_atte_" + atte + "_corr.2.weight"] = (C_FEAT, C_SE)
    return d


_BN_RE = re.compile(r"(\.bn[123]|glo_fc\.1|corr_atte\.[136])\.(weight|bias)$")


def _is_bn(name: str) -> bool:
    return _BN_RE.search(name) is not None


def make_head_params(seed: int = 0, randomize_bn: bool = True, dtype=torch.float32):
    """PyTorch-default-style init (U(-1/sqrt(fan_in), +1/sqrt(fan_in))) from PCG64.

    With `randomize_bn` the BN affine parameters and running buffers are perturbed
    so that eval-mode and backward tests are not degenerate (gamma=1, beta=0,
    mean=0, var=1 would hide mistakes in exactly those terms).
    """
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in head_param_shapes().items():
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros((), dtype=torch.int64)
            continue
        if name.endswith("running_mean"):
            v = rng.standard_normal(shape, dtype=np.float32) * 0.05 if randomize_bn else np.zeros(shape, np.float32)
        elif name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
        elif _is_bn(name):
            if name.endswith(".weight"):
                v = rng.uniform(0.6, 1.4, shape).astype(np.float32) if randomize_bn else np.ones(shape, np.float32)
            else:
                v = (rng.standard_normal(shape, dtype=np.float32) * 0.1) if randomize_bn else np.zeros(shape, np.float32)
        else:
            fan_in = shape[1] if len(shape) > 1 else None
            if fan_in is None:  # conv / linear bias: fan_in of the owning layer
                fan_in = C_FEAT
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, shape).astype(np.float32)
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return out


def make_head_input(B: int, T: int, seed: int = 123, dtype=torch.float32) -> torch.Tensor:
    """Structured layer4-like maps, NCHW [B*T, 2048, 16, 8].

    x = relu(mu[b,c] + 0.3 phi[b,t,c] + 0.5 sigma[b,t,h,w] + 0.5 eps): clips differ
    (so glo_fc's BatchNorm1d over B samples is well conditioned) and frames/pixels
    differ inside a clip (so the correlation map is not constant).
    """
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal((B, 1, C_FEAT, 1, 1), dtype=np.float32)
    phi = rng.standard_normal((B, T, C_FEAT, 1, 1), dtype=np.float32)
    sig = rng.standard_normal((B, T, 1, H_MAP, W_MAP), dtype=np.float32)
    eps = rng.standard_normal((B, T, C_FEAT, H_MAP, W_MAP), dtype=np.float32)
    x = np.maximum(mu + 0.3 * phi + 0.5 * sig + 0.5 * eps, 0.0).astype(np.float32)
    return torch.from_numpy(x.reshape(B * T, C_FEAT, H_MAP, W_MAP)).to(dtype)


def make_head_grads(B: int, T: int, seed: int = 1, dtype=torch.float32):
    """Upstream gradients on (f_uncorr [B,2048], f_corr [B,T,2048])."""
    rng = np.random.default_rng(seed)
    gu = rng.standard_normal((B, C_FEAT), dtype=np.float32)
    gc = rng.standard_normal((B, T, C_FEAT), dtype=np.float32)
    return torch.from_numpy(gu).to(dtype), torch.from_numpy(gc).to(dtype)


def make_eval_set(num_q: int, num_g_extra: int, dim: int, seed: int = 0, num_ids: int = 625,
                  num_cams: int = 6, noise: float = 0.3, distractors: float = 0.1,
                  missing_query_frac: float = 0.02):
    """Clustered re-ID features + ids shaped like ATTEvaluator.evaluate's inputs.

    Mirrors reid/evaluator/attevaluator.py:143-145: the gallery is
    `cat(queries, extra gallery)`, so every query's own copy (same pid, same cam)
    is in the gallery and must be removed as junk (eva_functions.py:153-154).
    A few queries get a pid that never appears elsewhere -> "no valid match"
    (eva_functions.py:159-161).  pid 0 rows act as distractors.
    Returns float32 features (unit norm) and int64 pid / camid arrays.
    """
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((num_ids + 1, dim), dtype=np.float32)
    q_pid = rng.integers(1, num_ids + 1, num_q)
    q_cam = rng.integers(0, num_cams, num_q)
    n_missing = int(round(missing_query_frac * num_q))
    if n_missing:
        # ids above num_ids never occur in the extra gallery
        q_pid[:n_missing] = num_ids + 1 + np.arange(n_missing)
    g_pid_x = rng.integers(1, num_ids + 1, num_g_extra)
    g_pid_x[rng.random(num_g_extra) < distractors] = 0
    g_cam_x = rng.integers(0, num_cams, num_g_extra)

    def feats(pid):
        base = cent[np.minimum(pid, num_ids)]
        # ids above num_ids share centroid `num_ids` but that is irrelevant: they have no positives
        f = base + noise * rng.standard_normal(base.shape, dtype=np.float32)
        f = f / np.linalg.norm(f, axis=1, keepdims=True)
        return f.astype(np.float32)

    qf = feats(q_pid)
    gf_x = feats(g_pid_x)
    gf = np.concatenate([qf, gf_x], 0)
    g_pid = np.concatenate([q_pid, g_pid_x])
    g_cam = np.concatenate([q_cam, g_cam_x])
    return qf, gf, q_pid.astype(np.int64), g_pid.astype(np.int64), q_cam.astype(np.int64), g_cam.astype(np.int64)


# ------------------------------------------------------------------------------------------------
# Eval feature tail: ResNet50_GRL_Model.corr_bn / uncorr_bn (grl_model.py:203-209) and the
# Siamese temporal self-attention (reid/models/Siamese.py:43-76; input_num=2048, output_num=512, mars_train.py:77)
# ------------------------------------------------------------------------------------------------
C_ATT = 512


def tail_param_shapes() -> "OrderedDict[str, tuple]":
    d: "OrderedDict[str, tuple]" = OrderedDict()

    def bn(prefix, n):
        d[prefix + ".weight"] = (n,)
        d[prefix + ".bias"] = (n,)
        d[prefix + ".running_mean"] = (n,)
        d[prefix + ".running_var"] = (n,)

    bn("corr_bn", C_FEAT)
    bn("uncorr_bn", C_FEAT)
    for name in ("featQ", "featK"):
        d["siamese." + name + ".weight"] = (C_ATT, C_FEAT)
        d["siamese." + name + ".bias"] = (C_ATT,)
        bn("siamese." + name + "_bn", C_ATT)
    return d


def make_tail_params(seed: int = 10, dtype=torch.float32):
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in tail_param_shapes().items():
        if name.endswith("running_mean"):
            v = rng.standard_normal(shape, dtype=np.float32) * 0.05
        elif name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif "bn" in name and name.endswith(".weight"):
            v = rng.uniform(0.6, 1.4, shape).astype(np.float32)
        elif "bn" in name:
            v = rng.standard_normal(shape, dtype=np.float32) * 0.1
        elif name.endswith(".weight"):
            v = (rng.standard_normal(shape, dtype=np.float32) * (2.0 / math.sqrt(C_FEAT))).astype(np.float32)
        else:
            v = rng.standard_normal(shape, dtype=np.float32) * 0.1
        out[name] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return out


def make_tail_input(n: int, T: int, seed: int = 11, dtype=torch.float32):
    """Head outputs (f_uncorr [n,2048], f_corr [n,T,2048]): positive, clip- and frame-dependent like pooled ReLU maps."""
    rng = np.random.default_rng(seed)
    base = np.abs(rng.standard_normal((n, 1, C_FEAT), dtype=np.float32))
    fc = np.abs(base + 0.4 * rng.standard_normal((n, T, C_FEAT), dtype=np.float32)).astype(np.float32)
    fu = np.abs(base[:, 0] + 0.4 * rng.standard_normal((n, C_FEAT), dtype=np.float32)).astype(np.float32)
    return torch.from_numpy(fu).to(dtype), torch.from_numpy(fc).to(dtype)


# ------------------------------------------------------------------------------------------------
# Loss neighbours: clip descriptors, identity labels and an OIM look-up table.
# ------------------------------------------------------------------------------------------------
def make_loss_inputs(B: int, D: int, C: int, seed: int = 0, n_ids: int = 8):
    """Unit-norm descriptors clustered by identity (like the model's normalised outputs), labels with repeats (the
    sampler draws several clips per identity: the OIM update must compose them in batch order), a row-normalised table.
    One identity appears once only (no positive) and two rows are duplicates (zero distance)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, n_ids, (B,), generator=g)
    ids[0] = n_ids                                   # singleton identity: no positive in the batch
    cent = torch.randn((n_ids + 1, D), generator=g)
    feat = torch.nn.functional.normalize(cent[ids] + 0.7 * torch.randn((B, D), generator=g), dim=1)
    feat[2] = feat[1]
    ids[2] = ids[1]
    lut = torch.nn.functional.normalize(torch.randn((C, D), generator=g), dim=1)
    targets = (ids * 7 + 3) % C                      # class indices spread over the table, repeats preserved
    return feat.float(), ids.long(), lut.float(), targets.long()


def make_siamese_inputs(n2: int, T: int, seed: int = 0):
    """Parameters of Siamese(2048, 512, 2) under the reference's key names (kaiming-style Q/K/classifier weights, non-trivial
    BN affine parameters and running buffers), normalised clip features x [2n, T, 2048] clustered by identity, seeded
    upstream gradients for both outputs, and identity labels [2n] (pairs (2i, 2i+1) share an identity half of the time)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *shape: torch.randn(shape, generator=g)
    p = {}
    for nm in ("featQ", "featK", "featV"):
        p[nm + ".weight"] = r(512, 2048) * (2.0 / 2048) ** 0.5
        p[nm + ".bias"] = 0.1 * r(512)
        p[nm + "_bn.weight"] = 1.0 + 0.2 * r(512)
        p[nm + "_bn.bias"] = 0.1 * r(512)
        p[nm + "_bn.running_mean"] = 0.1 * r(512)
        p[nm + "_bn.running_var"] = 1.0 + 0.2 * torch.rand(512, generator=g)
        p[nm + "_bn.num_batches_tracked"] = torch.tensor(3, dtype=torch.long)
    p["classifierBN.weight"] = 1.0 + 0.2 * r(2048)
    p["classifierBN.bias"] = 0.1 * r(2048)
    p["classifierBN.running_mean"] = 0.001 * torch.rand(2048, generator=g)
    p["classifierBN.running_var"] = 1e-6 * (1.0 + torch.rand(2048, generator=g))
    p["classifierBN.num_batches_tracked"] = torch.tensor(3, dtype=torch.long)
    p["classifierlinear.weight"] = 0.05 * r(2, 2048)
    p["classifierlinear.bias"] = 0.1 * r(2)
    n_ids = max(2, n2 // 4)
    ids = torch.randint(0, n_ids, (n2,), generator=g)
    ids[1::4] = ids[0::4][: ids[1::4].numel()]          # every other pair shares its identity
    cent = r(n_ids, 2048)
    x = torch.nn.functional.normalize(cent[ids].unsqueeze(1) + 0.5 * r(n2, 1, 2048) + 0.3 * r(n2, T, 2048), dim=2)
    d_cls = r(n2 // 2, n2 // 2, 2)
    d_out = r(n2, 2048)
    return p, x.float(), d_cls.float(), d_out.float(), ids.long()
