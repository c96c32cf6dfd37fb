"""GPU: the whole post-backbone training graph of SEQTrainer._forward (reid/train/trainer.py:107-166) composed from this
library's autograd nodes -- fused head -> corr_bn / uncorr_bn tail (PyTorch) -> Siamese pair classifier -> OIM (frame and clip
level, one shared table), batch-hard triplet, PairLoss -- runs forward + backward, gives finite gradients to the layer4 maps and
to every parameter, and its loss agrees with the fp64 oracle chain."""
import numpy as np
import pytest
import torch

from grl_b200 import synth

pytestmark = pytest.mark.gpu


def test_training_loss_graph_composes_and_matches_oracle():
    from grl_b200 import head
    from grl_b200.losses import OIMLoss, TripletLoss
    from grl_b200.siamese import PairLoss, Siamese
    from oracle import head_oracle as ho
    from oracle import loss_oracle as lo
    B, T, C = 4, 3, 40
    targets = torch.tensor([5, 5, 9, 17])                     # pairs (0,1) same identity, (2,3) different
    hp = synth.make_head_params(0)
    sp, _, _, _, _ = synth.make_siamese_inputs(B, T, 41)
    lut = torch.nn.functional.normalize(torch.randn((C, 2048), generator=torch.Generator().manual_seed(3)), dim=1)
    x = synth.make_head_input(B, T)

    model = head.ResNet50_GRL_Model(base=torch.nn.Identity())
    msd = model.state_dict()
    for k, v in hp.items():
        msd[k] = v
    model.load_state_dict(msd)
    sia = Siamese(2048, 512, 2)
    ssd = sia.state_dict()
    for k, v in sp.items():
        ssd[k] = v.clone()
    sia.load_state_dict(ssd)
    model, sia = model.cuda().train(), sia.cuda().train()
    crit_corr, crit_unc = OIMLoss(2048, C, scalar=30.0, momentum=0.5).cuda(), OIMLoss(2048, C, scalar=30.0, momentum=0.5).cuda()
    crit_corr.lut.copy_(lut)
    crit_unc.lut.copy_(lut)
    tri, ver = TripletLoss('soft', True), PairLoss()

    xin = x.cuda().requires_grad_(True)
    tg = targets.cuda()
    f_uncorr, f_corr, _, _, _ = model.head(xin, B, T)
    x_corr = torch.nn.functional.normalize(model.corr_bn(f_corr.view(B * T, 2048)).view(B, T, 2048), p=2, dim=2)   # grl_model.py:222-226
    x_uncorr = torch.nn.functional.normalize(model.uncorr_bn(f_uncorr), p=2, dim=1)
    loss_frame, _ = crit_corr(x_corr.view(B * T, -1), tg.unsqueeze(1).expand(B, T).reshape(-1))               # trainer.py:118-127
    tv = tg.view(B // 2, -1)
    target = torch.cat((tv[:, 0], tv[:, 1]))
    encode_scores, siamese_out = sia(x_corr)                                                                      # :137
    loss_vid, _ = crit_corr(siamese_out, target)
    loss_tri = tri(siamese_out, target).mean()
    n = B // 2
    score = torch.softmax(encode_scores.view(-1, 2), dim=-1).view(n, n, 2)[:, :, 1]
    loss_ver, prec = ver(score, tv[:, 0], tv[:, 1])
    loss_unc, _ = crit_unc(x_uncorr, tg)                      # the commented-out direct use of x_uncorr, trainer.py:113
    total = loss_unc + loss_frame + loss_vid + loss_ver * 20 + loss_tri                                           # :160-164
    total.backward()

    assert torch.isfinite(total) and torch.isfinite(xin.grad).all() and float(xin.grad.abs().max()) > 0
    for name, prm in list(model.named_parameters()) + list(sia.named_parameters()):
        if name.startswith("featV") or name.startswith("classifier."):
            continue                                          # unused by these forward paths, as in the reference
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), name
    assert not torch.equal(crit_corr.lut.cpu(), lut) and not torch.equal(crit_unc.lut.cpu(), lut)     # tables updated in backward

    # fp64 oracle chain, forward values
    p64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in hp.items()}
    for nm in ("corr_bn", "uncorr_bn"):                       # the tail BNs at their nn.BatchNorm1d defaults, like `model` above
        p64.update({nm + ".weight": torch.ones(2048, dtype=torch.float64), nm + ".bias": torch.zeros(2048, dtype=torch.float64),
                    nm + ".running_mean": torch.zeros(2048, dtype=torch.float64), nm + ".running_var": torch.ones(2048, dtype=torch.float64),
                    nm + ".num_batches_tracked": torch.tensor(0)})
    o = ho.ref_forward(p64, x.double(), B, T, True)
    xu64, xc64 = ho.ref_tail(p64, o["f_uncorr"], o["f_corr"], True)
    s64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sp.items()}
    cls64, out64 = lo.siamese_forward(s64, xc64, True)
    l_frame = lo.oim_loss(xc64.reshape(B * T, -1), targets.unsqueeze(1).expand(B, T).reshape(-1), lut.double(), 30.0, 0.5)[0]
    tgt = torch.cat((targets.view(n, -1)[:, 0], targets.view(n, -1)[:, 1]))
    l_vid = lo.oim_loss(out64, tgt, lut.double(), 30.0, 0.5)[0]
    l_tri = lo.triplet_loss(out64, tgt, 'soft')[0].mean()
    sc64 = torch.softmax(cls64.reshape(-1, 2), dim=-1).view(n, n, 2)[:, :, 1]
    l_ver = lo.pair_loss(sc64, targets.view(n, -1)[:, 0], targets.view(n, -1)[:, 1])[0]
    l_unc = lo.oim_loss(xu64, targets, lut.double(), 30.0, 0.5)[0]
    total64 = float(l_unc + l_frame + l_vid + l_ver * 20 + l_tri)
    parts = dict(frame=(float(loss_frame), float(l_frame)), vid=(float(loss_vid), float(l_vid)), tri=(float(loss_tri), float(l_tri)),
                 ver=(float(loss_ver), float(l_ver)), unc=(float(loss_unc), float(l_unc)))
    for k, (a, b) in parts.items():
        assert abs(a - b) < 1e-3 * max(1.0, abs(b)), (k, a, b)
    assert abs(float(total) - total64) < 1e-3 * abs(total64), (float(total), total64)
