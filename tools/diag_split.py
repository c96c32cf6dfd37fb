"""GPU diagnostic: stand-alone GCE / TRL operators vs the fused head and the fp64 oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import head, synth  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


B, T = 4, 3
N = B * T
sd = {k: v.cuda().contiguous() for k, v in synth.make_head_params(0).items()}
x = synth.make_head_input(B, T).cuda()
gu, gc = synth.make_head_grads(B, T)
fu, fc, cm, xu_f, xc_f, ws = head.head_forward_raw(sd, x, B, T, True, save=True, want_maps=True)
dx_f, g_f = head.head_backward_raw(sd, x, B, T, ws, gu.cuda(), gc.cuda())

sd2 = {k: v.cuda().contiguous() for k, v in synth.make_head_params(0).items()}
xu, xc, cm2, ws_g = head.gce_forward_raw(sd2, x, B, T, True, save=True)
print("xu", rel(xu, xu_f), "xc", rel(xc, xc_f), "cm", rel(cm2, cm))
fu2, fc2, ws_t = head.trl_forward_raw(sd2, xu, xc, B, T, True, save=True)
print("fu", rel(fu2, fu), "fc", rel(fc2, fc))
dxu, dxc, g_t = head.trl_backward_raw(sd2, B, T, ws_t, gu.cuda(), gc.cuda())
p64 = synth.make_head_params(0, dtype=torch.float64)
o = ho.plan_head(p64, synth.make_head_input(B, T, dtype=torch.float64), B, T, True, grads=(gu.double(), gc.double()))
print("dxu vs oracle", rel(dxu, ho.from_pm(o["dxu_pm"], N)), "dxc vs oracle", rel(dxc, ho.from_pm(o["dxc_pm"], N)))
for k in list(g_t)[:6]:
    print("  trl grad", k, rel(g_t[k], g_f[k]))
dx2, g_g = head.gce_backward_raw(sd2, B, T, ws_g, dxu, dxc, None)
print("dx split vs fused", rel(dx2, dx_f), " split vs oracle", rel(dx2, o["dx"]), " fused vs oracle", rel(dx_f, o["dx"]))
for k in g_g:
    print("  gce grad", k, rel(g_g[k], g_f[k]))
