"""GPU diagnostic: head forward (and backward when available) vs the fp64 oracle plan, per intermediate."""
import copy
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import _lib, head, synth  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

dev = torch.device("cuda")
DO_BWD = "--bwd" in sys.argv


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def line(name, r, tol=1e-3):
    print("  %-34s rel=%.3e %s" % (name, r, "" if r < tol else "  <-- BAD"), flush=True)
    return r < tol


def run(B, T, training):
    print("=== B=%d T=%d training=%s" % (B, T, training), flush=True)
    p64 = synth.make_head_params(0, dtype=torch.float64)
    x64 = synth.make_head_input(B, T, dtype=torch.float64)
    gu, gc = synth.make_head_grads(B, T, dtype=torch.float64)
    t0 = time.time()
    pref = copy.deepcopy(p64)
    with torch.no_grad():
        Xu, Xc, m, gctx = ho.plan_gce_forward(pref, x64, B, T, training)
        fu, fc, tctx = ho.plan_trl_forward(pref, Xu, Xc, B, T, training)
    print("  oracle fwd %.1fs" % (time.time() - t0), flush=True)
    sd = {k: v.float().to(dev).contiguous() for k, v in synth.make_head_params(0).items()}
    x = x64.float().to(dev)
    save = training
    f_uncorr, f_corr, corr_map, xu, xc, ws = head.head_forward_raw(sd, x, B, T, training, save, want_maps=True)
    torch.cuda.synchronize()
    ok = True
    N, P, R = B * T, B * T * 128, B * 128
    V = lambda name, dt, shape: head.ws_view(ws, B, T, save, name, dt, shape)
    planes = lambda name, shape: V(name + "_hi", torch.bfloat16, shape).float() + V(name + "_lo", torch.bfloat16, shape).float()
    ok &= line("xp planes", rel(planes("xp", (P, 2048)), gctx["X"]), 1e-4)
    ok &= line("g", rel(V("g", torch.float32, (B, 2048)), gctx["g"]))
    ok &= line("u", rel(V("u", torch.float32, (B, 1024)), gctx["u"]))
    ok &= line("glo", rel(V("glo", torch.float32, (B, 1024)), gctx["glo"]))
    ok &= line("y1 planes", rel(planes("y1", (P, 1024)), gctx["Y1"]))
    ok &= line("y2", rel(V("y2", torch.float32, (P, 256)), gctx["Y2"]))
    ok &= line("y3", rel(V("y3", torch.float32, (P,)), gctx["y3"]))
    ok &= line("m / corr_map", rel(corr_map, m))
    ok &= line("x_corr (nchw)", rel(xc, ho.from_pm(Xc, N)))
    ok &= line("x_uncorr (nchw)", rel(xu, ho.from_pm(Xu, N)))
    ok &= line("gc", rel(V("gc", torch.float32, (N, 2048)), tctx["Gc"].reshape(N, 2048)))
    f2 = V("f2", torch.float32, (P, 4096))
    ok &= line("f2 fwd", rel(f2[:, :2048], tctx["F2"][0]))
    ok &= line("f2 bwd", rel(f2[:, 2048:], tctx["F2"][1]))
    if save:
        for d in range(2):
            for i in (0, T - 1):
                st = tctx["steps"][d][i]
                mem = planes("mem", (T + 1, 2, R, 2048))
                ok &= line("d%d step%d M" % (d, i), rel(mem[i, d], st["M"]))
                ok &= line("d%d step%d f1" % (d, i), rel(V("f1", torch.float32, (T, 2, R, 2048))[i, d], st["F1"]))
                ok &= line("d%d step%d q" % (d, i), rel(V("se_q", torch.float32, (T, 2, B, 2048))[i, d], st["q"]))
                ok &= line("d%d step%d a" % (d, i), rel(V("se_a", torch.float32, (T, 2, B, 2048))[i, d], st["a"]))
                ok &= line("d%d step%d z" % (d, i), rel(planes("z", (T, 2, R, 2048))[i, d], st["Z"]))
                ok &= line("d%d step%d h1" % (d, i), rel(V("h1", torch.float32, (T, 2, R, 512))[i, d], st["H1"]))
                ok &= line("d%d step%d h1p" % (d, i), rel(planes("h1p", (T, 2, R, 512))[i, d], st["H1p"]))
                ok &= line("d%d step%d h2" % (d, i), rel(V("h2", torch.float32, (T, 2, R, 512))[i, d], st["H2"]))
                ok &= line("d%d step%d h3" % (d, i), rel(V("h3", torch.float32, (T, 2, R, 2048))[i, d], st["H3"]))
                ok &= line("d%d step%d Mn" % (d, i), rel(mem[i + 1, d], st["Mn"]))
    ok &= line("f_uncorr", rel(f_uncorr, fu))
    ok &= line("f_corr", rel(f_corr, fc))
    if training:
        worst = 0.0
        for k in pref:
            if "running" in k:
                worst = max(worst, rel(sd[k], pref[k]))
        ok &= line("BN running buffers (worst)", worst, 1e-5)
    if DO_BWD and training:
        t0 = time.time()
        with torch.no_grad():
            dXu, dXc, G = ho.plan_trl_backward(pref, tctx, gu, gc)
            dX, G2 = ho.plan_gce_backward(pref, gctx, dXu, dXc)
            G.update(G2)
        print("  oracle bwd %.1fs" % (time.time() - t0), flush=True)
        dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu.float().to(dev), gc.float().to(dev))
        torch.cuda.synchronize()
        ok &= line("dx", rel(dx, ho.from_pm(dX, N)), 2e-3)
        for k in head.head_param_names():
            g_ref = G[k].reshape(grads[k].shape)
            if float(g_ref.norm()) < 1e-9:
                r = float(grads[k].double().cpu().norm())      # exactly-zero gradients: absolute
                ok &= line("grad " + k + " (abs)", r, 1e-4)
            else:
                ok &= line("grad " + k, rel(grads[k], g_ref), 2e-3)
    print("  RESULT:", "OK" if ok else "MISMATCH", flush=True)
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(), _lib.load_library().grl_version().decode())
    allok = True
    for (B, T, tr) in ((2, 3, True), (4, 2, True), (3, 4, False)):
        try:
            allok &= run(B, T, tr)
        except Exception as e:  # noqa
            import traceback
            traceback.print_exc()
            allok = False
            if "CUDA" in str(e) or "cuda" in str(e):
                break
    print("DIAG HEAD DONE:", "ALL OK" if allok else "FAILURES")
