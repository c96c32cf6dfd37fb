"""CPU experiment: can the f1 / f2 attention convs -- 70 % of the head's FLOPs, feeding only
mean_hw((F1 - F2)^2) -> SE MLP -> sigmoid -- run as fewer tensor-core MMAs per product than the split-bf16 three?
Emulated on the fp64 oracle plan: the operands of exactly those six contractions (forward, dgrad, wgrad of f1 and f2) are
rounded the way each variant's planes would hold them, products and sums stay exact (fp64).  Reported: error of every head
output and gradient against the unmodified fp64 plan, next to the gate each one has in tests/test_gpu_head.py.

  python tools/exp_f1f2_precision.py [B T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import synth  # noqa: E402
from oracle import head_oracle as ho  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 3)


def bf16(x):
    return x.float().bfloat16().double()          # (inputs are O(1): the float cast does not round beyond bf16's 8 bits)


def fp16_scaled(x):
    """fp16 after a per-row power-of-two scaling (what f16_rows_kernel does for the retrieval operands): no overflow, exact descale"""
    m = x.abs().amax(dim=-1, keepdim=True).clamp_min(1e-300)
    s = torch.exp2(14 - torch.floor(torch.log2(m)))
    return (x * s).float().half().double() / s


def split(x):
    hi = bf16(x)
    return hi, bf16(x - hi)


def f16mm(a, b):
    return fp16_scaled(a) @ fp16_scaled(b.t()).t()


def x3(a, b):
    ah, al = split(a)
    bh, bl = split(b)
    return ah @ bh + al @ bh + ah @ bl


def mixed(fwd, dgrad, wgrad):
    table = {"fwd": fwd, "dgrad": dgrad, "wgrad": wgrad}
    return lambda a, b, role="fwd": table[role](a, b)


VARIANTS = {
    "3 MMA split-bf16 (today)": lambda a, b: (lambda ah, al, bh, bl: ah @ bh + al @ bh + ah @ bl)(*split(a), *split(b)),
    "2 MMA bf16 (hi*hi + lo*hi)": lambda a, b: (lambda ah, al, bh, bl: ah @ bh + al @ bh)(*split(a), *split(b)),
    "1 MMA fp16 (row-scaled)": lambda a, b: fp16_scaled(a) @ fp16_scaled(b.t()).t(),
    "1 MMA bf16": lambda a, b: bf16(a) @ bf16(b),
    "fwd x3, dgrad+wgrad fp16 x1": mixed(x3, f16mm, f16mm),
    "fwd x3, wgrad fp16 x1": mixed(x3, x3, f16mm),
    "fwd x3, dgrad fp16 x1": mixed(x3, f16mm, x3),
    "fwd fp16 x1, bwd x3": mixed(f16mm, x3, x3),
}


def run(mm, blk=torch.matmul):
    ho.MM_F12[0] = mm
    ho.MM_BLK[0] = blk
    p = synth.make_head_params(0, dtype=torch.float64)
    x = synth.make_head_input(B, T, dtype=torch.float64)
    gu, gc = synth.make_head_grads(B, T)
    return ho.plan_head(p, x, B, T, True, grads=(gu.double(), gc.double()))


rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-300))
ref = run(torch.matmul)
print("B=%d T=%d; error vs the unmodified fp64 plan (gates: outputs 1e-4, gradients 1e-3)" % (B, T))
for name, mm in VARIANTS.items() if "--all" in sys.argv else []:
    o = run(mm)
    worst = sorted(((rel(o["grads"][k], ref["grads"][k]), k) for k in ref["grads"] if float(ref["grads"][k].norm()) > 1e-9), reverse=True)
    print("%-30s f_corr %.1e  dx %.1e  worst gradients: %s" %
          (name, rel(o["f_corr"], ref["f_corr"]), rel(o["dx"], ref["dx"]), ", ".join("%s %.1e" % (k.split("block.")[-1], e) for e, k in worst[:4])), flush=True)
# ---- the memory block (conv1/2/3 of BasicBlock): f1/f2 as adopted (forward x3, gradients fp16 x1) + block variants
adopted = VARIANTS["fwd x3, dgrad+wgrad fp16 x1"]
print("memory block variants (f1/f2 as adopted):")
for name, blk in (("block fwd exact, wgrad fp16 x1", mixed(torch.matmul, torch.matmul, f16mm)),
                  ("block fwd exact, dgrad+wgrad fp16 x1", mixed(torch.matmul, f16mm, f16mm)), ("block all x3", mixed(x3, x3, x3)), ("block fwd x3, dgrad+wgrad fp16 x1", mixed(x3, f16mm, f16mm)),
                  ("block fwd x3, wgrad fp16 x1", mixed(x3, x3, f16mm)), ("block all fp16 x1", mixed(f16mm, f16mm, f16mm))):
    o = run(adopted, blk)
    worst = sorted(((rel(o["grads"][k], ref["grads"][k]), k) for k in ref["grads"] if float(ref["grads"][k].norm()) > 1e-9), reverse=True)
    print("%-38s f_uncorr %.1e f_corr %.1e  dx %.1e  worst gradients: %s" %
          (name, rel(o["f_uncorr"], ref["f_uncorr"]), rel(o["f_corr"], ref["f_corr"]), rel(o["dx"], ref["dx"]),
           ", ".join("%s %.1e" % (k.split("block.")[-1], e) for e, k in worst[:4])), flush=True)
ho.MM_F12[0] = torch.matmul
ho.MM_BLK[0] = torch.matmul
