"""GPU diagnostic: exercises the tcgen05 GEMM variants and the eval kernels, printing error
statistics for every case (does not stop at the first failure).  Run under gpurun."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import _lib  # noqa: E402

lib = _lib.load_library()
h = _lib.get_handle()
dev = torch.device("cuda")
print("version", lib.grl_version().decode(), "sms", lib.grl_num_sms(h), torch.cuda.get_device_name())


def gemm(A, B, a_mn=0, b_mn=0, bn=0, alpha=1.0, bias=None, row_scale=None, relu=0, C_init=None, stats=False, planes=False,
         batch=1):
    """A: [batch][M][K] (or [batch][K][M] if a_mn); B likewise.  Returns dict."""
    if a_mn:
        K, M = A.shape[-2:]
    else:
        M, K = A.shape[-2:]
    N = B.shape[-1] if b_mn else B.shape[-2]
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.lda, d.ldb, d.ldc = A.shape[-1], B.shape[-1], N
    d.c_bstride = M * N
    d.alpha = alpha
    d.relu = relu
    d.bn = bn
    Cout = torch.zeros((batch, M, N), device=dev) if C_init is None else C_init.clone()
    d.accumulate = 0 if C_init is None else 1
    out = {}
    if bias is not None:
        d.col_bias = bias.data_ptr()
    if row_scale is not None:
        d.row_scale = row_scale.data_ptr()
    mt = (M + 127) // 128
    if stats:
        out["sum"] = torch.zeros((batch, 4 * mt, N), device=dev)
        out["sq"] = torch.zeros((batch, 4 * mt, N), device=dev)
        d.col_sum, d.col_sq = out["sum"].data_ptr(), out["sq"].data_ptr()
    if planes:
        out["hi"] = torch.zeros((batch, M, N), dtype=torch.bfloat16, device=dev)
        out["lo"] = torch.zeros((batch, M, N), dtype=torch.bfloat16, device=dev)
        d.planes_hi, d.planes_lo = out["hi"].data_ptr(), out["lo"].data_ptr()
    wsb = lib.grl_gemm_workspace_bytes(C.byref(d))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    code = lib.grl_gemm_bf16x3(h, C.byref(d), A.data_ptr(), B.data_ptr(), Cout.data_ptr(), ws.data_ptr(), wsb,
                               _lib.stream_ptr())
    _lib.check(h, code, "gemm")
    torch.cuda.synchronize()
    out["C"] = Cout
    return out


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


results = []


def case(name, fn):
    try:
        t0 = time.time()
        r = fn()
        results.append((name, r))
        print("%-46s %s  (%.2fs)" % (name, "  ".join("%s=%.3e" % kv for kv in r.items()), time.time() - t0), flush=True)
    except Exception as e:  # noqa
        results.append((name, {"EXC": 1.0}))
        print("%-46s EXCEPTION %s" % (name, e), flush=True)
        traceback.print_exc()
        if "CUDA" in str(e) or "cuda" in str(e):
            print("fatal CUDA error; aborting diag")
            sys.exit(3)


def plain(M, N, K, bn, a_mn=0, b_mn=0, batch=1, identity=False):
    def f():
        g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
        A = torch.randn((batch, M, K), generator=g).to(dev)
        if identity:
            B = torch.eye(N, K).repeat(batch, 1, 1).to(dev)
        else:
            B = torch.randn((batch, N, K), generator=g).to(dev)
        ref = torch.matmul(A.double(), B.double().transpose(1, 2))
        Ain = A.transpose(1, 2).contiguous() if a_mn else A
        Bin = B.transpose(1, 2).contiguous() if b_mn else B
        o = gemm(Ain, Bin, a_mn, b_mn, bn, batch=batch)
        r = {"rel": rel(o["C"], ref)}
        if r["rel"] > 1e-3:
            r["maxabs"] = float((o["C"].double() - ref).abs().max())
            bad = ((o["C"].double() - ref).abs() > 1e-2 * ref.abs().max()).float()
            r["bad_frac"] = float(bad.mean())
            # which rows / cols are bad (layout debugging)
            rows = bad[0].mean(1).nonzero().flatten()[:8].tolist()
            cols = bad[0].mean(0).nonzero().flatten()[:8].tolist()
            print("    first bad rows", rows, "cols", cols)
            print("    got ", o["C"][0, :2, :6].tolist())
            print("    want", ref[0, :2, :6].tolist())
        return r
    return f


case("K-major 128x128x64 bn128 identity", plain(128, 128, 64, 128, identity=True))
case("K-major 128x128x64 bn128", plain(128, 128, 64, 128))
case("K-major 128x256x64 bn256", plain(128, 256, 64, 256))
case("K-major 128x128x256 bn128", plain(128, 128, 256, 128))
case("K-major 256x512x512 bn128", plain(256, 512, 512, 128))
case("K-major 256x512x512 bn256", plain(256, 512, 512, 256))
case("K-major 4096x2048x2048 bn256", plain(4096, 2048, 2048, 256))
case("K-major 4096x512x2048 bn128", plain(4096, 512, 2048, 128))
case("K-major ragged 1980x9330x2048 auto", plain(1980, 9330, 2048, 0))
case("K-major ragged 200x72x136 bn128", plain(200, 72, 136, 128))
case("K-major batch2 512x512x256 bn128", plain(512, 512, 256, 128, batch=2))
case("MN-major 128x128x64 bn128 identity", plain(128, 128, 64, 128, 1, 1, identity=True))
case("MN-major 128x128x64 bn128", plain(128, 128, 64, 128, 1, 1))
case("MN-major 128x256x128 bn256", plain(128, 256, 128, 256, 1, 1))
case("MN-major 512x2048x4096 bn256", plain(512, 2048, 4096, 256, 1, 1))
case("MN-major 2048x512x4096 bn128", plain(2048, 512, 4096, 128, 1, 1))
case("MN-major batch2 256x256x256 bn128", plain(256, 256, 256, 128, 1, 1, batch=2))
case("mixed (A K-major, B MN-major) 128x128x64 bn128", plain(128, 128, 64, 128, 0, 1))
case("mixed 256x512x512 bn256", plain(256, 512, 512, 256, 0, 1))
case("mixed 4096x512x2048 bn128", plain(4096, 512, 2048, 128, 0, 1))
case("mixed batch2 512x2048x512 auto", plain(512, 2048, 512, 0, 0, 1, batch=2))


def epilogue_case():
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 384, 512, 192
    A = torch.randn((1, M, K), generator=g).to(dev)
    B = torch.randn((1, N, K), generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    rs = torch.rand(M, generator=g).to(dev)
    C0 = torch.randn((1, M, N), generator=g).to(dev)
    ref = torch.relu(0.5 * torch.matmul(A.double(), B.double().transpose(1, 2)) * rs.double()[None, :, None] + bias.double())
    o = gemm(A, B, alpha=0.5, bias=bias, row_scale=rs, relu=1, stats=True, planes=True, bn=128)
    r = {"rel": rel(o["C"], ref)}
    r["sum"] = rel(o["sum"].sum(1), ref.sum(1))
    r["sq"] = rel(o["sq"].sum(1), (ref * ref).sum(1))
    r["planes"] = rel(o["hi"].float() + o["lo"].float(), ref)
    o2 = gemm(A, B, alpha=0.5, bias=bias, row_scale=rs, relu=1, C_init=C0, bn=256)
    r["accum"] = rel(o2["C"], ref + C0.double())
    return r


case("epilogue bias/rowscale/relu/stats/planes/accum", epilogue_case)


def timing():
    M, N, K = 4096, 2048, 2048
    A = torch.randn((1, M, K), device=dev)
    B = torch.randn((1, N, K), device=dev)
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, 1
    d.lda, d.ldb, d.ldc = K, K, N
    d.alpha = 1.0
    r = {}
    for bn in (128, 256):
        d.bn = bn
        Cout = torch.empty((M, N), device=dev)
        wsb = lib.grl_gemm_workspace_bytes(C.byref(d))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        for _ in range(3):
            lib.grl_gemm_bf16x3(h, C.byref(d), A.data_ptr(), B.data_ptr(), Cout.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.grl_gemm_bf16x3(h, C.byref(d), A.data_ptr(), B.data_ptr(), Cout.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        r["ms_bn%d(incl split)" % bn] = ms
        r["algTF_bn%d" % bn] = 2.0 * M * N * K / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3):
        torch.matmul(A[0], B[0].t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(A[0], B[0].t())
    e1.record()
    torch.cuda.synchronize()
    r["ms_torch_fp32"] = e0.elapsed_time(e1) / 10
    return r


case("timing 4096x2048x2048", timing)

nbad = sum(1 for _, r in results if r.get("EXC") or r.get("rel", 0) > 1e-3)
print("DIAG DONE: %d cases, %d bad" % (len(results), nbad))
