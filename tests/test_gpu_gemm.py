"""GPU parity: the exported contraction primitive grl_gemm_bf16x3 (split-bf16 tcgen05 GEMM; include/grl_b200.h) against fp64
matmul -- operand majors used by fprop / dgrad / wgrad, both N tiles, batches, ragged shapes, every epilogue option.
Bar: 1e-5 + 6e-9 * K relative: two bf16 planes per operand give ~2^-17 per product, and the tensor core's fp32 accumulation
truncates, which adds ~K/16 * 2^-24 (8e-6 at K = 2048, 1.5e-5 at K = 4096)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gemm(A, B, a_mn=0, b_mn=0, bn=0, alpha=1.0, bias=None, row_scale=None, relu=0, C_init=None, stats=False, planes=False, batch=1):
    from grl_b200 import _lib
    lib = _lib.load_library()
    h = _lib.get_handle()
    dev = A.device
    K, M = (A.shape[-2:] if a_mn else A.shape[-2:][::-1])
    N = B.shape[-1] if b_mn else B.shape[-2]
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.lda, d.ldb, d.ldc = A.shape[-1], B.shape[-1], N
    d.a_bstride, d.b_bstride, d.c_bstride = A.shape[-2] * A.shape[-1], B.shape[-2] * B.shape[-1], M * N
    d.alpha, d.relu, d.bn = alpha, relu, bn
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
    _lib.check(h, lib.grl_gemm_bf16x3(h, C.byref(d), A.data_ptr(), B.data_ptr(), Cout.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()),
               "grl_gemm_bf16x3")
    torch.cuda.synchronize()
    out["C"] = Cout
    return out


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def operands(batch, M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((batch, M, K), generator=g).cuda(), torch.randn((batch, N, K), generator=g).cuda()


@pytest.mark.parametrize("M,N,K,bn,a_mn,b_mn,batch", [
    (128, 128, 64, 128, 0, 0, 1), (256, 512, 512, 256, 0, 0, 1), (4096, 512, 2048, 0, 0, 0, 2),      # fprop shapes, both N tiles
    (1980, 9330, 2048, 0, 0, 0, 1), (200, 264, 72, 0, 0, 0, 1),                                       # ragged M / N / K (K-major)
    (512, 2048, 4096, 0, 1, 1, 2), (512, 512, 1024, 128, 1, 1, 1),                                    # wgrad: MN-major both
    (4096, 2048, 512, 0, 0, 1, 2), (300, 256, 128, 256, 0, 1, 1)])                                    # dgrad: K-major A, MN-major B
def test_gemm_matches_fp64(M, N, K, bn, a_mn, b_mn, batch):
    A, B = operands(batch, M, N, K, M * 7 + N * 3 + K)
    ref = torch.matmul(A.double(), B.double().transpose(1, 2))
    Ain = A.transpose(1, 2).contiguous() if a_mn else A
    Bin = B.transpose(1, 2).contiguous() if b_mn else B
    o = gemm(Ain, Bin, a_mn, b_mn, bn, batch=batch)
    assert rel(o["C"], ref) < 1e-5 + 6e-9 * K, rel(o["C"], ref)


def test_gemm_epilogue_options():
    M, N, K = 300, 512, 256
    A, B = operands(1, M, N, K, 5)
    g = torch.Generator().manual_seed(9)
    bias = torch.randn(N, generator=g).cuda()
    rs = torch.rand(M, generator=g).cuda() + 0.5
    C0 = torch.randn((1, M, N), generator=g).cuda()
    base = torch.matmul(A.double(), B.double().transpose(1, 2))
    ref = torch.relu(-0.5 * base * rs.double()[None, :, None] + bias.double()[None, None, :])
    o = gemm(A, B, alpha=-0.5, bias=bias, row_scale=rs, relu=1, stats=True, planes=True)
    assert rel(o["C"], ref) < 1e-5
    # per-(tile, warp) partial column sums of the stored value and of its square
    assert rel(o["sum"].sum(1), ref.sum(1)) < 1e-5 and rel(o["sq"].sum(1), (ref * ref).sum(1)) < 1e-5
    # the bf16 hi/lo planes re-split the stored value: hi + lo carries ~16 mantissa bits
    assert rel(o["hi"].float().double() + o["lo"].float().double(), ref) < 2e-5
    assert torch.equal(o["hi"], o["C"].to(torch.bfloat16))
    # C += v
    o2 = gemm(A, B, C_init=C0)
    assert rel(o2["C"], C0.double() + base) < 1e-5


def test_gemm_rejects_bad_arguments():
    from grl_b200 import _lib
    A, B = operands(1, 64, 64, 36, 1)                      # K-major rows must be 16-byte multiples: ld % 8 != 0
    with pytest.raises(_lib.GrlError):
        gemm(A, B)
    A, B = operands(1, 128, 128, 96, 2)                    # MN-major operands need K % 64 == 0
    with pytest.raises(_lib.GrlError):
        gemm(A.transpose(1, 2).contiguous(), B.transpose(1, 2).contiguous(), 1, 1)
