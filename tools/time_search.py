"""Per-stage device times of grl_sharded_topk on ONE GPU for the shard sizes a 1/2/4/8-way split of the 10k x 1M search gives
every rank (the collectives are absent; everything else a rank does is there).  python tools/time_search.py [ng ...]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import _lib, evaluator as ev  # noqa: E402

NQ, D, K = 10000, 2048, 100
ONCE = "--once" in sys.argv            # one warm-up + one profiled search per size (for ncu launch lists)
sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [1000000, 500000, 250000, 125000]
dev = torch.device("cuda", 0)
lib = _lib.load_library()
h = _lib.get_handle(dev)
q = torch.nn.functional.normalize(torch.randn((NQ, D), generator=torch.Generator().manual_seed(7))).to(dev)
out = (torch.empty((NQ, K), device=dev), torch.empty((NQ, K), dtype=torch.int64, device=dev))
stats = torch.zeros(8, dtype=torch.int32, device=dev)
for ng in sizes:
    g = torch.randn((ng, D), generator=torch.Generator(device=dev).manual_seed(1000), device=dev)
    g /= g.norm(dim=1, keepdim=True)
    pg = ev.PreparedGallery(g)
    for _ in range(1 if ONCE else 2):
        ev.sharded_topk(q, pg, K, 0, out=out, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(0 if ONCE else 5):
        ev.sharded_topk(q, pg, K, 0, out=out, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5 if not ONCE else 0.0
    lib.grl_search_profile(h, 1)
    lib.grl_profile_enable(h, 1)
    ev.sharded_topk(q, pg, K, 0, out=out, stats=stats)
    torch.cuda.synchronize()
    st = ev.search_stage_ms(dev)
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_longlong()
    lib.grl_profile_read(h, C.byref(g_ms), C.byref(g_fl), C.byref(g_n))
    lib.grl_profile_enable(h, 0)
    lib.grl_search_profile(h, 0)
    print(json.dumps({"ng": ng, "ms_per_search": round(ms, 3), "stages_ms": {k: round(v, 3) for k, v in st.items()},
                      "coarse_gemm_ms": round(g_ms.value, 3), "coarse_gemm_launches": g_n.value,
                      "coarse_gemm_tflops": round(g_fl.value / g_ms.value / 1e9, 1) if g_ms.value else None,
                      "stats": [int(v) for v in stats.cpu()[:4]]}), flush=True)
    del g, pg
    torch.cuda.empty_cache()
