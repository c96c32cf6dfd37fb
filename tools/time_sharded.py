"""GPU, torchrun: per-stage device times of evaluator.sharded_retrieve (10k queries x 1M gallery rows over WORLD_SIZE ranks)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grl_b200 import evaluator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NQ, NG, D, K = 10000, 1000000, 2048, 100
lo, n = evaluator.shard_bounds(NG, world, rank)
gen = torch.Generator(device=dev).manual_seed(1000 + rank)
gf = torch.randn((n, D), generator=gen, device=dev)
gf /= gf.norm(dim=1, keepdim=True)
qf = torch.nn.functional.normalize(torch.randn((NQ, D), generator=torch.Generator().manual_seed(7))).to(dev)
st = evaluator.CudaSearchStages
kp = st.kprime(K)
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


for rep in range(3):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    mark("start")
    cd, ci, gmax2, dirty = st.coarse(qf, gf, kp, lo, 0); mark("coarse")
    all_d = [torch.empty_like(cd) for _ in range(world)]; all_i = [torch.empty_like(ci) for _ in range(world)]
    dist.all_gather(all_d, cd); dist.all_gather(all_i, ci); mark("all_gather x2")
    dist.all_reduce(gmax2, op=dist.ReduceOp.MAX); dist.all_reduce(dirty, op=dist.ReduceOp.MAX); mark("all_reduce max x2")
    cdm, cim = st.merge(torch.stack(all_d), torch.stack(all_i)); mark("stack + merge")
    ed = st.rescore(qf, gf, cim, lo, 0); mark("rescore")
    dist.all_reduce(ed); mark("all_reduce sum")
    top_d, top_i, flags = st.finalize(qf, cdm, cim, ed, gmax2, dirty, K, 0); mark("finalize")
    nz = torch.nonzero(flags).flatten(); mark("nonzero (sync)")
    torch.cuda.synchronize()
if rank == 0:
    for (a, ea), (b, eb) in zip(marks, marks[1:]):
        print("%-20s %.3f ms" % (b, ea.elapsed_time(eb)))
    print("total %.3f ms, flagged %d" % (marks[0][1].elapsed_time(marks[-1][1]), int(nz.numel())))
dist.destroy_process_group()
