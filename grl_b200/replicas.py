"""Data-parallel glue for head replicas.

The reference trains under `nn.DataParallel` (mars_train.py:80): one process scatters the batch, replicates the module every
step and gathers the outputs; BatchNorm statistics stay per replica.  The B200-native arrangement is one process per GPU with
the head's parameter gradients summed over NCCL / NVLink while the next step already runs:

    sync = GradientAllReduce(head.head_param_names(), shapes, device)      # two flat fp32 buffers (112 MB each for the head)
    dx, grads = head.head_backward_raw(sd, x, B, T, ws, gu, gc, grads=sync.views())   # kernels write straight into the flat buffer
    sync.start()                       # all-reduce on a side stream, overlaps the following forward
    ...
    avg = sync.finish()                # {name: averaged gradient}, views of the reduced buffer

BatchNorm statistics are NOT synchronised (same semantics as the reference's per-replica BN).  Works with any
torch.distributed backend (NCCL on B200; gloo on CPU tensors in the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradientAllReduce(object):
    def __init__(self, names, shapes, device, group=None, average=True, dtype=torch.float32):
        self.names = list(names)
        self.shapes = [tuple(s) for s in shapes]
        self.device = torch.device(device)
        self.group = group
        self.average = average
        sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + (n + 63) // 64 * 64)      # 256-byte aligned slots
        self.flat = [torch.zeros(self.offsets[-1], dtype=dtype, device=self.device) for _ in range(2)]
        self.cur = 0
        self.work = [None, None]
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.done = [torch.cuda.Event() if self.cuda else None for _ in range(2)]

    def _views(self, b):
        return {k: self.flat[b][o:o + int(torch.Size(s).numel())].view(s) for k, s, o in zip(self.names, self.shapes, self.offsets)}

    def views(self):
        """Gradient tensors of the buffer the NEXT start() will reduce (write the step's gradients here).  Waits, in stream
        order, until the all-reduce that last used this buffer has finished."""
        b = self.cur
        if self.cuda and self.work[b] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[b])
        elif self.work[b] is not None:
            self.work[b].wait()
        self.work[b] = None
        return self._views(b)

    def start(self):
        """Launch the all-reduce of the current buffer (asynchronous on CUDA: a side stream ordered after the producing kernels)."""
        b = self.cur
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                if world > 1:
                    dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group)
                if self.average and world > 1:
                    self.flat[b].mul_(1.0 / world)
                self.done[b].record(self.stream)
            self.work[b] = True
        else:
            if world > 1:
                self.work[b] = dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self.work[b].wait()
                if self.average:
                    self.flat[b].mul_(1.0 / world)
            self.work[b] = None
        self.cur = 1 - b
        return b

    def finish(self, b=None):
        """Averaged gradients of the buffer given to the last start() (stream-ordered wait on CUDA)."""
        b = 1 - self.cur if b is None else b
        if self.cuda and self.work[b] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[b])
        return self._views(b)
