"""Data-parallel gradient all-reduce for the backbone: bucketed NCCL all-reduce over NVLink,
overlapped with backward.

Replaces what the reference gets from ``MMDistributedDataParallel`` (torch DDP) at
``mmdet/apis/train.py:91-99``: one process per GPU, gradients averaged across ranks every step,
``broadcast_buffers=False``.  Design (B200 / NVSwitch): gradients live as views into a few flat
fp32 buckets (~32 MB, sized for launch latency, not link count); parameters are assigned to buckets
in reverse execution order (stage 3 first) so a bucket completes early in backward; when the last
gradient of a bucket has been accumulated (``register_post_accumulate_grad_hook``) the bucket is
all-reduced (AVG) on a dedicated communication stream while backward keeps running on the compute
stream.  ``finish()`` makes the compute stream wait for the communication stream.
Works with backend "gloo" on CPU tensors for tests (synchronous there).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, broadcast_params: bool = True):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        if self.world > 1 and broadcast_params:
            for p in params:                      # DDP-constructor behaviour: rank 0's weights win
                dist.broadcast(p.data, src=0, group=process_group)
        cap = int(bucket_mb * 1024 * 1024 / 4)
        order = list(reversed(params))            # reverse execution order
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        cur, cur_n = [], 0
        groups = []
        for p in order:
            if cur and cur_n + p.numel() > cap:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        for bi, grp in enumerate(groups):
            n = sum(p.numel() for p in grp)
            flat = torch.zeros(n, dtype=torch.float32, device=grp[0].device)
            off = 0
            for p in grp:
                p.grad = flat[off:off + p.numel()].view_as(p)     # gradient-as-bucket-view
                off += p.numel()
                self._bucket_of[p] = bi
            self.buckets.append(flat)
        self._sizes = [len(gp) for gp in groups]
        self._pending = list(self._sizes)
        self._cuda = params[0].is_cuda if params else False
        self.comm_stream = torch.cuda.Stream() if self._cuda else None
        self._handles = []
        self._launched = 0
        for p in params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    # ------------------------------------------------------------------
    def _on_grad(self, p: torch.Tensor) -> None:
        bi = self._bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        self._launched += 1
        if self.world == 1:
            return
        flat = self.buckets[bi]
        if self._cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def finish(self) -> None:
        """Call after backward: all buckets reduced and visible to the compute stream."""
        for bi, left in enumerate(self._pending):
            if left != 0 and left != self._sizes[bi]:
                self._launch(bi)               # bucket with unused parameters this step
            elif left == self._sizes[bi] and self.world > 1:
                self._launch(bi)               # nothing arrived: still reduce (zeros) to keep ranks in lock-step
        if self._cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = list(self._sizes)

    def zero_grad(self) -> None:
        """Zero the flat buckets (keeps the grad views attached)."""
        for flat in self.buckets:
            flat.zero_()
