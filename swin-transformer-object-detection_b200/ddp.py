"""Data-parallel gradient all-reduce for the backbone: bucketed NCCL all-reduce over NVLink,
overlapped with backward.

Replaces what the reference gets from ``MMDistributedDataParallel`` (torch DDP) at
``mmdet/apis/train.py:91-99``: one process per GPU, gradients averaged across ranks every step,
``broadcast_buffers=False``.  Design (B200 / NVSwitch): a few flat fp32 buckets (~32 MB, sized for launch
latency, not link count); parameters are assigned to buckets in reverse execution order (stage 3 first) so a
bucket completes early in backward.  ``zero_grad()`` drops the gradients (``p.grad = None``) so autograd keeps the
tensor each backward kernel produced without a copy; when the last gradient of a bucket has arrived
(``register_post_accumulate_grad_hook``) ONE ``swin_grad_gather`` launch packs the bucket (torch DDP issues one copy
or add per parameter: 189 launches per step here), the parameters' ``.grad`` are re-pointed at their bucket views,
and the bucket is all-reduced (AVG) on a dedicated communication stream while backward keeps running on the compute
stream.  ``finish()`` makes the compute stream wait for the communication stream.  With a single process there is
nothing to exchange: gradients stay where the kernels wrote them (no bucket traffic at all).

Bucket order: the first backward records the order in which the gradients actually arrive (the output norm{i} of a
stage arrives with that stage, not with the other norms at the end of ``parameters()``); ``finish()`` of that first
step rebuilds the buckets in arrival order (as torch DDP does after its first iteration), with a deliberately SMALL last
bucket (``tail_kb``): the only all-reduce that cannot overlap backward is the one holding the last-arriving gradients
(PatchEmbed), so that one is sized for latency, a few hundred KB, instead of tens of MB.
Works with backend "gloo" on CPU tensors for tests (synchronous there, plain torch copies).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, broadcast_params: bool = True,
                 tail_kb: float = 256.0):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        if self.world > 1 and broadcast_params:
            with torch.no_grad():
                for p in params:                  # DDP-constructor behaviour: rank 0's weights win
                    dist.broadcast(p, src=0, group=process_group)
            # a c10d collective writes the parameter without bumping its version counter, so bf16 operand copies cached by an
            # earlier forward (functional._W16, keyed by parameter version) would survive on the ranks whose weights were just
            # replaced: drop them (tests/ddp_nccl_worker.py starts rank 1 from other weights and runs a forward first)
            try:
                from . import functional as _F
                for p in params:
                    _F._W16.pop(id(p), None)
            except ImportError:                   # gloo / CPU unit tests drive this class with plain torch modules
                pass
        self._cap = int(bucket_mb * 1024 * 1024 / 4)
        self._tail = int(tail_kb * 1024 / 4)
        self._cuda = params[0].is_cuda if params else False
        self.comm_stream = torch.cuda.Stream() if self._cuda else None
        self._arrival: List[torch.Tensor] = []
        self._ordered = False                     # buckets follow the measured arrival order
        self._build(list(reversed(params)))       # first guess: reverse registration order
        for p in params:
            p.grad = None
            p.register_post_accumulate_grad_hook(self._on_grad)

    def _build(self, order: List[torch.Tensor]) -> None:
        """Assign ``order`` (first-arriving first) to flat fp32 buckets: a small tail bucket for the last arrivals, the rest
        in ``bucket_mb`` chunks."""
        tail, tail_n = [], 0
        while len(order) > 1 and tail_n + order[-1].numel() <= self._tail:
            tail.insert(0, order.pop())
            tail_n += tail[0].numel()
        groups, cur, cur_n = [], [], 0
        for p in order:
            if cur and cur_n + p.numel() > self._cap:
                groups.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            groups.append(cur)
        if tail:
            groups.append(tail)
        self.groups = groups
        self.buckets: List[torch.Tensor] = []
        self._offsets: List[List[int]] = []
        self._views = {}
        self._bucket_of = {}
        for bi, grp in enumerate(groups):
            offs, off = [], 0
            for p in grp:
                offs.append(off)
                off += (p.numel() + 3) // 4 * 4               # 16-byte aligned slots (vectorised gather)
            flat = torch.zeros(off, dtype=torch.float32, device=grp[0].device)
            for p, o in zip(grp, offs):
                self._views[p] = flat[o:o + p.numel()].view_as(p)
                self._bucket_of[p] = bi
            self.buckets.append(flat)
            self._offsets.append(offs)
        self._sizes = [len(gp) for gp in groups]
        self._pending = list(self._sizes)
        self._done = [False] * len(groups)
        self._launched = 0

    # ------------------------------------------------------------------
    def _on_grad(self, p: torch.Tensor) -> None:
        if not self._ordered:
            self._arrival.append(p)
        bi = self._bucket_of[p]
        if self._done[bi]:
            # the bucket of this step is already being reduced on the communication stream: a second backward would
            # accumulate into memory NCCL is reading
            raise RuntimeError("BucketedGradAllReduce: a second backward reached an already-launched bucket; call finish() "
                               "(and zero_grad()) between backward passes, or accumulate with world_size == 1")
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _pack(self, bi: int) -> None:
        """Gradients of bucket ``bi`` -> the flat bucket (one kernel launch on CUDA); ``.grad`` becomes the bucket view."""
        flat, grp, offs = self.buckets[bi], self.groups[bi], self._offsets[bi]
        src, dst = [], []
        for p, o in zip(grp, offs):
            g = p.grad
            view = self._views[p]
            if g is None:
                view.zero_()                               # a parameter unused this step contributes zeros (its slot only:
                continue                                   # gradients accumulated in place in other slots are kept)
            if g.data_ptr() == view.data_ptr():
                continue                                   # accumulated in place into the bucket already
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            src.append(g.detach())
            dst.append(o)
        if src:
            if self._cuda:
                from . import ops
                ops.grad_gather(src, dst, flat)
            else:
                for g, o in zip(src, dst):
                    flat[o:o + g.numel()].copy_(g.reshape(-1))
        for p in grp:
            p.grad = self._views[p]

    def _launch(self, bi: int) -> None:
        self._launched += 1
        self._done[bi] = True
        if self.world == 1:
            return
        self._pack(bi)
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
        for bi in range(len(self.groups)):
            if not self._done[bi] and (self.world > 1 or self._pending[bi] != self._sizes[bi]):
                self._launch(bi)               # bucket with parameters unused this step (or none at all): still reduced, so
                                               # the ranks stay in lock-step
        if self._cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if not self._ordered:
            seen = set()
            order = [p for p in self._arrival if not (id(p) in seen or seen.add(id(p)))]
            self._arrival = []
            capturing = self._cuda and torch.cuda.is_current_stream_capturing()
            if len(order) == len(self.params) and not capturing:
                # every rank ran the same graph, so every rank measured the same order.  The gradients of THIS step stay
                # valid: .grad keeps the old bucket views (and their storage) alive until zero_grad().
                self._build(order)
                self._ordered = True
        self._pending = list(self._sizes)
        self._done = [False] * len(self.groups)

    def zero_grad(self) -> None:
        """Drop the gradients (``set_to_none``): the next backward stores each kernel's output tensor as ``.grad`` directly."""
        for p in self.params:
            p.grad = None
