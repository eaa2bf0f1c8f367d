"""Whole-step CUDA-graph capture for the backbone (forward + backward [+ gradient all-reduce]).

One Swin-T step enqueues ~1600 kernels; eager Python/ctypes dispatch costs ~30 ms of host time per step,
which would cap a 30-40 ms GPU step.  ``GraphedStep`` captures the step once (after a few eager warm-up
iterations on a side stream, as CUDA graphs require) and replays it with static input/output buffers, so the
host cost per step is one graph launch.  All kernels of libswin_b200.so are capture-safe: they only enqueue on
the current stream, never allocate or synchronise, and their TMA descriptors are by-value kernel parameters.

Constraints (the usual CUDA-graph ones): fixed shapes, and tensors the step reads must live in static buffers
(copy new data into ``static_inputs`` before ``replay``).  bf16 weight shadow copies are refreshed OUTSIDE the
graph (``refresh_weights``) because the graph holds pointers to them.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import functional as F_


class GraphedStep:
    def __init__(self, step_fn: Callable[[], torch.Tensor], warmup: int = 3):
        """``step_fn`` runs one full step on static tensors it closes over and returns a tensor (e.g. the loss)."""
        self.step_fn = step_fn
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.output = step_fn()

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.output


def refresh_weights(module: torch.nn.Module) -> None:
    """Re-cast, IN PLACE, the cached bf16 shadow copy of every parameter whose version changed (call after
    optimizer.step() when the step is replayed from a graph)."""
    from . import ops
    for p in module.parameters():
        hit = F_._W16.get(id(p))
        if hit is not None and hit[0]() is p and hit[1] != p._version:
            hit[2].copy_(p.detach())
            F_._W16[id(p)] = (hit[0], p._version, hit[2])
