"""CUDA-graph replay of a fixed-shape step through the public API.

The device time of a SubM layer-step at LiDAR sizes (~0.15 ms) is several times smaller than the
host time torch + Python need to issue its ~25 launches eagerly, so a step whose SHAPES repeat
(the same cloud evaluated many times, a static calibration batch, a benchmark) is best replayed as
one graph.  ``graph_capture`` wraps the boilerplate: warm-up on a side stream, capture, static
input buffers that later calls copy into.  Every kernel of this library is capturable except the
regular-conv rulebook, whose data-dependent output count is read back by the host
(``spx_conv_rulebook_stage1``; the reference syncs at the same point,
``spconv/csrc/sparse/indices.py:1454-1455``) -- SubM stacks, pooling on cached rulebooks and the
int8 path capture fine.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

import torch


class GraphedStep:
    """``step = GraphedStep(fn, example_inputs)`` then ``out = step(*inputs)``: inputs are copied into
    the captured static buffers (shapes / dtypes must match the examples), the graph is replayed and
    the STATIC output objects are returned (clone them to keep a result across replays)."""

    def __init__(self, fn: Callable[..., Any], example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        assert all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs), "CUDA tensor inputs only"
        self.fn = fn
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up configures kernels / allocator pools
            for _ in range(max(warmup, 1)):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.static_outputs = fn(*self.static_inputs)
        except Exception as e:
            torch.cuda.synchronize()
            raise RuntimeError(
                "graph_capture failed. A regular SparseConv / SparseMaxPool builds its rulebook with one host "
                "read-back (the output count) and cannot be captured; capture SubM-only stacks, or run the strided "
                f"layers eagerly around the captured part. Original error: {type(e).__name__}: {e}") from e

    def __call__(self, *inputs: torch.Tensor):
        assert len(inputs) == len(self.static_inputs), "same number of inputs as at capture time"
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                assert dst.shape == src.shape and dst.dtype == src.dtype, \
                    f"graph replay needs the captured shape {tuple(dst.shape)} / dtype, got {tuple(src.shape)}"
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs


def graph_capture(fn: Callable[..., Any], *example_inputs: torch.Tensor, warmup: int = 3) -> GraphedStep:
    """Capture ``fn(*example_inputs)`` (forward, or forward + backward) into a CUDA graph."""
    return GraphedStep(fn, list(example_inputs), warmup)
