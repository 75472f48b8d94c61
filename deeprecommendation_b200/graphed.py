"""CUDA-graph replay of an eval-mode forward with static shapes.

The scoring calls of this package are chains of 10-30 short kernels (and, multi-GPU, NCCL collectives); at batch 512
the Python/launch overhead of issuing them is comparable to their device time.  `GraphedForward` captures one call
into a CUDA graph (streams + graphs instead of a tracing compiler) and replays it with new input VALUES copied into
the captured input buffers.  Shapes, the model's parameters' storage and the graph object must stay the same; the
result tensor is overwritten by every replay."""
from __future__ import annotations

import torch


class GraphedForward:
    def __init__(self, fn, *example_inputs, warmup: int = 2):
        """fn(*tensors) -> tensor.  `example_inputs` are CUDA tensors whose shapes/dtypes fix the captured signature."""
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                      # lazy initialisations (cuFuncSetAttribute, NCCL channels, caches)
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        """inputs: CUDA or (pinned) host tensors of the captured shapes; host tensors are copied straight into the
        captured input buffers (one H2D copy, no staging)."""
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def replay(self):
        """re-run on whatever the captured input buffers (`static_in`) hold"""
        self.graph.replay()
        return self.static_out
