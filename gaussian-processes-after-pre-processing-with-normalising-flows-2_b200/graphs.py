"""CUDA-graph capture of the forward (density) pass: at the named batch sizes every kernel of the
stack runs for microseconds, so the step is launch-bound unless the whole layer sequence is replayed
as one graph (streams and graphs instead of a tracing compiler)."""
import torch

from . import _lib


class GraphedDensity:
    """Captures `model(x)` -> (z, bits/dim) for a fixed input shape.  `run(x)` copies x into the static
    input (device or pinned-host source), replays, and returns the static outputs."""

    def __init__(self, model, example_x, inject_noise=False, warmup=3):
        assert example_x.is_cuda
        self.model = model
        self.static_x = example_x.clone()
        self.static_noise = torch.rand_like(example_x) if inject_noise else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                     # fills the weight caches outside the graph
                model(self.static_x, noise=self.static_noise)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.LAUNCHES
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_z, self.static_nll, _ = model(self.static_x, noise=self.static_noise)
        self.flowk_launches = _lib.LAUNCHES - before      # flowk kernels per replay

    def run(self, x=None):
        if x is not None:
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_z, self.static_nll
