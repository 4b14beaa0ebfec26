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
        self.stream = torch.cuda.Stream()               # replays run here, so several instances can overlap
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(warmup):                     # fills the weight caches outside the graph
                model(self.static_x, noise=self.static_noise)
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.LAUNCHES
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=self.stream):
            self.static_z, self.static_nll, _ = model(self.static_x, noise=self.static_noise)
        self.flowk_launches = _lib.LAUNCHES - before      # flowk kernels per replay

    def run(self, x=None):
        """Replay on the current stream."""
        if x is not None:
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_z, self.static_nll

    def run_async(self, x=None, out=None):
        """Replay on this instance's own stream (ordered after the caller's current stream); optionally copy the
        per-image bits/dim into `out` (pinned host tensor).  Several instances replaying concurrently fill the SMs
        that one batch's deep-level kernels leave idle."""
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            if x is not None:
                self.static_x.copy_(x, non_blocking=True)
            self.graph.replay()
            if out is not None:
                out.copy_(self.static_nll, non_blocking=True)
        return self.static_z, self.static_nll


class DensityPipeline:
    """`depth` graph instances on separate streams, used round-robin: consecutive batches overlap on the GPU."""

    def __init__(self, model, example_x, depth=2):
        self.lanes = [GraphedDensity(model, example_x) for _ in range(depth)]
        self.i = 0
        self.flowk_launches = self.lanes[0].flowk_launches

    def submit(self, x, out=None):
        lane = self.lanes[self.i % len(self.lanes)]
        self.i += 1
        return lane.run_async(x, out)

    def drain(self):
        for lane in self.lanes:
            torch.cuda.current_stream().wait_stream(lane.stream)
