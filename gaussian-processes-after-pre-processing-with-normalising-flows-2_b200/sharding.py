"""Batch-sharded training / evaluation: one process per GPU, full parameter replica per rank, NCCL over
NVLink for the only two exchanges the path has: the gradient all-reduce and the bits/dim
sum.  Replaces the reference's single-process nn.DataParallel (marscf_main.py:326,379).

Works with any backend torch.distributed offers (tests run it on CPU with gloo, world_size 2)."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Process-group set-up from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def shard_batch(x, rank, world):
    """Rank r takes samples [ceil(r*B/G) ... ) with torch.tensor_split bounds: every sample goes to exactly one rank,
    the first B % G ranks take one extra (the reference's DataParallel scatters dim 0 into chunks the same way)."""
    n = x.shape[0]
    per, rem = divmod(n, world)
    lo = rank * per + min(rank, rem)
    return x[lo:lo + per + (1 if rank < rem else 0)]


def broadcast_module(module, src=0):
    """Make every replica identical to rank `src` (after ActNorm's data-dependent init: the reference's
    DataParallel keeps replica 0's statistics, marscf_main.py:326 / SURVEY.md section 7)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
    from . import _lib
    _lib.bump_generation()            # `.data` writes skip the version counter: invalidate every derived-weight cache


class GradientBuckets:
    """Flat fp32 gradient buckets, all-reduced asynchronously as soon as every gradient of a bucket has been
    accumulated (overlaps the exchange with the rest of backward).  Gradients are views into the flat buffers,
    so there is no packing copy."""

    def __init__(self, params, bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = []            # (flat tensor, [params])
        cur, cur_bytes = [], 0
        for p in reversed(self.params):          # backward produces gradients roughly in reverse order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._seal(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._seal(cur)
        self._pending = []
        self._ready = [0] * len(self.buckets)
        self.overlap = True          # False: no all-reduce from the hooks, call reduce_all() after the backward pass
        # NCCL averages inside the collective; gloo (CPU tests) sums and the mean is taken by a division afterwards
        self._avg = self.world > 1 and dist.get_backend() == "nccl"
        self._bucket_of = {}
        for bi, (_, ps) in enumerate(self.buckets):
            for p in ps:
                self._bucket_of[p] = bi
                p.register_post_accumulate_grad_hook(self._hook)

    def _seal(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=ps[0].dtype, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append((flat, list(ps)))

    def _hook(self, p):
        bi = self._bucket_of[p]
        self._ready[bi] += 1
        if self._ready[bi] == len(self.buckets[bi][1]):
            self._launch(bi)

    def _all_reduce(self, flat):
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, async_op=True)

    def _launch(self, bi):
        if self.world > 1 and self.overlap:
            self._pending.append(self._all_reduce(self.buckets[bi][0]))

    def reduce_all(self):
        """All-reduce every bucket now (used after a CUDA-graph replay of forward + backward) and average."""
        if self.world > 1:
            handles = [self._all_reduce(flat) for flat, _ in self.buckets]
            for h in handles:
                h.wait()
            if not self._avg:
                for flat, _ in self.buckets:
                    flat.div_(self.world)

    def zero(self):
        for flat, _ in self.buckets:
            flat.zero_()
        self._ready = [0] * len(self.buckets)

    def finish(self):
        """Wait for the outstanding all-reduces and turn sums into means."""
        for bi, n in enumerate(self._ready):       # parameters that got no gradient this step
            if n != len(self.buckets[bi][1]) and self.world > 1:
                self._pending.append(self._all_reduce(self.buckets[bi][0]))
        for h in self._pending:
            h.wait()
        self._pending = []
        if self.world > 1 and not self._avg:
            for flat, _ in self.buckets:
                flat.div_(self.world)

    def nbytes(self):
        return sum(f.numel() * 4 for f, _ in self.buckets)


class ShardedTrainer:
    """The reference's training step (marscf_main.py:302-303,331-347): Adamax(lr=1e-4), LambdaLR warm-up on the
    number of samples seen, loss = mean bits/dim - with the batch sharded over ranks.

    `use_graph` (CUDA only): the step is launch-bound in eager mode (thousands of small autograd kernels), so after
    `graph_after` eager steps the forward + backward and the optimizer update are captured into two CUDA graphs and
    replayed; the gradient all-reduce runs between them on the flat buckets."""

    def __init__(self, model, lr=1e-4, warm_up=10000, global_batch=None, bucket_bytes=32 << 20, use_graph=None,
                 graph_after=2):
        self.model = model
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = GradientBuckets(model.parameters(), bucket_bytes)
        on_cuda = next(model.parameters()).is_cuda
        self.use_graph = on_cuda if use_graph is None else (use_graph and on_cuda)
        self.base_lr, self.warm_up = lr, warm_up
        self.graph_allreduce = False
        self.fused = False
        if on_cuda:
            # one-launch Adamax over all parameter tensors (flowk.optim); the learning rate reaches the kernel through
            # a device scalar refreshed on the host, so the same update is captured when use_graph is on
            from .optim import FusedAdamax
            self.opt = FusedAdamax(model.parameters(), lr=lr)
            self.sched = None
            self.fused = True
        elif self.use_graph:
            self.lr_t = torch.tensor(lr, device=next(model.parameters()).device)
            self.opt = torch.optim.Adamax(model.parameters(), lr=self.lr_t, capturable=True, foreach=True)
            self.sched = None
        else:
            self.opt = torch.optim.Adamax(model.parameters(), lr=lr)
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda s: min(1., s / warm_up))
        self.global_step = 0
        self._lr_samples = 0           # the sample count the schedule sees (one step behind, see _set_lr)
        self.global_batch = global_batch
        self.graph_after = graph_after
        self._calls = 0
        self._graphs = None
        if self.sched is None:
            self._set_lr()             # LambdaLR's constructor applies lambda(0) = 0 to the first step, too

    def _set_lr(self):
        """LambdaLR(min(1, samples / warm_up)) stepped like marscf_main.py:345-347: `scheduler.step(global_step)` runs
        BEFORE `global_step += batch_size`, so step k uses the sample count after step k-2 (the first two steps run
        with lr = 0)."""
        seen = self._lr_samples
        if self.sched is not None:
            self.sched.last_epoch = seen - 1
            self.sched.step()
        elif self.fused:
            for group in self.opt.param_groups:
                group["lr"] = self.base_lr * min(1., seen / self.warm_up)
        else:
            self.lr_t.fill_(self.base_lr * min(1., seen / self.warm_up))

    def _eager_step(self, x_local):
        self.buckets.zero()
        _, nll, _ = self.model(x_local)
        loss = nll.mean()
        loss.backward()
        self.buckets.finish()
        self.opt.step()
        return loss.detach()

    def _capture(self, x_local):
        # NCCL collectives are capturable: with FLOWK_GRAPH_ALLREDUCE=1 (default) the bucket all-reduces fired by the
        # gradient hooks become nodes of the forward+backward graph, on NCCL's stream, overlapped with the rest of the
        # backward pass; =0 keeps them outside (one reduce_all() between the two graphs).
        self.graph_allreduce = self.world > 1 and os.environ.get("FLOWK_GRAPH_ALLREDUCE", "1") != "0"
        self.buckets.overlap = self.graph_allreduce
        static_x = x_local.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        torch.cuda.synchronize()
        fb, up = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(fb, stream=side):
            self.buckets.zero()
            _, nll, _ = self.model(static_x)
            loss = nll.mean()
            loss.backward()
            if self.graph_allreduce:
                self.buckets.finish()
        self.buckets.overlap = False
        if self.fused:
            self.opt.prepare_step()              # host side (step count, lr scalar); the graph holds only the kernel
            with torch.cuda.graph(up, stream=side):
                self.opt.apply()
        else:
            with torch.cuda.graph(up, stream=side):
                self.opt.step()
        self._graphs = (fb, up, static_x, loss.detach())

    def step(self, x_local):
        self._calls += 1
        if not self.use_graph or self._calls <= self.graph_after:
            loss = self._eager_step(x_local)
        else:
            if self._graphs is None:
                self._capture(x_local)
            fb, up, static_x, loss = self._graphs
            if x_local.shape != static_x.shape:           # ragged last batch: the graph is shape-specialised
                self.buckets.overlap = True               # hooks fire the all-reduces again (off while replaying graphs)
                loss = self._eager_step(x_local)
                self.buckets.overlap = False
                from . import _lib
                _lib.bump_generation()
                self._advance(x_local)
                return loss
            static_x.copy_(x_local, non_blocking=True)
            fb.replay()
            if not self.graph_allreduce:
                self.buckets.reduce_all()
            if self.fused and not getattr(self, "_first_replay_done", False):
                self._first_replay_done = True   # prepare_step() already ran for this step inside _capture
            elif self.fused:
                self.opt.prepare_step()
            up.replay()
            from . import _lib
            _lib.bump_generation()    # a graph replay updates the parameters without touching their version counters
        self._advance(x_local)
        return loss

    def _advance(self, x_local):
        self._lr_samples = self.global_step
        self.global_step += self.global_batch or x_local.shape[0] * self.world
        self._set_lr()


def mean_bits_per_dim(nll_local):
    """Global mean of per-sample bits/dim: one 2-element all-reduce (sum, count)."""
    acc = torch.stack([nll_local.sum(), torch.tensor(float(nll_local.numel()), device=nll_local.device)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc)
    return acc[0] / acc[1]
