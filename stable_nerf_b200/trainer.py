"""Training-step driver for the hot path: one call = near/far -> march -> field -> composite -> L1 -> backward
(-> gradient all-reduce when ray-sharded over several GPUs).

This is the call pattern of the reference's ``forward_iteration`` (train.py:61-70: ``nerf.render(rays, bg_color=1,
max_steps=...)`` followed by an L1 loss and ``backward``) packaged so that the whole step can be captured ONCE in a
CUDA graph and replayed: after the reference-style warm-up has produced ``mean_count`` (nerf/renderer.py:321-325) the
step has no host synchronisation, so the ~40 launches of a step cost one graph launch on the host.

Ray-sharded data parallelism: every rank owns a contiguous shard of the step's rays, the
occupancy bitfield is broadcast from rank 0, and the flat gradients (colour MLP, sigma MLP, hash table) are summed
with NCCL.  The reference itself does not synchronise NeRF gradients (it unwraps the model from DDP, train.py:188);
the oracle for the sharded step is the single-GPU step on the concatenated batch.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) shard of n rays for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, world_size=None, group=None, average=False):
    """Sum (optionally average) the gradients of `params` across ranks, smallest tensors first so that the MLP
    buckets are on the wire while the big hash-table bucket is still being scattered."""
    if not (dist.is_available() and dist.is_initialized()):
        return []
    world = dist.get_world_size(group) if world_size is None else world_size
    if world == 1:
        return []
    grads = sorted((p.grad for p in params if p.grad is not None and p.grad.numel() > 0), key=lambda g: g.numel())
    handles = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads]
    for h in handles:
        h.wait()
    if average:
        for g in grads:
            g.div_(world)
    return grads


def broadcast_occupancy(model, src=0, group=None):
    """Make every rank march the same grid: broadcast bitfield, grid and the python-side running estimates."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.broadcast(model.density_bitfield, src=src, group=group)
    dist.broadcast(model.density_grid, src=src, group=group)
    meta = torch.tensor([float(model.mean_density), float(model.mean_count), float(model.iter_density)],
                        dtype=torch.float64, device=model.density_grid.device)
    dist.broadcast(meta, src=src, group=group)
    model.mean_density, model.mean_count, model.iter_density = float(meta[0]), int(meta[1]), int(meta[2])


class TrainStep:
    """fwd + L1 + bwd (+ all-reduce) of one ray batch, eager or as a replayed CUDA graph.

    ``step(rays_o, rays_d, target)`` accepts device tensors [N,3], [N,3], [N,C]; ``step_from_host`` takes pinned host
    tensors and includes the H2D copies and the D2H read of the loss (the end-to-end form bench.py times).
    """

    def __init__(self, model, n_rays, max_steps=1024, bg_color=1, T_thresh=1e-4, use_graph=True, world_size=1,
                 loss_scale=1.0):
        self.model, self.n_rays, self.max_steps = model, int(n_rays), int(max_steps)
        self.bg_color, self.T_thresh, self.world_size = bg_color, T_thresh, world_size
        self.loss_scale = loss_scale
        dev = next(model.parameters()).device
        C = model.channel_dim
        self.rays_o = torch.zeros(self.n_rays, 3, device=dev)
        self.rays_d = torch.zeros(self.n_rays, 3, device=dev)
        self.target = torch.zeros(self.n_rays, C, device=dev)
        self.loss = torch.zeros((), device=dev)
        self.loss_host = torch.zeros((), pin_memory=True) if dev.type == "cuda" else torch.zeros(())
        self.use_graph = use_graph
        self.graph = None
        self.params = [p for p in model.parameters() if p.numel() > 0]
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        model.grads_in_place = True  # the field backward accumulates straight into these .grad tensors

    def _body(self):
        m = self.model
        for p in self.params:
            p.grad.zero_()
        out = m.render(self.rays_o[None], self.rays_d[None], bg_color=self.bg_color, max_steps=self.max_steps,
                       T_thresh=self.T_thresh)
        loss = (out['image'].view(-1, m.channel_dim) - self.target).abs().mean()  # utils/loss_utils.py:9-10 l1_loss
        (loss * self.loss_scale).backward()
        self.loss.copy_(loss.detach())

    def warmup(self, rays_o, rays_d, target, iters=3):
        """Reference-style first steps on the synchronising path, then ``mean_count`` from the measured sample
        counts (nerf/renderer.py:321-325) so that later steps never read the device."""
        self.rays_o.copy_(rays_o)
        self.rays_d.copy_(rays_d)
        self.target.copy_(target)
        m = self.model
        m.train()
        for _ in range(iters):
            self._body()
        total_step = min(16, m.local_step)
        m.mean_count = int(m.step_counter[:total_step, 0].sum().item() / total_step)
        m.local_step = 0
        for _ in range(2):  # steady-state path once eagerly: sizes every workspace before a capture
            self._body()
        torch.cuda.synchronize()
        if self.use_graph:
            self._capture()

    def _capture(self):
        m = self.model
        local_step = m.local_step
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        m.local_step = local_step  # the graph is recorded for one step_counter row; keep replaying that row
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.launch_count() - before
        m.local_step = local_step + 1

    def step(self, rays_o=None, rays_d=None, target=None):
        if rays_o is not None:
            self.rays_o.copy_(rays_o, non_blocking=True)
            self.rays_d.copy_(rays_d, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        if self.world_size > 1:
            allreduce_gradients(self.params, self.world_size)
        return self.loss

    def step_from_host(self, rays_o_pinned, rays_d_pinned, target_pinned):
        """End-to-end form: pinned host inputs -> device, one step, loss back to the host (synchronises)."""
        self.step(rays_o_pinned, rays_d_pinned, target_pinned)
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)
