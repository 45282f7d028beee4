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
import os

import torch
import torch.distributed as dist

from . import _lib
from .field import _precision_code
from .raymarching import _pad_up, march_train_workspace_bytes


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) shard of n rays for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, world_size=None, group=None, average=False, async_op=True):
    """Sum (optionally average) the gradients of `params` across ranks, smallest tensors first so that the MLP
    buckets are on the wire while the big hash-table bucket is still being scattered."""
    if not (dist.is_available() and dist.is_initialized()):
        return []
    world = dist.get_world_size(group) if world_size is None else world_size
    if world == 1:
        return []
    grads = sorted((p.grad for p in params if p.grad is not None and p.grad.numel() > 0), key=lambda g: g.numel())
    if async_op:
        handles = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in grads]
        for h in handles:
            h.wait()
    else:  # stream-ordered (what a CUDA-graph capture needs)
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
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
                 loss_scale=1.0, fused=True, perturb=False, dt_gamma=0, optimizer=None, overlap_allreduce=False,
                 scatter_groups=2, group=None, exchange="auto", exchange_timeout_ms=0, pipeline=False,
                 overlap_exchange="auto", overlap_split_level=8, overlap_side_ctas=16):
        self.model, self.n_rays, self.max_steps = model, int(n_rays), int(max_steps)
        self.bg_color, self.T_thresh, self.world_size = bg_color, T_thresh, world_size
        self.loss_scale = loss_scale
        self.fused, self.perturb, self.dt_gamma = bool(fused), bool(perturb), dt_gamma
        # pipeline=True: what a step leaves to do once its gradients exist -- the gradient exchange between the ranks and
        # the optimiser update -- is deferred to the START of the next step, on a second stream beside that step's ray
        # march (near/far + march need the rays and the occupancy grid, not the parameters), and joined before the
        # hash-grid gather.  A step is still march -> field -> composite -> loss -> backward -> exchange -> update, one
        # exchange and one update per step; they just run under the next march (~80 us of the step at 4096 rays).
        # ``finish()`` applies what the last step left pending.  Parameters / summed gradients read between steps are one
        # update behind until ``finish()``.
        self.pipeline = bool(pipeline)
        self._pipe_stream = None
        self._dry = True  # warm-up / dry-run bodies: never step the optimiser
        self._capturing = False
        self.first_epoch_fused = True  # the mean_count <= 0 steps take the fused body too (one 4-byte read per step)
        self.fuse_tail = True  # whole steps: composite forward + L1 + composite backward in one launch
        self.zero_in_backward = True  # the gradients' zero fills belong to the field backward call
        self._bufs = None
        self._mark = None
        self._graph_fwd = self._graph_bwd = None
        # optional optimiser (stable_nerf_b200.optim.FusedAdam/W or any torch optimiser), stepped after every step();
        # one that zeroes the gradients inside its step spares the step's own 49 MB memset
        self.optimizer = optimizer
        self._opt_zeroes = bool(getattr(optimizer, "zero_grad_in_step", False))
        # NCCL exchange only, opt-in: the table scatter-add runs per group of levels OUTSIDE the graph, and each group's
        # slice of the gradient starts its all-reduce while the next group is still being scattered
        # (trainer._scatter_and_reduce).  Off by default: at 2 ranks it measured slower than scatter-then-all-reduce.
        self.overlap_allreduce = bool(overlap_allreduce) and world_size > 1
        self.scatter_groups = int(scatter_groups)
        self.group = group  # process group of the gradient exchange (None = the default group)
        # In-graph exchange (NVLS / peer kernel) split in two against the table scatter-add: the FINE levels
        # [overlap_split_level, L) -- the bulk of the bytes, the tail of the gradient arena -- are scattered first and their
        # slice is exchanged on a high-priority second stream (its own flag channel) while the coarse levels are still being
        # scattered; the rest (MLP gradients + coarse levels) follows on the step's stream.  "auto": on for the NVLS kernel
        # (few CTAs: it runs beside the scatter-add), off for the peer kernel (2 B200: 0.718 whole, 0.737 split).
        # overlap_split_level: an int, or several cut levels (up to 3: one flag channel per exchanged group) -- measured on
        # 4 B200: one cut at 8: 0.735 ms/step, cuts (10, 6): 0.742, (11, 8, 5): 0.762, whole: 0.774.
        # overlap_side_ctas: CTAs of the exchanges that run beside a scatter-add -- 8 B200, cut at 8: 0.7446 / 0.7359 /
        # 0.7327 ms/step at 32 / 16 / 8 CTAs; 4 B200: 0.7351 / 0.7291-0.7320 / 0.7374 / 0.8050 at 32 / 16 / 8 / 4 (alone,
        # the whole-arena exchange is fastest at 32).  16: within 0.5 % of the best on both and away from the cliff below 8.
        self.overlap_exchange = overlap_exchange
        self.overlap_split_levels = ([int(v) for v in overlap_split_level] if isinstance(overlap_split_level, (list, tuple))
                                     else [int(overlap_split_level)])
        self.overlap_split_level = self.overlap_split_levels[0]
        self.overlap_side_ctas = int(overlap_side_ctas)
        self._ex_stream = None
        dev = next(model.parameters()).device
        C = model.channel_dim
        # one allocation [rays_o | rays_d | target] (and a pinned host mirror of it): a step's inputs arrive in ONE H2D copy
        n3 = self.n_rays * 3
        self._inputs = torch.zeros(2 * n3 + self.n_rays * C, device=dev)
        self.rays_o = self._inputs[:n3].view(self.n_rays, 3)
        self.rays_d = self._inputs[n3:2 * n3].view(self.n_rays, 3)
        self.target = self._inputs[2 * n3:].view(self.n_rays, C)
        self._staging = None
        self.loss = torch.zeros((), device=dev)
        self.loss_host = torch.zeros((), pin_memory=True) if dev.type == "cuda" else torch.zeros(())
        self.use_graph = use_graph
        self.graph = None
        self.params = [p for p in model.parameters() if p.numel() > 0]
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        model.grads_in_place = True  # the field backward accumulates straight into these .grad tensors
        # Gradient exchange between ranks, one kernel inside the step's graph (stable_nerf_b200.p2p / csrc/p2p_reduce.cu):
        # "nvls" = reduced inside the NVSwitch (multicast), "p2p" = peer loads / stores in rank order (bit-reproducible);
        # "nccl" = NCCL all-reduce after the replay; "auto" = nvls, else p2p, else nccl, whatever the node supports.
        self.exchange, self.exchange_kind, self.exchange_error = None, ("none" if world_size == 1 else "nccl"), None
        self.exchange_timeout_ms = int(exchange_timeout_ms)
        if exchange not in ("auto", "nvls", "p2p", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'nvls', 'p2p' or 'nccl', got {exchange!r}")
        if exchange in ("p2p", "nvls") and world_size > 1 and dev.type != "cuda":
            raise RuntimeError(f"exchange={exchange!r} needs CUDA devices (NVLink); use 'auto' or 'nccl'")
        if world_size > 1 and exchange in ("auto", "nvls", "p2p") and dev.type == "cuda":
            try:
                self._setup_p2p(dev, {"auto": "auto", "nvls": "nvls", "p2p": "peer"}[exchange])
            except RuntimeError as e:
                if exchange != "auto":
                    raise
                self.exchange_error = str(e)

    def _apply_pending(self):
        """gradient exchange + optimiser update of the gradients the previous step produced (current stream)"""
        if self.exchange is not None:
            self.exchange.all_reduce()
        if self.optimizer is not None and (self._capturing or not self._dry):
            self.optimizer.step()

    def _pipelined(self):
        return self.pipeline and not (self.world_size > 1 and self.exchange is None)  # (NCCL after the replay: not deferred)

    def finish(self):
        """pipeline=True: apply the exchange / update the last ``step()`` left pending.  No-op otherwise."""
        if self._pipelined() and self.fused and self.model.mean_count > 0:
            self._dry = False
            self._apply_pending()

    def _setup_p2p(self, dev, algo="auto"):
        """Move the parameters' .grad into one peer-mapped arena: [colour MLP | sigma MLP | table], 16-byte aligned, so
        that "everything but the fine levels of the table" is one contiguous range at the front."""
        from .p2p import P2PExchange
        m = self.model
        order = [p for p in (getattr(m, "color_net", None), getattr(m, "sigma_net", None)) if p is not None]
        order = [mod.params for mod in order if any(mod.params is q for q in self.params)]
        order += [p for p in self.params if not any(p is q for q in order)]
        offs, n = {}, 0
        for p in order:
            offs[id(p)] = n
            n += (p.numel() + 3) // 4 * 4
        ex = P2PExchange(n, dev, group=self.group, algo=algo, timeout_ms=self.exchange_timeout_ms)
        for p in order:
            o = offs[id(p)]
            p.grad = ex.tensor[o:o + p.numel()].view_as(p)
        self.exchange, self.exchange_kind, self._ex_off = ex, ("nvls" if ex.algo == "nvls" else "p2p"), offs
        if ex.nvls_error and ex.algo != "nvls":
            self.exchange_error = f"NVLS unavailable: {ex.nvls_error}"
        # The scatter-add stays inside the field backward and the exchange follows it as one launch over the whole arena.
        # Exchanging the fine levels' slice on a (high-priority) side stream while the coarse levels are scattered was
        # measured at 2 ranks and bought nothing (763 vs 768 us/step), like the NCCL variant of the same overlap
        # (838 us/step against 805 for scatter-then-all-reduce): the two kernels do not run side by side.
        self.overlap_allreduce = False

    def _check_arena(self):
        """The exchange sums the arena, not whatever .grad points at: refuse to run if a gradient was re-created
        elsewhere (e.g. ``zero_grad(set_to_none=True)`` followed by autograd allocating a fresh tensor)."""
        base = self.exchange.tensor.data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * self._ex_off[id(p)]:
                raise RuntimeError("a parameter's .grad no longer lives in the peer-mapped gradient arena; keep the "
                                   "tensors TrainStep installed (zero them in place, do not set them to None)")

    def _body(self):
        m = self.model
        if self.fused and m.training and m.bg_radius <= 0 and hasattr(m, "fdesc"):
            if m.mean_count > 0:
                return self._body_fused()
            if self.first_epoch_fused:
                # before update_extra_state has produced mean_count the reference sizes the step from the
                # measured sample total (raymarching.py:196-229: N*max_steps rows zero-filled, .item(), empty_cache).
                # Same step here as the same straight sequence of C-ABI calls, with that ONE 4-byte read after the
                # count pass -- no N*max_steps buffers, no autograd graph.
                return self._body_fused(first_epoch=True)
        for p in self.params:
            p.grad.zero_()
        out = m.render(self.rays_o[None], self.rays_d[None], bg_color=self.bg_color, max_steps=self.max_steps,
                       T_thresh=self.T_thresh, perturb=self.perturb, dt_gamma=self.dt_gamma)
        loss = (out['image'].view(-1, m.channel_dim) - self.target).abs().mean()  # utils/loss_utils.py:9-10 l1_loss
        (loss * self.loss_scale).backward()
        self.loss.copy_(loss.detach())
        if self.exchange is not None:
            self.exchange.all_reduce()

    def _fused_buffers(self, M):
        """Every array of the fused step, allocated once per (M): a captured graph replays on fixed addresses."""
        if self._bufs is not None and self._bufs["M"] == M:
            return self._bufs
        m, N, C = self.model, self.n_rays, self.model.channel_dim
        dev = self.rays_o.device
        lib = _lib.load()
        prec = _precision_code(m.precision)
        f32 = dict(dtype=torch.float32, device=dev)
        e = lambda *shape: torch.empty(*shape, **f32)
        b = dict(M=M, nears=e(N), fars=e(N), noises=torch.zeros(N, **f32), rays=torch.empty(N, 3, dtype=torch.int32, device=dev),
                 n_samples=torch.empty(1, dtype=torch.int32, device=dev), xyzs=e(M, 3), dirs=e(M, 3), deltas=e(M, 2),
                 sigmas=e(M), rgbs=e(M, C), ws=e(N), depth=e(N), image=e(N, C), g_img=e(N, C), g_ws=e(N), g_sig=e(M),
                 g_rgb=e(M, C), pred=e(N, C), depth_norm=e(N), tail_counter=torch.zeros(1, dtype=torch.int32, device=dev))
        b["march_ws_bytes"] = march_train_workspace_bytes(lib, N, self.max_steps)
        b["march_ws"] = torch.empty(b["march_ws_bytes"], dtype=torch.uint8, device=dev)
        b["field_ws_bytes"] = max(lib.snerf_field_workspace_bytes(m.fdesc, M, prec, 0),
                                  lib.snerf_field_workspace_bytes(m.fdesc, M, prec, 1))
        b["field_ws"] = torch.empty(max(b["field_ws_bytes"], 256), dtype=torch.uint8, device=dev)
        b["saved_bytes"] = lib.snerf_field_saved_bytes(m.fdesc, M, prec)
        b["saved"] = torch.empty(b["saved_bytes"], dtype=torch.uint8, device=dev) if b["saved_bytes"] else None
        if (self.overlap_allreduce or self._overlap_exchange_on(M)) and prec == _lib.PRECISION_BF16:
            b["d_enc"] = e(M, m.fdesc.grid.n_levels * m.fdesc.grid.n_features)
        bg = self.bg_color
        b["bg"] = bg.to(**f32).contiguous().view(-1) if torch.is_tensor(bg) else None
        b["bg_scalar"] = 1.0 if bg is None else (0.0 if torch.is_tensor(bg) else float(bg))
        self._bufs = b
        return b

    def _body_fused(self, phase="both", first_epoch=False):
        """The same step as a straight sequence of C-ABI calls: no autograd graph, no torch glue kernels.
        phase: "both" (a whole step), or "forward" / "backward" for callers that put their own differentiable stage
        between the rendered image and the NeRF's backward (``forward()`` / ``backward(grad_image)``).
        near/far -> march (count, scan, write) -> field forward -> composite forward -> L1 loss + its gradient
        (one launch: background blend, depth normalisation, loss, d loss/d image, d loss/d weights_sum) ->
        composite backward -> field backward (accumulates straight into the parameters' .grad).
        Call pattern of train.py:61-70 / nerf/renderer.py:75-114."""
        m, N, C = self.model, self.n_rays, self.model.channel_dim
        lib = _lib.load()
        P, S, chk = _lib.ptr, _lib.stream(), _lib.check
        prec = _precision_code(m.precision)
        if first_epoch:
            # rows are known after the count pass; buffers for the largest total seen so far (in 64 Ki-row steps)
            M = 0
            b = self._bufs if self._bufs is not None else self._fused_buffers(1 << 16)
        else:
            M = _pad_up(int(m.mean_count), 128)
            b = self._fused_buffers(M)
        sp, cp = m.sigma_net.params, m.color_net.params
        nm = m.sigma_net.n_mlp
        mark = self._mark or (lambda name: None)
        if phase == "backward":
            return self._fused_backward(b, M, mark)
        mark("start")
        pipelined = self._pipelined() and phase == "both"
        if pipelined:  # the previous step's exchange + update on a second stream, beside this step's march
            if self._pipe_stream is None:
                self._pipe_stream = torch.cuda.Stream(device=self.rays_o.device)
            cur = torch.cuda.current_stream()
            self._pipe_stream.wait_stream(cur)
            with torch.cuda.stream(self._pipe_stream):
                self._apply_pending()
        if not self._opt_zeroes and not pipelined:
            # The field's gradients (table 46.5 MiB, MLP weights 0.4 MB) are zero-filled by the field backward itself
            # (SNERF_BWD_ZERO_*), on the library's side stream under the colour kernel; anything else is zeroed here.
            # (Zeroing everything on a side stream under the march was measured: +5 us/step -- the fills slow the
            # latency-bound march down by more than their own 12 us.)
            for p in self.params:
                if not (self.zero_in_backward and (p is sp or p is cp)):
                    p.grad.zero_()
        counter = m.step_counter[m.local_step % 16]
        counter.zero_()
        m.local_step += 1
        if self.perturb:
            b["noises"].uniform_()
        mark("zero_grads")
        geom = (float(m.bound), float(self.dt_gamma), int(self.max_steps), N, int(m.cascade), int(m.grid_size))
        chk(lib.snerf_march_rays_train_count_aabb(P(self.rays_o), P(self.rays_d), P(m.density_bitfield), P(m.aabb_train),
                                                  float(m.min_near), *geom, P(b["nears"]), P(b["fars"]), P(counter),
                                                  P(b["noises"]), P(b["march_ws"]), b["march_ws_bytes"], S),
            "near/far + march count")
        if first_epoch:
            total = int(counter[0].item())  # the step's one read-back (raymarching.py:223)
            M = _pad_up(max(total, 1), 128)
            if b["M"] < M:
                b = self._fused_buffers(_pad_up(M, 1 << 16))
                # (the count pass wrote nears/fars and its workspace into the old buffers: run it again into the new ones)
                counter.zero_()
                chk(lib.snerf_march_rays_train_count_aabb(P(self.rays_o), P(self.rays_d), P(m.density_bitfield), P(m.aabb_train),
                                                          float(m.min_near), *geom, P(b["nears"]), P(b["fars"]), P(counter),
                                                          P(b["noises"]), P(b["march_ws"]), b["march_ws_bytes"], S),
                    "near/far + march count")
        chk(lib.snerf_march_rays_train_write(P(self.rays_o), P(self.rays_d), P(m.density_bitfield), *geom, M,
                                             P(b["nears"]), P(b["fars"]), P(b["xyzs"]), P(b["dirs"]), P(b["deltas"]),
                                             P(b["rays"]), P(b["noises"]), 1, P(b["n_samples"]), P(b["march_ws"]),
                                             b["march_ws_bytes"], S), "march write")
        mark("march")
        if pipelined:  # join: the parameters are updated, the gradients consumed
            torch.cuda.current_stream().wait_stream(self._pipe_stream)
            if not self._opt_zeroes:
                for p in self.params:
                    if not (self.zero_in_backward and (p is sp or p is cp)):
                        p.grad.zero_()
        spd = sp.detach()
        chk(lib.snerf_field_forward(m.fdesc, P(b["xyzs"]), P(b["dirs"]), M, P(spd[nm:]), P(spd[:nm]), P(cp.detach()), prec,
                                    P(b["sigmas"]), P(b["rgbs"]), P(b["saved"]), b["saved_bytes"], P(b["field_ws"]),
                                    b["field_ws_bytes"], S), "field forward")
        if m.density_scale != 1:
            b["sigmas"].mul_(m.density_scale)
        mark("field_fwd")
        if phase == "both" and self.fuse_tail:
            # composite forward + blend/L1 + composite backward as ONE launch: the L1 gradient of a ray depends on that ray
            # only, so the warp that composited it carries straight on into its backward (csrc/composite.cu)
            chk(lib.snerf_composite_l1_train(P(b["sigmas"]), P(b["rgbs"]), P(b["deltas"]), P(b["rays"]), M, N,
                                             float(self.T_thresh), C, P(self.target), P(b["bg"]), b["bg_scalar"],
                                             float(self.loss_scale) / (N * C), P(b["nears"]), P(b["fars"]), P(b["ws"]),
                                             P(b["depth"]), P(b["image"]), P(b["pred"]), P(b["depth_norm"]), P(self.loss),
                                             P(b["g_sig"]), P(b["g_rgb"]), P(b["n_samples"]), P(b["tail_counter"]), S),
                "composite + L1 + composite backward")
            mark("composite_l1_fused")
            self.outputs = {"image": b["pred"], "depth": b["depth_norm"], "weights_sum": b["ws"]}
            return self._fused_backward(b, M, mark, composite=False, exchange=not pipelined)
        chk(lib.snerf_composite_rays_train_forward(P(b["sigmas"]), P(b["rgbs"]), P(b["deltas"]), P(b["rays"]), M, N,
                                                   float(self.T_thresh), C, P(b["ws"]), P(b["depth"]), P(b["image"]), S),
            "composite forward")
        mark("composite_fwd")
        chk(lib.snerf_l1_loss_backward(P(b["image"]), P(b["ws"]), P(self.target), P(b["bg"]), b["bg_scalar"], N, C,
                                       float(self.loss_scale) / (N * C), P(self.loss), P(b["g_img"]), P(b["g_ws"]),
                                       P(b["pred"]), P(b["depth"]), P(b["nears"]), P(b["fars"]), P(b["depth_norm"]), S),
            "l1 loss")
        mark("loss")
        self.outputs = {"image": b["pred"], "depth": b["depth_norm"], "weights_sum": b["ws"]}
        if phase == "forward":
            return
        self._fused_backward(b, M, mark)

    def _overlap_exchange_on(self, M=None):
        if self.exchange is None or _precision_code(self.model.precision) != _lib.PRECISION_BF16:
            return False
        L = int(self.model.fdesc.grid.n_levels)
        if not any(0 < v < L for v in self.overlap_split_levels):
            return False
        if self.overlap_exchange == "auto":
            # measured on 8 B200 (NVLS): 4096 rays per rank 0.7765 -> 0.746 ms/step; 32768 rays per rank (2.9 M rows: the
            # scatter-add is 1 ms, the exchange 0.15) 4.43 -> 4.46-4.51 ms/step -- two launches of a long scatter-add cost
            # more than the exchange they hide, so "auto" splits short steps only
            if M is None:
                M = self._bufs["M"] if self._bufs is not None else _pad_up(max(int(self.model.mean_count), 0), 128)
            return self.exchange_kind == "nvls" and 0 < M <= (1 << 20)
        return bool(self.overlap_exchange)

    def _scatter_and_exchange(self, b, M):
        """Table scatter-add in groups of levels, fine levels first, with each finished group's slice of the arena exchanged
        beside the next group's scatter-add (second stream, its own flag channel); the last group's slice travels with the
        MLP gradients on the step's stream.  Everything is recorded into the step's graph."""
        m = self.model
        lib = _lib.load()
        P, S, chk = _lib.ptr, _lib.stream(), _lib.check
        g = m.fdesc.grid
        nm, F, L = m.sigma_net.n_mlp, int(g.n_features), int(g.n_levels)
        sp = m.sigma_net.params
        grad_table = sp.grad[nm:]
        ex = self.exchange
        cuts = sorted({v for v in self.overlap_split_levels if 0 < v < L}, reverse=True)[:_lib.SNERF_P2P_CHANNELS - 1]
        table0 = self._ex_off[id(sp)] + nm  # first float of the table in the arena
        if self._ex_stream is None:
            self._ex_stream = torch.cuda.Stream(device=self.rays_o.device, priority=-1)
        cur = torch.cuda.current_stream()
        hi_level, hi_float = L, ex.n_floats
        for k, lvl in enumerate(cuts):
            chk(lib.snerf_hashgrid_backward_levels(g, P(b["xyzs"]), float(m.bound), P(b["d_enc"]), M, P(grad_table), lvl,
                                                   hi_level, S), "scatter levels")
            lo_float = table0 + int(g.offset[lvl]) * F
            assert lo_float % 4 == 0
            self._ex_stream.wait_stream(cur)
            with torch.cuda.stream(self._ex_stream):
                ex.all_reduce(lo=lo_float, hi=hi_float, channel=1 + k, n_ctas=self.overlap_side_ctas)
            hi_level, hi_float = lvl, lo_float
        chk(lib.snerf_hashgrid_backward_levels(g, P(b["xyzs"]), float(m.bound), P(b["d_enc"]), M, P(grad_table), 0, hi_level, S),
            "scatter coarse levels")
        ex.all_reduce(lo=0, hi=hi_float, channel=0)
        cur.wait_stream(self._ex_stream)

    def _fused_backward(self, b, M, mark, composite=True, exchange=True):
        m, N, C = self.model, self.n_rays, self.model.channel_dim
        lib = _lib.load()
        P, S, chk = _lib.ptr, _lib.stream(), _lib.check
        prec = _precision_code(m.precision)
        sp, cp = m.sigma_net.params, m.color_net.params
        nm = m.sigma_net.n_mlp
        spd = sp.detach()
        if composite:
            chk(lib.snerf_composite_rays_train_backward_ex(P(b["g_ws"]), P(b["g_img"]), P(b["sigmas"]), P(b["rgbs"]),
                                                           P(b["deltas"]), P(b["rays"]), P(b["ws"]), P(b["image"]), M, N,
                                                           float(self.T_thresh), C, P(b["g_sig"]), P(b["g_rgb"]),
                                                           P(b["n_samples"]), S), "composite backward")
        if m.density_scale != 1:
            b["g_sig"].mul_(m.density_scale)
        if composite:
            mark("composite_bwd")
        flags = (_lib.BWD_ZERO_TABLE_GRAD | _lib.BWD_ZERO_W_GRADS) if (self.zero_in_backward and not self._opt_zeroes) else 0
        split = exchange and self.exchange is not None and b.get("d_enc") is not None and self._overlap_exchange_on(M)
        d_enc = b.get("d_enc") if (split or self.exchange is None) else None  # (stop before the scatter-add only for a split)
        chk(lib.snerf_field_backward_ex(m.fdesc, P(b["xyzs"]), P(b["dirs"]), M, P(spd[nm:]), P(spd[:nm]), P(cp.detach()),
                                        P(b["g_sig"]), P(b["g_rgb"]), prec, P(sp.grad[nm:]), P(sp.grad[:nm]), P(cp.grad),
                                        P(b["saved"]), b["saved_bytes"], P(b["field_ws"]), b["field_ws_bytes"],
                                        P(d_enc), flags, S), "field backward")
        if split:
            self._scatter_and_exchange(b, M)
            mark("field_bwd")
            return
        mark("field_bwd")
        if self.exchange is not None and exchange:  # all ranks' gradients summed in place, same stream: part of the captured step
            self.exchange.all_reduce()

    def profile_stages(self, iters=10):
        """Device time of each stage of the fused step (CUDA events on the launch stream, eager launches): returns
        ({stage: ms}, n_samples, M)."""
        acc = {}
        fuse_tail = self.fuse_tail
        # first the three kernels of the step's tail one by one (their own rooflines), then the one-launch form the step runs
        for fuse in (False, True):
            self.fuse_tail = fuse
            for it in range(iters + 2):
                evs = []

                def mark(name):
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    evs.append((name, e))
                self._mark = mark
                try:
                    self._body_fused()
                finally:
                    self._mark = None
                torch.cuda.synchronize()
                if it >= 2:
                    for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
                        if not fuse or n1 == "composite_l1_fused":
                            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / iters
        self.fuse_tail = fuse_tail
        n_samples = int(self._bufs["n_samples"].item())
        return acc, n_samples, self._bufs["M"]

    def profile_field_kernels(self, iters=10):
        """Device time (us) of each kernel of the bf16 field calls on the last step's samples, one kernel per launch
        through the stage mask of the debug build -- libsnerf_b200_dbg.so, the same sources compiled with
        -DSNERF_DEBUG_HOOKS -- (CUDA events on the launch stream).  Leaves the gradients meaningless."""
        m, b = self.model, self._bufs
        lib = _lib.load_debug()  # the stage mask is a hook of the debug build (same sources, same kernels)
        P, S, chk = _lib.ptr, _lib.stream(), _lib.check
        prec = _precision_code(m.precision)
        if prec != _lib.PRECISION_BF16 or b is None:
            return {}
        M, nm = b["M"], m.sigma_net.n_mlp
        sp, cp = m.sigma_net.params, m.color_net.params
        spd = sp.detach()

        def fwd():
            chk(lib.snerf_field_forward(m.fdesc, P(b["xyzs"]), P(b["dirs"]), M, P(spd[nm:]), P(spd[:nm]), P(cp.detach()),
                                        prec, P(b["sigmas"]), P(b["rgbs"]), P(b["saved"]), b["saved_bytes"],
                                        P(b["field_ws"]), b["field_ws_bytes"], S), "field forward")

        def bwd():
            chk(lib.snerf_field_backward(m.fdesc, P(b["xyzs"]), P(b["dirs"]), M, P(spd[nm:]), P(spd[:nm]), P(cp.detach()),
                                         P(b["g_sig"]), P(b["g_rgb"]), prec, P(sp.grad[nm:]), P(sp.grad[:nm]), P(cp.grad),
                                         P(b["saved"]), b["saved_bytes"], P(b["field_ws"]), b["field_ws_bytes"], S),
                "field backward")
        stages = [("pack_weights_fwd", 1, fwd), ("hashgrid_gather", 2, fwd), ("sigma_net_fwd", 4, fwd),
                  ("color_net_fwd", 8, fwd), ("pack_weights_bwd", 16, bwd), ("color_net_bwd", 32, bwd),
                  ("sigma_net_bwd", 64, bwd), ("hashgrid_scatter", 128, bwd)]
        out = {}
        try:
            for name, mask, fn in stages:
                lib.snerf_debug_set_field_stage_mask(mask)
                for _ in range(2):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                out[name] = e0.elapsed_time(e1) / iters * 1e3
        finally:
            lib.snerf_debug_set_field_stage_mask(0xffffffff)
        return out

    def warmup(self, rays_o, rays_d, target, iters=3, batches=None):
        """Reference-style first steps on the synchronising path, then ``mean_count`` from the measured sample
        counts (nerf/renderer.py:321-325) so that later steps never read the device.  ``batches``: optional list of
        (rays_o, rays_d, target) triples, one per warm-up step in turn -- ``mean_count`` is then the mean over different
        ray batches, as it is in training, instead of one batch's exact count."""
        self.rays_o.copy_(rays_o)
        self.rays_d.copy_(rays_d)
        self.target.copy_(target)
        m = self.model
        m.train()
        self._dry = True
        if self.optimizer is not None and hasattr(self.optimizer, "init_state"):
            self.optimizer.init_state()  # moments allocated before any capture
        for it in range(iters):
            if batches:
                o, d, t = batches[it % len(batches)]
                self.rays_o.copy_(o)
                self.rays_d.copy_(d)
                self.target.copy_(t)
            self._body()
        total_step = min(16, m.local_step)
        self.warmup_counts = [int(c) for c in m.step_counter[:total_step, 0].tolist()]  # sample totals of the warm-up steps
        m.mean_count = int(sum(self.warmup_counts) / total_step)
        m.local_step = 0
        for _ in range(2):  # steady-state path once eagerly: sizes every workspace before a capture
            self._body()
        torch.cuda.synchronize()
        if self.use_graph:
            self._capture()
        self._zero_grads_after_dry_runs()

    def _reset_pipeline(self):
        """The first pipelined step applies 'the previous step's' exchange and update before any real gradient exists:
        make that a no-op (zero gradients sum to zero; the optimiser skips one application and does not count it)."""
        if not self._pipelined():
            return
        for p in self.params:
            p.grad.zero_()
        if self.optimizer is not None:
            if not hasattr(self.optimizer, "skip_next") or not getattr(self.optimizer, "capturable", False):
                raise RuntimeError("pipeline=True needs an optimiser whose step can be captured: FusedAdam(..., capturable=True)")
            self.optimizer.skip_next()
        torch.cuda.synchronize()

    def _zero_grads_after_dry_runs(self):
        """With an optimiser that zeroes the gradients inside its own step (``zero_grad_in_step``) the fused body zeroes
        nothing, so the bodies run by warmup / capture / launch counting -- none followed by an optimiser step -- leave
        their gradients accumulated (and, ray-sharded, all-reduced again and again).  Start the first real step from zero."""
        if self._opt_zeroes:
            for p in self.params:
                p.grad.zero_()
            torch.cuda.synchronize()
        self._reset_pipeline()

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
        # The gradient all-reduce stays an eager call after the replay: capturing the NCCL collectives into the graph
        # measured no gain at N=2 (0.931 vs 0.941 ms/step) and made process teardown hang on this torch/NCCL pair.
        self.allreduce_in_graph = False
        self.graph = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            with torch.cuda.graph(self.graph):
                self._body()
        finally:
            self._capturing = False
        before = _lib.launch_count()
        m.local_step = local_step
        self._mark = None
        self._count_launches()
        self.launches_per_step = _lib.launch_count() - before
        m.local_step = local_step + 1

    def _count_launches(self):
        """One eager pass of the body, to count the library launches of a step (the graph replays the same ones)."""
        self._body()
        torch.cuda.synchronize()

    def step(self, rays_o=None, rays_d=None, target=None):
        if rays_o is not None:
            self.rays_o.copy_(rays_o, non_blocking=True)
            self.rays_d.copy_(rays_d, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        if self.exchange is not None:
            self._check_arena()
            self.exchange.raise_on_error()  # a rank gave up waiting in an earlier step: those gradients were never summed
        self._dry = False
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        if self.world_size > 1 and self.exchange is None:
            if self._bufs is not None and self._bufs.get("d_enc") is not None:
                self._scatter_and_reduce()
            else:
                allreduce_gradients(self.params, self.world_size, group=self.group)
        if self.optimizer is not None and not (self._pipelined() and self.fused and self.model.mean_count > 0):
            self.optimizer.step()
        return self.loss

    def sample_overflow(self):
        """(samples the last step's march produced, rows M the step's buffers hold).  The step is captured for one M =
        pad128(mean_count); like in the reference (raymarching.cu:417) rays whose samples do not fit are dropped
        silently, so a caller that changes the occupancy grid checks this (one 4-byte read, synchronises) or calls
        ``refresh()``."""
        if self._bufs is None:
            return 0, 0
        return int(self._bufs["n_samples"].item()), int(self._bufs["M"])

    def refresh(self):
        """Call after ``model.update_extra_state()``: the reference re-estimates ``mean_count`` there
        (nerf/renderer.py:321-325) and sizes the next steps' buffers from it; a captured step keeps the M it was recorded
        with, so when pad128(mean_count) has changed the buffers are re-allocated and the graph is captured again.
        Returns True when it re-captured."""
        m = self.model
        if not (self.fused and m.mean_count > 0):
            return False
        M = _pad_up(int(m.mean_count), 128)
        if self._bufs is not None and self._bufs["M"] == M:
            return False
        self.finish()  # (pipeline: the last step's exchange / update must not be lost in the dry runs below)
        self._dry = True
        self.graph = self._graph_fwd = self._graph_bwd = None
        local_step = m.local_step
        for _ in range(2):
            self._body()
        torch.cuda.synchronize()
        if self.use_graph:
            self._capture()
        m.local_step = local_step
        self._zero_grads_after_dry_runs()
        return True

    def _scatter_and_reduce(self):
        """Table scatter-add in groups of levels, each group's (contiguous) slice of the gradient all-reduced while the
        next group is scattered.  Fine levels go first: their slices are the big ones (524 288 entries per level) and
        their scatter is the quick one (no run merging); the coarse group's slice is extended over the sigma MLP's
        gradient, which precedes the table in the same tensor, so the whole exchange is 1 + n_groups collectives
        (measured at N=2: one 49 MB all-reduce 117 us, four quarters 193 us -- few, large calls).
        dist.all_reduce(async_op) runs on NCCL's own stream after everything queued so far on the current stream,
        i.e. right after its group's scatter launch."""
        m, b = self.model, self._bufs
        lib = _lib.load()
        P, S, chk = _lib.ptr, _lib.stream(), _lib.check
        g = m.fdesc.grid
        nm, F, L = m.sigma_net.n_mlp, g.n_features, g.n_levels
        grad = m.sigma_net.params.grad
        handles = [dist.all_reduce(m.color_net.params.grad, group=self.group, async_op=True)]
        n_groups = max(1, min(self.scatter_groups, L))
        bounds = [round(k * L / n_groups) for k in range(n_groups + 1)]
        for k in range(n_groups - 1, -1, -1):
            lb, le = bounds[k], bounds[k + 1]
            chk(lib.snerf_hashgrid_backward_levels(g, P(b["xyzs"]), float(m.bound), P(b["d_enc"]), b["M"], P(grad[nm:]),
                                                   lb, le, S), "scatter levels")
            lo = 0 if lb == 0 else nm + g.offset[lb] * F
            hi = nm + (g.offset[le] * F if le < L else g.n_entries * F)
            handles.append(dist.all_reduce(grad[lo:hi], group=self.group, async_op=True))
        for h in handles:
            h.wait()

    def forward(self, rays_o=None, rays_d=None, target=None):
        """First half of a step, for training loops that feed the rendered image into a further differentiable stage
        (train.py:61-99: the NeRF's latent image conditions the SD U-Net): march -> field -> composite -> blend, L1 loss
        and its gradient.  Returns ``{'image' [N,C], 'depth' [N], 'weights_sum' [N]}`` (views of persistent buffers)."""
        if not (self.fused and self.model.mean_count > 0):
            raise RuntimeError("forward()/backward() need the fused path: call warmup() first")
        if rays_o is not None:
            self.rays_o.copy_(rays_o, non_blocking=True)
            self.rays_d.copy_(rays_d, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        if self.use_graph:
            if self._graph_fwd is None:
                self._graph_fwd = self._capture_phase("forward")
            self._graph_fwd.replay()
        else:
            self._body_fused("forward")
        return self.outputs

    def backward(self, grad_image=None):
        """Second half: backward of the path.  grad_image [N,C] (optional) is an external d loss / d image -- e.g. what
        autograd returns for the SD loss w.r.t. ``forward()['image']`` -- added to the L1 loss's own gradient
        (image = composite + (1 - weights_sum) * bg, so it also feeds d loss / d weights_sum)."""
        b = self._bufs
        if self.exchange is not None:
            self._check_arena()
            self.exchange.raise_on_error()
        if grad_image is not None:
            g = grad_image.detach().to(torch.float32).reshape(self.n_rays, self.model.channel_dim)
            b["g_img"].add_(g)
            bg = b["bg"] if b["bg"] is not None else b["bg_scalar"]
            b["g_ws"].sub_((g * bg).sum(-1))
        if self.use_graph:
            if self._graph_bwd is None:
                self._graph_bwd = self._capture_phase("backward")
            self._graph_bwd.replay()
        else:
            self._body_fused("backward")
        if self.world_size > 1 and self.exchange is None:
            if b.get("d_enc") is not None:
                self._scatter_and_reduce()
            else:
                allreduce_gradients(self.params, self.world_size, group=self.group)
        if self.optimizer is not None:
            self.optimizer.step()
        return self.loss

    def _capture_phase(self, phase):
        m = self.model
        local_step = m.local_step
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._body_fused(phase)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        m.local_step = local_step
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body_fused(phase)
        if phase == "forward":
            m.local_step = local_step  # replays keep writing the step_counter row recorded here
        if phase == "backward":
            self._zero_grads_after_dry_runs()  # the dry run above accumulated a second backward of the same forward
        return g

    def pinned_inputs(self):
        """(rays_o, rays_d, target) views into ONE pinned host buffer laid out like the device inputs.  Filled by the
        data loader and passed to ``step_from_host``, they travel in a single 36+4C B/ray copy instead of three."""
        if self._staging is None:
            self._staging = torch.zeros(self._inputs.numel(), pin_memory=self._inputs.is_cuda)
        n3 = self.n_rays * 3
        st = self._staging
        return (st[:n3].view(self.n_rays, 3), st[n3:2 * n3].view(self.n_rays, 3), st[2 * n3:].view(self.n_rays, -1))

    def new_pinned_batch(self):
        """A pinned host buffer laid out like the device inputs, and its (rays_o, rays_d, target) views: fill the views,
        pass the buffer to ``step_from_packed`` -- a data loader keeps a few of these in rotation."""
        st = torch.zeros(self._inputs.numel(), pin_memory=self._inputs.is_cuda)
        n3 = self.n_rays * 3
        return st, (st[:n3].view(self.n_rays, 3), st[n3:2 * n3].view(self.n_rays, 3), st[2 * n3:].view(self.n_rays, -1))

    def step_from_packed(self, packed, read_loss=False, next_packed=None):
        """One step on a batch packed as [rays_o | rays_d | target] (device tensor, or pinned host tensor: ONE copy of
        36+4C B/ray either way).
        read_loss=True: also bring the loss back to the host and wait for it -- the end-to-end form (synchronises).
        read_loss="lagged": the data-loader form of the same.  Every step still copies its inputs from the host and reads
        its loss back, but nothing waits on the step just launched: the loss of THIS step is copied to a pinned slot
        asynchronously and the call returns the loss of the PREVIOUS step (None for the first), so the device never idles
        between steps; ``next_packed`` (pinned host tensor) starts the next step's input copy on a copy stream while this
        step computes.  ``drain()`` returns the last step's loss."""
        if read_loss == "lagged":
            return self._step_lagged(packed, next_packed)
        self._inputs.copy_(packed, non_blocking=True)
        self.step()
        if not read_loss:
            return self.loss
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if self.exchange is not None:
            self.exchange.raise_on_error()
        return float(self.loss_host)

    def _step_lagged(self, packed, next_packed):
        dev = self._inputs.device
        if getattr(self, "_lag", None) is None:
            self._lag = dict(copy_stream=torch.cuda.Stream(device=dev), staged=torch.empty_like(self._inputs), staged_src=None,
                             staged_ev=torch.cuda.Event(), slots=[torch.zeros((), pin_memory=True) for _ in range(2)],
                             evs=[torch.cuda.Event(), torch.cuda.Event()], k=0, pending=None, consumed=torch.cuda.Event())
        L = self._lag
        cur = torch.cuda.current_stream()
        if L["staged_src"] is packed:  # its H2D copy was started while the previous step ran
            cur.wait_event(L["staged_ev"])
            self._inputs.copy_(L["staged"], non_blocking=True)
        else:
            self._inputs.copy_(packed, non_blocking=True)
        L["consumed"].record(cur)
        self.step()
        slot = L["k"] % 2
        L["slots"][slot].copy_(self.loss, non_blocking=True)
        L["evs"][slot].record(cur)
        L["staged_src"] = None
        if next_packed is not None:
            with torch.cuda.stream(L["copy_stream"]):
                L["copy_stream"].wait_event(L["consumed"])  # the staging buffer's previous content has been consumed
                L["staged"].copy_(next_packed, non_blocking=True)
                L["staged_ev"].record(L["copy_stream"])
            L["staged_src"] = next_packed
        prev, L["pending"] = L["pending"], slot
        L["k"] += 1
        if prev is None:
            return None
        L["evs"][prev].synchronize()  # the previous step's loss has landed; the step just launched keeps running
        if self.exchange is not None:
            self.exchange.raise_on_error()
        return float(L["slots"][prev])

    def drain(self):
        """Loss of the last ``read_loss="lagged"`` step (waits for it)."""
        L = getattr(self, "_lag", None)
        if L is None or L["pending"] is None:
            return None
        L["evs"][L["pending"]].synchronize()
        v, L["pending"] = float(L["slots"][L["pending"]]), None
        return v

    def step_from_host(self, rays_o_pinned, rays_d_pinned, target_pinned):
        """End-to-end form: pinned host inputs -> device, one step, loss back to the host (synchronises)."""
        if self._staging is not None and rays_o_pinned.data_ptr() == self._staging.data_ptr() and \
                rays_d_pinned.data_ptr() == self._staging.data_ptr() + 4 * self.n_rays * 3 and \
                target_pinned.data_ptr() == self._staging.data_ptr() + 8 * self.n_rays * 3:
            self._inputs.copy_(self._staging, non_blocking=True)
            self.step()
        else:
            self.step(rays_o_pinned, rays_d_pinned, target_pinned)
        self.loss_host.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if self.exchange is not None:
            self.exchange.raise_on_error()  # the step has completed: its own exchange is covered too
        return float(self.loss_host)
