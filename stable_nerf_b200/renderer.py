"""``NeRFRenderer`` with the reference's interface (nerf/renderer.py:8-334), driving libsnerf_b200 kernels.

Same constructor arguments, buffers (``aabb_train``, ``aabb_infer``, ``density_grid``, ``density_bitfield``,
``step_counter``), python-side state (``mean_density``, ``iter_density``, ``mean_count``, ``local_step``) and
methods (``render``, ``run_cuda``, ``update_extra_state``, ``mark_untrained_grid``, ``reset_extra_state``), so a
reference ``state_dict`` of these buffers loads unchanged.

Differences that do not change results:
* the occupancy sweep enumerates cells directly in Morton order (``morton3D_invert(arange)``) and writes the grid
  contiguously, instead of building xyz meshgrids and scattering through ``morton3D`` indices;
* the inference loop compacts the alive list on the device (``compact_rays``) and reads one int per iteration instead
  of running ``rays_alive[rays_alive >= 0]`` (a nonzero + gather + sync);
* ``min_n_step`` (default 1 = the reference's schedule ``n_step = max(min(N // n_alive, 8), 1)``, nerf/renderer.py:146)
  can raise the number of samples marched per alive ray and loop iteration: with 4 a frame needs 52 instead of 122
  iterations (28 vs 40 ms at 800x800) because march_rays re-reads every ray's state and diverges on its slowest lane in
  each of them.  Like in the reference, ``rays_t`` is re-accumulated from the deltas between iterations, so a different
  schedule moves sample positions by an ulp: images agree to ~1e-6, not bit for bit -- hence opt-in;
* the module does not force ``.cuda()`` in the constructor (nerf/renderer.py:30); buffers move with ``.to(device)``.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, raymarching


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, channel_dim=3, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self.channel_dim = channel_dim
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius
        self.min_n_step = 1  # inference: lower bound of the samples marched per alive ray and iteration (1 = reference)
        self.native_loop = True  # inference loop issued by the library (snerf_render_rays) when the field is the fused one
        self._host_count = None

        aabb = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32)
        self.register_buffer('aabb_train', aabb)
        self.register_buffer('aabb_infer', aabb.clone())
        cells = self.cascade * self.grid_size ** 3
        self.register_buffer('density_grid', torch.zeros(self.cascade, self.grid_size ** 3, dtype=torch.float32))
        self.register_buffer('density_bitfield', torch.zeros(cells // 8, dtype=torch.uint8))
        self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))  # 16 steps of history
        self.mean_density = 0
        self.iter_density = 0
        self.mean_count = 0
        self.local_step = 0
        self.error_map = None

    # ---- subclass interface (nerf/renderer.py:50-58)
    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def reset_extra_state(self):
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    # ---- rendering
    def _finish(self, image, depth, weights_sum, nears, fars, bg_color, prefix):
        # background blend and depth normalisation (nerf/renderer.py:111-114, :164-167)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return image.view(*prefix, self.channel_dim), depth.view(*prefix)

    @torch.no_grad()
    def _render_native(self, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh):
        """The inference loop below as ONE library call (``snerf_render_rays``, csrc/render_loop.cu): the same kernels and
        the reference's schedule iteration for iteration, with the launches and the per-iteration count read issued from
        native code instead of through ~15 Python wrapper calls per iteration.  Needs the fused field (a subclass that
        overrides ``forward`` sets ``native_loop = False`` and gets the generic loop)."""
        from .field import _precision_code
        lib = _lib.load()
        P = _lib.ptr
        dev = rays_o.device
        N = rays_o.shape[0]
        prec = _precision_code(self.precision)
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, self.channel_dim, dtype=torch.float32, device=dev)
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else None
        mns = max(int(self.min_n_step), 1)
        nbytes = lib.snerf_render_rays_workspace_bytes(self.fdesc, N, mns, prec)
        ws = _lib.workspace.get("render_loop", nbytes, dev)
        if self._host_count is None:
            self._host_count = torch.zeros(1, dtype=torch.int32).pin_memory()
        st = _lib.RenderStats()
        sp = self.sigma_net.params.detach()
        nm = self.sigma_net.n_mlp
        _lib.check(lib.snerf_render_rays(
            self.fdesc, P(rays_o), P(rays_d), N, P(self.density_bitfield), int(self.cascade), int(self.grid_size),
            float(self.bound), float(dt_gamma), int(max_steps), P(nears), P(fars), P(noises), P(sp[nm:]), P(sp[:nm]),
            P(self.color_net.params.detach()), prec, float(self.density_scale), float(T_thresh), mns, P(weights_sum),
            P(depth), P(image), ctypes.c_void_p(self._host_count.data_ptr()), ctypes.byref(st), P(ws), nbytes, _lib.stream()),
            "render_rays")
        self.last_render_stats = {"iterations": int(st.iterations), "rows": int(st.rows), "samples": int(st.samples)}
        return weights_sum, depth, image

    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, **kwargs):
        """rays_o, rays_d [B,N,3] -> {'image' [B,N,C], 'depth' [B,N], 'weights_sum' [B*N] (training only)}."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device

        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)

        if self.bg_radius > 0:
            # the reference calls an undefined self.background here (nerf/renderer.py:88)
            raise NotImplementedError("bg_radius > 0 is a dead path in the reference (no background model exists)")
        if bg_color is None:
            bg_color = 1

        results = {}
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
                self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs = self(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            weights_sum, depth, image = raymarching.composite_rays_train(
                sigmas.to(torch.float32), rgbs.to(torch.float32), deltas, rays, T_thresh, self.channel_dim)
            results['weights_sum'] = weights_sum
        else:
            if self.native_loop and hasattr(self, "fdesc") and N > 0:
                weights_sum, depth, image = self._render_native(rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps,
                                                                T_thresh)
                image, depth = self._finish(image, depth, weights_sum, nears, fars, bg_color, prefix)
                results['depth'] = depth
                results['image'] = image
                return results
            weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
            depth = torch.zeros(N, dtype=torch.float32, device=device)
            image = torch.zeros(N, self.channel_dim, dtype=torch.float32, device=device)
            rays_alive = torch.arange(N, dtype=torch.int32, device=device)
            spare = torch.empty_like(rays_alive)
            count = torch.empty(1, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            n_alive, step = N, 0
            stats = {"iterations": 0, "rows": 0, "samples": 0}  # network-evaluated rows incl. / excl. alignment padding
            self.last_render_stats = stats
            while step < max_steps and n_alive > 0:
                n_step = max(min(N // n_alive, 8), int(self.min_n_step), 1)
                xyzs, dirs, deltas = raymarching.march_rays(
                    n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, self.density_bitfield, self.cascade,
                    self.grid_size, nears, fars, 128, perturb if step == 0 else False, dt_gamma, max_steps)
                stats["iterations"] += 1
                stats["rows"] += xyzs.shape[0]
                stats["samples"] += n_alive * n_step
                sigmas, rgbs = self(xyzs, dirs)
                if self.density_scale != 1:
                    sigmas = self.density_scale * sigmas
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                                           image, T_thresh, self.channel_dim)
                spare, count = raymarching.compact_rays(rays_alive, n_alive, out=spare, count=count)
                rays_alive, spare = spare, rays_alive
                n_alive = int(count.item())
                step += n_step

        image, depth = self._finish(image, depth, weights_sum, nears, fars, bg_color, prefix)
        results['depth'] = depth
        results['image'] = image
        return results

    def render(self, rays_o, rays_d, **kwargs):
        return self.run_cuda(rays_o, rays_d, **kwargs)

    # ---- occupancy grid maintenance
    def _cell_coords(self, indices):
        """int32 [n,3] grid coordinates of Morton indices."""
        return raymarching.morton3D_invert(indices.to(torch.int32))

    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """Cells no camera sees get density -1 (nerf/renderer.py:174-234).  poses [B,4,4] cam2world; intrinsic
        (fx, fy, cx, cy).  One launch per cascade (``snerf_mark_untrained_grid``, csrc/grid_update.cu): every cell runs
        the reference's frustum test over all cameras and stops at the first that sees it; ``S`` (the reference's
        chunking against OOM) has no effect on the result and is ignored."""
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        device = self.density_grid.device
        _lib.require_cuda(self.density_grid)
        poses = poses.to(device=device, dtype=torch.float32).contiguous().view(-1, 4, 4)
        fx, fy, cx, cy = intrinsic
        n_marked = torch.zeros(1, dtype=torch.int32, device=device)
        _lib.check(_lib.load().snerf_mark_untrained_grid(
            _lib.ptr(poses), poses.shape[0], float(cx / fx), float(cy / fy), float(self.bound), int(self.cascade),
            int(self.grid_size), _lib.ptr(self.density_grid), _lib.ptr(n_marked), _lib.stream()), "mark_untrained_grid")
        print(f'[mark untrained grid] {int(n_marked.item())} from {self.grid_size ** 3 * self.cascade}')

    def _query_sigma(self, cas_xyzs):
        """density of the sample points, before ``density_scale`` (nerf/renderer.py:268 / :302)"""
        return self.density(cas_xyzs)['sigma'].reshape(-1).detach().to(torch.float32)

    # Source of the update's random numbers.  The reference draws them with torch.rand_like / torch.randint on the device
    # (nerf/renderer.py:266, :280, :284, :300); a test replaces these two methods to feed the draws of a reference run.
    def _jitter_noise(self, cells, first, n, device):
        """uniforms [n,3] for the cells' jitter, or None = counter-based uniforms generated inside the kernel from a seed
        drawn from torch's generator (no 25 MB noise tensor per cascade)."""
        return None

    def _randint(self, high, shape, device):
        return torch.randint(0, high, shape, device=device)

    def _cell_points(self, cells, first, n, cas, device):
        """jittered sample positions [n,3] of Morton cells of cascade cas (nerf/renderer.py:259-266)"""
        noise = self._jitter_noise(cells, first, n, device)
        seed = 0 if noise is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        xyzs = torch.empty(n, 3, dtype=torch.float32, device=device)
        if noise is not None:
            noise = noise.to(device=device, dtype=torch.float32).contiguous()
        if cells is not None:
            cells = cells.to(torch.int32).contiguous()
        _lib.check(_lib.load().snerf_grid_cell_points(_lib.ptr(cells), int(first), int(n), int(cas), float(self.bound),
                                                      int(self.grid_size), _lib.ptr(noise), seed, _lib.ptr(xyzs),
                                                      _lib.stream()), "grid_cell_points")
        return xyzs

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """EMA update of the occupancy grid + bitfield + running sample-count estimate (nerf/renderer.py:236-327).

        Full sweep (the first 16 calls): per cascade one launch turns the Morton cells into jittered sample points, the
        density query runs on them (hash-grid gather + sigma net on the tensor cores), and ONE call
        (``snerf_grid_ema_update``: two launches) does max(grid*decay, sigma), the mean of the clamped grid, the threshold
        min(mean, density_thresh) and the bitfield without a host round trip; the only read-back is ``mean_density`` at
        the end (a python attribute in the reference too).  ``S`` chunks the sweep (S**3 cells per density query)."""
        device = self.density_grid.device
        _lib.require_cuda(self.density_grid)
        lib = _lib.load()
        H3 = self.grid_size ** 3
        tmp_grid = torch.empty_like(self.density_grid)

        if self.iter_density < 16:  # full sweep: every cell of every cascade, Morton order = the grid's own order
            chunk = min(max(int(S), 1) ** 3, H3)
            for start in range(0, H3, chunk):
                n = min(chunk, H3 - start)
                for cas in range(self.cascade):
                    tmp_grid[cas, start:start + n] = self._query_sigma(self._cell_points(None, start, n, cas, device))
        else:  # partial update: H^3/4 uniform cells + H^3/4 occupied cells per cascade (nerf/renderer.py:277-303)
            tmp_grid.fill_(-1)
            n = H3 // 4
            for cas in range(self.cascade):
                coords = self._randint(self.grid_size, (n, 3), device)
                indices = raymarching.morton3D(coords).long()
                occ = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                if occ.shape[0] > 0:  # (the reference raises on an empty grid: randint(0, 0))
                    occ = occ[self._randint(occ.shape[0], [n], device).long()]
                    indices = torch.cat([indices, occ], dim=0)
                sig = self._query_sigma(self._cell_points(indices, 0, indices.shape[0], cas, device))
                # `tmp_grid[cas, indices] = sigmas` with duplicate indices: on the reference's device the winner of a
                # duplicate is arbitrary; here the LAST occurrence wins (what a sequential assignment gives), always
                order = torch.arange(indices.shape[0], device=device)
                last = torch.full((H3,), -1, dtype=torch.long, device=device).scatter_reduce_(0, indices, order, "amax")
                hit = last >= 0
                tmp_grid[cas, hit] = sig[last[hit]]

        # EMA, mean, threshold, bitfield (nerf/renderer.py:310-319)
        nbytes = lib.snerf_grid_ema_workspace_bytes(tmp_grid.numel())
        ws = getattr(self, "_ema_ws", None)
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = self._ema_ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        out = torch.empty(2, dtype=torch.float32, device=device)
        _lib.check(lib.snerf_grid_ema_update(_lib.ptr(self.density_grid), _lib.ptr(tmp_grid), tmp_grid.numel(),
                                             float(self.density_scale), float(decay), float(self.density_thresh),
                                             _lib.ptr(out), _lib.ptr(self.density_bitfield), _lib.ptr(ws), nbytes,
                                             _lib.stream()), "grid_ema_update")
        self.mean_density = out[0].item()
        self.iter_density += 1

        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0
