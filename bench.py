#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: train rays/s (fwd+bwd) on the cfg2 workload of BASELINE.json
("Blender-shaped synthetic scene 800x800, hash-grid + cuda_ray occupancy marching, 4096-ray train step on 1 B200").

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (CPU arm: the oracle port timed on the host cores)

A step = near/far -> occupancy march -> hash-grid encode + sigma/colour MLP -> alpha compositing -> L1 loss ->
backward of all of it (composite bwd, MLP dgrad/wgrad, hash-grid scatter-add) [-> NCCL all-reduce of the gradients
when N>1].  Weak scaling: every rank marches its own 4096-ray shard.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
N_BATCHES = 8  # distinct seeded ray batches rotated through the timed loops (training draws new rays every step)
MAX_STEPS = 1024
CHANNELS = 3
TABLE_SCALE = 1e4  # hash table U(-1e-4,1e-4) * 1e4: non-degenerate densities


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed regions run.  NVML in a thread (2 ms period: a 30-step
    timed region is only ~20 ms, too short for nvidia-smi's 200 ms loop to land a sample in it); nvidia-smi -lms as
    the fall-back when NVML cannot be opened.  `mark()` brackets the timed regions: the summary is taken over the
    samples inside them (or, if none landed inside, over every sample of the sampler's lifetime, and says so)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0, uuid=None):
        self.index, self.uuid = index, uuid
        self.rows = []          # (t, sm_mhz, set of reasons)
        self.marks = []         # (t0, t1) of the timed regions
        self.sm_max, self.source = None, None
        self.proc = self.thread = None
        self._stop = threading.Event()

    # ---- NVML
    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            u = self.uuid if str(self.uuid).startswith("GPU-") else "GPU-" + str(self.uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(u)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByUUID(u.encode())
        else:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip()]
            phys = int(ids[self.index]) if ids and all(v.strip().isdigit() for v in ids) and self.index < len(ids) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # raises here, not in the thread, if unsupported
        return pynvml, h

    def _nvml_loop(self, nv, h):
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(reasons_fn(h))
                except Exception:
                    mask = 0
                self.rows.append((time.perf_counter(), mhz, {n for n, b in self.BITS if mask & b}))
            except Exception:
                pass
            self._stop.wait(0.002)

    # ---- nvidia-smi fall-back
    def _smi_loop(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 8 and r[0].replace(".", "").isdigit():
                if self.sm_max is None and r[1].replace(".", "").isdigit():
                    self.sm_max = float(r[1])
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                self.rows.append((time.perf_counter(), float(r[0]),
                                  {n for k, n in enumerate(names) if r[4 + k].lower().startswith("active")}))

    def __enter__(self):
        try:
            nv, h = self._nvml_open()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            try:
                sel = [f"--id={self.uuid if str(self.uuid).startswith('GPU-') else 'GPU-' + str(self.uuid)}"] if self.uuid else [f"--id={self.index}"]
                self.proc = subprocess.Popen(["nvidia-smi", *sel, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
                self.thread = threading.Thread(target=self._smi_loop, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = self.thread = None
        # do not start the timed region before the sampler delivers (nvidia-smi needs ~1 s to come up)
        t_end = time.perf_counter() + 5.0
        while self.thread is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)
        return self

    def mark(self):
        """Context manager bracketing one timed region."""
        sampler = self

        class _M:
            def __enter__(self):
                self.t0 = time.perf_counter()

            def __exit__(self, *a):
                sampler.marks.append((self.t0, time.perf_counter()))
        return _M()

    def __exit__(self, *a):
        if self.source == "nvidia-smi" and self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        return False

    def summary(self):
        try:
            return self._summary()
        except Exception as e:  # the sampler must never take the bench line down with it
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "source": self.source,
                    "error": repr(e)[:200]}

    def _summary(self):
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "source": self.source}
        inside = [r for r in rows if any(t0 <= r[0] <= t1 for t0, t1 in self.marks)]
        window = "timed regions" if inside else "sampler lifetime (no sample landed inside a timed region)"
        use = inside or rows
        sm = [r[1] for r in use]
        reasons = sorted(set().union(*[r[2] for r in use]))
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "window": window, "source": self.source}


def workload(n_rays, seed):
    from stable_nerf_b200 import synthetic as syn
    grid = syn.occupancy_grid(lego_like=True, seed=0)
    bitfield = syn.pack_bitfield(grid)
    rays_o, rays_d = syn.train_batch(n_rays, seed=seed)
    target = np.random.default_rng(seed + 1).random((n_rays, CHANNELS), dtype=np.float32)  # torch.rand-like targets
    return bitfield, rays_o, rays_d, target


# ------------------------------------------------------------------------------------------------ CPU arms

def oracle_step(orc, fd, bitfield, rays_o, rays_d, target, sp, cp, n_mlp):
    """The same step as the GPU arm, on the CPU oracle.  Returns (n_samples, loss)."""
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    xyzs, dirs, deltas, rays, counter = orc.march_rays_train(rays_o, rays_d, 1.0, bitfield, 1, 128, nears, fars,
                                                             max_steps=MAX_STEPS)
    sig, rgb = orc.field_forward(fd, xyzs, dirs, sp[n_mlp:], sp[:n_mlp], cp)
    ws, depth, image = orc.composite_rays_train_forward(sig, rgb, deltas, rays, 1e-4)
    pred = image + (1 - ws)[:, None]
    loss = float(np.abs(pred - target).mean())
    g_img = (np.sign(pred - target) / pred.size).astype(np.float32)
    gs, gr = orc.composite_rays_train_backward(-g_img.sum(-1), g_img, sig, rgb, deltas, rays, ws, image, 1e-4)
    orc.field_backward(fd, xyzs, dirs, sp[n_mlp:], sp[:n_mlp], cp, gs, gr)
    return int(counter[0]), loss


def cpu_setup():
    from oracle import oracle as orc
    from stable_nerf_b200 import NeRFNetwork
    orc.set_threads(0)
    model = NeRFNetwork(channel_dim=CHANNELS)
    sp = model.sigma_net.params.detach().numpy().copy()
    n_mlp = model.sigma_net.n_mlp
    sp[n_mlp:] *= TABLE_SCALE
    cp = model.color_net.params.detach().numpy().copy()
    fd = orc.copy_desc(model.fdesc, orc.FieldDesc)
    return orc, fd, sp, cp, n_mlp


def cpu_baseline(budget_s=20.0):
    """oracle port on the host cores over a bounded sample of the cfg2 batch (reported baseline only)."""
    orc, fd, sp, cp, n_mlp = cpu_setup()
    bitfield, rays_o, rays_d, target = workload(RAYS_PER_GPU, 0)
    t0 = time.perf_counter()
    oracle_step(orc, fd, bitfield, rays_o[:32], rays_d[:32], target[:32], sp, cp, n_mlp)
    t32 = time.perf_counter() - t0
    n = int(min(RAYS_PER_GPU, max(32, 32 * budget_s / max(t32, 1e-3) // 32 * 32)))
    t0 = time.perf_counter()
    ns, _ = oracle_step(orc, fd, bitfield, rays_o[:n], rays_d[:n], target[:n], sp, cp, n_mlp)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "rays/s", "cores": orc.get_threads(), "kind": "port",
            "sample": f"first {n} of the {RAYS_PER_GPU} cfg2 rays ({ns} samples), one fwd+bwd step in {dt:.1f} s on "
                      f"the C oracle (OpenMP); the reference ships no CPU path (SURVEY R2)"}


def run_reference_arm(args):
    """--impl reference: the reference has no CPU implementation of this path (nerf/renderer.py:330-334 is CUDA only)
    and its tiny-cuda-nn dependency is absent, so the timed thing is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, fd, sp, cp, n_mlp = cpu_setup()
    bitfield, rays_o, rays_d, target = workload(RAYS_PER_GPU, 0)
    t0 = time.perf_counter()
    oracle_step(orc, fd, bitfield, rays_o[:32], rays_d[:32], target[:32], sp, cp, n_mlp)
    t32 = time.perf_counter() - t0
    budget = 150.0 / (args.steps + args.warmup)
    n = int(min(RAYS_PER_GPU, max(32, 32 * budget / max(t32, 1e-3) // 32 * 32)))
    for w in range(args.warmup):
        oracle_step(orc, fd, bitfield, rays_o[:n], rays_d[:n], target[:n], sp, cp, n_mlp)
    t0 = time.perf_counter()
    ns = 0
    for k in range(args.steps):
        lo = (k * n) % (RAYS_PER_GPU - n + 1)
        s, _ = oracle_step(orc, fd, bitfield, rays_o[lo:lo + n], rays_d[lo:lo + n], target[lo:lo + n], sp, cp, n_mlp)
        ns += s
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cb = {"value": value, "unit": "rays/s", "cores": orc.get_threads(), "kind": "port",
          "sample": f"{n} of the {RAYS_PER_GPU} cfg2 rays per step ({ns // max(args.steps, 1)} samples/step)"}
    print(json.dumps({
        "impl": "reference", "metric": "train rays/s (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 4096-ray train step, 800x800 Blender-shaped synthetic scene, hash grid + occupancy "
                               "marching, max_steps 1024 (CPU arm: bounded sample per step)"},
        "cpu_baseline": cb, "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "samples_per_s": ns / dt}))


# ------------------------------------------------------------------------------------------------ GPU arm

def l2_peaks(dev, n_entries=6098120):
    """What this GPU's L2 sustains for the hash-grid kernels' access pattern (csrc/probes/l2_probe.cu, libsnerf_probes.so --
    measurement only, not the product library): random 8-byte gathers / red.global.add.v2.f32 over a table the size of the
    hash table (46.5 MiB, L2-resident).  GB/s of USEFUL bytes (8 per access; the L2 moves a 32-byte sector for each)."""
    import ctypes
    import torch
    path = os.path.join(ROOT, "stable_nerf_b200", "libsnerf_probes.so")
    if not os.path.exists(path):
        return {"error": "libsnerf_probes.so not built"}
    lib = ctypes.CDLL(path)
    table = torch.zeros(n_entries * 2, device=dev)
    sink = torch.zeros(4, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    threads, per = 148 * 2048 * 4, 64
    out = {}
    for name, call in (("gather", lambda: lib.snerf_probe_l2_gather(ctypes.c_void_p(table.data_ptr()), n_entries, threads, per,
                                                                     ctypes.c_void_p(sink.data_ptr()), stream)),
                       ("reduce", lambda: lib.snerf_probe_l2_reduce(ctypes.c_void_p(table.data_ptr()), n_entries, threads, per,
                                                                     stream))):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            rc = call()
        e1.record()
        torch.cuda.synchronize()
        if rc != 0:
            return {"error": f"probe launch failed ({rc})"}
        out[name + "_GBs"] = threads * per * 8.0 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out["how"] = (f"{threads} threads x {per} random 8-byte accesses over {n_entries} float2 entries (46.5 MiB), 8 in flight per "
                  "thread, best-effort peak of useful bytes; CUDA events, 5 launches")
    return out


def ref_gpu_kernels():
    """BASELINE.md section 3 'B-ref-GPU': the unmodified reference kernels (oracle/_ref/_raymarching.so) next to ours, in a
    subprocess (scripts/ref_gpu_kernels.py)."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ref_gpu_kernels.py")], capture_output=True, text=True,
                           timeout=240)
        for line in r.stdout.splitlines():
            if line.startswith("REF_GPU_KERNELS "):
                return json.loads(line[len("REF_GPU_KERNELS "):])
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:  # an extra of the report must not take the bench line down
        return {"error": repr(e)[:200]}


def stage_times(model, ts, iters=10):
    """Device time of each stage of one step (CUDA events on the launch stream), eager, steady-state sizes."""
    import torch
    from stable_nerf_b200 import raymarching as rm
    m = model
    dev = ts.rays_o.device
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc, n_samples, M = {}, 0, 0
    aabb = m.aabb_train
    for it in range(iters + 2):
        marks = [("start", ev())]
        marks[0][1].record()

        def mark(name):
            e = ev()
            e.record()
            marks.append((name, e))
        nears, fars = rm.near_far_from_aabb(ts.rays_o, ts.rays_d, aabb, m.min_near)
        mark("near_far")
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        xyzs, dirs, deltas, rays = rm.march_rays_train(ts.rays_o, ts.rays_d, m.bound, m.density_bitfield, m.cascade,
                                                       m.grid_size, nears, fars, counter, m.mean_count, False, 128,
                                                       False, 0, ts.max_steps)
        mark("march")
        sig, rgb = m(xyzs, dirs)
        mark("field_fwd")
        ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, rays, ts.T_thresh, m.channel_dim)
        mark("composite_fwd")
        pred = image + (1 - ws).unsqueeze(-1)
        loss = (pred - ts.target).abs().mean()
        g_sig, g_rgb = torch.autograd.grad(loss, [sig, rgb], retain_graph=True)
        mark("loss+composite_bwd")
        torch.autograd.backward([sig, rgb], [g_sig, g_rgb])
        mark("field_bwd")
        torch.cuda.synchronize()
        if it >= 2:
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / iters
        n_samples, M = int(counter[0].item()), xyzs.shape[0]
    return acc, n_samples, M


def render_bench(model, dev, world, rank, frames, barrier, max_over_ranks, min_n_step=1):
    """cfg3: full-frame 800x800 inference render (march_rays / field / composite_rays / compact_rays loop,
    nerf/renderer.py:117-167), pixel rows sharded over the ranks.  Returns the `render` object of the JSON line."""
    import torch
    from stable_nerf_b200 import synthetic as syn
    ro, rd = syn.full_frame()
    # pixels interleaved over the ranks (rank r renders pixels r, r+W, ...): every shard is a uniform subsample of the
    # image, so the ranks carry equal work (contiguous row blocks leave the object to the middle ranks)
    ro, rd = (torch.from_numpy(np.ascontiguousarray(a[rank::world])).to(dev)[None] for a in (ro, rd))
    was_training = model.training
    model.eval()
    model.min_n_step = min_n_step
    with torch.no_grad():
        model.render(ro, rd, bg_color=1, max_steps=MAX_STEPS)  # warm-up frame (sizes the workspaces)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rows = samples = iters = 0
        for _ in range(frames):
            out = model.render(ro, rd, bg_color=1, max_steps=MAX_STEPS)
            st = model.last_render_stats
            rows, samples, iters = rows + st["rows"], samples + st["samples"], iters + st["iterations"]
        e1.record()
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    upd_ms = None
    if min_n_step == 1:  # cfg3 "with density-grid update": one full sweep (2 097 152 density queries), nerf/renderer.py:236-327
        state = (model.density_grid.clone(), model.density_bitfield.clone(), model.mean_density, model.iter_density,
                 model.mean_count, model.local_step)
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        model.iter_density = 0
        model.update_extra_state()
        model.iter_density = 0
        u0.record()
        model.update_extra_state()
        u1.record()
        torch.cuda.synchronize()
        upd_ms = u0.elapsed_time(u1)
        model.density_grid.copy_(state[0])
        model.density_bitfield.copy_(state[1])
        model.mean_density, model.iter_density, model.mean_count, model.local_step = state[2:]
    tot = torch.tensor([rows, samples], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot)
    model.train(was_training)
    model.min_n_step = 1
    return {"workload": "cfg3: 800x800 full-frame inference render, eval loop with on-device compaction, T_thresh 1e-4, "
                        f"max_steps {MAX_STEPS}, pixels interleaved over {world} GPU(s), "
                        + ("the reference's loop schedule (n_step from 1)" if min_n_step == 1 else
                           f"at least {min_n_step} samples per ray and iteration (images equal to 1e-6)"),
            "samples_per_s": float(tot[1]) / (ms * 1e-3), "rows_per_s_incl_padding": float(tot[0]) / (ms * 1e-3),
            "ms_per_frame": ms / frames, "frames": frames, "rays": 640000, "loop_iterations_per_frame": iters / frames,
            "samples_per_frame": float(tot[1]) / frames, "unit": "samples/s", "density_grid_update_ms": upd_ms}


def first_epoch_profile(model, ts, iters=10):
    """cfg2 on the reference's first-epoch path: before update_extra_state has produced
    mean_count, march_rays_train sizes its outputs from the measured sample total -- one D2H read per step
    (raymarching.py:196-229) -- so the step cannot be a graph; it runs through NeRFNetwork.render + autograd, eagerly.
    This is the path TrainStep.warmup() itself takes for its first steps."""
    import torch
    mean_count, local_step, bufs = model.mean_count, model.local_step, ts._bufs
    model.mean_count = 0
    res = {}
    try:
        for fused in (True, False):
            ts.first_epoch_fused = fused
            for _ in range(3):
                ts._body()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                ts._body()
            e1.record()
            torch.cuda.synchronize()
            res[fused] = e0.elapsed_time(e1) / iters
    finally:
        ts.first_epoch_fused = True
        model.mean_count, model.local_step, ts._bufs = mean_count, local_step, bufs
    ms = res[True]
    return {"workload": "cfg2 before mean_count exists (SURVEY Q8): the fused step issued eagerly, rows sized from the sample total "
                        "read back after the count pass (one 4-byte D2H per step, no graph)",
            "ms_per_step": ms, "rays_per_s": RAYS_PER_GPU / (ms * 1e-3),
            "through_render_and_autograd_ms": res[False]}


def large_batch_profile(model, dev, pk, n_rays=1 << 18):
    """cfg5's per-step ray count on ONE GPU (2^18 rays, ~22 M samples): at this size the byte-moving kernels are no
    longer launch-bound, so their HBM fractions mean something.  Eager fused step, CUDA events per stage / kernel."""
    import torch
    from stable_nerf_b200.trainer import TrainStep
    _, rays_o, rays_d, target = workload(n_rays, seed=7)
    model.mean_count, model.local_step = 0, 0
    ts = TrainStep(model, n_rays, max_steps=MAX_STEPS, use_graph=False)
    d = [torch.from_numpy(a).to(dev) for a in (rays_o, rays_d, target)]
    ts.warmup(*d, iters=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ts.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    st, ns, M = ts.profile_stages(iters=3)
    kt = ts.profile_field_kernels(iters=3)
    C = CHANNELS
    work = {
        "march": ("hbm", st.get("march", 0) * 1e3, 48.0 * n_rays + 32.0 * ns),
        "composite_fwd": ("hbm", st.get("composite_fwd", 0) * 1e3, (12 + 4 * C) * ns + (20 + 4 * C) * n_rays),
        "composite_bwd": ("hbm", st.get("composite_bwd", 0) * 1e3, (16 + 8 * C) * ns + (20 + 8 * C) * n_rays),
        "hashgrid_gather": ("l2_gather", kt.get("hashgrid_gather", 0), 1024.0 * M),
        "hashgrid_scatter": ("l2_reduce", kt.get("hashgrid_scatter", 0), 1024.0 * M),
        "sigma_net_fwd": ("tensor", kt.get("sigma_net_fwd", 0), 2.0 * 38912 * M),
        "color_net_fwd": ("tensor", kt.get("color_net_fwd", 0), 2.0 * 55296 * M),
        "color_net_bwd": ("tensor", kt.get("color_net_bwd", 0), 4.0 * 55296 * M),  # dgrad + wgrad (recompute not credited)
        "sigma_net_bwd": ("tensor", kt.get("sigma_net_bwd", 0), 4.0 * 38912 * M),
    }
    roof = {}
    for k, (b, us, w) in work.items():
        if us > 0:
            a = w / (us * 1e-6) / (1e12 if b == "tensor" else 1e9)
            peak = pk["tf_sust"] if b == "tensor" else (pk["hbm"] if b == "hbm" else pk.get(b))
            roof[k] = {"bound": b, "us": round(us, 1), "achieved": round(a, 1), "unit": "TFLOP/s" if b == "tensor" else "GB/s",
                       "frac": round(a / peak, 4) if peak else None}
    del ts
    torch.cuda.empty_cache()
    return {"workload": f"cfg5 per-step batch on one GPU: {n_rays} rays, {ns} samples, eager fused step (no graph)",
            "rays": n_rays, "samples": ns, "ms_per_step": ms, "rays_per_s": n_rays / (ms * 1e-3),
            "samples_per_s": ns / (ms * 1e-3), "stage_rooflines": roof}


def cfg5_profile(args, dev, world, rank, bitfield, barrier, max_over_ranks, n_total=1 << 18):
    """BASELINE.json configs[4]: 2^18 rays per step split evenly over the ranks (strong scaling), fwd + bwd + gradient
    exchange as one replayed graph per rank; next to it the same 2^18-ray step on ONE GPU (every rank runs it on its own
    device at the same time; rank 0's time is reported), which is what the efficiency is measured against."""
    import torch
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.trainer import TrainStep, shard_range

    def make_model():
        torch.manual_seed(0)
        m = NeRFNetwork(channel_dim=CHANNELS, precision=args.precision).to(dev)
        with torch.no_grad():
            m.sigma_net.params[m.sigma_net.n_mlp:] *= TABLE_SCALE
        m.density_bitfield.copy_(torch.from_numpy(bitfield))
        m.train()
        return m

    def timed_steps(ts, packed, steps):
        for k in range(3):
            ts.step_from_packed(packed[k % len(packed)])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            ts.step_from_packed(packed[k % len(packed)])
        ts.finish()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    full = [workload(n_total, seed=7000 + k) for k in range(2)]  # the same two global batches on every rank
    lo, hi = shard_range(n_total, rank, world)
    out = {"workload": f"cfg5: {n_total} rays per step over {world} GPUs ({hi - lo} per rank), fwd + L1 + bwd + gradient exchange, "
                       "one CUDA graph per rank, 2 global batches rotated"}
    # ---- sharded
    model = make_model()
    ts = TrainStep(model, hi - lo, max_steps=MAX_STEPS, use_graph=not args.no_graph, world_size=world, loss_scale=1.0 / world,
                   exchange=args.exchange, pipeline=args.pipeline,
                   overlap_exchange={"auto": "auto", "on": True, "off": False}[args.overlap_exchange],
                   overlap_split_level=[int(v) for v in str(args.overlap_split_level).split(",")],
                   overlap_side_ctas=args.overlap_side_ctas)
    shards = [tuple(torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in b[1:]) for b in full]
    packed = [torch.cat([t.reshape(-1) for t in s]) for s in shards]
    ts.warmup(*shards[0], iters=2, batches=shards)
    steps = 10
    ms_w = timed_steps(ts, packed, steps)
    ex_us = None
    if ts.exchange is not None:  # the exchange alone: same number of calls on every rank
        for _ in range(3):
            ts.exchange.all_reduce()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(10):
            ts.exchange.all_reduce()
        x1.record()
        barrier()
        ex_us = max_over_ranks(x0.elapsed_time(x1)) / 10 * 1e3
        calls, waits = ts.exchange.status()
        out["exchange_status"] = {"calls": calls, "timeouts": waits}
    out.update({"ms_per_step": ms_w, "rays_per_s": n_total / (ms_w * 1e-3), "exchange_kind": ts.exchange_kind,
                "exchange_us": ex_us, "exchange_bytes": 4 * sum(p.numel() for p in ts.params),
                "samples_per_rank": int(sum(ts.warmup_counts) / len(ts.warmup_counts))})
    if ts.exchange is not None:
        for p in ts.params:
            p.grad = None
        ex, ts.exchange = ts.exchange, None
        del ts
        ex.close()
    del model, shards, packed
    torch.cuda.empty_cache()
    # ---- the same step on one GPU
    model1 = make_model()
    ts1 = TrainStep(model1, n_total, max_steps=MAX_STEPS, use_graph=not args.no_graph)
    fulls = [tuple(torch.from_numpy(a).to(dev) for a in b[1:]) for b in full]
    packed1 = [torch.cat([t.reshape(-1) for t in s]) for s in fulls]
    ts1.warmup(*fulls[0], iters=2, batches=fulls)
    ms_1 = timed_steps(ts1, packed1, 5)
    out["one_gpu"] = {"ms_per_step": ms_1, "rays_per_s": n_total / (ms_1 * 1e-3)}
    out["speedup_vs_one_gpu"] = ms_1 / ms_w
    out["efficiency_vs_one_gpu"] = ms_1 / ms_w / world
    del ts1, model1
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from stable_nerf_b200 import NeRFNetwork, _lib
    from stable_nerf_b200.trainer import TrainStep, broadcast_occupancy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # measured on this pool (scripts/allreduce_probe.py, 8 ranks, 49 MB fp32): Ring 185 us, NCCL's default choice
        # 220 us, and for the two half-size slices of the overlapped exchange 107 vs 171 us each
        os.environ.setdefault("NCCL_ALGO", "Ring")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()  # fail loudly when the extension is missing

    # N_BATCHES distinct ray batches per rank (different pixels, different views): the timed loops rotate through them, so
    # the sample total, the padding / overflow against M and the L2 contents change from step to step as in training
    batches = [workload(RAYS_PER_GPU, seed=rank * 1000 + k) for k in range(N_BATCHES)]
    bitfield, rays_o, rays_d, target = batches[0]
    model = NeRFNetwork(channel_dim=CHANNELS, precision=args.precision).to(dev)
    with torch.no_grad():
        model.sigma_net.params[model.sigma_net.n_mlp:] *= TABLE_SCALE
    model.density_bitfield.copy_(torch.from_numpy(bitfield))
    broadcast_occupancy(model)
    model.train()

    # (--pipeline: the gradient exchange of step k beside step k+1's ray march, TrainStep(pipeline=True).  Measured on 2 and 8
    # B200 it does not pay for the exchange -- 0.731 vs 0.718 and 0.812 vs 0.808 ms/step -- so it is off by default; the
    # optimiser update does profit from it, see with_optimizer)
    ts = TrainStep(model, RAYS_PER_GPU, max_steps=MAX_STEPS, use_graph=not args.no_graph, world_size=world,
                   loss_scale=1.0 / world, exchange=args.exchange, pipeline=world > 1 and args.pipeline,
                   overlap_exchange={"auto": "auto", "on": True, "off": False}[args.overlap_exchange],
                   overlap_split_level=[int(v) for v in str(args.overlap_split_level).split(",")],
                   overlap_side_ctas=args.overlap_side_ctas)
    d_batches = [tuple(torch.from_numpy(a).to(dev) for a in b[1:]) for b in batches]
    d_packed = [torch.cat([t.reshape(-1) for t in b]) for b in d_batches]   # [rays_o | rays_d | target], resident in HBM
    h_packed = []                                                           # the same in pinned host memory (one H2D copy)
    for b in batches:
        st, views = ts.new_pinned_batch()
        for h, a in zip(views, b[1:]):
            h.copy_(torch.from_numpy(a))
        h_packed.append(st)
    d_o, d_d, d_t = d_batches[0]
    # reference-style warm-up over the rotating batches: mean_count = mean of their sample totals (nerf/renderer.py:321-325)
    ts.warmup(d_o, d_d, d_t, iters=N_BATCHES, batches=d_batches)
    batch_counts = list(ts.warmup_counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: inputs already in HBM, K steps, CUDA events, max over ranks
    for k in range(args.warmup):
        ts.step_from_packed(d_packed[k % N_BATCHES])
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    with ClockSampler(local, uuid) as clocks:
        barrier()
        with clocks.mark():
            e0.record()
            for k in range(args.steps):
                ts.step_from_packed(d_packed[k % N_BATCHES])  # one 147 KB device copy + the step (a graph replay)
            ts.finish()  # (pipelined exchange of the last step)
            e1.record()
            barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = (ts.launches_per_step if ts.graph is not None else (_lib.launch_count() - launches0) // args.steps)
        M_step = ts._bufs["M"] if ts._bufs is not None else 0
        n_samples = int(round(sum(batch_counts) / len(batch_counts)))

        # ---- the same batch replayed every step (what round 1 reported as `value`): exactly-sized M, warm L2
        ts.step_from_packed(d_packed[0])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            ts.step()
        ts.finish()
        f1.record()
        barrier()
        frozen_ms = max_over_ranks(f0.elapsed_time(f1))

        # ---- end to end, the data-loader form: every step copies ITS inputs from pinned host memory (147 KB, started on a
        # copy stream while the previous step computes) and brings ITS loss back to the host (4 B, into a pinned slot); the
        # host reads the loss of step k-1 while step k runs, so the device never waits for the host.  Wall clock between
        # device-complete points, all copies and the last loss inside.
        for k in range(args.warmup):  # (also creates the copy stream, the staging buffer and the pinned loss slots)
            ts.step_from_packed(h_packed[k % N_BATCHES], read_loss="lagged", next_packed=h_packed[(k + 1) % N_BATCHES])
        ts.drain()
        barrier()
        with clocks.mark():
            t0 = time.perf_counter()
            for k in range(args.steps):
                ts.step_from_packed(h_packed[k % N_BATCHES], read_loss="lagged", next_packed=h_packed[(k + 1) % N_BATCHES])
            loss = ts.drain()
            ts.finish()
            barrier()
            e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        # ---- the synchronous form of the same (H2D, step, loss D2H, host waits for the loss before the next step)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            loss_sync = ts.step_from_packed(h_packed[k % N_BATCHES], read_loss=True)
        ts.finish()
        barrier()
        e2e_sync_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)

    total_rays = RAYS_PER_GPU * world
    out = {
        "metric": "train rays/s (fwd+bwd)", "value": total_rays * args.steps / (ms * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 4096-ray train step per GPU, 800x800 Blender-shaped synthetic scene (lego-like "
                               "occupancy, C=1, H=128), hash grid 16x2 T=2^19 + occupancy marching, max_steps 1024, "
                               "channel_dim 3, fwd + L1 + bwd",
                   "rays_per_gpu": RAYS_PER_GPU, "samples_per_step_per_gpu": n_samples, "mlp_precision": args.precision,
                   "batches": f"{N_BATCHES} distinct seeded ray batches per rank rotated through the timed loops; rank 0's "
                              f"sample totals {batch_counts}; rows per step M = pad128(mean) = {M_step}: "
                              f"{sum(c > M_step for c in batch_counts)} of them overflow M and lose their last rays as in "
                              "the reference (raymarching.cu:417)",
                   "cuda_graph": ts.graph is not None, "parallelism": f"ray-sharded dp{world}",
                   "exchange_schedule": ("the exchange of step k runs beside the ray march of step k+1 (second stream inside the "
                                         "graph, joined before the hash-grid gather); K exchanges for K steps, the last one by "
                                         "finish() inside the timed region") if ts._pipelined() else (
                                         f"split against the table scatter-add: levels [{ts.overlap_split_level}, L) are scattered first "
                                         f"({ts.overlap_side_ctas} CTAs for that exchange) "
                                         "and their slice of the arena is exchanged on a second stream (flag channel 1) beside the "
                                         "coarse levels' scatter-add; MLP gradients + coarse levels follow on the step's stream"
                                         if ts._overlap_exchange_on() else "after the backward, same stream"),
                   "gradient_exchange": {"none": "none (one rank)",
                                         "nvls": "one kernel per rank, reduced inside the NVSwitch (multimem.ld_reduce / "
                                                 "multimem.st on a multicast mapping of the gradient arenas), inside the "
                                                 "step's CUDA graph (csrc/p2p_reduce.cu)",
                                         "p2p": "one kernel per rank over NVLink peer memory, inside "
                                         "the step's CUDA graph (csrc/p2p_reduce.cu)",
                                         "nccl": "NCCL all-reduce after the graph replay"}[ts.exchange_kind]
                                        + (f" [peer memory unavailable: {ts.exchange_error}]" if ts.exchange_error else ""),
                   "l2": "not flushed between steps: the step streams params 49 MB + grads 49 MB + samples and "
                         "activations (> 126 MB L2 together); the hash table is meant to stay L2-resident across steps"},
        "e2e": {"value": total_rays * args.steps / (e2e_ms * 1e-3), "unit": "rays/s",
                "h2d_bytes_per_step": int(h_packed[0].numel() * 4), "d2h_bytes_per_step": 4,
                "form": "TrainStep.step_from_packed(read_loss='lagged', next_packed=...): per step one H2D copy of that step's "
                        "inputs (copy stream, under the previous step) and one D2H of that step's loss, read by the host one "
                        "step later; synchronous form (host waits for every loss): "
                        f"{total_rays * args.steps / (e2e_sync_ms * 1e-3):.0f} rays/s"},
        "frozen_batch": {"ms_per_step": frozen_ms / args.steps, "rays_per_s": total_rays * args.steps / (frozen_ms * 1e-3),
                         "note": "one batch replayed every step with no input copy (round 1's `value`): optimistic"},
        "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
        "samples_per_s": n_samples * world * args.steps / (ms * 1e-3), "loss": loss,
        "clocks": clocks.summary(),
    }

    if rank == 0 and world == 1 and not args.no_stages and ts.fused:
        # the same step with the fused Adam update (test_nerf.py:52 hyper-parameters): fwd + bwd + optimiser.
        # (a) the update inside the step's graph, beside the next step's ray march (pipeline=True); (b) stepped eagerly
        # after the replay (round 1's form)
        from stable_nerf_b200.optim import FusedAdam
        saved_params = [p.detach().clone() for p in ts.params]
        opt = FusedAdam(model.get_params(1e-4), betas=(0.9, 0.99), eps=1e-15, capturable=True, zero_grad_in_step=True)
        ts_opt = TrainStep(model, RAYS_PER_GPU, max_steps=MAX_STEPS, use_graph=not args.no_graph, optimizer=opt, pipeline=True)
        local_step, mean_count = model.local_step, model.mean_count
        model.mean_count = 0
        ts_opt.warmup(d_o, d_d, d_t, iters=N_BATCHES, batches=d_batches)
        for k in range(3):
            ts_opt.step_from_packed(d_packed[k % N_BATCHES])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for k in range(args.steps):
            ts_opt.step_from_packed(d_packed[k % N_BATCHES])
        ts_opt.finish()
        a1.record()
        torch.cuda.synchronize()
        ms_opt = a0.elapsed_time(a1) / args.steps
        del ts_opt
        ts.optimizer = FusedAdam(model.get_params(1e-4), betas=(0.9, 0.99), eps=1e-15)
        for k in range(3):
            ts.step_from_packed(d_packed[k % N_BATCHES])
        torch.cuda.synchronize()
        a0.record()
        for k in range(args.steps):
            ts.step_from_packed(d_packed[k % N_BATCHES])
        a1.record()
        torch.cuda.synchronize()
        ms_opt_eager = a0.elapsed_time(a1) / args.steps
        ts.optimizer = None
        with torch.no_grad():  # the measurements that follow run on the parameters the step was timed with
            for p, q in zip(ts.params, saved_params):
                p.copy_(q)
        model.local_step, model.mean_count = local_step, mean_count
        out["with_optimizer"] = {"ms_per_step": ms_opt, "rays_per_s": RAYS_PER_GPU / (ms_opt * 1e-3),
                                 "extra_us_over_value": round((ms_opt - ms / args.steps) * 1e3, 1),
                                 "optimizer": "FusedAdam betas=(0.9,0.99) eps=1e-15 over 12.29 M fp32 params (344 MB/step), "
                                              "capturable: its launches are part of the step's CUDA graph, run beside the "
                                              "next step's ray march, and leave the gradients zeroed (no memset)",
                                 "stepped_eagerly_after_the_replay_ms": ms_opt_eager}
    if world > 1:
        # what a reader needs to trust a multi-GPU number: which exchange ran, that no wait ran out, and that every rank
        # holds the same gradients after it
        calls, waits = ts.exchange.status() if ts.exchange is not None else (0, 0)
        cs = torch.stack([p.grad.double().sum() for p in ts.params]).sum().reshape(1)
        cs_hi, cs_lo = cs.clone(), cs.clone()
        dist.all_reduce(cs_hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs_lo, op=dist.ReduceOp.MIN)
        w = torch.tensor([float(waits)], dtype=torch.float64, device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        out["exchange_status"] = {"kind": ts.exchange_kind, "calls_rank0": calls, "timeouts_max_over_ranks": int(w.item()),
                                  "gradient_checksum": float(cs_hi.item()),
                                  "gradient_checksum_equal_on_all_ranks": bool(cs_hi.item() == cs_lo.item()),
                                  "note": ts.exchange_error}
    if not args.no_render:
        out["render"] = render_bench(model, dev, world, rank, 3, barrier, max_over_ranks)
        out["render_min_n_step_4"] = render_bench(model, dev, world, rank, 3, barrier, max_over_ranks, min_n_step=4)
    if rank == 0 and world == 1 and not args.no_stages:
        pk = peaks()
        local_step = model.local_step
        st, ns, M = ts.profile_stages() if ts.fused else stage_times(model, ts)
        model.local_step = local_step
        out["stages_ms"] = {k: round(v, 4) for k, v in st.items()}
        C = CHANNELS
        alg = {  # algorithmic bytes / flops per launch (SURVEY section 8d figures)
            "near_far": ("hbm", 32.0 * RAYS_PER_GPU),
            "march": ("hbm", (32.0 + 48.0) * RAYS_PER_GPU + 32.0 * ns),  # near/far folded into the count launch
            "composite_fwd": ("hbm", (12 + 4 * C) * ns + (20 + 4 * C) * RAYS_PER_GPU),
            "composite_bwd": ("hbm", (16 + 8 * C) * ns + (20 + 8 * C) * RAYS_PER_GPU),
            "loss+composite_bwd": ("hbm", (16 + 8 * C) * ns + (20 + 8 * C) * RAYS_PER_GPU),
            "field_fwd": ("tensor", 188416.0 * M),
            "field_bwd": ("tensor", 565248.0 * M),  # recompute + dgrad + wgrad
        }
        # per-kernel device times of the field calls (one kernel per launch through the library's stage mask)
        kt = ts.profile_field_kernels() if ts.fused else {}
        out["field_kernels_us"] = {k: round(v, 2) for k, v in kt.items()}
        # algorithmic work per launch, SURVEY section 8d per-sample figures x M.  Tensor kernels: flops; the BACKWARD
        # kernels are credited with dgrad + wgrad = 4 MACs-worth per weight (565 248 - 188 416 flop/sample over both nets) --
        # their on-chip forward recompute is work the design chose to do, not algorithmic work (it is reported separately
        # as `issued`: 6 x MACs).  Hash-grid kernels: the 1024 B/sample of table gathers / reductions are L2 traffic and
        # are measured against the L2 peaks below; what must cross HBM is the sample's own rows.
        kalg = {
            "hashgrid_gather": ("l2_gather", 1024.0 * M),
            "sigma_net_fwd": ("tensor", 2.0 * 38912 * M),
            "color_net_fwd": ("tensor", 2.0 * 55296 * M),
            "color_net_bwd": ("tensor", 4.0 * 55296 * M),
            "sigma_net_bwd": ("tensor", 4.0 * 38912 * M),
            "hashgrid_scatter": ("l2_reduce", 1024.0 * M),
        }
        issued = {"color_net_bwd": 6.0 * 55296 * M, "sigma_net_bwd": 6.0 * 38912 * M}
        l2 = l2_peaks(dev)
        out["l2_peaks"] = l2
        pk["l2_gather"], pk["l2_reduce"] = l2.get("gather_GBs"), l2.get("reduce_GBs")
        ncu_traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_kernel_traffic.json")
        if os.path.exists(tpath):
            ncu_traffic = json.load(open(tpath))
        cands = {k: (v * 1e-3, kalg[k]) for k, v in kt.items() if k in kalg}
        for k in ("near_far", "march", "composite_fwd", "composite_bwd"):
            if k in st:
                cands[k] = (st[k], alg[k])
        def roof(bound, work, t_ms):
            t = t_ms * 1e-3
            if bound == "tensor":
                return work / t / 1e12, pk["tf_sust"], "TFLOP/s"
            peak = pk["hbm"] if bound == "hbm" else pk.get(bound)
            return work / t / 1e9, peak, "GB/s"
        dom = max(cands, key=lambda k: cands[k][0])
        t_ms, (bound, work) = cands[dom]
        ach, peak, unit = roof(bound, work, t_ms)
        out["roofline"] = {"kernel": dom, "bound": "tensor" if bound == "tensor" else "hbm", "achieved": ach, "peak": peak,
                           "unit": unit, "frac": (ach / peak) if peak else None, "traffic": ncu_traffic.get(dom),
                           "peak_source": pk["src"] if bound in ("tensor", "hbm") else "measured in this run (l2_peaks)",
                           "launch_us": t_ms * 1e3,
                           "accounting": "backward kernels: 4 x MACs (dgrad + wgrad, SURVEY 8d); forward recompute not credited"
                                         if dom in issued else "SURVEY 8d per-sample figure x samples"}
        if dom in issued:
            out["roofline"]["issued_frac"] = issued[dom] / (t_ms * 1e-3) / 1e12 / pk["tf_sust"]
        if bound.startswith("l2"):
            out["roofline"]["bound_detail"] = f"L2-resident table: {bound} peak of this run"
        # What actually binds the backward MLP kernels is the SM's shared-memory data pipe (128 B/clk/SM: tensor-core
        # operand fetches + the epilogues' LDS/STS + the bulk copies of the weight ring all go through it), not the tensor
        # pipe.  Bytes per 128-sample tile from the kernel's static schedule (checked against the ncu
        # wavefront counters of profiles/r2b_ncu_full_bwd_smem_pipe.csv: tc 910 / 653 KB, TMA 228 / 164 KB per tile as
        # modelled, LSU 524 / 459 KB incl. ~100 KB of bank-conflict replays), times the tiles of the busiest CTA, against
        # 128 B/clk at the SM clock sampled in this run.
        smem_tile = {"color_net_bwd": {"mma_operands": 908, "epilogue_lsu": 414, "weight_ring": 232},
                     "sigma_net_bwd": {"mma_operands": 652, "epilogue_lsu": 332, "weight_ring": 164}}
        smem_ncu = {"color_net_bwd": 0.850, "sigma_net_bwd": 0.838}  # tc + lsu + tma wavefronts, % of peak, under ncu
        def smem_pipe(k, tm_ms):
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            tiles = -(-(-(-M // 128)) // n_sm)  # ceil(ceil(M/128) / SMs): the busiest CTA's tiles
            kb = sum(smem_tile[k].values())
            mhz = (out.get("clocks") or {}).get("sm_mhz") or 1965.0
            peak = 128.0 * mhz * 1e6 / 1e9  # GB/s per SM
            ach = tiles * kb * 1024.0 / (tm_ms * 1e-3) / 1e9
            return {"bound": "shared-memory data pipe (128 B/clk/SM)", "KB_per_tile": smem_tile[k], "tiles_per_cta": tiles,
                    "achieved_GBs_per_sm": round(ach, 1), "peak_GBs_per_sm": round(peak, 1), "frac_model": round(ach / peak, 4),
                    "frac_ncu_wavefronts": smem_ncu[k], "source": "profiles/r2b_ncu_full_bwd_smem_pipe.csv"}
        if dom in smem_tile:
            out["roofline"]["smem_pipe"] = smem_pipe(dom, t_ms)
        out["stage_rooflines"] = {}
        for k, (tm, (b_, w)) in cands.items():
            if tm > 0:
                a_, p_, u_ = roof(b_, w, tm)
                out["stage_rooflines"][k] = {"bound": b_, "achieved": round(a_, 2), "unit": u_,
                                             "frac": round(a_ / p_, 4) if p_ else None}
                if k in issued:
                    out["stage_rooflines"][k]["issued_frac_incl_recompute"] = round(issued[k] / (tm * 1e-3) / 1e12 / pk["tf_sust"], 4)
                    out["stage_rooflines"][k]["smem_pipe"] = smem_pipe(k, tm)
                if k in ("hashgrid_gather", "hashgrid_scatter"):
                    own = (12 + 64) * M if k == "hashgrid_gather" else (12 + 128) * M  # what must cross HBM: the sample's rows
                    out["stage_rooflines"][k].update({
                        "served_by": "L2-resident table (46.5 MiB of the 126 MB L2); peak = l2_peaks of this run",
                        "hbm_compulsory_GBs": round(own / (tm * 1e-3) / 1e9, 2),
                        "hbm_compulsory_frac": round(own / (tm * 1e-3) / 1e9 / pk["hbm"], 4),
                        "dram_bytes_ncu": ncu_traffic.get(k)})
        try:
            out["first_epoch_path"] = first_epoch_profile(model, ts)
        except Exception as e:  # an extra line of the report must not take the bench line down
            out["first_epoch_path"] = {"error": repr(e)[:200]}
    if rank == 0 and world == 1 and not args.no_large and not args.no_stages and ts.fused:
        out["large_batch"] = large_batch_profile(model, dev, pk)
    if rank == 0 and world == 1 and not args.no_ref_kernels:
        torch.cuda.synchronize()
        out["ref_gpu_kernels"] = ref_gpu_kernels()
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline()
    if world > 1 and not args.no_cfg5:
        try:
            if ts.exchange is not None:  # release the cfg2 step's arena before the cfg5 step maps its own
                for p in ts.params:
                    p.grad = None
                ex, ts.exchange = ts.exchange, None
                ex.close()
            out["cfg5"] = cfg5_profile(args, dev, world, rank, bitfield, barrier, max_over_ranks)
        except Exception as e:  # collective code: an exception here is raised on every rank or on none
            out["cfg5"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SNERF_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-stages", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-large", action="store_true")
    ap.add_argument("--no-ref-kernels", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--overlap-exchange", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--overlap-split-level", default="8", help="cut level(s) of the split exchange, e.g. 8 or 10,6")
    ap.add_argument("--overlap-side-ctas", type=int, default=16)
    ap.add_argument("--exchange", default="auto", choices=["auto", "nvls", "p2p", "nccl"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
