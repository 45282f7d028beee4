"""Golden trace of every ``_raymarching`` backend call the UNMODIFIED reference makes (b).

The reference reaches its kernels through ``submodules/raymarching/raymarching.py`` (autograd wrappers, :9-12 import the
pybind11 module ``_raymarching``) called from ``nerf/renderer.py`` (``run_cuda``, :70-172).  Here both files are imported
UNMODIFIED from /root/reference and run on the CPU of the build container over a stand-in ``_raymarching`` module whose ten
functions (raymarching.h:7-18 signatures: preallocated output tensors filled in place) are bound to the C oracle -- itself
pinned bit-exactly against the reference kernels (tests/test_oracle.py).  ``Tensor.cuda()`` is an identity; the field is an
analytic one made of separately rounded fp32 torch ops (same bits on CPU and GPU).  A recorder notes, for every backend
call, the function name, every argument (tensors by value BEFORE the call) and the output tensors AFTER the call.

What runs, for each scene of tests/trace_scene.py (s1: one cascade, the reference's defaults; s2: bound 2 = two cascades,
dt_gamma 1/128, 4 channels, density_scale 0.5): one training ``run_cuda`` on the first-epoch path (mean_count = 0:
N*max_steps rows, raymarching.py:196-229), its backward (composite_rays_train_backward), one training ``run_cuda`` with a
too-small ``mean_count`` (overflowing rays are dropped, raymarching.cu:417), one full inference ``run_cuda`` (the n_step
schedule of nerf/renderer.py:130, every iteration), for s2 also a perturbed training and inference run; and direct
wrapper calls of sph_from_ray / morton3D / morton3D_invert / packbits.

``tests/test_backend_trace.py`` replays the trace on the GPU against ``stable_nerf_b200.backend`` (the same ten functions
over libsnerf_b200.so) and against ``stable_nerf_b200.raymarching`` + ``NeRFRenderer`` (the drop-in operator surface).

Run:  python tests/golden/make_golden_backend_trace.py   ->  tests/golden/backend_trace.npz
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from oracle import oracle as orc  # noqa: E402
from trace_scene import SCENES, analytic_field, scene_inputs  # noqa: E402

torch.set_num_threads(1)
torch.Tensor.cuda = lambda self, *a, **k: self

# ------------------------------------------------------------------------------- oracle-backed `_raymarching` stand-in
# (signatures: submodules/raymarching/src/raymarching.h:7-18; outputs are filled in place)


def _np(t):
    return t.detach().contiguous().numpy()


def _put(dst, arr):
    dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view_as(dst))


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    n, f = orc.near_far_from_aabb(_np(rays_o), _np(rays_d), _np(aabb), min_near)
    _put(nears, n), _put(fars, f)


def sph_from_ray(rays_o, rays_d, radius, N, coords):
    _put(coords, orc.sph_from_ray(_np(rays_o), _np(rays_d), radius))


def morton3D(coords, N, indices):
    _put(indices, orc.morton3D(_np(coords)))


def morton3D_invert(indices, N, coords):
    _put(coords, orc.morton3D_invert(_np(indices)))


def packbits(grid, N, density_thresh, bitfield):
    _put(bitfield, orc.packbits(_np(grid).reshape(-1), np.float32(density_thresh)))


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                     counter, noises):
    x, d, dl, r, cnt = orc.march_rays_train(_np(rays_o), _np(rays_d), bound, _np(grid), C, H, _np(nears), _np(fars),
                                            noises=_np(noises), dt_gamma=dt_gamma, max_steps=max_steps, M=M)
    # (the oracle returns all M rows: what the kernel produced, zeros elsewhere -- the wrapper's buffers are zero-initialised)
    _put(xyzs, x), _put(dirs, d), _put(deltas, dl), _put(rays, r)
    counter += torch.from_numpy(cnt)


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, channel_dim, weights_sum, depth, image):
    ws, dp, im = orc.composite_rays_train_forward(_np(sigmas), _np(rgbs).reshape(-1, channel_dim), _np(deltas), _np(rays),
                                                  T_thresh)
    _put(weights_sum, ws), _put(depth, dp), _put(image, im)


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                  T_thresh, channel_dim, grad_sigmas, grad_rgbs):
    gs, gr = orc.composite_rays_train_backward(_np(grad_weights_sum), _np(grad_image).reshape(-1, channel_dim), _np(sigmas),
                                               _np(rgbs).reshape(-1, channel_dim), _np(deltas), _np(rays), _np(weights_sum),
                                               _np(image).reshape(-1, channel_dim), T_thresh)
    _put(grad_sigmas, gs), _put(grad_rgbs, gr)


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
               xyzs, dirs, deltas, noises):
    x, d, dl = orc.march_rays(n_alive, n_step, _np(rays_alive), _np(rays_t), _np(rays_o), _np(rays_d), bound, _np(grid), C,
                              H, _np(nears), _np(fars), noises=_np(noises), dt_gamma=dt_gamma, max_steps=max_steps,
                              M=xyzs.shape[0])
    _put(xyzs, x), _put(dirs, d), _put(deltas, dl)


def composite_rays(n_alive, n_step, T_thresh, channel_dim, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                   image):
    ra, rt = _np(rays_alive).copy(), _np(rays_t).copy()
    ws, dp, im = _np(weights_sum).copy(), _np(depth).copy(), _np(image).reshape(-1, channel_dim).copy()
    orc.composite_rays(n_alive, n_step, ra, rt, _np(sigmas), _np(rgbs).reshape(-1, channel_dim), _np(deltas), ws, dp, im,
                       T_thresh)
    _put(rays_alive, ra), _put(rays_t, rt), _put(weights_sum, ws), _put(depth, dp), _put(image, im)


# which positional arguments a function writes (raymarching.h: the non-const tensors)
OUTPUTS = {
    "near_far_from_aabb": [5, 6], "sph_from_ray": [4], "morton3D": [2], "morton3D_invert": [2], "packbits": [3],
    "march_rays_train": [12, 13, 14, 15, 16], "composite_rays_train_forward": [8, 9, 10],
    "composite_rays_train_backward": [12, 13], "march_rays": [14, 15, 16],
    "composite_rays": [4, 5, 9, 10, 11],
}

store, trace = {}, []


def keep(arr):
    arr = np.ascontiguousarray(arr)
    key = "a" + hashlib.sha1(arr.tobytes() + str(arr.dtype).encode() + str(arr.shape).encode()).hexdigest()[:14]
    store.setdefault(key, arr.copy())
    return key


def recorded(name, fn):
    def call(*args):
        rec = {"fn": name, "args": [], "outs": {}}
        for a in args:
            if torch.is_tensor(a):
                rec["args"].append({"t": keep(_np(a))})
            elif isinstance(a, bool):
                rec["args"].append({"b": a})
            elif isinstance(a, int):
                rec["args"].append({"i": a})
            else:
                rec["args"].append({"f": float(a)})
        fn(*args)
        for k in OUTPUTS[name]:
            rec["outs"][str(k)] = keep(_np(args[k]))
        trace.append(rec)
    return call


backend = types.ModuleType("_raymarching")
for _name in OUTPUTS:
    setattr(backend, _name, recorded(_name, globals()[_name]))
sys.modules["_raymarching"] = backend

import submodules.raymarching as ref_rm  # noqa: E402  (the unmodified reference wrapper, bound to the stand-in above)
import nerf.renderer as ref_renderer  # noqa: E402      (the unmodified reference renderer)

assert ref_renderer.raymarching is ref_rm


class Field(ref_renderer.NeRFRenderer):
    def forward(self, x, d):
        return analytic_field(x, d, self.channel_dim)


def mark(label):
    trace.append({"fn": "#", "label": label})


results, inputs = {}, {}
totals = {}
for name, SC in SCENES.items():
    inp = scene_inputs(name)
    inputs[name] = inp
    kw = dict(bg_color=SC["bg_color"], max_steps=SC["max_steps"], T_thresh=SC["T_thresh"], dt_gamma=SC["dt_gamma"])

    def model(cls=Field):
        mm = cls(bound=SC["bound"], channel_dim=SC["channel_dim"], density_scale=SC["density_scale"])
        mm.density_bitfield.copy_(torch.from_numpy(inp["bitfield"]))
        return mm

    m = model()
    R = lambda k, v: results.__setitem__(f"{name}_{k}", v)  # noqa: E731

    # ---- training, first-epoch path (mean_count = 0)
    m.train()
    mark(f"{name}:train_first_epoch")
    o, d = torch.from_numpy(inp["train_o"])[None], torch.from_numpy(inp["train_d"])[None]
    out = m.run_cuda(o, d, **kw)
    R("train_image", _np(out["image"])), R("train_depth", _np(out["depth"])), R("train_weights_sum", _np(out["weights_sum"]))
    R("train_counter", _np(m.step_counter[0]))
    total = int(_np(m.step_counter[0])[0])
    totals[name] = total

    # ---- its backward through composite_rays_train (the wrapper's autograd Function): d sum(image * w)
    mark(f"{name}:train_backward")
    probe = {}

    class FieldGrad(Field):
        def forward(self, x, dd):
            s, c = analytic_field(x, dd, self.channel_dim)
            s, c = s.clone().requires_grad_(True), c.clone().requires_grad_(True)
            probe["s"], probe["c"] = s, c
            return s, c

    mg = model(FieldGrad)
    mg.train()
    out = mg.run_cuda(o, d, **kw)
    w_img = torch.from_numpy(inp["loss_weights"]).view_as(out["image"])
    (out["image"] * w_img).sum().backward()
    R("train_grad_sigmas", _np(probe["s"].grad)), R("train_grad_rgbs", _np(probe["c"].grad))

    # ---- training with an under-estimated mean_count: rays whose samples do not fit are dropped (raymarching.cu:417)
    mark(f"{name}:train_mean_count")
    m.mean_count = max(total * 3 // 4, 1)
    out = m.run_cuda(o, d, **kw)
    R("train_mc_image", _np(out["image"])), R("train_mc_depth", _np(out["depth"]))
    m.mean_count = 0

    # ---- inference loop, every iteration
    mark(f"{name}:eval")
    m.eval()
    eo, ed = torch.from_numpy(inp["eval_o"])[None], torch.from_numpy(inp["eval_d"])[None]
    kw_eval = dict(kw, T_thresh=SC["T_thresh_eval"])
    with torch.no_grad():
        out = m.run_cuda(eo, ed, **kw_eval)
    R("eval_image", _np(out["image"])), R("eval_depth", _np(out["depth"]))

    # ---- perturbed runs (the wrapper draws the noises with torch.rand: they are recorded as inputs of the backend calls,
    # so only the call-by-call replay covers these)
    if SC["perturbed"]:
        torch.manual_seed(1234)
        mark(f"{name}:train_perturbed")
        m.train()
        m.run_cuda(o, d, perturb=True, **kw)
        mark(f"{name}:eval_perturbed")
        m.eval()
        with torch.no_grad():
            m.run_cuda(eo, ed, perturb=True, **kw_eval)

# ---- the small operators through the unmodified wrapper
mark("utils")
inp = inputs["s1"]
coords = torch.from_numpy(inp["coords"])
idx = ref_rm.morton3D(coords)
back = ref_rm.morton3D_invert(idx)
assert torch.equal(back.int(), coords.int())
grid = torch.from_numpy(inp["grid_values"])
ref_rm.packbits(grid, 0.5)
ref_rm.sph_from_ray(torch.from_numpy(inp["train_o"]), torch.from_numpy(inp["train_d"]), 2.0)

n_calls = sum(1 for r in trace if r["fn"] != "#")
by_fn = {}
for r in trace:
    by_fn[r["fn"]] = by_fn.get(r["fn"], 0) + 1
print("backend calls:", n_calls, by_fn)
print("train samples per scene", totals)
out_npz = dict(store)
out_npz["trace_json"] = np.frombuffer(json.dumps(trace).encode(), np.uint8)
for k, v in results.items():
    out_npz["result_" + k] = v
for name, inp in inputs.items():
    for k in ("bitfield", "train_o", "train_d", "eval_o", "eval_d", "loss_weights"):
        out_npz[f"input_{name}_{k}"] = inp[k]
path = os.path.join(HERE, "backend_trace.npz")
np.savez_compressed(path, **out_npz)
print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")
