"""The CUDA path against the LIVE reference at BASELINE.json's sizes: oracle/_ref/_raymarching.so is the reference's
unmodified submodules/raymarching/src/raymarching.cu compiled for sm_100a (oracle/build_ref.sh), run here side by side
with libsnerf_b200 on the same inputs (the committed goldens are 96-192 rays; these are the full configurations):

  cfg2  4096 rays of an 800x800 view, channel_dim 3, max_steps 1024
  cfg4  two 64x64 views = 8192 rays, channel_dim 4, max_steps 256, focal 3.058 (the reference's degenerate intrinsics, Q13)
  cfg3  800x800 = 640 000 rays, the first three iterations of the inference loop (nerf/renderer.py:136-162)

Bars: near/far, per-ray sample counts, xyzs/dirs/deltas, rays_t and the alive lists bit-exact (the reference's sample
offsets come from atomicAdd and are canonicalised to ray order); compositing sums and gradients <= 1e-4
relative (max-norm).  Skipped when the reference extension was not built (oracle/_ref is git-ignored; build() makes it
where /root/reference exists and it travels to the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    path = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(path, "_raymarching.so")):
        pytest.skip("oracle/_ref/_raymarching.so not built (needs /root/reference at build time)")
    sys.path.insert(0, path)
    try:
        import _raymarching
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"reference extension not importable: {e!r}")
    return _raymarching


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def scene(cuda):
    from stable_nerf_b200 import synthetic as syn
    bitfield = torch.from_numpy(syn.pack_bitfield(syn.occupancy_grid(lego_like=True, seed=0))).to(cuda)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=cuda)
    return bitfield, aabb


def rays_for(cfg):
    from stable_nerf_b200 import synthetic as syn
    if cfg == "cfg2":
        ro, rd = syn.train_batch(4096, seed=0)
        return ro, rd, 3, 1024
    if cfg == "cfg4":
        pix = np.arange(64 * 64)
        od = [syn.rays_from_pixels(p, 3.058, 3.058, 32.0, 32.0, pix % 64, pix // 64) for p in syn.orbit_poses(2, seed=4)]
        return (np.concatenate([o for o, _ in od]).astype(np.float32), np.concatenate([d for _, d in od]).astype(np.float32),
                4, 256)
    ro, rd = syn.full_frame()
    return ro, rd, 3, 1024


@pytest.mark.parametrize("cfg", ["cfg2", "cfg4"])
def test_training_path_against_the_live_reference(cfg, ref, built_lib, cuda):
    from stable_nerf_b200 import _lib
    lib, P, S, chk = built_lib, _lib.ptr, _lib.stream(), _lib.check
    bitfield, aabb = scene(cuda)
    ro, rd, C, max_steps = rays_for(cfg)
    N = ro.shape[0]
    rays_o, rays_d = torch.from_numpy(ro).to(cuda), torch.from_numpy(rd).to(cuda)
    # ---- near/far
    n_r, f_r = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    n_o, f_o = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    ref.near_far_from_aabb(rays_o, rays_d, aabb, N, 0.2, n_r, f_r)
    chk(lib.snerf_near_far_from_aabb(P(rays_o), P(rays_d), P(aabb), N, 0.2, P(n_o), P(f_o), S), "near_far")
    assert torch.equal(n_r, n_o) and torch.equal(f_r, f_o)
    # ---- training march, perturbed (noise per ray), reference wrapper semantics: N*max_steps rows
    noises = torch.rand(N, generator=torch.Generator().manual_seed(3)).to(cuda)
    Mr = N * max_steps
    xr, dr, dlr = torch.zeros(Mr, 3, device=cuda), torch.zeros(Mr, 3, device=cuda), torch.zeros(Mr, 2, device=cuda)
    rays_r = torch.empty(N, 3, dtype=torch.int32, device=cuda)
    counter_r = torch.zeros(2, dtype=torch.int32, device=cuda)
    ref.march_rays_train(rays_o, rays_d, bitfield, 1.0, 0.0, max_steps, N, 1, 128, Mr, n_r, f_r, xr, dr, dlr, rays_r, counter_r, noises)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    ws_bytes = lib.snerf_march_rays_train_workspace_bytes_ex(N, max_steps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    geom = (1.0, 0.0, max_steps, N, 1, 128)
    chk(lib.snerf_march_rays_train_count(P(rays_o), P(rays_d), P(bitfield), *geom, P(n_o), P(f_o), P(counter), P(noises), P(ws),
                                         ws_bytes, S), "count")
    total = int(counter[0].item())
    assert total == int(counter_r[0].item()) and int(counter[1].item()) == int(counter_r[1].item()) == N
    M = (total + 127) // 128 * 128
    xo, do_, dlo = torch.empty(M, 3, device=cuda), torch.empty(M, 3, device=cuda), torch.empty(M, 2, device=cuda)
    rays = torch.empty(N, 3, dtype=torch.int32, device=cuda)
    n_samples = torch.empty(1, dtype=torch.int32, device=cuda)
    chk(lib.snerf_march_rays_train_write(P(rays_o), P(rays_d), P(bitfield), *geom, M, P(n_o), P(f_o), P(xo), P(do_), P(dlo), P(rays),
                                         P(noises), 1, P(n_samples), P(ws), ws_bytes, S), "write")
    torch.cuda.synchronize()
    # canonical comparison: reference rows sorted by ray id, samples of ray r at its own offset
    order = torch.argsort(rays_r[:, 0].long())
    rr = rays_r[order].long()
    assert torch.equal(rr[:, 0], torch.arange(N, device=cuda)) and torch.equal(rays[:, 0].long(), rr[:, 0])
    assert torch.equal(rr[:, 2], rays[:, 2].long()), "per-ray sample counts"
    counts = rays[:, 2].long()
    assert int(counts.sum()) == total and total > 0  # (cfg4 with the degenerate focal: almost every ray misses the box)
    ray_of = torch.repeat_interleave(torch.arange(N, device=cuda), counts)
    local = torch.arange(total, device=cuda) - rays[:, 1].long()[ray_of]
    idx_r = rr[:, 1][ray_of] + local
    assert torch.equal(xr[idx_r], xo[:total]) and torch.equal(dr[idx_r], do_[:total]) and torch.equal(dlr[idx_r], dlo[:total])
    assert float(xo[total:].abs().max() if M > total else 0) == 0.0
    # ---- compositing forward / backward on the canonical packing (both read the same rays table)
    g = torch.Generator().manual_seed(7)
    sig = (torch.rand(M, generator=g) ** 3 * 40.0).to(cuda)
    rgb = torch.rand(M, C, generator=g).to(cuda)
    out_r = [torch.empty(N, device=cuda), torch.empty(N, device=cuda), torch.empty(N, C, device=cuda)]
    out_o = [torch.empty_like(t) for t in out_r]
    ref.composite_rays_train_forward(sig, rgb, dlo, rays, M, N, 1e-4, C, *out_r)
    chk(lib.snerf_composite_rays_train_forward(P(sig), P(rgb), P(dlo), P(rays), M, N, 1e-4, C, *(P(t) for t in out_o), S), "cf")
    for a, b, name in zip(out_o, out_r, ("weights_sum", "depth", "image")):
        assert rel(a, b) <= 1e-4, name
    assert float(out_r[0].max()) > 0.9, "opaque rays exercise the early termination"
    g_ws, g_img = torch.randn(N, generator=g).to(cuda), torch.randn(N, C, generator=g).to(cuda)
    gs_r, gr_r = torch.zeros(M, device=cuda), torch.zeros(M, C, device=cuda)
    gs_o, gr_o = torch.full((M,), float("nan"), device=cuda), torch.full((M, C), float("nan"), device=cuda)
    ref.composite_rays_train_backward(g_ws, g_img, sig, rgb, dlo, rays, out_r[0], out_r[2], M, N, 1e-4, C, gs_r, gr_r)
    chk(lib.snerf_composite_rays_train_backward_ex(P(g_ws), P(g_img), P(sig), P(rgb), P(dlo), P(rays), P(out_r[0]), P(out_r[2]), M, N,
                                                   1e-4, C, P(gs_o), P(gr_o), P(n_samples), S), "cb")
    assert rel(gs_o, gs_r) <= 1e-4 and rel(gr_o, gr_r) <= 1e-4


def test_cfg3_inference_loop_against_the_live_reference(ref, built_lib, cuda):
    """800x800 frame, the first three iterations of the reference loop: march_rays -> (synthetic field values) ->
    composite_rays -> rays_alive[rays_alive >= 0], each side with its own state, compared after every step."""
    from stable_nerf_b200 import _lib, raymarching as rm
    lib, P, S, chk = built_lib, _lib.ptr, _lib.stream(), _lib.check
    bitfield, aabb = scene(cuda)
    ro, rd, C, max_steps = rays_for("cfg3")
    N = ro.shape[0]
    rays_o, rays_d = torch.from_numpy(ro).to(cuda), torch.from_numpy(rd).to(cuda)
    nears, fars = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
    ref.near_far_from_aabb(rays_o, rays_d, aabb, N, 0.2, nears, fars)
    st = {}
    for side in ("ref", "ours"):
        st[side] = dict(alive=torch.arange(N, dtype=torch.int32, device=cuda), t=nears.clone(), ws=torch.zeros(N, device=cuda),
                        depth=torch.zeros(N, device=cuda), image=torch.zeros(N, C, device=cuda))
    g = torch.Generator().manual_seed(5)
    step = 0
    for it in range(3):
        n_alive = st["ref"]["alive"].shape[0]
        assert n_alive == st["ours"]["alive"].shape[0] and n_alive > 0
        n_step = max(min(N // n_alive, 8), 1)
        Mi = n_alive * n_step
        Mi += 128 - Mi % 128  # the reference's align always adds (raymarching.py:331-332)
        noise = torch.rand(n_alive, generator=g).to(cuda) if it == 0 else torch.zeros(n_alive, device=cuda)
        r, o = st["ref"], st["ours"]
        xr, dr, dlr = torch.zeros(Mi, 3, device=cuda), torch.zeros(Mi, 3, device=cuda), torch.zeros(Mi, 2, device=cuda)
        ref.march_rays(n_alive, n_step, r["alive"], r["t"], rays_o, rays_d, 1.0, 0.0, max_steps, 1, 128, bitfield, nears, fars,
                       xr, dr, dlr, noise)
        xo, do_, dlo = (torch.full((Mi, k), float("nan"), device=cuda) for k in (3, 3, 2))
        chk(lib.snerf_march_rays_ex(n_alive, n_step, P(o["alive"]), P(o["t"]), P(rays_o), P(rays_d), 1.0, 0.0, max_steps, 1, 128,
                                    P(bitfield), P(nears), P(fars), P(xo), P(do_), P(dlo), P(noise), Mi, S), "march_rays")
        assert torch.equal(xr, xo) and torch.equal(dr, do_) and torch.equal(dlr, dlo), f"iteration {it}: samples"
        sig = (torch.rand(Mi, generator=g) ** 2 * 60.0).to(cuda)
        rgb = torch.rand(Mi, C, generator=g).to(cuda)
        ref.composite_rays(n_alive, n_step, 1e-4, C, r["alive"], r["t"], sig, rgb, dlr, r["ws"], r["depth"], r["image"])
        chk(lib.snerf_composite_rays(n_alive, n_step, 1e-4, C, P(o["alive"]), P(o["t"]), P(sig), P(rgb), P(dlo), P(o["ws"]),
                                     P(o["depth"]), P(o["image"]), S), "composite_rays")
        assert torch.equal(r["alive"], o["alive"]), f"iteration {it}: termination flags"
        assert torch.equal(r["t"], o["t"]), f"iteration {it}: rays_t"
        for k in ("ws", "depth", "image"):
            assert rel(o[k], r[k]) <= 1e-4, f"iteration {it}: {k}"
        r["alive"] = r["alive"][r["alive"] >= 0]                         # nerf/renderer.py:158
        spare, count = rm.compact_rays(o["alive"], n_alive)              # the device-side form of the same
        o["alive"] = spare[:int(count.item())].clone()
        assert torch.equal(r["alive"], o["alive"]), f"iteration {it}: alive list"
        step += n_step
    assert st["ref"]["alive"].shape[0] < N, "some rays terminated within three iterations"
