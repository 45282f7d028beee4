"""CPU checks of the oracle itself: against the committed golden vectors produced by the UNMODIFIED reference kernels
(tests/golden/*.npz, generated on a B200 by tests/golden/make_golden.py), against independent numpy restatements and
through size-independent properties."""
import os

import numpy as np
import pytest

from scenarios import SCENARIOS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(sc):
    path = os.path.join(GOLDEN_DIR, sc.name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet")
    return np.load(path)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and (a.view(np.uint8) == b.view(np.uint8)).all()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_march_bit_exact_vs_reference_golden(sc):
    from oracle import oracle as orc
    g = golden(sc)
    inp = sc.inputs()
    nears, fars = orc.near_far_from_aabb(inp["rays_o"], inp["rays_d"], inp["aabb"], sc.min_near)
    assert bits_equal(nears, g["nears"]) and bits_equal(fars, g["fars"])
    x, d, dl, rays, counter = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"], sc.cascades,
                                                   sc.H, nears, fars, inp["noises"], sc.dt_gamma, sc.max_steps)
    assert bits_equal(counter, g["counter"])
    assert bits_equal(rays[:, 2], g["counts"]), "per-ray sample counts"
    assert bits_equal(x, g["xyzs"]) and bits_equal(d, g["dirs"]) and bits_equal(dl, g["deltas"])


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_composite_and_inference_vs_reference_golden(sc):
    from oracle import oracle as orc
    g = golden(sc)
    N, counts = sc.n_rays, g["counts"]
    total = int(counts.sum())
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.arange(N, dtype=np.int32), offsets, counts], -1)
    sig, rgb = sc.sample_values(total)
    g_ws, g_img = sc.upstream(N)
    ws, depth, image = orc.composite_rays_train_forward(sig, rgb, g["deltas"], rays, sc.t_thresh)
    assert rel_err(ws, g["comp_ws"]) <= 1e-4 and rel_err(depth, g["comp_depth"]) <= 1e-4
    assert rel_err(image, g["comp_image"]) <= 1e-4
    gs, gr = orc.composite_rays_train_backward(g_ws, g_img, sig, rgb, g["deltas"], rays, g["comp_ws"], g["comp_image"],
                                               sc.t_thresh)
    assert rel_err(gs, g["comp_gs"]) <= 2e-4 and rel_err(gr, g["comp_gr"]) <= 1e-4
    # inference iterations
    inp = sc.inputs()
    alive = np.arange(N, dtype=np.int32)
    rays_t = g["nears"].copy()
    ws_i, dep_i, img_i = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, sc.channels), np.float32)
    for it in range(2):
        if f"inf{it}_xyzs" not in g:
            break
        n_alive = alive.shape[0]
        assert np.array_equal(alive, g[f"inf{it}_alive_in"])
        noises = inp["noises"][:n_alive] if it == 0 else np.zeros(n_alive, np.float32)
        x, d, dl = orc.march_rays(n_alive, 4, alive, rays_t, inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"],
                                  sc.cascades, sc.H, g["nears"], g["fars"], noises, sc.dt_gamma, sc.max_steps)
        assert bits_equal(x, g[f"inf{it}_xyzs"]) and bits_equal(d, g[f"inf{it}_dirs"]) and bits_equal(dl, g[f"inf{it}_deltas"])
        orc.composite_rays(n_alive, 4, alive, rays_t, g[f"inf{it}_sigmas"], g[f"inf{it}_rgbs"], dl, ws_i, dep_i, img_i, 1e-2)
        assert np.array_equal(alive, g[f"inf{it}_alive_out"])
        assert bits_equal(rays_t, g[f"inf{it}_rays_t"])
        assert rel_err(ws_i, g[f"inf{it}_ws"]) <= 1e-4 and rel_err(img_i, g[f"inf{it}_image"]) <= 1e-4
        assert rel_err(dep_i, g[f"inf{it}_depth"]) <= 1e-4
        alive = orc.compact_rays(alive)


def test_small_utils_vs_reference_golden():
    from oracle import oracle as orc
    sc = SCENARIOS[0]
    g = golden(sc)
    assert bits_equal(orc.packbits(g["pack_in"], 0.01), g["pack_out"])
    assert bits_equal(orc.morton3D(g["morton_in"]), g["morton_out"])
    assert bits_equal(orc.morton3D_invert(g["morton_out"]), g["morton_back"])
    inp = sc.inputs()
    assert rel_err(orc.sph_from_ray(inp["rays_o"], inp["rays_d"], 4.0), g["sph"]) <= 1e-5


# ------------------------------------------------------------------ independent restatements / properties (no goldens)

def test_morton_and_packbits_against_numpy():
    from oracle import oracle as orc
    from stable_nerf_b200 import synthetic as syn
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (4096, 3)).astype(np.int32)
    idx = orc.morton3D(c)
    ref = np.zeros(4096, np.int64)
    for b in range(10):
        for k in range(3):
            ref |= ((c[:, k].astype(np.int64) >> b) & 1) << (3 * b + k)
    assert np.array_equal(idx.astype(np.int64) & 0xffffffff, ref & 0xffffffff)
    assert np.array_equal(orc.morton3D_invert(idx), c)
    grid = rng.random(8 * 1000).astype(np.float32)
    assert np.array_equal(orc.packbits(grid, 0.5), syn.pack_bitfield(grid, 0.5))
    assert orc.packbits(np.full(8, 0.5, np.float32), 0.5)[0] == 0  # strict '>'


def test_near_far_edge_cases():
    from oracle import oracle as orc
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    o = np.array([[0, 0, 3], [0, 0, 3], [3, 3, 3], [0, 0, 0.5], [0, 0, 3]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0.2, 0, -1]], np.float32)
    n, f = orc.near_far_from_aabb(o, d, aabb, 0.2)
    assert n[0] == 2 and f[0] == 4                      # straight through the box
    assert n[2] == np.finfo(np.float32).max == f[2]     # miss -> FLT_MAX (raymarching.cu:123)
    assert n[3] == np.float32(0.2) and f[3] == 1.5      # origin inside: near clamped to min_near
    assert n[1] == np.float32(0.2) and f[1] == -2       # pointing away: far < near, no samples are produced
    assert n[4] == 2 and f[4] == 4                      # oblique hit, z slab decides


def test_march_properties():
    from oracle import oracle as orc
    sc = SCENARIOS[2]  # cascades, dt_gamma, perturb
    inp = sc.inputs()
    nears, fars = orc.near_far_from_aabb(inp["rays_o"], inp["rays_d"], inp["aabb"], sc.min_near)
    x, d, dl, rays, counter = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"], sc.cascades,
                                                   sc.H, nears, fars, inp["noises"], sc.dt_gamma, sc.max_steps)
    total = int(counter[0])
    assert total == rays[:, 2].sum() > 0 and counter[1] == sc.n_rays
    assert (rays[:, 2] <= sc.max_steps).all()
    assert (np.abs(x) <= sc.bound).all()
    dt_min, dt_max = 2 * np.sqrt(3) / sc.max_steps, 2 * np.sqrt(3) * 2 ** (sc.cascades - 1) / sc.H
    assert (dl[:, 0] >= dt_min * (1 - 1e-6)).all() and (dl[:, 0] <= dt_max * (1 + 1e-6)).all()
    assert (dl[:, 1] >= dl[:, 0] * (1 - 1e-5)).all(), "t advances by at least dt between consecutive samples"
    # every sample's direction equals its ray's direction; rays pointing away / missing have no samples
    owner = np.repeat(np.arange(sc.n_rays), rays[:, 2])
    assert np.array_equal(d, inp["rays_d"][owner])
    assert rays[2, 2] == 0 and rays[3, 2] == 0
    # M too small: overflowing rays keep (offset,count) but write nothing (raymarching.cu:417)
    M = total // 3
    x2, _, dl2, rays2, _ = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"], sc.cascades,
                                                sc.H, nears, fars, inp["noises"], sc.dt_gamma, sc.max_steps, M=M)
    assert np.array_equal(rays, rays2)
    fits = rays[:, 1] + rays[:, 2] <= M
    last = (rays[fits, 1] + rays[fits, 2]).max()
    assert bits_equal(x2[:last], x[:last]) and (x2[last:] == 0).all()
    # inference marching from t=near reproduces the first n_step training samples of each ray
    n_step = 4
    alive = np.arange(sc.n_rays, dtype=np.int32)
    xi, di, dli = orc.march_rays(sc.n_rays, n_step, alive, nears, inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"],
                                 sc.cascades, sc.H, nears, fars, inp["noises"], sc.dt_gamma, sc.max_steps)
    for n in range(sc.n_rays):
        k = min(n_step, rays[n, 2])
        assert bits_equal(xi[n * n_step:n * n_step + k], x[rays[n, 1]:rays[n, 1] + k])
        assert (dli[n * n_step + k:(n + 1) * n_step] == 0).all()


def numpy_composite(sig, rgb, dl, rays, T_thresh):
    N, C = rays.shape[0], rgb.shape[1]
    ws, depth, image = np.zeros(N), np.zeros(N), np.zeros((N, C))
    for n, off, cnt in rays:
        T, t = 1.0, 0.0
        for i in range(off, off + cnt):
            a = 1 - np.exp(-float(sig[i]) * float(dl[i, 0]))
            w = a * T
            image[n] += w * rgb[i]
            t += dl[i, 1]
            depth[n] += w * t
            ws[n] += w
            T *= 1 - a
            if T < T_thresh:
                break
    return ws, depth, image


def test_composite_against_float64_numpy_and_finite_differences():
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    counts = np.array([0, 5, 40, 1, 17, 64], np.int32)
    N, total = counts.shape[0], int(counts.sum())
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.arange(N, dtype=np.int32)[::-1].copy(), offsets, counts], -1)  # ids need not be sorted
    sig = (rng.random(total) * 8).astype(np.float32)
    rgb = rng.random((total, 4)).astype(np.float32)
    dl = np.stack([np.full(total, 0.05), rng.random(total) * 0.1 + 0.05], -1).astype(np.float32)
    ws, depth, image = orc.composite_rays_train_forward(sig, rgb, dl, rays, 1e-4)
    rws, rdepth, rimage = numpy_composite(sig, rgb, dl, rays, 1e-4)
    assert rel_err(ws, rws) < 1e-5 and rel_err(depth, rdepth) < 1e-5 and rel_err(image, rimage) < 1e-5
    assert ws[rays[0, 0]] == 0 and (image[rays[0, 0]] == 0).all()  # empty ray
    # backward == d/d(sigma, rgb) of  sum(g_ws*ws) + sum(g_img*image)   (float64 finite differences)
    g_ws, g_img = rng.standard_normal(N).astype(np.float32), rng.standard_normal((N, 4)).astype(np.float32)
    gs, gr = orc.composite_rays_train_backward(g_ws, g_img, sig, rgb, dl, rays, ws, image, 1e-4)

    def L(s, c):
        a, _, b = numpy_composite(s, c, dl, rays, 0.0)
        return (g_ws * a).sum() + (g_img * b).sum()
    for i in rng.integers(0, total, 12):
        e = np.zeros(total); e[i] = 1e-4
        num = (L(sig + e, rgb) - L(sig - e, rgb)) / 2e-4
        assert abs(num - gs[i]) <= 2e-3 * max(1.0, abs(num)), (i, num, gs[i])
        e2 = np.zeros((total, 4)); e2[i, 1] = 1e-3
        num = (L(sig, rgb + e2) - L(sig, rgb - e2)) / 2e-3
        assert abs(num - gr[i, 1]) <= 1e-3 * max(1.0, abs(num))


@pytest.mark.parametrize("color_in_pad", [1.0, 0.0])
def test_field_oracle_against_torch_float64(color_in_pad):
    """hash grid + SH + MLP restated independently with torch float64 ops; fp32 oracle must agree to 1e-5.
    color_in_pad: the colour net's 32nd input -- 1.0 (tiny-cuda-nn's Identity-encoding padding, the default) and 0.0
    (a manual zero pad, nerf/network.py:54); with 1.0 first-layer column 31 acts as a bias and receives a gradient."""
    import torch
    from oracle import oracle as orc
    from stable_nerf_b200 import synthetic as syn
    from stable_nerf_b200.config import BaseNeRFConfig
    from stable_nerf_b200.field import make_field_desc, mlp_layer_shapes
    f = make_field_desc(BaseNeRFConfig().as_dict(), 3, 15, 1.0, color_in_pad=color_in_pad)
    of = orc.copy_desc(f, orc.FieldDesc)
    assert of.color_in_pad == color_in_pad
    ss, sc_ = mlp_layer_shapes(32, 128, 3), mlp_layer_shapes(32, 128, 4)
    ws, table, wc = syn.field_params(38912, f.grid.n_entries * 2, 55296, shapes_sigma=ss, shapes_color=sc_, table_scale=1.0)
    rng = np.random.default_rng(0)
    M = 300
    x = rng.uniform(-1, 1, (M, 3)).astype(np.float32)
    d = rng.standard_normal((M, 3)); d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    sig, rgb, geo = orc.field_forward(of, x, d, table, ws, wc, want_geo=True)

    # positions are fp32 by definition of the path: x01 and pos = fma(x01, scale, 0.5) are rounded to fp32 exactly
    # as the oracle/kernels do (at the finest level one fp32 ulp of pos is 6e-5 of a cell); the rest is float64
    x01 = torch.from_numpy(((x + np.float32(1.0)) / np.float32(2.0)).astype(np.float32)).double()
    tab = torch.from_numpy(table).double().view(-1, 2).requires_grad_(True)
    feats = []
    g = f.grid
    for l in range(16):
        pos = (x01 * float(np.float32(g.scale[l])) + 0.5).float().double()
        fl = torch.floor(pos)
        w = pos - fl
        c = fl.long()
        acc = 0
        for corner in range(8):
            b = [(corner >> k) & 1 for k in range(3)]
            wt = 1
            for k in range(3):
                wt = wt * (w[:, k] if b[k] else 1 - w[:, k])
            ix, iy, iz = [(c[:, k] + b[k]) for k in range(3)]
            if g.hashed[l]:
                idx = ((ix * 1) ^ ((iy * 2654435761) & 0xffffffff) ^ ((iz * 805459861) & 0xffffffff)) & 0xffffffff
            else:
                idx = ix + iy * g.resolution[l] + iz * g.resolution[l] ** 2
            idx = idx % g.size[l] + g.offset[l]
            acc = acc + wt[:, None] * tab[idx]
        feats.append(acc)
    enc = torch.cat(feats, -1)
    assert rel_err(orc.hashgrid_forward(of.grid, x01.float().numpy(), table), enc.detach().numpy()) < 1e-5

    def mlp(h, W, shapes):
        off = 0
        for i, (o, k) in enumerate(shapes):
            h = h @ W[off:off + o * k].view(o, k).T
            off += o * k
            if i < len(shapes) - 1:
                h = torch.relu(h)
        return h
    Ws = torch.from_numpy(ws).double().requires_grad_(True)
    Wc = torch.from_numpy(wc).double().requires_grad_(True)
    hs = mlp(enc, Ws, ss)
    sigma_t = torch.relu(hs[:, 0])
    sh = torch.from_numpy(orc.sh4_forward(((d + 1) / 2).astype(np.float32))).double()
    cin = torch.cat([sh, hs[:, 1:16], torch.full((M, 1), color_in_pad).double()], -1)
    rgb_t = torch.sigmoid(mlp(cin, Wc, sc_)[:, :3])
    assert rel_err(sig, sigma_t.detach().numpy()) < 1e-5 and rel_err(rgb, rgb_t.detach().numpy()) < 1e-5
    assert rel_err(geo, hs[:, 1:16].detach().numpy()) < 1e-5
    # SH-4 sanity: l=0 constant and orthogonality-free identity sum_m Y_1m^2 = 3/(4pi)
    assert np.allclose(sh[:, 0], 0.28209479177387814) and np.allclose((sh[:, 1:4] ** 2).sum(-1), 3 / (4 * np.pi), atol=1e-6)
    g_sig = rng.standard_normal(M).astype(np.float32)
    g_rgb = rng.standard_normal((M, 3)).astype(np.float32)
    ((sigma_t * torch.from_numpy(g_sig)).sum() + (rgb_t * torch.from_numpy(g_rgb)).sum()).backward()
    gt, gws, gwc = orc.field_backward(of, x, d, table, ws, wc, g_sig, g_rgb)
    assert rel_err(gws, Ws.grad.numpy()) < 1e-4 and rel_err(gwc, Wc.grad.numpy()) < 1e-4
    assert rel_err(gt, tab.grad.view(-1).numpy()) < 1e-4
    col31 = gwc[:128 * 32].reshape(128, 32)[:, 31]  # d loss / d (first-layer column 31): the bias when the pad is 1
    assert (np.abs(col31).max() > 0) == (color_in_pad != 0.0)
    # bf16 emulation stays within the stated bf16 tolerance of the fp32 result
    sig_b, rgb_b = orc.field_forward(of, x, d, table, ws, wc, emulate_bf16=True)
    assert rel_err(sig_b, sig) < 2e-2 and rel_err(rgb_b, rgb) < 2e-2


def test_trunc_exp_oracle():
    from oracle import oracle as orc
    x = np.array([-30, -15, 0, 1, 15, 30], np.float32)
    assert np.allclose(orc.trunc_exp_forward(x), np.exp(x), rtol=1e-6)
    assert np.allclose(orc.trunc_exp_backward(np.ones(6, np.float32), x), np.exp(np.clip(x, -15, 15)), rtol=1e-6)
