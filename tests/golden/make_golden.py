"""Generate golden vectors by running the UNMODIFIED reference kernels (oracle/_ref/_raymarching.so, built from
/root/reference/submodules/raymarching/src by oracle/build_ref.sh) on a CUDA device.

Run on a GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.  The reference's sample order is atomics-ordered
(raymarching.cu:406-407), so march outputs are canonicalised to ray order before they are stored.
Reads nothing under /root/reference at run time.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _raymarching as ref  # noqa: E402  the reference extension
from scenarios import SCENARIOS  # noqa: E402


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def canonical_march(rays, xyzs, dirs, deltas, N):
    """reference rays rows (id, offset, count) in atomic order -> per-ray-ordered arrays + counts."""
    rays = rays.cpu().numpy()
    order = np.argsort(rays[:, 0], kind="stable")
    rays = rays[order]
    assert (rays[:, 0] == np.arange(N)).all()
    counts = rays[:, 2].astype(np.int64)
    idx = np.concatenate([np.arange(o, o + c) for o, c in zip(rays[:, 1], counts)]) if counts.sum() else np.zeros(0, np.int64)
    x, d, dl = xyzs.cpu().numpy()[idx], dirs.cpu().numpy()[idx], deltas.cpu().numpy()[idx]
    return counts.astype(np.int32), x, d, dl


def run(sc):
    inp = sc.inputs()
    N = sc.n_rays
    rays_o, rays_d = T(inp["rays_o"]), T(inp["rays_d"])
    bitfield, aabb, noises = T(inp["bitfield"]), T(inp["aabb"]), T(inp["noises"])
    out = {}
    nears = torch.empty(N, device="cuda")
    fars = torch.empty(N, device="cuda")
    ref.near_far_from_aabb(rays_o, rays_d, aabb, N, sc.min_near, nears, fars)
    out["nears"], out["fars"] = nears.cpu().numpy(), fars.cpu().numpy()

    # --- training march (reference wrapper semantics: M = N*max_steps zero-filled, raymarching.py:193-218)
    M = N * sc.max_steps
    xyzs = torch.zeros(M, 3, device="cuda")
    dirs = torch.zeros(M, 3, device="cuda")
    deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ref.march_rays_train(rays_o, rays_d, bitfield, sc.bound, sc.dt_gamma, sc.max_steps, N, sc.cascades, sc.H, M, nears,
                         fars, xyzs, dirs, deltas, rays, counter, noises)
    torch.cuda.synchronize()
    out["counter"] = counter.cpu().numpy()
    counts, x, d, dl = canonical_march(rays, xyzs, dirs, deltas, N)
    out["counts"], out["xyzs"], out["dirs"], out["deltas"] = counts, x, d, dl
    total = int(counts.sum())
    assert total == int(out["counter"][0])

    # --- compositing on the canonical packing
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    crays = np.stack([np.arange(N, dtype=np.int32), offsets, counts], -1)
    sig, rgb = sc.sample_values(total)
    g_ws, g_img = sc.upstream(N)
    t_sig, t_rgb, t_dl, t_rays = T(sig), T(rgb), T(dl), T(crays)
    ws = torch.empty(N, device="cuda")
    depth = torch.empty(N, device="cuda")
    image = torch.empty(N, sc.channels, device="cuda")
    ref.composite_rays_train_forward(t_sig, t_rgb, t_dl, t_rays, total, N, sc.t_thresh, sc.channels, ws, depth, image)
    gs = torch.zeros(total, device="cuda")
    gr = torch.zeros(total, sc.channels, device="cuda")
    ref.composite_rays_train_backward(T(g_ws), T(g_img), t_sig, t_rgb, t_dl, t_rays, ws, image, total, N, sc.t_thresh,
                                      sc.channels, gs, gr)
    out["comp_ws"], out["comp_depth"], out["comp_image"] = ws.cpu().numpy(), depth.cpu().numpy(), image.cpu().numpy()
    out["comp_gs"], out["comp_gr"] = gs.cpu().numpy(), gr.cpu().numpy()

    # --- inference: two iterations of march_rays / composite_rays with n_step = 4 (nerf/renderer.py:136-162)
    n_step = 4
    alive = torch.arange(N, dtype=torch.int32, device="cuda")
    rays_t = nears.clone()
    iws = torch.zeros(N, device="cuda")
    idepth = torch.zeros(N, device="cuda")
    iimage = torch.zeros(N, sc.channels, device="cuda")
    rng = np.random.default_rng(sc.seed + 17)
    for it in range(2):
        n_alive = alive.shape[0]
        Mi = n_alive * n_step
        ix = torch.zeros(Mi, 3, device="cuda")
        idr = torch.zeros(Mi, 3, device="cuda")
        idl = torch.zeros(Mi, 2, device="cuda")
        inoise = noises[:n_alive].contiguous() if it == 0 else torch.zeros(n_alive, device="cuda")
        ref.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, sc.bound, sc.dt_gamma, sc.max_steps, sc.cascades,
                       sc.H, bitfield, nears, fars, ix, idr, idl, inoise)
        isig = (rng.random(Mi, dtype=np.float32) ** 3 * 200.0).astype(np.float32)
        irgb = rng.random((Mi, sc.channels), dtype=np.float32)
        out[f"inf{it}_alive_in"] = alive.cpu().numpy()
        out[f"inf{it}_xyzs"], out[f"inf{it}_dirs"], out[f"inf{it}_deltas"] = ix.cpu().numpy(), idr.cpu().numpy(), idl.cpu().numpy()
        out[f"inf{it}_sigmas"], out[f"inf{it}_rgbs"] = isig, irgb
        ref.composite_rays(n_alive, n_step, 1e-2, sc.channels, alive, rays_t, T(isig), T(irgb), idl, iws, idepth, iimage)
        out[f"inf{it}_alive_out"] = alive.cpu().numpy()
        out[f"inf{it}_rays_t"] = rays_t.cpu().numpy()
        out[f"inf{it}_ws"], out[f"inf{it}_depth"], out[f"inf{it}_image"] = iws.cpu().numpy(), idepth.cpu().numpy(), iimage.cpu().numpy()
        alive = alive[alive >= 0]  # nerf/renderer.py:158
        if alive.shape[0] == 0:
            break

    # --- small utilities
    rng = np.random.default_rng(sc.seed + 19)
    g = (rng.random(8 * 4096, dtype=np.float32) * 0.02).astype(np.float32)
    bits = torch.empty(4096, dtype=torch.uint8, device="cuda")
    ref.packbits(T(g), 4096, 0.01, bits)
    out["pack_in"], out["pack_out"] = g, bits.cpu().numpy()
    coords = rng.integers(0, 128, size=(2048, 3)).astype(np.int32)
    ind = torch.empty(2048, dtype=torch.int32, device="cuda")
    ref.morton3D(T(coords), 2048, ind)
    back = torch.empty(2048, 3, dtype=torch.int32, device="cuda")
    ref.morton3D_invert(ind, 2048, back)
    out["morton_in"], out["morton_out"], out["morton_back"] = coords, ind.cpu().numpy(), back.cpu().numpy()
    sph = torch.empty(N, 2, device="cuda")
    ref.sph_from_ray(rays_o, rays_d, 4.0, N, sph)
    out["sph"] = sph.cpu().numpy()
    torch.cuda.synchronize()
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    for sc in SCENARIOS:
        out = run(sc)
        path = os.path.join(outdir, sc.name + ".npz")
        np.savez_compressed(path, **out)
        print(f"[golden] {sc.name}: N={sc.n_rays} samples={int(out['counter'][0])} -> {path} "
              f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
