"""Seeded synthetic "Blender-shaped" workloads: occupancy grids, orbit cameras and rays.

Host-side numpy only.  Used by bench.py, __graft_entry__.smoke() and the tests so that every leg (CUDA path, CPU
oracle, reference kernels) sees byte-identical inputs.  Ray generation follows the reference's ``get_rays``
convention (utils/graphics_utils.py:22-24, :75-83: pixel centre +0.5, normalised directions, d = dir @ R^T) and
its orbit-camera construction (``rand_poses``, utils/graphics_utils.py:91-125).
"""
import math

import numpy as np

BLENDER_RADIUS = 4.0311 * 0.33          # camera distance after nerf_matrix_to_ngp scaling (graphics_utils.py:129)
BLENDER_FOCAL_800 = 0.5 * 800 / math.tan(0.5 * 0.6911112)  # camera_angle_x of the Blender scenes -> 1111.11 px


def _morton_invert(idx):
    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    idx = idx.astype(np.uint32)
    return np.stack([compact(idx), compact(idx >> 1), compact(idx >> 2)], axis=-1).astype(np.int64)


def occupancy_grid(H=128, cascades=1, bound=1.0, lego_like=False, seed=0):
    """Density grid f32 [cascades, H^3] in Morton order: 1.0 inside (sphere r=0.5) U (box .35x.15x.35), else 0.

    ``lego_like`` punches seeded Bernoulli(0.5) holes at 8^3-cell granularity.  ~7 % of cascade 0 is occupied.
    """
    idx = np.arange(H ** 3, dtype=np.uint32)
    coords = _morton_invert(idx)                                   # [H^3, 3]
    grid = np.zeros((cascades, H ** 3), dtype=np.float32)
    rng = np.random.default_rng(seed)
    for cas in range(cascades):
        b = min(2.0 ** cas, bound)
        centres = ((2 * coords + 1) / H - 1.0) * b                 # cell centres in world units
        sphere = (centres ** 2).sum(-1) < 0.5 ** 2
        box = (np.abs(centres[:, 0]) < 0.35) & (np.abs(centres[:, 1]) < 0.15) & (np.abs(centres[:, 2]) < 0.35)
        occ = sphere | box
        if lego_like:
            holes = rng.random((H // 8, H // 8, H // 8)) < 0.5
            blk = coords // 8
            occ &= ~holes[blk[:, 0], blk[:, 1], blk[:, 2]]
        grid[cas] = occ.astype(np.float32)
    return grid


def pack_bitfield(grid, thresh=0.01):
    """uint8 bitfield, bit i of byte n = grid.flat[8n+i] > thresh (raymarching.cu:268-301)."""
    bits = (grid.reshape(-1) > thresh)
    return np.packbits(bits.reshape(-1, 8), axis=-1, bitorder="little").reshape(-1)


def orbit_poses(n, radius=BLENDER_RADIUS, elev_deg=(0.0, 60.0), seed=0):
    """[n,4,4] cam2world look-at-origin poses on an orbit (utils/graphics_utils.py:91-125 construction)."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(rng.uniform(elev_deg[0], elev_deg[1], n))
    phi = rng.uniform(0.0, 2 * np.pi, n)
    theta = np.pi / 2 - elev
    centers = np.stack([radius * np.sin(theta) * np.sin(phi), radius * np.cos(theta),
                        radius * np.sin(theta) * np.cos(phi)], -1)

    def normalize(v):
        return v / (np.linalg.norm(v, axis=-1, keepdims=True) + 1e-10)

    forward = -normalize(centers)
    up = np.tile(np.array([0.0, -1.0, 0.0]), (n, 1))
    right = normalize(np.cross(forward, up))
    up = normalize(np.cross(right, forward))
    poses = np.tile(np.eye(4), (n, 1, 1))
    poses[:, :3, :3] = np.stack([right, up, forward], axis=-1)
    poses[:, :3, 3] = centers
    return poses.astype(np.float32)


def rays_from_pixels(pose, fx, fy, cx, cy, px, py):
    """Rays through pixel centres (px+0.5, py+0.5) of one camera: float32 rays_o, rays_d [n,3]."""
    pose = pose.astype(np.float32)
    i = px.astype(np.float32) + np.float32(0.5)
    j = py.astype(np.float32) + np.float32(0.5)
    d = np.stack([(i - np.float32(cx)) / np.float32(fx), (j - np.float32(cy)) / np.float32(fy), np.ones_like(i)], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = (d @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, np.ascontiguousarray(rays_d)


def train_batch(n_rays, H=800, W=800, focal=BLENDER_FOCAL_800, n_views=1, seed=0, radius=BLENDER_RADIUS):
    """cfg2-style batch: ``n_rays`` random pixels spread over ``n_views`` orbit views of an HxW image."""
    rng = np.random.default_rng(seed + 1000)
    poses = orbit_poses(n_views, radius=radius, seed=seed)
    per = [n_rays // n_views + (1 if v < n_rays % n_views else 0) for v in range(n_views)]
    os_, ds_ = [], []
    for v, n in enumerate(per):
        inds = rng.integers(0, H * W, size=n)
        o, d = rays_from_pixels(poses[v], focal, focal, W / 2, H / 2, inds % W, inds // W)
        os_.append(o)
        ds_.append(d)
    return np.concatenate(os_), np.concatenate(ds_)


def full_frame(H=800, W=800, focal=BLENDER_FOCAL_800, seed=0, radius=BLENDER_RADIUS):
    """cfg3-style: every pixel of one orbit view, row-major."""
    pose = orbit_poses(1, radius=radius, seed=seed)[0]
    inds = np.arange(H * W)
    return rays_from_pixels(pose, focal, focal, W / 2, H / 2, inds % W, inds // W)


def field_params(n_sigma_mlp, n_table, n_color, seed=1337, table_scale=1e-4, shapes_sigma=None, shapes_color=None):
    """Deterministic numpy parameters: Xavier-uniform matrices, U(-table_scale, table_scale) table."""
    rng = np.random.default_rng(seed)

    def xavier(shapes, n):
        if shapes is None:
            return (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(0.1)
        out = np.empty(n, np.float32)
        off = 0
        for (o, i) in shapes:
            a = math.sqrt(6.0 / (i + o))
            out[off:off + o * i] = (rng.random(o * i, dtype=np.float32) * 2 - 1) * np.float32(a)
            off += o * i
        return out

    w_sigma = xavier(shapes_sigma, n_sigma_mlp)
    w_color = xavier(shapes_color, n_color)
    table = (rng.random(n_table, dtype=np.float32) * 2 - 1) * np.float32(table_scale)
    return w_sigma, table, w_color
