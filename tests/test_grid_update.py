"""Occupancy-grid maintenance against vectors produced by the reference's OWN python:
tests/golden/grid_update.npz holds what ``NeRFRenderer.mark_untrained_grid`` / ``update_extra_state``
(nerf/renderer.py:174-327, run unmodified on the CPU by tests/golden/make_golden_grid.py) leave in ``density_grid``,
``density_bitfield``, ``mean_density``, ``iter_density``, ``mean_count`` for an analytic density and seeded draws.

Bars: the -1 marks equal the reference's on every cell except the listed last-bit-ambiguous ones (17 + 8 of 6.3 M);
density_grid bit-exact (sha256 of the fp32 bytes); bitfield bit-exact; mean_density to 1e-6 relative (the reference sums
in fp32 with torch's reduction order, here block sums in double); iter_density / mean_count / local_step exact.
CPU tests pin the oracle's restatement; GPU tests pin the kernels through the drop-in NeRFRenderer."""
import hashlib
import os

import numpy as np
import pytest
import torch

from grid_stub import SCENARIOS, analytic_sigma, scenario_poses

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grid_update.npz")
H = 128
H3 = H ** 3


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def morton_lin():
    """linear index (x*H + y)*H + z of Morton cell i: the reference enumerates cells in meshgrid order
    (nerf/renderer.py:249-257), the kernels in Morton order; its noise row for cell i is noise[lin[i]]."""
    from oracle import oracle as orc
    c = orc.morton3D_invert(np.arange(H3, dtype=np.int32)).astype(np.int64)
    return (c[:, 0] * H + c[:, 1]) * H + c[:, 2]


class ReferenceDraws:
    """Replays the reference's draws from torch's CPU generator: per full-sweep cascade one rand_like([H^3,3]) (:266); per
    partial-update cascade randint(0,H,(N,3)), randint(0,n_occ,[N]), rand_like([2N,3]) (:280,:284,:300)."""

    def __init__(self, seed, lin):
        torch.manual_seed(seed)
        self.lin = torch.from_numpy(lin)

    def jitter(self, cells, first, n):
        if cells is None:  # full sweep, one chunk: Morton cells first..first+n of the cascade
            assert first == 0 and n == H3
            return torch.rand(H3, 3)[self.lin]
        return torch.rand(n, 3)

    def randint(self, high, shape):
        return torch.randint(0, high, tuple(shape))


def check_step(g, name, k, grid, bits, mean, state):
    assert abs(mean - float(g[f"{name}_step{k}_mean"])) <= 1e-6 * float(g[f"{name}_step{k}_mean"]), "mean_density"
    sample = np.asarray(grid, np.float32).reshape(-1)[::4099]
    assert np.array_equal(sample, g[f"{name}_step{k}_sample"]), \
        f"density_grid sample differs, max abs {np.abs(sample - g[f'{name}_step{k}_sample']).max()}"
    assert np.array_equal(sha(np.asarray(grid, np.float32)), g[f"{name}_step{k}_sha256"]), "density_grid (sha256 of the fp32 bytes)"
    assert np.array_equal(np.asarray(bits), g[f"{name}_step{k}_bits"]), "density_bitfield"
    assert list(state) == list(g[f"{name}_step{k}_state"]), "iter_density, mean_count, local_step"


def check_untrained(g, name, grid):
    got = np.asarray(grid).reshape(-1) == -1
    want = np.unpackbits(g[f"{name}_untrained_bits"], bitorder="little").astype(bool)
    diff = np.nonzero(got != want)[0]
    assert np.isin(diff, g[f"{name}_ambiguous"]).all(), f"{diff.size} cells differ outside the ambiguous set"
    return want


@pytest.mark.parametrize("sc", SCENARIOS, ids=[s["name"] for s in SCENARIOS])
def test_oracle_grid_maintenance_matches_reference_python(sc):
    from oracle import oracle as orc
    g = np.load(GOLD)
    name, bound = sc["name"], sc["bound"]
    C = 1 + int(np.ceil(np.log2(bound)))
    lin = morton_lin()
    grid = np.zeros((C, H3), np.float32)
    orc.mark_untrained_grid(scenario_poses(sc).numpy(), sc["intrinsic"], bound, C, H, grid)
    want = check_untrained(g, name, grid)
    grid[:] = np.where(want.reshape(C, H3), np.float32(-1), np.float32(0))  # continue from the reference's marks
    iter_density, mean_count, local_step = 0, 0, 0
    for k, step in enumerate(sc["steps"]):
        iter_density = step.get("iter_density", iter_density)
        draws = ReferenceDraws(step["seed"], lin)
        tmp = np.full((C, H3), -1, np.float32)
        for cas in range(C):
            if iter_density < 16:
                xyz = orc.grid_cell_points(None, 0, H3, cas, bound, H, draws.jitter(None, 0, H3).numpy())
                tmp[cas] = analytic_sigma(torch.from_numpy(xyz), sc["amp"], sc["r2"]).numpy()
            else:
                N = H3 // 4
                idx = orc.morton3D(draws.randint(H, (N, 3)).int().numpy()).astype(np.int64)
                occ = np.nonzero(grid[cas] > 0)[0]
                idx = np.concatenate([idx, occ[draws.randint(occ.shape[0], [N]).numpy()]])
                xyz = orc.grid_cell_points(idx.astype(np.int32), 0, idx.shape[0], cas, bound, H,
                                           draws.jitter(idx, 0, idx.shape[0]).numpy())
                tmp[cas, idx] = analytic_sigma(torch.from_numpy(xyz), sc["amp"], sc["r2"]).numpy()  # numpy: last wins
        flat = grid.reshape(-1)
        mean, thresh, bits = orc.grid_ema_update(flat, tmp, 0.95, 0.01, tmp_scale=sc["density_scale"])
        iter_density += 1
        if step.get("counts"):
            mean_count = int(sum(step["counts"]) / len(step["counts"]))
        check_step(g, name, k, grid, bits, mean, (iter_density, mean_count, 0))


def make_renderer(sc, dev, lin):
    from stable_nerf_b200.renderer import NeRFRenderer

    class Field(NeRFRenderer):
        draws = None

        def density(self, x):
            return {"sigma": analytic_sigma(x, sc["amp"], sc["r2"])}

        def _jitter_noise(self, cells, first, n, device):
            return self.draws.jitter(cells, first, n).to(device)

        def _randint(self, high, shape, device):
            return self.draws.randint(high, shape).to(device)

    return Field(bound=sc["bound"], density_scale=sc["density_scale"]).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("sc", SCENARIOS, ids=[s["name"] for s in SCENARIOS])
def test_renderer_grid_maintenance_matches_reference_python(sc, built_lib, cuda):
    """the drop-in NeRFRenderer (kernels of csrc/grid_update.cu) replays the reference run: same draws, same densities"""
    g = np.load(GOLD)
    name = sc["name"]
    lin = morton_lin()
    m = make_renderer(sc, cuda, lin)
    m.mark_untrained_grid(scenario_poses(sc), sc["intrinsic"])
    want = check_untrained(g, name, m.density_grid.cpu().numpy())
    m.density_grid.copy_(torch.from_numpy(np.where(want, np.float32(-1), np.float32(0)).reshape(m.density_grid.shape)))
    for k, step in enumerate(sc["steps"]):
        if step.get("iter_density") is not None:
            m.iter_density = step["iter_density"]
        if step.get("counts") is not None:
            m.step_counter.zero_()
            c = torch.tensor(step["counts"], dtype=torch.int32, device=cuda)
            m.step_counter[:c.shape[0], 0] = c
            m.local_step = int(c.shape[0])
        m.draws = ReferenceDraws(step["seed"], lin)
        m.update_extra_state()
        check_step(g, name, k, m.density_grid.cpu().numpy(), m.density_bitfield.cpu().numpy(), m.mean_density,
                   (m.iter_density, m.mean_count, m.local_step))


@pytest.mark.gpu
def test_grid_kernels_vs_oracle_and_device_noise(built_lib, cuda):
    """kernels against the oracle on random inputs (incl. negative / zero cells, a ragged last block is impossible: n is
    a multiple of 8) and the in-kernel counter-based jitter: uniform, inside the cell, different per seed."""
    from oracle import oracle as orc
    from stable_nerf_b200 import _lib
    lib, P, S = built_lib, _lib.ptr, _lib.stream()
    rng = np.random.default_rng(0)
    n = 8 * 4099
    grid = rng.uniform(-0.2, 1.0, n).astype(np.float32)
    grid[rng.random(n) < 0.2] = -1
    tmp = rng.uniform(-0.5, 2.0, n).astype(np.float32)
    g_o = grid.copy()
    mean_o, thresh_o, bits_o = orc.grid_ema_update(g_o, tmp, 0.95, 0.01, tmp_scale=0.5)
    tg, tt = torch.from_numpy(grid).to(cuda), torch.from_numpy(tmp).to(cuda)
    nb = lib.snerf_grid_ema_workspace_bytes(n)
    ws = torch.zeros(nb, dtype=torch.uint8, device=cuda)
    out = torch.empty(2, device=cuda)
    bits = torch.empty(n // 8, dtype=torch.uint8, device=cuda)
    for rep in range(2):  # the workspace is left ready for the next call
        tg.copy_(torch.from_numpy(grid))
        _lib.check(lib.snerf_grid_ema_update(P(tg), P(tt), n, 0.5, 0.95, 0.01, P(out), P(bits), P(ws), nb, S), "ema")
        assert np.array_equal(tg.cpu().numpy(), g_o) and np.array_equal(bits.cpu().numpy(), bits_o)
        assert abs(float(out[0]) - mean_o) <= 1e-6 * abs(mean_o) and float(out[1]) == pytest.approx(thresh_o, rel=1e-6)
    # cell points: explicit noise == oracle, bit for bit
    cells = rng.integers(0, H3, 5000).astype(np.int32)
    noise = rng.random((5000, 3), dtype=np.float32)
    for cas, bound in ((0, 1.0), (1, 2.0), (2, 3.5)):
        xyz_o = orc.grid_cell_points(cells, 0, 5000, cas, bound, H, noise)
        xyz = torch.empty(5000, 3, device=cuda)
        tc, tn = torch.from_numpy(cells).to(cuda), torch.from_numpy(noise).to(cuda)
        _lib.check(lib.snerf_grid_cell_points(P(tc), 0, 5000, cas, bound, H, P(tn), 0, P(xyz), S), "points")
        assert np.array_equal(xyz.cpu().numpy(), xyz_o)
    # in-kernel jitter: inside the cell, mean 0.5, differs between seeds, reproducible for one seed
    def dev_points(seed):
        xyz = torch.empty(H3, 3, device=cuda)
        _lib.check(lib.snerf_grid_cell_points(None, 0, H3, 0, 1.0, H, None, seed, P(xyz), S), "points")
        return xyz.cpu().numpy()
    a, b, a2 = dev_points(1), dev_points(2), dev_points(1)
    centre = orc.grid_cell_points(None, 0, H3, 0, 1.0, H, np.full((H3, 3), 0.5, np.float32))
    half = 1.0 / H
    u = (a - centre) / (2 * half) + 0.5
    assert np.array_equal(a, a2) and not np.array_equal(a, b)
    assert u.min() >= -1e-4 and u.max() <= 1 + 1e-4 and abs(u.mean() - 0.5) < 2e-3 and abs(u.std() - 12 ** -0.5) < 2e-3
    assert abs(np.corrcoef(u[:-1, 0], u[1:, 0])[0, 1]) < 5e-3 and abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 5e-3
    # mark_untrained_grid against the oracle on random poses: identical except where the oracle itself is one ulp from flipping
    poses = scenario_poses(SCENARIOS[0]).numpy()
    g1 = np.zeros((2, H3), np.float32)
    n_o = orc.mark_untrained_grid(poses, (40.0, 45.0, 30.0, 34.0), 2.0, 2, H, g1)
    tg = torch.zeros(2, H3, device=cuda)
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    tp = torch.from_numpy(poses).to(cuda)
    _lib.check(lib.snerf_mark_untrained_grid(P(tp), poses.shape[0], 30.0 / 40.0, 34.0 / 45.0, 2.0, 2, H, P(tg), P(cnt), S), "mark")
    assert np.array_equal(tg.cpu().numpy(), g1) and int(cnt.item()) == n_o
