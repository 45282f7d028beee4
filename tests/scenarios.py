"""Seeded parity scenarios shared by the golden-vector generator (tests/golden/make_golden.py, runs the UNMODIFIED
reference kernels on a GPU box), the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np

from stable_nerf_b200 import synthetic as syn


class Scenario:
    def __init__(self, name, n_rays, bound, cascades, max_steps, dt_gamma, perturb, lego, channels, img=64, seed=0,
                 t_thresh=1e-4, radius=None):
        self.name, self.n_rays, self.bound, self.cascades = name, n_rays, float(bound), cascades
        self.max_steps, self.dt_gamma, self.perturb, self.lego = max_steps, float(dt_gamma), perturb, lego
        self.channels, self.img, self.seed, self.t_thresh = channels, img, seed, float(t_thresh)
        self.radius = syn.BLENDER_RADIUS if radius is None else radius
        self.H = 128
        self.min_near = 0.2

    def inputs(self):
        """dict of numpy inputs, fully determined by the scenario's seeds."""
        grid = syn.occupancy_grid(self.H, self.cascades, self.bound, lego_like=self.lego, seed=self.seed)
        bitfield = syn.pack_bitfield(grid, 0.01)
        focal = syn.BLENDER_FOCAL_800 * self.img / 800.0
        rays_o, rays_d = syn.train_batch(self.n_rays, self.img, self.img, focal, n_views=3, seed=self.seed,
                                         radius=self.radius)
        rng = np.random.default_rng(self.seed + 7)
        if self.n_rays >= 8:  # a few hostile rays: axis-parallel (zero components -> inf reciprocals) and a miss
            rays_d[0] = np.array([0.0, 0.0, -1.0], np.float32)
            rays_o[0] = np.array([0.05, 0.02, 1.3], np.float32)
            rays_d[1] = np.array([1.0, 0.0, 0.0], np.float32)
            rays_o[1] = np.array([-1.4, 0.1, -0.2], np.float32)
            rays_d[2] = np.array([0.0, 1.0, 0.0], np.float32)
            rays_o[2] = np.array([3.0, -2.0, 0.0], np.float32)  # misses the box
            rays_d[3] = -rays_d[3]                               # points away
        noises = rng.random(self.n_rays, dtype=np.float32) if self.perturb else np.zeros(self.n_rays, np.float32)
        aabb = np.array([-self.bound] * 3 + [self.bound] * 3, np.float32)
        return dict(grid=grid, bitfield=bitfield, rays_o=rays_o, rays_d=rays_d, noises=noises, aabb=aabb)

    def sample_values(self, M):
        """seeded sigma / rgb / upstream gradients for compositing parity."""
        rng = np.random.default_rng(self.seed + 11)
        sigmas = (rng.random(M, dtype=np.float32) ** 3 * 60.0).astype(np.float32)
        rgbs = rng.random((M, self.channels), dtype=np.float32)
        return sigmas, rgbs

    def upstream(self, N):
        rng = np.random.default_rng(self.seed + 13)
        g_ws = rng.standard_normal(N).astype(np.float32)
        g_img = rng.standard_normal((N, self.channels)).astype(np.float32)
        return g_ws, g_img


SCENARIOS = [
    Scenario("blender_c1", 192, 1, 1, 128, 0.0, False, False, 3),
    Scenario("lego_c1_perturb", 160, 1, 1, 256, 0.0, True, True, 4, t_thresh=1e-2),
    Scenario("bound2_c2_gamma", 160, 2, 2, 192, 1.0 / 128, True, True, 3, radius=2.2),
    Scenario("bound4_c3", 96, 4, 3, 64, 0.0, False, False, 1, radius=3.0),
    Scenario("bound1p5_c2", 96, 1.5, 2, 96, 0.0, False, True, 2, radius=1.8),
]


def by_name(name):
    return next(s for s in SCENARIOS if s.name == name)
