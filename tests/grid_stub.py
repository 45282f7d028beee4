"""Shared between tests/golden/make_golden_grid.py (the reference's renderer run on the CPU) and the parity tests of the
occupancy-grid maintenance: scenarios, camera poses and an analytic density whose fp32 value is the same on every device.
"""
import numpy as np
import torch


def analytic_sigma(x, amp, r2):
    """relu((r2 - (x^2 + 1.5 y^2 + 0.75 z^2)) * amp) from separately rounded fp32 tensor ops (one torch kernel per
    multiply / add: nothing is contracted into an FMA), so CPU and CUDA give the same bits."""
    x = x.to(torch.float32)
    x2 = x * x
    s = x2[:, 0] + x2[:, 1] * 1.5
    s = s + x2[:, 2] * 0.75
    return torch.relu((r2 - s) * amp)


# bound 1 (one cascade): dense blob, mean density above density_thresh -> the threshold is 0.01;
# bound 2 (two cascades), density_scale 0.5, faint blob: mean below 0.01 -> the threshold is the mean itself.
SCENARIOS = [
    dict(name="b1", bound=1, density_scale=1, amp=30.0, r2=0.30, n_poses=5, radius=1.6, intrinsic=(70.0, 64.0, 32.0, 30.0),
         steps=[dict(seed=11), dict(seed=12, counts=[1000, 1200, 1100]), dict(seed=13, iter_density=16, counts=[900, 950])]),
    dict(name="b2", bound=2, density_scale=0.5, amp=0.05, r2=0.20, n_poses=3, radius=2.5, intrinsic=(50.0, 50.0, 32.0, 32.0),
         steps=[dict(seed=21)]),
]


def scenario_poses(sc):
    """cam2world matrices on a circle around the origin, looking at it (deterministic, no RNG)."""
    n, r = sc["n_poses"], sc["radius"]
    poses = []
    for k in range(n):
        az = 2 * np.pi * k / n + 0.3
        el = 0.35 + 0.2 * k
        c = np.array([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)])
        fwd = -c / np.linalg.norm(c)           # camera looks along +z of its own frame towards the origin
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, fwd, c
        poses.append(m)
    return torch.from_numpy(np.stack(poses).astype(np.float32))
