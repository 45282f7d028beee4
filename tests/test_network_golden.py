"""NeRFNetwork.forward / density / color against the REFERENCE's own ``nerf/network.py`` run on the CPU
(tests/golden/make_golden_network.py: the unmodified module over a stand-in ``tinycudann`` -- oracle hash grid / SH-4 and a
torch fp32 MLP).  Pins what nerf/network.py itself does around the tiny-cuda-nn modules (normalisation, output slicing,
[SH | geo] concatenation, activations, the density dict, masked colour queries); tiny-cuda-nn's internals stay unpinned.

Bars: fp32 path 1e-4 relative (max-norm); bf16 tensor-core path 2e-2 (the stated bf16 tolerance)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network.npz")


def _params(C, z):
    from stable_nerf_b200 import synthetic as syn
    from stable_nerf_b200.config import BaseNeRFConfig
    from stable_nerf_b200.field import make_field_desc, mlp_layer_shapes
    f = make_field_desc(BaseNeRFConfig().as_dict(), C, 15, 1.0)
    ss, sc = mlp_layer_shapes(32, 128, 3), mlp_layer_shapes(32, 128, 4)
    ws, table, wc = syn.field_params(sum(o * i for o, i in ss), f.grid.n_entries * 2, sum(o * i for o, i in sc), shapes_sigma=ss,
                                     shapes_color=sc, table_scale=1.0, seed=int(z[f"c{C}_param_seed"]))
    return ws, table, wc


def test_parameter_recipe_reproduces_the_generators(tmp_path):
    """(CPU) the weights / table the goldens were computed with are regenerated from the stored seed"""
    z = np.load(GOLDEN)
    for C in (3, 4):
        ws, table, wc = _params(C, z)
        assert np.array_equal(table[::100003], z[f"c{C}_table_probe"])
        assert np.array_equal(ws[::1009], z[f"c{C}_w_sigma_probe"]) and np.array_equal(wc[::1009], z[f"c{C}_w_color_probe"])


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("C", [3, 4])
def test_network_matches_the_reference_module(C, precision, tol, built_lib, cuda):
    from stable_nerf_b200 import NeRFNetwork
    z = np.load(GOLDEN)
    ws, table, wc = _params(C, z)
    m = NeRFNetwork(channel_dim=C, precision=precision).to(cuda)
    m.eval()
    with torch.no_grad():
        m.sigma_net.params.copy_(torch.from_numpy(np.concatenate([ws, table])))
        m.color_net.params.copy_(torch.from_numpy(wc))
        x, d = torch.from_numpy(z[f"c{C}_x"]).to(cuda), torch.from_numpy(z[f"c{C}_d"]).to(cuda)
        sigma, color = m(x, d)
        assert sigma.dtype == torch.float32 and color.dtype == torch.float32 and color.shape == (x.shape[0], C)
        assert _rel(sigma.cpu().numpy(), z[f"c{C}_sigma"]) <= tol, "forward sigma"
        assert _rel(color.cpu().numpy(), z[f"c{C}_color"]) <= tol, "forward colour"
        dens = m.density(x)
        assert set(dens) == {"sigma", "geo_feat"} and dens["geo_feat"].shape == (x.shape[0], 15)
        assert _rel(dens["sigma"].cpu().numpy(), z[f"c{C}_density_sigma"]) <= tol, "density sigma"
        assert _rel(dens["geo_feat"].float().cpu().numpy(), z[f"c{C}_geo_feat"]) <= tol, "geometry features"
        col = m.color(x, d, geo_feat=dens["geo_feat"])
        assert _rel(col.cpu().numpy(), z[f"c{C}_color_fn"]) <= tol, "color()"
        if C == 3:
            mask = torch.from_numpy(z["c3_mask"]).to(cuda)
            got = m.color(x, d, mask=mask, geo_feat=dens["geo_feat"])
            assert _rel(got.cpu().numpy(), z["c3_color_masked"]) <= tol, "masked color()"
            assert np.array_equal(got.cpu().numpy()[~z["c3_mask"]], z["c3_color_masked"][~z["c3_mask"]])  # untouched rows are zero
            empty = m.color(x, d, mask=torch.zeros_like(mask), geo_feat=dens["geo_feat"])
            assert np.array_equal(empty.cpu().numpy(), z["c3_color_empty_mask"])
