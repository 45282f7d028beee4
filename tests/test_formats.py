"""Formats either side of the path: the reference's nerf.pth (train.py:307,472) and the
latent + ray-direction block handed to the diffusion side (train.py:72-82).  The packing oracle is pinned against
vectors made with the reference's own tensor expressions (tests/golden/make_golden_extras.py)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras.npz")


def small_model(channel_dim=4, seed_shift=0):
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.config import BaseNeRFConfig
    cfg = BaseNeRFConfig().as_dict()
    cfg["encoding_sigma"]["n_levels"] = 4
    cfg["encoding_sigma"]["log2_hashmap_size"] = 12
    m = NeRFNetwork(config=cfg, channel_dim=channel_dim, bound=1)
    if seed_shift:
        g = torch.Generator().manual_seed(seed_shift)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=g))
            m.density_grid.copy_(torch.rand(m.density_grid.shape, generator=g))
            m.density_bitfield.copy_(torch.randint(0, 256, m.density_bitfield.shape, generator=g, dtype=torch.uint8))
            m.step_counter.copy_(torch.randint(0, 1000, (16, 2), generator=g, dtype=torch.int32))
    return m


def same_state(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)


# ---------------------------------------------------------------------------------------------------------- nerf.pth

def test_reference_layout_names_and_segments():
    from stable_nerf_b200.checkpoint import describe_params, reference_state_layout
    m = small_model()
    lay = reference_state_layout(m)
    assert list(lay) == ["aabb_train", "aabb_infer", "density_grid", "density_bitfield", "step_counter",
                         "sigma_net.params", "encoder_dir.params", "color_net.params"]  # nerf/renderer.py:32-45, network.py:23-37
    segs = describe_params(m)
    sig = [s for s in segs if s[0] == "sigma_net.params"]
    assert [s[3] for s in sig[:4]] == [(128, 8), (128, 128), (128, 128), (16, 128)]  # 4 levels x 2 features in
    end = max(off + shp[0] * shp[1] for _, _, off, shp in sig)
    assert end == lay["sigma_net.params"][0][0]  # segments tile the flat tensor exactly
    offs = sorted((off, shp[0] * shp[1]) for _, _, off, shp in sig)
    assert all(offs[i][0] + offs[i][1] == offs[i + 1][0] for i in range(len(offs) - 1))
    col = [s for s in segs if s[0] == "color_net.params"]
    assert sum(s[3][0] * s[3][1] for s in col) == lay["color_net.params"][0][0]


def test_state_dict_roundtrip_with_prefixes_and_half(tmp_path):
    from stable_nerf_b200.checkpoint import load_reference_checkpoint, load_reference_state_dict, save_reference_checkpoint
    src, dst = small_model(seed_shift=5), small_model()
    assert not same_state(src, dst)
    save_reference_checkpoint(src, tmp_path / "nerf_sd.pth")
    assert load_reference_checkpoint(dst, tmp_path / "nerf_sd.pth") == ([], [])
    assert same_state(src, dst)
    # DDP / accelerate / torch.compile prefixes, wrapped dict, fp16 parameters (widened to fp32)
    wrapped = {"model": {"module._orig_mod." + k: (v.half() if v.dtype == torch.float32 and "params" in k else v)
                         for k, v in src.state_dict().items()}}
    dst2 = small_model()
    from stable_nerf_b200.checkpoint import extract_state
    tensors, plain = extract_state(wrapped)
    load_reference_state_dict(dst2, tensors, plain=plain)
    assert dst2.sigma_net.params.dtype == torch.float32
    assert torch.equal(dst2.sigma_net.params, src.sigma_net.params.half().float())
    assert torch.equal(dst2.density_bitfield, src.density_bitfield)


def test_pickled_reference_module_loads_without_its_classes(tmp_path):
    """train.py:307 pickles the module object; its classes (nerf.network.NeRFNetwork, tinycudann.modules.*) do not exist
    in this process.  Build such a pickle from look-alike classes, remove them, and load it."""
    from stable_nerf_b200.checkpoint import load_reference_checkpoint
    src = small_model(seed_shift=9)
    fake_tcnn, fake_nerf = types.ModuleType("tinycudann.modules"), types.ModuleType("nerf.network")
    pkgs = {"tinycudann": types.ModuleType("tinycudann"), "tinycudann.modules": fake_tcnn,
            "nerf": types.ModuleType("nerf"), "nerf.network": fake_nerf}

    class Module(torch.nn.Module):
        def __init__(self, params):
            super().__init__()
            self.params = torch.nn.Parameter(params.clone())
            self.loss_scale = 128.0

    class NeRFNetwork(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for k in ("aabb_train", "aabb_infer", "density_grid", "density_bitfield", "step_counter"):
                self.register_buffer(k, getattr(src, k).clone())
            self.mean_density, self.iter_density, self.mean_count, self.local_step = 0.0123, 17, 211, 3
            self.sigma_net = Module(src.sigma_net.params.detach())
            self.encoder_dir = Module(torch.empty(0))
            self.color_net = Module(src.color_net.params.detach())

    Module.__module__, Module.__qualname__ = "tinycudann.modules", "Module"
    NeRFNetwork.__module__, NeRFNetwork.__qualname__ = "nerf.network", "NeRFNetwork"
    fake_tcnn.Module, fake_nerf.NeRFNetwork = Module, NeRFNetwork
    saved = {k: sys.modules.get(k) for k in pkgs}
    sys.modules.update(pkgs)
    try:
        torch.save(NeRFNetwork(), tmp_path / "nerf.pth")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "tinycudann" not in sys.modules and "nerf.network" not in sys.modules
    dst = small_model()
    assert load_reference_checkpoint(dst, tmp_path / "nerf.pth") == ([], [])
    assert same_state(src, dst)
    assert (dst.mean_density, dst.iter_density, dst.mean_count, dst.local_step) == (0.0123, 17, 211, 3)


def test_mismatched_checkpoint_is_refused_untouched(tmp_path):
    from stable_nerf_b200.checkpoint import CheckpointError, load_reference_checkpoint, load_reference_state_dict
    src, dst = small_model(channel_dim=3, seed_shift=2), small_model(channel_dim=4)
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    sd = src.state_dict()
    sd["sigma_net.params"] = sd["sigma_net.params"][:-8]  # another hash-table size
    with pytest.raises(CheckpointError, match="sigma_net.params.*config mismatch"):
        load_reference_state_dict(dst, sd)
    sd = src.state_dict()
    del sd["color_net.params"]
    sd["extra.weight"] = torch.zeros(3)
    with pytest.raises(CheckpointError, match="missing.*color_net.params.*unexpected.*extra.weight"):
        load_reference_state_dict(dst, sd)
    assert all(torch.equal(before[k], v) for k, v in dst.state_dict().items())  # nothing was written
    assert load_reference_state_dict(dst, sd, strict=False) == (["color_net.params"], ["extra.weight"])
    (tmp_path / "junk.pth").write_bytes(b"not a checkpoint")
    with pytest.raises(CheckpointError, match="cannot read checkpoint"):
        load_reference_checkpoint(dst, tmp_path / "junk.pth")
    torch.save({"lr": 0.1}, tmp_path / "empty.pth")
    with pytest.raises(CheckpointError):
        load_reference_checkpoint(dst, tmp_path / "empty.pth")


# ------------------------------------------------------------------------------------- latent block for the SD side

def test_oracle_pack_sd_condition_matches_reference_golden():
    from oracle import oracle as orc
    g = np.load(GOLD)
    out = orc.pack_sd_condition(g["sd_latent"], g["sd_dirs"])
    assert out.shape == (2, 7, 64)
    assert np.array_equal(out.reshape(g["sd_cond"].shape), g["sd_cond"])  # copies and one exact affine map: bit-exact


@pytest.mark.gpu
def test_pack_sd_condition_cuda_vs_golden_and_autograd(built_lib, cuda):
    from stable_nerf_b200.sd_bridge import pack_sd_condition
    g = np.load(GOLD)
    lat = torch.from_numpy(g["sd_latent"]).to(cuda).requires_grad_(True)
    dirs = torch.from_numpy(g["sd_dirs"]).to(cuda)
    cond = pack_sd_condition(lat, dirs, encoder_output_dim=8)
    assert cond.shape == (2, 7, 8, 8)
    assert np.array_equal(cond.detach().cpu().numpy(), g["sd_cond"])
    cond.backward(torch.from_numpy(g["sd_gout"]).to(cuda))
    assert np.array_equal(lat.grad.cpu().numpy(), g["sd_glatent"])
    with pytest.raises(RuntimeError):
        pack_sd_condition(lat[:, :63], dirs[:, :63])  # 63 rays per view is not a square


@pytest.mark.gpu
@pytest.mark.parametrize("B,E,C", [(1, 64, 4), (3, 16, 3), (2, 5, 1)])
def test_sd_image_embeds_vs_reference_expression(B, E, C, built_lib, cuda):
    """train.py:75-82 spelled with torch ops (view / permute / cat) against the two-launch packer, values and grads."""
    from oracle import oracle as orc
    from stable_nerf_b200.sd_bridge import sd_image_embeds
    gen = torch.Generator().manual_seed(B * 100 + E)
    pred = torch.rand(B, E * E, C, generator=gen).to(cuda).requires_grad_(True)
    ref_lt = torch.randn(B, C, E, E, generator=gen).to(cuda)
    t_d, r_d = (torch.nn.functional.normalize(torch.randn(B, E * E, 3, generator=gen), dim=-1).to(cuda) for _ in range(2))
    out = sd_image_embeds(pred, t_d, ref_lt, r_d, encoder_output_dim=E)
    pred2 = pred.detach().clone().requires_grad_(True)
    top = torch.cat([pred2.view(B, C, E, E) * 2 - 1, t_d.permute(0, 2, 1).view(B, 3, E, E)], dim=1)
    bot = torch.cat([ref_lt, r_d.permute(0, 2, 1).view(B, 3, E, E)], dim=1)
    want = torch.cat([top, bot], dim=0)
    assert out.shape == (2 * B, C + 3, E, E) and torch.equal(out, want)
    assert np.array_equal(out[:B].detach().cpu().numpy().reshape(B, C + 3, E * E),
                          orc.pack_sd_condition(pred.detach().cpu().numpy(), t_d.cpu().numpy()))
    gout = torch.randn(out.shape, generator=gen).to(cuda)
    out.backward(gout)
    want.backward(gout)
    assert torch.equal(pred.grad, pred2.grad)
    # what SDNetwork.forward does next (stable_diffusion/network.py:194-197): each block flattens to (C+3)*E*E
    assert out.view(-1, (C + 3) * E * E).shape == (2 * B, (C + 3) * E * E)


def test_crafted_pickle_cannot_reach_code(tmp_path):
    """The loader resolves an explicit allowlist of (module, name) pairs only: a file that REDUCEs builtins.eval /
    os.system / torch.hub.load gets inert stubs back (nothing runs) and is refused for holding no tensors."""
    import pickle

    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.checkpoint import CheckpointError, _Unpickler, load_reference_checkpoint
    import io

    class Evil:
        def __init__(self, mod, name, *args):
            self.fn, self.args = (mod, name), args

        def __reduce__(self):
            import importlib
            return getattr(importlib.import_module(self.fn[0]), self.fn[1]), self.args
    marker = tmp_path / "pwned"
    for mod, name, args in (("builtins", "eval", (f"open({str(marker)!r}, 'w').write('x')",)),
                            ("os", "system", (f"touch {marker}",)),
                            ("builtins", "getattr", (object, "__subclasses__"))):
        blob = pickle.dumps(Evil(mod, name, *args))
        obj = _Unpickler(io.BytesIO(blob)).load()
        assert not marker.exists(), f"{mod}.{name} ran"
        assert type(obj).__mro__[1].__name__ == "_Stub"
    path = tmp_path / "evil.pth"
    path.write_bytes(pickle.dumps(Evil("builtins", "eval", (f"open({str(marker)!r}, 'w').write('x')",))))
    with pytest.raises((CheckpointError, Exception)):
        load_reference_checkpoint(NeRFNetwork(), str(path))
    assert not marker.exists()
