"""Host-side logic (CPU): hash-grid level table, config schema, module state layout, synthetic workloads."""
import math

import numpy as np
import torch


def test_grid_level_table_matches_survey_appendix_a():
    from stable_nerf_b200.config import BaseNeRFConfig
    from stable_nerf_b200.field import make_grid_desc
    cfg = BaseNeRFConfig().as_dict()
    g = make_grid_desc(cfg["encoding_sigma"])
    res = [16, 23, 31, 43, 59, 81, 112, 154, 213, 295, 407, 562, 777, 1073, 1483, 2048]
    size = [4096, 12168, 29792, 79512, 205384] + [524288] * 11
    assert [g.resolution[l] for l in range(16)] == res
    assert [g.size[l] for l in range(16)] == size
    assert [g.hashed[l] for l in range(16)] == [0] * 5 + [1] * 11
    assert g.n_entries == 6098120 and g.n_entries * g.n_features == 12196240
    assert abs(g.scale[15] - 2047.0) < 1e-3 and g.scale[0] == 15.0
    assert [g.offset[l] for l in range(16)] == list(np.cumsum([0] + size[:-1]))


def test_config_dict_equals_reference_values():
    """nerf/config.py:47-72 (the reference file cannot be imported on python >= 3.11)."""
    from stable_nerf_b200.config import BaseNeRFConfig
    d = BaseNeRFConfig().as_dict()
    assert d["encoding_sigma"] == dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                                       base_resolution=16, per_level_scale=float(np.exp2(np.log2(2048 / 16) / 15)))
    assert d["network_sigma"] == dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=128,
                                      n_hidden_layers=3)
    assert d["encoding_dir"] == dict(otype="SphericalHarmonics", degree=4)
    assert d["network_color"]["n_hidden_layers"] == 4 and d["network_color"]["n_neurons"] == 128
    # two instances do not share mutable state
    a, b = BaseNeRFConfig(), BaseNeRFConfig()
    a.network_sigma.n_neurons = 64
    assert b.network_sigma.n_neurons == 128


def test_network_parameters_and_buffers():
    from stable_nerf_b200 import NeRFNetwork
    m = NeRFNetwork(channel_dim=4, bound=2)
    assert m.cascade == 2 and m.grid_size == 128
    sd = m.state_dict()
    assert sd["density_grid"].shape == (2, 128 ** 3) and sd["density_grid"].dtype == torch.float32
    assert sd["density_bitfield"].shape == (2 * 128 ** 3 // 8,) and sd["density_bitfield"].dtype == torch.uint8
    assert sd["step_counter"].shape == (16, 2) and sd["step_counter"].dtype == torch.int32
    assert sd["aabb_train"].tolist() == [-2, -2, -2, 2, 2, 2] and sd["aabb_infer"].tolist() == [-2, -2, -2, 2, 2, 2]
    assert sd["sigma_net.params"].numel() == 38912 + 12196240
    assert sd["color_net.params"].numel() == 55296
    assert sd["encoder_dir.params"].numel() == 0
    assert sum(p.numel() for p in m.parameters()) == 12290448
    m2 = NeRFNetwork(channel_dim=4, bound=2)
    m2.load_state_dict(sd)  # round-trips
    assert torch.equal(m.sigma_net.params, m2.sigma_net.params)  # seeded init is deterministic
    tab = m.sigma_net.table().detach()
    assert tab.abs().max() <= 1e-4 and tab.abs().mean() > 4e-5
    # Xavier bound of the first sigma-net matrix [128,32]
    w0 = m.sigma_net.mlp().detach()[:128 * 32]
    assert w0.abs().max() <= math.sqrt(6.0 / (128 + 32)) + 1e-6
    m.reset_extra_state()
    assert m.mean_count == 0 and m.local_step == 0 and m.iter_density == 0


def test_unsupported_configs_are_refused():
    import pytest
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.config import BaseNeRFConfig
    cfg = BaseNeRFConfig().as_dict()
    cfg["network_sigma"]["activation"] = "Sigmoid"
    with pytest.raises(ValueError):
        NeRFNetwork(config=cfg)
    cfg = BaseNeRFConfig().as_dict()
    cfg["encoding_dir"]["degree"] = 3
    with pytest.raises(ValueError):
        NeRFNetwork(config=cfg)


def test_align_padding_always_adds():
    from stable_nerf_b200.raymarching import _pad_up
    assert _pad_up(100, 128) == 128 and _pad_up(128, 128) == 256 and _pad_up(0, 128) == 128  # raymarching.py:201-202
    assert _pad_up(77, -1) == 77


def test_synthetic_workloads_are_deterministic_and_shaped():
    from stable_nerf_b200 import synthetic as syn
    g = syn.occupancy_grid()
    assert g.shape == (1, 128 ** 3) and 0.05 < g.mean() < 0.09
    bf = syn.pack_bitfield(g)
    assert bf.shape == (128 ** 3 // 8,) and bf.dtype == np.uint8
    o1, d1 = syn.train_batch(4096)
    o2, d2 = syn.train_batch(4096)
    assert np.array_equal(o1, o2) and np.array_equal(d1, d2)
    assert o1.shape == (4096, 3) and o1.dtype == np.float32
    assert np.allclose(np.linalg.norm(d1, axis=-1), 1, atol=1e-5)
    assert np.allclose(np.linalg.norm(o1, axis=-1), syn.BLENDER_RADIUS, atol=1e-4)
    assert abs(syn.BLENDER_FOCAL_800 - 1111.11) < 0.01
    fo, fd = syn.full_frame(64, 64, 88.0)
    assert fo.shape == (4096, 3)
    # the centre pixel looks at the origin
    c = fd[32 * 64 + 32]
    assert np.dot(c, -fo[0] / np.linalg.norm(fo[0])) > 0.999


def test_train_step_exchange_argument_on_cpu():
    """The peer-memory gradient exchange needs CUDA devices: on the CPU an explicit request is refused, 'auto' keeps the
    torch.distributed all-reduce, and the gradients stay ordinary tensors."""
    import pytest
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.config import BaseNeRFConfig
    from stable_nerf_b200.trainer import TrainStep
    cfg = BaseNeRFConfig().as_dict()
    cfg["encoding_sigma"]["n_levels"] = 2
    cfg["encoding_sigma"]["log2_hashmap_size"] = 10
    m = NeRFNetwork(config=cfg)
    with pytest.raises(RuntimeError, match="needs CUDA"):
        TrainStep(m, 16, world_size=2, use_graph=False, exchange="p2p")
    with pytest.raises(ValueError):
        TrainStep(m, 16, world_size=2, use_graph=False, exchange="mpi")
    ts = TrainStep(m, 16, world_size=2, use_graph=False, exchange="auto")
    assert ts.exchange is None and ts.exchange_kind == "nccl" and not ts.overlap_allreduce
    assert TrainStep(m, 16, world_size=1, use_graph=False).exchange_kind == "none"
    assert all(p.grad is not None and p.grad.shape == p.shape for p in ts.params)


def test_split_exchange_policy_and_slices():
    """TrainStep(overlap_exchange=...): "auto" splits the in-graph exchange against the table scatter-add only for the
    NVSwitch kernel, the bf16 path and steps of at most 2^20 rows; the fine group's slice starts at a 16-byte aligned float
    of the arena and the groups cover the table without a gap (host logic only: no exchange is made)."""
    import torch
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.trainer import TrainStep

    class FakeExchange:  # stands where P2PExchange would be; never called
        n_floats = 0

    m = NeRFNetwork(precision="bf16")
    ts = TrainStep(m, 64, use_graph=False)
    assert ts.overlap_split_levels == [8] and ts.overlap_side_ctas == 16 and not ts._overlap_exchange_on(356352)  # no exchange
    ts.exchange, ts.exchange_kind = FakeExchange(), "nvls"
    assert ts._overlap_exchange_on(356352) and ts._overlap_exchange_on(1 << 20)
    assert not ts._overlap_exchange_on((1 << 20) + 128) and not ts._overlap_exchange_on(0)
    ts.exchange_kind = "p2p"
    assert not ts._overlap_exchange_on(356352)
    ts.overlap_exchange = True
    assert ts._overlap_exchange_on(356352) and ts._overlap_exchange_on(5 << 20)
    ts.overlap_exchange = False
    assert not ts._overlap_exchange_on(356352)
    ts.overlap_exchange, ts.exchange_kind = "auto", "nvls"
    m.precision = "fp32"
    assert not ts._overlap_exchange_on(356352)
    m.precision = "bf16"
    ts2 = TrainStep(m, 64, use_graph=False, overlap_split_level=(10, 6, 99, 0))
    assert ts2.overlap_split_levels == [10, 6, 99, 0]
    g = m.fdesc.grid
    L, F, nm = int(g.n_levels), int(g.n_features), m.sigma_net.n_mlp
    cuts = sorted({v for v in ts2.overlap_split_levels if 0 < v < L}, reverse=True)
    assert cuts == [10, 6]
    # arena order of TrainStep._setup_p2p: [colour MLP | sigma MLP | table]; slices [lo(level), end) per cut, then the rest
    table0 = (m.color_net.params.numel() + 3) // 4 * 4 + nm
    hi = table0 + int(g.n_entries) * F
    for lvl in cuts:
        lo = table0 + int(g.offset[lvl]) * F
        assert lo % 4 == 0 and table0 < lo < hi
        hi = lo
    assert hi > table0  # the coarse group is not empty and travels with the MLP gradients
