"""GPU parity of the raymarching operators through the C ABI / python surface:
CUDA kernels vs the CPU oracle, vs the committed golden vectors of the UNMODIFIED reference kernels, and -- when the
prebuilt oracle/_ref/_raymarching.so travelled to the box -- vs the reference kernels run live.

Bars: bit-exact for near/far, marching (per-ray counts and per-ray sample bytes), Morton, packbits, compaction;
1e-4 relative for compositing (fp32 re-association + ex2.approx)."""
import os

import numpy as np
import pytest
import torch

from scenarios import SCENARIOS

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4


def dev_t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def assert_bits_equal(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = a.view(np.uint8) == b.view(np.uint8)
    assert same.all(), f"{what}: {np.count_nonzero(~same)} differing bytes (first at {np.argwhere(~same)[0]})"


def assert_close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max error {err:.3e} of scale {scale:.3e} exceeds {rtol}"


def run_train_march(sc, dev, inp):
    from stable_nerf_b200 import raymarching as rm
    o, d = dev_t(inp["rays_o"], dev), dev_t(inp["rays_d"], dev)
    nears, fars = rm.near_far_from_aabb(o, d, dev_t(inp["aabb"], dev), sc.min_near)
    return o, d, nears, fars


def cuda_march_with_noises(sc, dev, inp, nears, fars, M=None, keep_positions=True):
    """call the C ABI directly so the seeded noises are used (the python op draws torch.rand when perturb=True).
    keep_positions: workspace with room for the per-sample positions (the write pass expands them) or the minimal one
    (the write pass marches again); both must give the same bits."""
    from stable_nerf_b200 import _lib
    from stable_nerf_b200._lib import check, ptr, stream
    lib = _lib.load()
    o, d = dev_t(inp["rays_o"], dev), dev_t(inp["rays_d"], dev)
    bf, noises = dev_t(inp["bitfield"], dev), dev_t(inp["noises"], dev)
    N = sc.n_rays
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    nb = (lib.snerf_march_rays_train_workspace_bytes_ex(N, sc.max_steps) if keep_positions
          else lib.snerf_march_rays_train_workspace_bytes(N))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    geom = (sc.bound, sc.dt_gamma, sc.max_steps, N, sc.cascades, sc.H)
    check(lib.snerf_march_rays_train_count(ptr(o), ptr(d), ptr(bf), *geom, ptr(nears), ptr(fars), ptr(counter),
                                           ptr(noises), ptr(ws), nb, stream()), "count")
    total = int(counter[0].item())
    if M is None:
        M = total + 37  # some padding rows to check zero fill
    xyzs = torch.full((M, 3), float("nan"), device=dev)
    dirs = torch.full((M, 3), float("nan"), device=dev)
    deltas = torch.full((M, 2), float("nan"), device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    nsamp = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.snerf_march_rays_train_write(ptr(o), ptr(d), ptr(bf), *geom, M, ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs),
                                           ptr(deltas), ptr(rays), ptr(noises), 1, ptr(nsamp), ptr(ws), nb, stream()),
          "write")
    torch.cuda.synchronize()
    assert int(nsamp.item()) == total
    return counter.cpu().numpy(), xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy(), total


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_near_far_and_march_train_vs_oracle(sc, built_lib, cuda):
    from oracle import oracle as orc
    inp = sc.inputs()
    o, d, nears, fars = run_train_march(sc, cuda, inp)
    on, of = orc.near_far_from_aabb(inp["rays_o"], inp["rays_d"], inp["aabb"], sc.min_near)
    assert_bits_equal(nears.cpu().numpy(), on, "nears")
    assert_bits_equal(fars.cpu().numpy(), of, "fars")

    counter, xyzs, dirs, deltas, rays, total = cuda_march_with_noises(sc, cuda, inp, nears, fars)
    again = cuda_march_with_noises(sc, cuda, inp, nears, fars, keep_positions=False)  # second-march write pass
    for a, b, name in zip((counter, xyzs, dirs, deltas, rays), again, ("counter", "xyzs", "dirs", "deltas", "rays")):
        assert_bits_equal(a, b, name + " (expand vs re-march write pass)")
    # thread-per-ray grain (what large batches use): the threshold is a compile-time constant of the product library, so the
    # other grain runs through the debug build (same sources, settable tunables) and is compared with the product's bits
    from stable_nerf_b200 import _lib
    with _lib.debug_library() as dbg:
        dbg.snerf_debug_set_march_warp_max_rays(0)
        try:
            for keep in (True, False):
                again = cuda_march_with_noises(sc, cuda, inp, nears, fars, keep_positions=keep)
                for a, b, name in zip((counter, xyzs, dirs, deltas, rays), again, ("counter", "xyzs", "dirs", "deltas", "rays")):
                    assert_bits_equal(a, b, name + f" (thread-per-ray grain, keep_positions={keep})")
        finally:
            dbg.snerf_debug_set_march_warp_max_rays(49152)
    ox, od, odl, orays, ocounter = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"],
                                                        sc.cascades, sc.H, on, of, inp["noises"], sc.dt_gamma,
                                                        sc.max_steps)
    assert counter[0] == ocounter[0] and counter[1] == sc.n_rays
    assert_bits_equal(rays, orays, "rays (id, offset, count)")
    assert_bits_equal(xyzs[:total], ox, "xyzs")
    assert_bits_equal(dirs[:total], od, "dirs")
    assert_bits_equal(deltas[:total], odl, "deltas")
    # rows past the packed samples are zero-filled by the kernel
    assert (xyzs[total:] == 0).all() and (dirs[total:] == 0).all() and (deltas[total:] == 0).all()
    assert total > 0


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_march_train_vs_reference_golden(sc, built_lib, cuda):
    path = os.path.join(GOLDEN_DIR, sc.name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet (tests/golden/make_golden.py)")
    g = np.load(path)
    inp = sc.inputs()
    o, d, nears, fars = run_train_march(sc, cuda, inp)
    assert_bits_equal(nears.cpu().numpy(), g["nears"], "nears vs reference")
    assert_bits_equal(fars.cpu().numpy(), g["fars"], "fars vs reference")
    counter, xyzs, dirs, deltas, rays, total = cuda_march_with_noises(sc, cuda, inp, nears, fars)
    assert_bits_equal(counter, g["counter"], "counter")
    assert_bits_equal(rays[:, 2], g["counts"], "per-ray sample counts vs reference")
    assert_bits_equal(xyzs[:total], g["xyzs"], "xyzs vs reference")
    assert_bits_equal(dirs[:total], g["dirs"], "dirs vs reference")
    assert_bits_equal(deltas[:total], g["deltas"], "deltas vs reference")


def test_march_train_overflow_drops_rays(built_lib, cuda):
    """rays whose segment does not fit in M are dropped (raymarching.cu:417) and their rows are zero."""
    from oracle import oracle as orc
    sc = SCENARIOS[0]
    inp = sc.inputs()
    o, d, nears, fars = run_train_march(sc, cuda, inp)
    full = cuda_march_with_noises(sc, cuda, inp, nears, fars)
    total = full[5]
    M = total // 2
    counter, xyzs, dirs, deltas, rays, _ = cuda_march_with_noises(sc, cuda, inp, nears, fars, M=M)
    ox, od, odl, orays, _ = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"], sc.cascades,
                                                 sc.H, nears.cpu().numpy(), fars.cpu().numpy(), inp["noises"],
                                                 sc.dt_gamma, sc.max_steps, M=M)
    assert_bits_equal(rays, orays, "rays")
    assert_bits_equal(xyzs, ox, "xyzs with dropped rays")
    assert_bits_equal(deltas, odl, "deltas with dropped rays")


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_composite_train_fwd_bwd(sc, built_lib, cuda):
    from oracle import oracle as orc
    from stable_nerf_b200 import raymarching as rm
    inp = sc.inputs()
    on, of = orc.near_far_from_aabb(inp["rays_o"], inp["rays_d"], inp["aabb"], sc.min_near)
    ox, od, odl, orays, ocounter = orc.march_rays_train(inp["rays_o"], inp["rays_d"], sc.bound, inp["bitfield"],
                                                        sc.cascades, sc.H, on, of, inp["noises"], sc.dt_gamma,
                                                        sc.max_steps)
    total, N = int(ocounter[0]), sc.n_rays
    M = total + 128  # padded rows
    sig, rgb = sc.sample_values(M)
    sig[total:] = 0
    dl = np.zeros((M, 2), np.float32)
    dl[:total] = odl
    g_ws, g_img = sc.upstream(N)
    ws, depth, image = orc.composite_rays_train_forward(sig, rgb, dl, orays, sc.t_thresh)
    gs, gr = orc.composite_rays_train_backward(g_ws, g_img, sig, rgb, dl, orays, ws, image, sc.t_thresh)

    t_sig = dev_t(sig, cuda).requires_grad_(True)
    t_rgb = dev_t(rgb, cuda).requires_grad_(True)
    t_rays = dev_t(orays, cuda)
    t_rays._snerf_n_samples = torch.tensor([total], dtype=torch.int32, device=cuda)
    c_ws, c_depth, c_image = rm.composite_rays_train(t_sig, t_rgb, dev_t(dl, cuda), t_rays, sc.t_thresh, sc.channels)
    assert_close(c_ws.detach().cpu().numpy(), ws, "weights_sum")
    assert_close(c_depth.detach().cpu().numpy(), depth, "depth")
    assert_close(c_image.detach().cpu().numpy(), image, "image")
    torch.autograd.backward([c_ws, c_image], [dev_t(g_ws, cuda), dev_t(g_img, cuda)])
    assert_close(t_sig.grad.cpu().numpy(), gs, "grad_sigmas", rtol=2e-4)
    assert_close(t_rgb.grad.cpu().numpy(), gr, "grad_rgbs")
    assert (t_sig.grad[total:] == 0).all() and (t_rgb.grad[total:] == 0).all()

    # the plain C-ABI entry (no n_samples hint) must give the same result
    from stable_nerf_b200 import _lib
    from stable_nerf_b200._lib import check, ptr, stream
    gs2 = torch.full((M,), float("nan"), device=cuda)
    gr2 = torch.full((M, sc.channels), float("nan"), device=cuda)
    keep = [dev_t(g_ws, cuda), dev_t(g_img, cuda), dev_t(dl, cuda)]  # ptr() of a temporary would dangle
    check(_lib.load().snerf_composite_rays_train_backward(
        ptr(keep[0]), ptr(keep[1]), ptr(t_sig.detach()), ptr(t_rgb.detach()), ptr(keep[2]),
        ptr(t_rays), ptr(c_ws.detach()), ptr(c_image.detach()), M, N, sc.t_thresh, sc.channels, ptr(gs2), ptr(gr2),
        stream()), "bwd")
    assert torch.equal(gs2, t_sig.grad) and torch.equal(gr2, t_rgb.grad)

    path = os.path.join(GOLDEN_DIR, sc.name + ".npz")
    if os.path.exists(path):  # reference kernels' own outputs (same seeded sigma/rgb over the exact total)
        g = np.load(path)
        sig_g, rgb_g = sc.sample_values(total)
        c2 = rm.composite_rays_train(dev_t(sig_g, cuda), dev_t(rgb_g, cuda), dev_t(odl, cuda), dev_t(orays, cuda),
                                     sc.t_thresh, sc.channels)
        assert_close(c2[0].cpu().numpy(), g["comp_ws"], "weights_sum vs reference")
        assert_close(c2[1].cpu().numpy(), g["comp_depth"], "depth vs reference")
        assert_close(c2[2].cpu().numpy(), g["comp_image"], "image vs reference")


@pytest.mark.parametrize("sc", SCENARIOS, ids=lambda s: s.name)
def test_inference_march_composite_compact(sc, built_lib, cuda):
    """two iterations of the eval loop body (nerf/renderer.py:136-162) vs the oracle."""
    from oracle import oracle as orc
    from stable_nerf_b200 import _lib, raymarching as rm
    from stable_nerf_b200._lib import check, ptr, stream
    lib = _lib.load()
    inp = sc.inputs()
    N, n_step, C = sc.n_rays, 4, sc.channels
    on, of = orc.near_far_from_aabb(inp["rays_o"], inp["rays_d"], inp["aabb"], sc.min_near)
    o, d = dev_t(inp["rays_o"], cuda), dev_t(inp["rays_d"], cuda)
    bf, nears, fars = dev_t(inp["bitfield"], cuda), dev_t(on, cuda), dev_t(of, cuda)
    # oracle state
    alive_o = np.arange(N, dtype=np.int32)
    rays_t_o = on.copy()
    ws_o, dep_o, img_o = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, C), np.float32)
    # cuda state
    alive = torch.arange(N, dtype=torch.int32, device=cuda)
    rays_t = nears.clone()
    ws, dep, img = torch.zeros(N, device=cuda), torch.zeros(N, device=cuda), torch.zeros(N, C, device=cuda)
    rng = np.random.default_rng(sc.seed + 17)
    for it in range(3):
        n_alive = alive_o.shape[0]
        if n_alive == 0:
            break
        noises = inp["noises"][:n_alive].copy() if it == 0 else np.zeros(n_alive, np.float32)
        M = n_alive * n_step + 128
        ox, od, odl = orc.march_rays(n_alive, n_step, alive_o, rays_t_o, inp["rays_o"], inp["rays_d"], sc.bound,
                                     inp["bitfield"], sc.cascades, sc.H, on, of, noises, sc.dt_gamma, sc.max_steps, M=M)
        xyzs = torch.full((M, 3), float("nan"), device=cuda)
        dirs = torch.full((M, 3), float("nan"), device=cuda)
        deltas = torch.full((M, 2), float("nan"), device=cuda)
        t_noise = dev_t(noises, cuda)
        check(lib.snerf_march_rays_ex(n_alive, n_step, ptr(alive), ptr(rays_t), ptr(o), ptr(d), sc.bound, sc.dt_gamma,
                                      sc.max_steps, sc.cascades, sc.H, ptr(bf), ptr(nears), ptr(fars), ptr(xyzs),
                                      ptr(dirs), ptr(deltas), ptr(t_noise), M, stream()), "march_rays")
        assert_bits_equal(xyzs.cpu().numpy(), ox, f"it{it} xyzs")
        assert_bits_equal(dirs.cpu().numpy(), od, f"it{it} dirs")
        assert_bits_equal(deltas.cpu().numpy(), odl, f"it{it} deltas")
        sig = (rng.random(M, dtype=np.float32) ** 3 * 200.0).astype(np.float32)
        rgb = rng.random((M, C), dtype=np.float32)
        orc.composite_rays(n_alive, n_step, alive_o, rays_t_o, sig, rgb, odl, ws_o, dep_o, img_o, 1e-2)
        rm.composite_rays(n_alive, n_step, alive, rays_t, dev_t(sig, cuda), dev_t(rgb, cuda), deltas, ws, dep, img, 1e-2, C)
        assert np.array_equal(alive[:n_alive].cpu().numpy(), alive_o[:n_alive]), f"it{it} termination flags"
        assert_close(ws.cpu().numpy(), ws_o, f"it{it} weights_sum")
        assert_close(dep.cpu().numpy(), dep_o, f"it{it} depth")
        assert_close(img.cpu().numpy(), img_o, f"it{it} image")
        assert_bits_equal(rays_t.cpu().numpy(), rays_t_o, f"it{it} rays_t")
        # compaction == rays_alive[rays_alive >= 0]
        expect = orc.compact_rays(alive_o, n_alive)
        out, count = rm.compact_rays(alive, n_alive)
        k = int(count.item())
        assert k == expect.shape[0]
        assert np.array_equal(out[:k].cpu().numpy(), expect)
        assert torch.equal(out[:k], alive[:n_alive][alive[:n_alive] >= 0])
        alive = out[:k].contiguous()
        alive_o = expect


def test_compact_edge_cases(built_lib, cuda):
    from stable_nerf_b200 import raymarching as rm
    for n, frac in ((1, 0.0), (1, 1.0), (31, 0.5), (256, 0.0), (257, 1.0), (100000, 0.3), (640000, 0.9)):
        g = torch.Generator(device="cpu").manual_seed(n)
        ids = torch.arange(n, dtype=torch.int32)
        dead = torch.rand(n, generator=g) < frac
        ids[dead] = -1
        ids = ids.to(cuda)
        out, count = rm.compact_rays(ids)
        k = int(count.item())
        ref = ids[ids >= 0]
        assert k == ref.shape[0]
        assert torch.equal(out[:k], ref)
    out, count = rm.compact_rays(torch.empty(0, dtype=torch.int32, device=cuda))
    assert int(count.item()) == 0


def test_small_utils_vs_oracle_and_golden(built_lib, cuda):
    from oracle import oracle as orc
    from stable_nerf_b200 import raymarching as rm
    rng = np.random.default_rng(5)
    grid = (rng.random((2, 128 ** 3 // 64), dtype=np.float32) * 0.02).astype(np.float32)
    bits = rm.packbits(dev_t(grid, cuda), 0.01)
    assert_bits_equal(bits.cpu().numpy(), orc.packbits(grid, 0.01), "packbits")
    # reuse of a caller-provided bitfield (raymarching.py:147-151)
    buf = torch.zeros_like(bits)
    assert rm.packbits(dev_t(grid, cuda), 0.01, buf).data_ptr() == buf.data_ptr() and torch.equal(buf, bits)
    coords = rng.integers(0, 1024, size=(5000, 3)).astype(np.int32)
    ind = rm.morton3D(dev_t(coords, cuda))
    assert_bits_equal(ind.cpu().numpy(), orc.morton3D(coords), "morton3D")
    back = rm.morton3D_invert(ind)
    assert_bits_equal(back.cpu().numpy(), coords, "morton3D_invert round trip")
    sc = SCENARIOS[0]
    inp = sc.inputs()
    sph = rm.sph_from_ray(dev_t(inp["rays_o"], cuda), dev_t(inp["rays_d"], cuda), 4.0)
    assert_close(sph.cpu().numpy(), orc.sph_from_ray(inp["rays_o"], inp["rays_d"], 4.0), "sph_from_ray", rtol=1e-5)
    path = os.path.join(GOLDEN_DIR, sc.name + ".npz")
    if os.path.exists(path):
        g = np.load(path)
        assert_bits_equal(rm.packbits(dev_t(g["pack_in"].reshape(1, -1), cuda), 0.01).cpu().numpy(), g["pack_out"],
                          "packbits vs reference")
        assert_bits_equal(rm.morton3D(dev_t(g["morton_in"], cuda)).cpu().numpy(), g["morton_out"], "morton vs reference")
        assert_close(sph.cpu().numpy(), g["sph"], "sph vs reference", rtol=1e-5)
    # empty inputs
    e = torch.empty(0, 3, device=cuda)
    n, f = rm.near_far_from_aabb(e, e, torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda), 0.2)
    assert n.shape[0] == 0 and f.shape[0] == 0


def test_argument_errors(built_lib, cuda):
    from stable_nerf_b200 import raymarching as rm
    sig = torch.zeros(128, device=cuda)
    rgb = torch.zeros(128, 5, device=cuda)
    dl = torch.zeros(128, 2, device=cuda)
    rays = torch.zeros(4, 3, dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="channel_dim"):
        rm.composite_rays_train(sig, rgb, dl, rays, 1e-4, 5)
    o = torch.zeros(4, 3, device=cuda)
    n = torch.zeros(4, device=cuda)
    with pytest.raises(RuntimeError, match="power of two"):
        rm.march_rays_train(o, o + 1, 1.0, torch.zeros(100, dtype=torch.uint8, device=cuda), 1, 100, n, n + 1)


def test_full_size_properties(built_lib, cuda):
    """cfg2 size (4096 rays, max_steps 1024): size-independent properties + equality with the oracle's counts."""
    from oracle import oracle as orc
    from stable_nerf_b200 import raymarching as rm, synthetic as syn
    grid = syn.occupancy_grid(lego_like=True)
    bf = syn.pack_bitfield(grid)
    ro, rd = syn.train_batch(4096)
    o, d = dev_t(ro, cuda), dev_t(rd, cuda)
    nears, fars = rm.near_far_from_aabb(o, d, torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 1.0, dev_t(bf, cuda), 1, 128, nears, fars, counter, -1, False,
                                                   128, False, 0, 1024)
    total = int(counter[0].item())
    r = rays.cpu().numpy()
    assert (r[:, 0] == np.arange(4096)).all()
    assert (np.cumsum(r[:, 2]) - r[:, 2] == r[:, 1]).all(), "offsets are the exclusive scan of counts in ray order"
    assert r[:, 2].sum() == total and xyzs.shape[0] == total + (128 - total % 128)
    on, of = orc.near_far_from_aabb(ro, rd, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    _, _, odl, orays, oc = orc.march_rays_train(ro, rd, 1.0, bf, 1, 128, on, of, max_steps=1024)
    assert int(oc[0]) == total
    assert_bits_equal(r, orays, "rays at cfg2 size")
    assert_bits_equal(deltas[:total].cpu().numpy(), odl, "deltas at cfg2 size")
    # every sample lies in an occupied cell of the bitfield
    x = xyzs[:total].cpu().numpy()
    cell = np.clip((x + np.float32(1.0)) * np.float32(64.0), 0, 127).astype(np.uint32)  # same fp32 steps as the kernel
    idx = orc.morton3D(cell.astype(np.int32)).astype(np.uint32)
    assert ((bf[idx >> 3] >> (idx & 7)) & 1).all()
    # linearity of compositing in rgb, and weights_sum + T_final == 1 bound
    sig = torch.rand(xyzs.shape[0], device=cuda) * 30
    rgb = torch.rand(xyzs.shape[0], 3, device=cuda)
    w1, d1, i1 = rm.composite_rays_train(sig, rgb, deltas, rays, 1e-4, 3)
    w2, d2, i2 = rm.composite_rays_train(sig, rgb * 2, deltas, rays, 1e-4, 3)
    assert torch.allclose(i2, i1 * 2, rtol=1e-5, atol=1e-6) and torch.equal(w1, w2)
    assert (w1 <= 1 + 1e-5).all() and (w1 >= 0).all()


def test_cfg5_size_thread_grain_and_fused_near_far(built_lib, cuda):
    """cfg5 size (2^18 rays): the thread-per-ray grain (what batches above 49 152 rays use) against the warp-per-ray
    grain on the same rays (counts, offsets, positions bit for bit), sample-in-occupied-cell property, and the count
    entry point with near/far folded in against the stand-alone near_far_from_aabb."""
    from oracle import oracle as orc
    from stable_nerf_b200 import raymarching as rm, synthetic as syn
    from stable_nerf_b200._lib import check, ptr, stream
    lib = built_lib
    N, max_steps = 1 << 18, 1024
    bf_np = syn.pack_bitfield(syn.occupancy_grid(lego_like=True))
    ro, rd = syn.train_batch(N, n_views=8, seed=4)
    o, d, bf = dev_t(ro, cuda), dev_t(rd, cuda), dev_t(bf_np, cuda)
    aabb = torch.tensor([-1., -1, -1, 1, 1, 1], device=cuda)
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    noises = torch.zeros(N, device=cuda)
    nb = lib.snerf_march_rays_train_workspace_bytes_ex(N, max_steps)
    geom = (1.0, 0.0, max_steps, N, 1, 128)
    out = {}
    from stable_nerf_b200 import _lib as libmod
    # "thread": the product library as it stands (2^18 rays are above its compile-time crossover of 49 152);
    # "warp": the debug build of the same sources with the crossover raised
    for name, thr in (("warp", 1 << 30), ("thread", None)):
        ctx = libmod.debug_library() if thr is not None else None
        lib = ctx.__enter__() if ctx is not None else built_lib
        if thr is not None:
            lib.snerf_debug_set_march_warp_max_rays(thr)
        try:
            ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
            counter = torch.zeros(2, dtype=torch.int32, device=cuda)
            n2, f2 = torch.empty(N, device=cuda), torch.empty(N, device=cuda)
            check(lib.snerf_march_rays_train_count_aabb(ptr(o), ptr(d), ptr(bf), ptr(aabb), 0.2, *geom, ptr(n2), ptr(f2),
                                                        ptr(counter), ptr(noises), ptr(ws), nb, stream()), "count_aabb")
            total = int(counter[0].item())
            M = total + 128 - total % 128
            xyzs, dirs, deltas = torch.empty(M, 3, device=cuda), torch.empty(M, 3, device=cuda), torch.empty(M, 2, device=cuda)
            rays = torch.empty(N, 3, dtype=torch.int32, device=cuda)
            check(lib.snerf_march_rays_train_write(ptr(o), ptr(d), ptr(bf), *geom, M, ptr(n2), ptr(f2), ptr(xyzs), ptr(dirs),
                                                   ptr(deltas), ptr(rays), ptr(noises), 1, None, ptr(ws), nb, stream()), "write")
            torch.cuda.synchronize()
            out[name] = (total, rays.cpu().numpy(), xyzs.cpu().numpy(), deltas.cpu().numpy(), n2.cpu().numpy(), f2.cpu().numpy())
        finally:
            if ctx is not None:
                lib.snerf_debug_set_march_warp_max_rays(49152)
                ctx.__exit__(None, None, None)
    tw, tt = out["warp"], out["thread"]
    assert tw[0] == tt[0] and tw[0] > 10_000_000
    for a, b, what in zip(tw[1:], tt[1:], ("rays", "xyzs", "deltas", "nears", "fars")):
        assert_bits_equal(a, b, what + " (warp vs thread grain at 2^18 rays)")
    assert_bits_equal(tw[4], nears.cpu().numpy(), "nears (folded vs stand-alone)")
    assert_bits_equal(tw[5], fars.cpu().numpy(), "fars (folded vs stand-alone)")
    r = tw[1]
    assert (np.cumsum(r[:, 2].astype(np.int64)) - r[:, 2] == r[:, 1]).all()
    x = tw[2][:tw[0]]
    cell = np.clip((x + np.float32(1.0)) * np.float32(64.0), 0, 127).astype(np.uint32)
    idx = orc.morton3D(cell.astype(np.int32)).astype(np.uint32)
    assert ((bf_np[idx >> 3] >> (idx & 7)) & 1).all()
