"""Section-8f neighbours of the path: ray generation (utils/graphics_utils.py:6-88) and the Adam/AdamW step
(train.py:183, test_nerf.py:52).  The numpy oracle is pinned against vectors produced by the reference's own python
(tests/golden/make_golden_extras.py); the CUDA kernels are compared with the oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extras.npz")
ADAM_CASES = {"adam": (False, dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.0)),
              "adam_wd": (False, dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)),
              "adamw": (True, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2))}


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / (np.abs(b).max() + 1e-30))


def test_oracle_get_rays_matches_reference_golden():
    from oracle import oracle as orc
    g = np.load(GOLD)
    H, W = g["HW"]
    ro, rd = orc.get_rays(g["poses"], g["intrinsics"], int(W), n=int(H * W))
    assert np.array_equal(g["inds"][0], np.arange(H * W))
    assert np.array_equal(ro, g["rays_o"])
    assert rel(rd, g["rays_d"]) <= 1e-6  # torch.norm / matmul round differently from numpy in the last ulp


@pytest.mark.parametrize("name", list(ADAM_CASES))
def test_oracle_adam_matches_torch_golden(name):
    from oracle import oracle as orc
    g = np.load(GOLD)
    decoupled, kw = ADAM_CASES[name]
    p, m, v = g["adam_p0"].copy(), np.zeros(64, np.float32), np.zeros(64, np.float32)
    for k, grad in enumerate(g["adam_grads"]):
        p, m, v = orc.adam_step(p, grad, m, v, k + 1, decoupled=decoupled, **kw)
        assert rel(p, g[name + "_traj"][k]) <= 2e-6, f"step {k + 1}"


@pytest.mark.gpu
def test_get_rays_cuda_vs_oracle_and_golden(built_lib, cuda):
    from oracle import oracle as orc
    from stable_nerf_b200.graphics_utils import get_rays
    g = np.load(GOLD)
    H, W = (int(x) for x in g["HW"])
    poses = torch.from_numpy(g["poses"]).to(cuda)
    out = get_rays(poses, g["intrinsics"], H, W, N=-1)
    assert out["rays_o"].shape == (2, H * W, 3) and out["inds"].shape == (2, H * W)
    assert np.array_equal(out["rays_o"].cpu().numpy(), g["rays_o"])
    assert rel(out["rays_d"].cpu().numpy(), g["rays_d"]) <= 1e-6
    # random subset, patch and error-map sampling: the rays must be the ones of the returned pixel indices
    torch.manual_seed(3)
    for kw in (dict(N=100), dict(N=64, patch_size=4), dict(N=50, error_map=torch.rand(2, 128 * 128))):
        o = get_rays(poses, g["intrinsics"], H, W, **kw)
        inds = o["inds"].cpu().numpy()
        assert inds.min() >= 0 and inds.max() < H * W and o["rays_d"].shape[:2] == inds.shape
        ro, rd = orc.get_rays(g["poses"], g["intrinsics"], W, inds=inds)
        assert np.array_equal(o["rays_o"].cpu().numpy(), ro) and rel(o["rays_d"].cpu().numpy(), rd) <= 1e-6
        assert ("inds_coarse" in o) == ("error_map" in kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ADAM_CASES))
def test_fused_adam_vs_torch_and_golden(name, built_lib, cuda):
    from stable_nerf_b200.optim import FusedAdam, FusedAdamW
    g = np.load(GOLD)
    decoupled, kw = ADAM_CASES[name]
    p = torch.nn.Parameter(torch.from_numpy(g["adam_p0"]).to(cuda))
    opt = (FusedAdamW if decoupled else FusedAdam)([p], **kw)
    for k, grad in enumerate(g["adam_grads"]):
        p.grad = torch.from_numpy(grad).to(cuda)
        opt.step()
        assert rel(p.detach().cpu().numpy(), g[name + "_traj"][k]) <= 2e-6, f"step {k + 1}"
    # a field-sized tensor against torch.optim on the device, with the gradient zeroed by the step
    n = 1 << 20
    gen = torch.Generator(device="cpu").manual_seed(5)
    p0 = torch.randn(n, generator=gen).to(cuda)
    a, b = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    ours = (FusedAdamW if decoupled else FusedAdam)([a], zero_grad_in_step=True, **kw)
    ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)([b], **kw)
    for k in range(3):
        gk = torch.randn(n, generator=gen).to(cuda)
        a.grad, b.grad = gk.clone(), gk.clone()
        ours.step()
        ref.step()
        assert float(a.grad.abs().max()) == 0.0
    assert rel(a.detach().cpu().numpy(), b.detach().cpu().numpy()) <= 2e-6
    assert ours.state_dict()["state"][0]["step"] == 3
    # capturable form (step count and bias corrections on the device): same trajectory; a skipped application does not
    # count; the whole step replays as a CUDA graph
    p = torch.nn.Parameter(torch.from_numpy(g["adam_p0"]).to(cuda))
    opt = (FusedAdamW if decoupled else FusedAdam)([p], capturable=True, **kw)
    p.grad = torch.zeros_like(p)
    opt.skip_next()
    opt.step()  # skipped: nothing moves, nothing counts
    assert np.array_equal(p.detach().cpu().numpy(), g["adam_p0"]) and opt.steps_applied() == [0]
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        p.grad.copy_(torch.from_numpy(g["adam_grads"][0]).to(cuda))
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    assert rel(p.detach().cpu().numpy(), g[name + "_traj"][0]) <= 2e-6
    with torch.cuda.graph(graph):
        opt.step()
    for k, grad in list(enumerate(g["adam_grads"]))[1:]:
        p.grad.copy_(torch.from_numpy(grad).to(cuda))
        graph.replay()
        assert rel(p.detach().cpu().numpy(), g[name + "_traj"][k]) <= 2e-6, f"captured step {k + 1}"
    assert opt.steps_applied() == [len(g["adam_grads"])]


@pytest.mark.gpu
@pytest.mark.parametrize("C,bg", [(3, 1.0), (4, [0.2, 0.4, 0.6, 1.0]), (1, 0.0)])
def test_l1_loss_backward_kernel_vs_oracle_and_autograd(C, bg, built_lib, cuda):
    """snerf_l1_loss_backward against the numpy oracle and against torch autograd of
    ((image + (1 - ws) * bg) - target).abs().mean() (nerf/renderer.py:111, utils/loss_utils.py:9-10)."""
    from oracle import oracle as orc
    from stable_nerf_b200._lib import check, ptr, stream
    rng = np.random.default_rng(C)
    N = 5000
    image, ws = rng.random((N, C), dtype=np.float32), rng.random(N, dtype=np.float32)
    target = rng.random((N, C), dtype=np.float32)
    target[:7] = (image + (1 - ws)[:, None] * np.broadcast_to(np.float32(bg), (C,)))[:7]  # exact zeros: sign(0) = 0
    depth, nears = rng.random(N, dtype=np.float32) * 3, np.full(N, 0.2, np.float32)
    fars = nears + 1 + rng.random(N, dtype=np.float32)
    scale = 0.5 / (N * C)
    loss_o, gi_o, gw_o, pred_o, dn_o = orc.l1_loss_backward(image, ws, target, bg, scale, depth, nears, fars)
    d = {k: torch.from_numpy(v).to(cuda) for k, v in dict(image=image, ws=ws, target=target, depth=depth, nears=nears, fars=fars).items()}
    bg_t = torch.tensor(bg, dtype=torch.float32, device=cuda) if isinstance(bg, list) else None
    loss = torch.zeros((), device=cuda)
    gi, gw, pred, dn = torch.empty(N, C, device=cuda), torch.empty(N, device=cuda), torch.empty(N, C, device=cuda), torch.empty(N, device=cuda)
    check(built_lib.snerf_l1_loss_backward(ptr(d["image"]), ptr(d["ws"]), ptr(d["target"]), ptr(bg_t),
                                           0.0 if isinstance(bg, list) else float(bg), N, C, scale, ptr(loss), ptr(gi), ptr(gw),
                                           ptr(pred), ptr(d["depth"]), ptr(d["nears"]), ptr(d["fars"]), ptr(dn), stream()), "l1")
    torch.cuda.synchronize()
    assert abs(float(loss) - loss_o) <= 1e-6 * loss_o
    assert np.array_equal(gi.cpu().numpy(), gi_o) and rel(gw.cpu().numpy(), gw_o) <= 1e-6
    assert rel(pred.cpu().numpy(), pred_o) <= 1e-6 and rel(dn.cpu().numpy(), dn_o) <= 1e-6
    img_t, ws_t = d["image"].clone().requires_grad_(True), d["ws"].clone().requires_grad_(True)
    bgv = bg_t if bg_t is not None else float(bg)
    l = ((img_t + (1 - ws_t).unsqueeze(-1) * bgv) - d["target"]).abs().mean() * 0.5
    l.backward()
    assert rel(gi.cpu().numpy(), img_t.grad.cpu().numpy()) <= 1e-6 and rel(gw.cpu().numpy(), ws_t.grad.cpu().numpy()) <= 1e-5
