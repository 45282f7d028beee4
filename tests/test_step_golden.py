"""One whole training step and the eval / test steps against the REFERENCE's python stack run on the CPU
(tests/golden/make_golden_step.py: nerf/network.py train_step / eval_step / test_step -> nerf/renderer.py run_cuda ->
submodules/raymarching wrappers -> utils/loss_utils.py l1_loss -> backward, all unmodified, over the stand-ins of
tests/ref_standins.py).  Three cases: one cascade / 3 channels; two cascades, dt_gamma, density_scale 0.5 / 4 channels;
BASELINE.json's configs[3] (two 64x64 views = 8192 rays, 4 channels, max_steps 256: 181 756 samples) and configs[1], the
headline workload (4096 rays of an 800x800 view, max_steps 1024), both at full size.

Compared: rendered image, loss, sample counter, d loss / d (sigma MLP, colour MLP, hash table) -- through this repo's
``NeRFNetwork.train_step`` + autograd AND through the fused ``TrainStep`` (the product's training path) -- and the images /
depths of eval_step (native inference loop) and test_step.  Bars: fp32 path 1e-4 relative (max-norm; loss 1e-5); the bf16
tensor-core path within its stated tolerance (image 2e-2, gradients 30 % of the largest entry and cosine >= 0.99).
Measured on a B200: fp32 gradients 1e-7 .. 5e-6 (autograd and fused alike); bf16 MLP gradients 2e-3 .. 1e-2, table 4e-2 / 1.2e-1
per entry, gradient norms within 0.3 %."""
import os

import numpy as np
import pytest
import torch

from trace_scene import STEP_SCENES as SCENES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step.npz")
# cfg4: 8192 rays (two 64x64 views), max_steps 256 = BASELINE configs[3]; cfg2: 4096 rays, max_steps 1024 = configs[1], the
# headline workload -- both at full size
CASES = [("s1", 3), ("s2", 4), ("cfg4", 4), ("cfg2", 3)]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _model(name, C, precision, z, dev):
    from stable_nerf_b200 import NeRFNetwork, synthetic as syn
    from stable_nerf_b200.field import mlp_layer_shapes
    SC = SCENES[name]
    m = NeRFNetwork(channel_dim=C, bound=SC["bound"], density_scale=SC["density_scale"], precision=precision).to(dev)
    ss, sc = mlp_layer_shapes(32, 128, 3), mlp_layer_shapes(32, 128, 4)
    ws, table, wc = syn.field_params(sum(o * i for o, i in ss), m.fdesc.grid.n_entries * 2, sum(o * i for o, i in sc),
                                     shapes_sigma=ss, shapes_color=sc, seed=int(z[f"{name}_param_seed"]))
    table = (table * np.float32(z[f"{name}_table_scale"])).astype(np.float32)
    with torch.no_grad():
        m.sigma_net.params.copy_(torch.from_numpy(np.concatenate([ws, table])))
        m.color_net.params.copy_(torch.from_numpy(wc))
    m.density_bitfield.copy_(torch.from_numpy(z[f"{name}_input_bitfield"]))
    return m


def _images(name, C, n_rays, z):
    """targets of the training step: regenerated from the stored seed (first draw of the generator's rng), checked against
    the stored (strided, for the large case) copy"""
    img = np.random.default_rng(int(z[f"{name}_images_seed"])).random((1, n_rays, C), dtype=np.float32)
    stride = 16 if n_rays > 1000 else 1
    assert np.array_equal(img[:, ::stride], z[f"{name}_images"])
    return img, stride


def _check_grads(name, m, z, tol, report):
    nm = m.sigma_net.n_mlp
    gs, gc = m.sigma_net.params.grad.detach().cpu().numpy(), m.color_net.params.grad.detach().cpu().numpy()
    G = lambda k: z[f"{name}_{k}"]  # noqa: E731
    errs = {
        "w_sigma": _rel(gs[:nm][::7], G("grad_w_sigma_probe")), "w_color": _rel(gc[::7], G("grad_w_color_probe")),
        "table": _rel(gs[nm:][G("grad_table_idx")], G("grad_table_val")),
        "w_sigma_norm": abs(np.linalg.norm(gs[:nm].astype(np.float64)) / float(G("grad_w_sigma_norm")) - 1),
        "w_color_norm": abs(np.linalg.norm(gc.astype(np.float64)) / float(G("grad_w_color_norm")) - 1),
        "table_norm": abs(np.linalg.norm(gs[nm:].astype(np.float64)) / float(G("grad_table_norm")) - 1),
    }
    print(report, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v <= tol, f"{report}: grad {k} {v:.3e} > {tol}"
    if tol <= 1e-3:
        assert int(np.count_nonzero(gs[nm:])) == int(G("grad_table_nnz")), "touched table entries"


@pytest.mark.parametrize("name,C", CASES)
def test_train_step_and_backward_match_the_reference_stack(name, C, built_lib, cuda):
    z = np.load(GOLDEN)
    SC = SCENES[name]
    m = _model(name, C, "fp32", z, cuda)
    kw = dict(max_steps=SC["max_steps"], T_thresh=SC["T_thresh"], dt_gamma=SC["dt_gamma"])
    o = torch.from_numpy(z[f"{name}_input_train_o"]).to(cuda)[None]
    d = torch.from_numpy(z[f"{name}_input_train_d"]).to(cuda)[None]
    img, stride = _images(name, C, o.shape[1], z)
    images = torch.from_numpy(img).to(cuda)
    m.train()
    pred, gt, losses = m.train_step({"rays_o": o, "rays_d": d, "images": images},
                                    loss_fns={"l1": lambda a, b: torch.abs(a - b).mean()}, **kw)
    loss = losses["l1"]
    loss.backward()
    assert np.array_equal(m.step_counter[0].cpu().numpy(), z[f"{name}_train_counter"])
    assert _rel(pred.detach().cpu().numpy()[:, ::stride], z[f"{name}_train_pred"]) <= 1e-4, "rendered image"
    assert abs(loss.item() - float(z[f"{name}_train_loss"])) <= 1e-5 * float(z[f"{name}_train_loss"]), "L1 loss"
    _check_grads(name, m, z, 1e-4, f"{name} autograd fp32")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 0.30)])
@pytest.mark.parametrize("name,C", CASES)
def test_fused_train_step_matches_the_reference_stack(name, C, precision, tol, built_lib, cuda):
    """the product's training path (TrainStep: the step as a straight sequence of C-ABI calls, here on its first-epoch
    branch: rows sized from the measured sample total) on the reference's batch"""
    from stable_nerf_b200.trainer import TrainStep
    z = np.load(GOLDEN)
    SC = SCENES[name]
    m = _model(name, C, precision, z, cuda)
    m.train()
    o = torch.from_numpy(z[f"{name}_input_train_o"]).to(cuda)
    d = torch.from_numpy(z[f"{name}_input_train_d"]).to(cuda)
    img, stride = _images(name, C, o.shape[0], z)
    images = torch.from_numpy(img).to(cuda)[0]
    bg = 1 if C == 3 else torch.ones(C, device=cuda)  # nerf/network.py:136-140
    ts = TrainStep(m, o.shape[0], max_steps=SC["max_steps"], bg_color=bg, T_thresh=SC["T_thresh"], dt_gamma=SC["dt_gamma"],
                   use_graph=False)
    ts.rays_o.copy_(o), ts.rays_d.copy_(d), ts.target.copy_(images)
    ts._body()  # mean_count == 0: the fused first-epoch body
    torch.cuda.synchronize()
    assert np.array_equal(m.step_counter[0].cpu().numpy(), z[f"{name}_train_counter"])
    ltol = 1e-5 if precision == "fp32" else 2e-2
    assert abs(float(ts.loss) - float(z[f"{name}_train_loss"])) <= ltol * float(z[f"{name}_train_loss"]), "L1 loss"
    assert _rel(ts.outputs["image"].cpu().numpy().reshape(1, -1, C)[:, ::stride], z[f"{name}_train_pred"]) <= (1e-4 if precision == "fp32" else 2e-2)
    _check_grads(name, m, z, tol, f"{name} fused {precision}")
    if precision == "bf16":  # direction of the gradient (stated bf16 tolerance)
        nm = m.sigma_net.n_mlp
        g = m.sigma_net.params.grad.detach().cpu().numpy()[nm:][z[f"{name}_grad_table_idx"]].astype(np.float64)
        r = z[f"{name}_grad_table_val"].astype(np.float64)
        assert float(g @ r / (np.linalg.norm(g) * np.linalg.norm(r))) >= 0.99


@pytest.mark.parametrize("name,C", CASES)
def test_eval_and_test_steps_match_the_reference_stack(name, C, built_lib, cuda):
    z = np.load(GOLDEN)
    SC = SCENES[name]
    m = _model(name, C, "fp32", z, cuda)
    m.eval()
    kw = dict(max_steps=SC["max_steps"], T_thresh=SC["T_thresh_eval"], dt_gamma=SC["dt_gamma"])
    hw = SC["eval_hw"]
    eo = torch.from_numpy(z[f"{name}_input_eval_o"]).to(cuda)[None]
    ed = torch.from_numpy(z[f"{name}_input_eval_d"]).to(cuda)[None]
    eimg = torch.from_numpy(z[f"{name}_eval_images"]).to(cuda)
    with torch.no_grad():
        pred_rgb, pred_depth, gt_rgb, losses = m.eval_step({"rays_o": eo, "rays_d": ed, "images": eimg},
                                                           loss_fns={"l1": lambda a, b: torch.abs(a - b).mean()}, **kw)
        t_rgb, t_depth = m.test_step({"rays_o": eo, "rays_d": ed, "H": hw, "W": hw}, bg_color=SC["bg_color"], **kw)
    assert pred_rgb.shape == (1, hw, hw, C) and pred_depth.shape == (1, hw, hw)
    assert _rel(pred_rgb.cpu().numpy(), z[f"{name}_eval_pred"]) <= 1e-4 and _rel(pred_depth.cpu().numpy(), z[f"{name}_eval_depth"]) <= 1e-4
    assert abs(float(losses["l1"]) - float(z[f"{name}_eval_loss"])) <= 1e-5 * float(z[f"{name}_eval_loss"])
    assert _rel(t_rgb.cpu().numpy(), z[f"{name}_test_pred"]) <= 1e-4 and _rel(t_depth.cpu().numpy(), z[f"{name}_test_depth"]) <= 1e-4
