"""Golden vectors for one whole training step and the eval / test steps of the REFERENCE's python stack (SURVEY section 8:
train.py:61-70 call pattern; nerf/network.py:128-226 train_step / eval_step / test_step -> nerf/renderer.py run_cuda ->
submodules/raymarching wrappers -> utils/loss_utils.py l1_loss -> backward), every one of those files imported UNMODIFIED
from /root/reference and run on the CPU over the stand-ins of tests/ref_standins.py (oracle-backed ``_raymarching``,
stand-in ``tinycudann`` with this repo's parameter layout).  Stored: rendered image / depth, the L1 loss, and the gradient
of the loss with respect to EVERY parameter -- both MLPs in full norm + strided probes, the hash table as norm + the values
at a strided subset of its non-zero entries.  ``tests/test_step_golden.py`` repeats the step on the GPU through this repo's
``NeRFNetwork.train_step`` + autograd and through the fused ``TrainStep``.

Run:  python tests/golden/make_golden_step.py   ->  tests/golden/step.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

from oracle import oracle as orc  # noqa: E402
import ref_standins  # noqa: E402
from stable_nerf_b200 import synthetic as syn  # noqa: E402
from trace_scene import STEP_SCENES as SCENES, scene_inputs  # noqa: E402

torch.set_num_threads(1)
ref_standins.install(orc)

import nerf.network as ref_network  # noqa: E402   (unmodified)
from utils.loss_utils import l1_loss  # noqa: E402  (unmodified)

out = {}
torch.set_num_threads(8)  # (the MLP matmuls of the 8192-ray case; every comparison of the test has a 1e-4 bar)
for name, C, table_scale in (("s1", 3, 1e4), ("s2", 4, 1e4), ("cfg4", 4, 1e4), ("cfg2", 3, 1e4)):
    SC = SCENES[name]
    inp = scene_inputs(name)
    m = ref_network.NeRFNetwork(channel_dim=C, bound=SC["bound"], density_scale=SC["density_scale"])
    seed = 9000 + C + (100 if name == "cfg2" else 0)
    ws, table, wc = syn.field_params(m.sigma_net.n_mlp, m.sigma_net.gdesc.n_entries * 2, m.color_net.params.numel(),
                                     shapes_sigma=m.sigma_net.shapes, shapes_color=m.color_net.shapes, seed=seed)
    table = (table * np.float32(table_scale)).astype(np.float32)  # U(-1e-4, 1e-4) * 1e4: non-degenerate densities
    with torch.no_grad():
        m.sigma_net.params.copy_(torch.from_numpy(np.concatenate([ws, table])))
        m.color_net.params.copy_(torch.from_numpy(wc))
    m.density_bitfield.copy_(torch.from_numpy(inp["bitfield"]))
    kw = dict(max_steps=SC["max_steps"], T_thresh=SC["T_thresh"], dt_gamma=SC["dt_gamma"])
    img_seed = 31 + C + (100 if name == "cfg2" else 0)
    rng = np.random.default_rng(img_seed)
    o, d = torch.from_numpy(inp["train_o"])[None], torch.from_numpy(inp["train_d"])[None]
    images = torch.from_numpy(rng.random((1, o.shape[1], C), dtype=np.float32))
    big = o.shape[1] > 1000  # store seeds / strided probes instead of full arrays

    # ---- train_step + backward (train.py:61-70)
    m.train()
    pred, gt, losses = m.train_step({"rays_o": o, "rays_d": d, "images": images}, loss_fns={"l1": l1_loss}, **kw)
    loss = losses["l1"]
    loss.backward()
    gs, gc = m.sigma_net.params.grad.numpy(), m.color_net.params.grad.numpy()
    nm = m.sigma_net.n_mlp
    gtab = gs[nm:]
    nz = np.nonzero(gtab)[0]
    sel = nz[::max(1, nz.size // 4000)]
    P = lambda k, v: out.__setitem__(f"{name}_{k}", v)  # noqa: E731
    P("param_seed", np.int64(seed)), P("table_scale", np.float32(table_scale))
    P("images_seed", np.int64(img_seed)), P("images", images.numpy()[:, ::(16 if big else 1)])
    P("train_pred", pred.detach().numpy()[:, ::(16 if big else 1)]), P("train_loss", np.float32(loss.item()))
    P("train_counter", m.step_counter[0].numpy().copy())
    P("grad_w_sigma_probe", gs[:nm][::7].copy()), P("grad_w_sigma_norm", np.float64(np.linalg.norm(gs[:nm].astype(np.float64))))
    P("grad_w_color_probe", gc[::7].copy()), P("grad_w_color_norm", np.float64(np.linalg.norm(gc.astype(np.float64))))
    P("grad_table_idx", sel.astype(np.int64)), P("grad_table_val", gtab[sel].copy())
    P("grad_table_norm", np.float64(np.linalg.norm(gtab.astype(np.float64)))), P("grad_table_nnz", np.int64(nz.size))
    print(f"{name}: loss {loss.item():.6f} samples {int(m.step_counter[0, 0])} |g_ws| {np.linalg.norm(gs[:nm]):.3e} "
          f"|g_wc| {np.linalg.norm(gc):.3e} |g_table| {np.linalg.norm(gtab):.3e} nnz {nz.size}")

    # ---- eval_step / test_step (full small frame)
    m.eval()
    hw = SC["eval_hw"]
    eo, ed = torch.from_numpy(inp["eval_o"])[None], torch.from_numpy(inp["eval_d"])[None]
    eimg = torch.from_numpy(rng.random((1, hw, hw, C), dtype=np.float32))
    with torch.no_grad():
        pred_rgb, pred_depth, gt_rgb, elosses = m.eval_step({"rays_o": eo, "rays_d": ed, "images": eimg},
                                                            loss_fns={"l1": l1_loss}, **dict(kw, T_thresh=SC["T_thresh_eval"]))
        t_rgb, t_depth = m.test_step({"rays_o": eo, "rays_d": ed, "H": hw, "W": hw}, bg_color=SC["bg_color"],
                                     **dict(kw, T_thresh=SC["T_thresh_eval"]))
    P("eval_images", eimg.numpy()), P("eval_pred", pred_rgb.numpy()), P("eval_depth", pred_depth.numpy())
    P("eval_loss", np.float32(elosses["l1"].item())), P("test_pred", t_rgb.numpy()), P("test_depth", t_depth.numpy())
    for k in ("bitfield", "train_o", "train_d", "eval_o", "eval_d"):
        P("input_" + k, inp[k])
path = os.path.join(HERE, "step.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
