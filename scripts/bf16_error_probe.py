"""Measured error of the bf16 tensor-core field path against the fp32 oracle and the bf16-emulating oracle (the numbers
behind the tolerances written in tests/test_gpu_field.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_gpu_field as T
from oracle import oracle as orc
from stable_nerf_b200 import synthetic as syn
from stable_nerf_b200.config import BaseNeRFConfig
from stable_nerf_b200.field import make_field_desc, mlp_layer_shapes
dev = torch.device("cuda:0")
for C, M in ((3, 300), (4, 2500), (3, 6000)):
    f = make_field_desc(BaseNeRFConfig().as_dict(), C, 15, 1.0)
    ws, table, wc = syn.field_params(38912, f.grid.n_entries * 2, 55296, shapes_sigma=mlp_layer_shapes(32, 128, 3),
                                     shapes_color=mlp_layer_shapes(32, 128, 4), table_scale=1.0, seed=1337 + C)
    of = orc.copy_desc(f, orc.FieldDesc)
    x, dirs = T.sample_points(M, seed=C * 77 + M)
    rng = np.random.default_rng(M)
    g_sig = rng.standard_normal(M).astype(np.float32); g_rgb = rng.standard_normal((M, C)).astype(np.float32)
    g_sig[M // 2:] = 0; g_rgb[M // 2:] = 0
    sig, rgb, (gt, gws, gwc) = T.run_cuda_field(f, x, dirs, ws, table, wc, 1, g_sig, g_rgb, dev)
    sig_e, rgb_e = orc.field_forward(of, x, dirs, table, ws, wc, emulate_bf16=True)
    sig_o, rgb_o = orc.field_forward(of, x, dirs, table, ws, wc)
    gt_e, gws_e, gwc_e = orc.field_backward(of, x, dirs, table, ws, wc, g_sig, g_rgb, emulate_bf16=True)
    gt_o, gws_o, gwc_o = orc.field_backward(of, x, dirs, table, ws, wc, g_sig, g_rgb)
    print(f"C={C} M={M}: fwd vs emu sigma {T.rel_err(sig, sig_e):.2e} rgb {T.rel_err(rgb, rgb_e):.2e} | vs fp32 sigma {T.rel_err(sig, sig_o):.2e} rgb {T.rel_err(rgb, rgb_o):.2e}")
    for name, a, e, o in (("w_sigma", gws, gws_e, gws_o), ("w_color", gwc, gwc_e, gwc_o), ("table", gt, gt_e, gt_o)):
        a64, o64 = a.astype(np.float64), o.astype(np.float64)
        cos = float(np.dot(a64, o64) / (np.linalg.norm(a64) * np.linalg.norm(o64) + 1e-300))
        l2 = float(np.linalg.norm(a64 - o64) / (np.linalg.norm(o64) + 1e-300))
        print(f"   grad {name:8s}: vs emu max-rel {T.rel_err(a, e):.2e} | vs fp32 max-rel {T.rel_err(a, o):.2e}  rel-L2 {l2:.2e}  1-cos {1-cos:.2e}   (emu-oracle vs fp32: max-rel {T.rel_err(e, o):.2e})")
