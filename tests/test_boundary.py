"""The drop-in boundary (CPU-only checks): the C-ABI library loads, exports every symbol include/snerf.h declares,
the ctypes table matches the header, the python surface has the reference's names/signatures, and the product
package never touches oracle/."""
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(debug=False):
    """symbols include/snerf.h declares: the product surface, or (debug=True) the block under #ifdef SNERF_DEBUG_HOOKS"""
    text = open(os.path.join(ROOT, "include", "snerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    m = re.search(r"#ifdef SNERF_DEBUG_HOOKS(.*?)#endif", text, flags=re.S)
    assert m, "include/snerf.h keeps the measurement hooks under #ifdef SNERF_DEBUG_HOOKS"
    text = m.group(1) if debug else text[:m.start()] + text[m.end():]
    return sorted(set(re.findall(r"\b(snerf_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    import subprocess
    from stable_nerf_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"libsnerf_b200.so does not export {s}"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and include/snerf.h disagree"
    assert built_lib.snerf_version() >= 100
    assert b"channel_dim" in built_lib.snerf_error_string(-2)
    # the product library carries no measurement hooks, no probe code and no settable tunables: nothing but the surface
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted({l.split()[-1] for l in exported.splitlines() if " T " in l and l.split()[-1].startswith("snerf_")})
    assert exported == syms, set(exported) ^ set(syms)
    # the debug build is the same surface plus exactly the hooks the header declares under SNERF_DEBUG_HOOKS
    dbg_syms = header_symbols(debug=True)
    assert sorted(_lib.DEBUG_SIGNATURES) == dbg_syms and all(s.startswith(("snerf_debug_", "snerf_tc_")) for s in dbg_syms)
    dbg = ctypes.CDLL(_lib.DBG_LIB_PATH)
    for s in syms + dbg_syms:
        assert hasattr(dbg, s), f"libsnerf_b200_dbg.so does not export {s}"


def test_ctypes_arity_matches_header():
    from stable_nerf_b200 import _lib
    text = open(os.path.join(ROOT, "include", "snerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in {**_lib.SIGNATURES, **_lib.DEBUG_SIGNATURES}.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table {len(args)}"


def test_struct_layouts_match_header():
    import ctypes
    from oracle import oracle as orc
    from stable_nerf_b200 import _lib
    assert ctypes.sizeof(_lib.GridDesc) == 16 + 5 * 16 * 4
    assert ctypes.sizeof(_lib.FieldDesc) == ctypes.sizeof(_lib.GridDesc) + 7 * 4  # ... bound, color_in_pad
    assert ctypes.sizeof(orc.GridDesc) == ctypes.sizeof(_lib.GridDesc)
    assert ctypes.sizeof(orc.FieldDesc) == ctypes.sizeof(_lib.FieldDesc)


def test_python_surface_matches_reference_names():
    import stable_nerf_b200 as pkg
    from stable_nerf_b200 import raymarching as rm
    for name in ("near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                 "composite_rays_train", "march_rays", "composite_rays", "compact_rays"):
        assert callable(getattr(rm, name)), name
    # positional signatures of the forward()s (raymarching.py:22,55,85,108,132,164,241,300,354)
    def params(fn):
        return [p for p in inspect.signature(fn).parameters][1:]
    assert params(rm._near_far_from_aabb.forward) == ["rays_o", "rays_d", "aabb", "min_near"]
    assert params(rm._march_rays_train.forward) == ["rays_o", "rays_d", "bound", "density_bitfield", "C", "H", "nears",
                                                    "fars", "step_counter", "mean_count", "perturb", "align",
                                                    "force_all_rays", "dt_gamma", "max_steps"]
    assert params(rm._composite_rays_train.forward) == ["sigmas", "rgbs", "deltas", "rays", "T_thresh", "num_channels"]
    assert params(rm._march_rays.forward) == ["n_alive", "n_step", "rays_alive", "rays_t", "rays_o", "rays_d", "bound",
                                              "density_bitfield", "C", "H", "near", "far", "align", "perturb", "dt_gamma",
                                              "max_steps"]
    assert params(rm._composite_rays.forward) == ["n_alive", "n_step", "rays_alive", "rays_t", "sigmas", "rgbs", "deltas",
                                                  "weights_sum", "depth", "image", "T_thresh", "num_channels"]
    d = inspect.signature(rm.march_rays_train).parameters
    assert d["mean_count"].default == -1 and d["align"].default == -1 and d["max_steps"].default == 1024
    assert inspect.signature(rm._composite_rays.forward).parameters["T_thresh"].default == 1e-2
    assert inspect.signature(rm._composite_rays_train.forward).parameters["T_thresh"].default == 1e-4
    for name in ("NeRFRenderer", "NeRFNetwork", "trunc_exp", "BaseNeRFConfig"):
        assert hasattr(pkg, name)
    sig = inspect.signature(pkg.NeRFRenderer.run_cuda).parameters
    assert [p for p in sig][1:9] == ["rays_o", "rays_d", "dt_gamma", "bg_color", "perturb", "force_all_rays", "max_steps",
                                     "T_thresh"]
    assert sig["max_steps"].default == 1024 and sig["T_thresh"].default == 1e-4
    for m in ("render", "run_cuda", "update_extra_state", "mark_untrained_grid", "reset_extra_state", "density", "color",
              "get_params", "train_step", "eval_step", "test_step"):
        assert callable(getattr(pkg.NeRFNetwork, m)), m


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "stable_nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for bad in ("import oracle", "from oracle", "libsnerf_oracle", "snerf_oracle.h", "orc_"):
                    hits = [l for l in text.splitlines() if bad in l and not l.lstrip().startswith(("//", "#", "*", '"'))
                            and "oracle/" not in l]
                    assert not hits, f"{fn} references the oracle: {hits[:2]}"


def test_missing_extension_fails_loudly(monkeypatch):
    import pytest
    from stable_nerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsnerf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected_without_cuda():
    import pytest
    import torch
    from stable_nerf_b200 import raymarching as rm
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(Exception):
        rm.composite_rays_train(torch.zeros(4), torch.zeros(4, 3), torch.zeros(4, 2), torch.zeros(1, 3, dtype=torch.int32))
