"""ctypes binding of libsnerf_b200.so (C ABI declared in include/snerf.h).

There is no CPU fallback and no other backend: if the shared library is missing or a call returns a non-zero
code, a RuntimeError is raised.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C stable_nerf_b200/csrc``.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnerf_b200.so")
DBG_LIB_PATH = os.path.join(_HERE, "libsnerf_b200_dbg.so")

SNERF_MAX_LEVELS = 16
SNERF_MAX_CHANNELS = 4
PRECISION_FP32 = 0
PRECISION_BF16 = 1
BWD_ZERO_TABLE_GRAD = 1  # snerf_field_backward_ex flags (include/snerf.h)
BWD_ZERO_W_GRADS = 2


class GridDesc(ctypes.Structure):
    """snerf_grid_desc (include/snerf.h)."""
    _fields_ = [
        ("n_levels", c_uint32), ("n_features", c_uint32), ("n_entries", c_uint32), ("reserved", c_uint32),
        ("scale", c_float * SNERF_MAX_LEVELS),
        ("resolution", c_uint32 * SNERF_MAX_LEVELS),
        ("offset", c_uint32 * SNERF_MAX_LEVELS),
        ("size", c_uint32 * SNERF_MAX_LEVELS),
        ("hashed", c_uint32 * SNERF_MAX_LEVELS),
    ]


class FieldDesc(ctypes.Structure):
    """snerf_field_desc (include/snerf.h)."""
    _fields_ = [
        ("grid", GridDesc),
        ("width", c_uint32), ("n_hidden_sigma", c_uint32), ("n_hidden_color", c_uint32),
        ("geo_feat_dim", c_uint32), ("channel_dim", c_uint32), ("bound", c_float), ("color_in_pad", c_float),
    ]


SNERF_P2P_MAX_RANKS = 16
SNERF_P2P_HANDLE_BYTES = 64
SNERF_P2P_CHANNELS = 4
SNERF_P2P_EMULATE_RANKS = 1


class P2PPeers(ctypes.Structure):
    """snerf_p2p_peers (include/snerf.h): arenas and flag blocks of all ranks as mapped in this process."""
    _fields_ = [("buf", c_void_p * SNERF_P2P_MAX_RANKS), ("flags", c_void_p * SNERF_P2P_MAX_RANKS),
                ("mc_buf", c_void_p), ("host_error", c_void_p), ("timeout_ms", c_uint32), ("flags_word", c_uint32)]


class RenderStats(ctypes.Structure):
    """snerf_render_stats (include/snerf.h)."""
    _fields_ = [("iterations", c_uint32), ("reserved", c_uint32), ("rows", c_uint64), ("samples", c_uint64)]


_P = c_void_p
_U = c_uint32
_F = c_float
_S = c_void_p  # stream

# name -> (restype, argtypes); mirrors include/snerf.h one to one
SIGNATURES = {
    "snerf_version": (c_int, []),
    "snerf_error_string": (c_char_p, [c_int]),
    "snerf_launch_count": (c_uint64, []),
    "snerf_near_far_from_aabb": (c_int, [_P, _P, _P, _U, _F, _P, _P, _S]),
    "snerf_sph_from_ray": (c_int, [_P, _P, _F, _U, _P, _S]),
    "snerf_morton3D": (c_int, [_P, _U, _P, _S]),
    "snerf_morton3D_invert": (c_int, [_P, _U, _P, _S]),
    "snerf_packbits": (c_int, [_P, _U, _F, _P, _S]),
    "snerf_march_rays_train_workspace_bytes": (c_size_t, [_U]),
    "snerf_march_rays_train_workspace_bytes_ex": (c_size_t, [_U, _U]),
    "snerf_march_rays_train_count": (c_int, [_P, _P, _P, _F, _F, _U, _U, _U, _U, _P, _P, _P, _P, _P, c_size_t, _S]),
    "snerf_march_rays_train_count_aabb": (c_int, [_P, _P, _P, _P, _F, _F, _F, _U, _U, _U, _U, _P, _P, _P, _P, _P, c_size_t,
                                                  _S]),
    "snerf_march_rays_train_write": (c_int, [_P, _P, _P, _F, _F, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, c_int,
                                             _P, _P, c_size_t, _S]),
    "snerf_march_rays_ex": (c_int, [_U, _U, _P, _P, _P, _P, _F, _F, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _U, _S]),
    "snerf_march_rays_train": (c_int, [_P, _P, _P, _F, _F, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                       c_size_t, _S]),
    "snerf_composite_rays_train_forward": (c_int, [_P, _P, _P, _P, _U, _U, _F, _U, _P, _P, _P, _S]),
    "snerf_composite_rays_train_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _F, _U, _P, _P, _S]),
    "snerf_composite_rays_train_backward_ex": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _F, _U, _P, _P, _P, _S]),
    "snerf_l1_loss_backward": (c_int, [_P, _P, _P, _P, _F, _U, _U, _F, _P, _P, _P, _P, _P, _P, _P, _P, _S]),
    "snerf_march_rays": (c_int, [_U, _U, _P, _P, _P, _P, _F, _F, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _S]),
    "snerf_composite_rays": (c_int, [_U, _U, _F, _U, _P, _P, _P, _P, _P, _P, _P, _P, _S]),
    "snerf_compact_rays_workspace_bytes": (c_size_t, [_U]),
    "snerf_compact_rays": (c_int, [_P, _U, _P, _P, _P, c_size_t, _S]),
    "snerf_hashgrid_forward": (c_int, [POINTER(GridDesc), _P, _P, _U, _P, _S]),
    "snerf_hashgrid_backward": (c_int, [POINTER(GridDesc), _P, _P, _U, _P, _S]),
    "snerf_sh4_forward": (c_int, [_P, _U, _P, _S]),
    "snerf_mlp_sigma_params": (c_uint32, [POINTER(FieldDesc)]),
    "snerf_mlp_color_params": (c_uint32, [POINTER(FieldDesc)]),
    "snerf_field_workspace_bytes": (c_size_t, [POINTER(FieldDesc), _U, c_int, c_int]),
    "snerf_field_saved_bytes": (c_size_t, [POINTER(FieldDesc), _U, c_int]),
    "snerf_field_forward": (c_int, [POINTER(FieldDesc), _P, _P, _U, _P, _P, _P, c_int, _P, _P, _P, c_size_t, _P, c_size_t,
                                    _S]),
    "snerf_field_density": (c_int, [POINTER(FieldDesc), _P, _U, _P, _P, c_int, _P, _P, _P, c_size_t, _S]),
    "snerf_field_backward": (c_int, [POINTER(FieldDesc), _P, _P, _U, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_size_t,
                                     _P, c_size_t, _S]),
    "snerf_get_rays": (c_int, [_P, _F, _F, _F, _F, _U, _P, _U, _U, c_int, _P, _P, _S]),
    "snerf_pack_sd_condition": (c_int, [_P, _P, _U, _U, _U, _F, _F, _P, _S]),
    "snerf_pack_sd_condition_backward": (c_int, [_P, _U, _U, _U, _F, _P, _S]),
    "snerf_adam_step": (c_int, [_P, _P, _P, _P, _U, _F, _F, _F, _F, _F, c_int, _U, c_int, _S]),
    "snerf_adam_advance": (c_int, [_P, _F, _F, _F, _S]),
    "snerf_adam_step_dev": (c_int, [_P, _P, _P, _P, _U, _F, _F, _F, _F, _F, c_int, _P, c_int, _S]),
    "snerf_field_backward_ex": (c_int, [POINTER(FieldDesc), _P, _P, _U, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_size_t,
                                        _P, c_size_t, _P, _U, _S]),
    "snerf_hashgrid_backward_levels": (c_int, [POINTER(GridDesc), _P, _F, _P, _U, _P, _U, _U, _S]),
    "snerf_mark_untrained_grid": (c_int, [_P, _U, _F, _F, ctypes.c_double, _U, _U, _P, _P, _S]),
    "snerf_grid_cell_points": (c_int, [_P, _U, _U, _U, ctypes.c_double, _U, _P, c_uint64, _P, _S]),
    "snerf_grid_ema_workspace_bytes": (c_size_t, [_U]),
    "snerf_grid_ema_update": (c_int, [_P, _P, _U, _F, _F, _F, _P, _P, _P, c_size_t, _S]),
    "snerf_composite_l1_train": (c_int, [_P, _P, _P, _P, _U, _U, _F, _U, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                         _P, _P, _S]),
    "snerf_render_rays_workspace_bytes": (c_size_t, [POINTER(FieldDesc), _U, _U, c_int]),
    "snerf_render_rays": (c_int, [POINTER(FieldDesc), _P, _P, _U, _P, _U, _U, _F, _F, _U, _P, _P, _P, _P, _P, _P, c_int, _F,
                                  _F, _U, _P, _P, _P, _P, POINTER(RenderStats), _P, c_size_t, _S]),
    "snerf_p2p_flag_bytes": (c_size_t, []),
    "snerf_p2p_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "snerf_p2p_free": (c_int, [_P]),
    "snerf_p2p_export": (c_int, [_P, _P]),
    "snerf_p2p_open": (c_int, [_P, POINTER(c_void_p)]),
    "snerf_p2p_close": (c_int, [_P]),
    "snerf_p2p_allreduce": (c_int, [POINTER(P2PPeers), _U, _U, c_size_t, c_size_t, _U, _U, _S]),
    "snerf_p2p_status": (c_int, [_P, _U, POINTER(c_uint32), POINTER(c_uint32)]),
    "snerf_mc_supported": (c_int, []),
    "snerf_mc_granularity": (c_size_t, [_U, c_size_t]),
    "snerf_mc_arena_create": (c_int, [c_size_t, c_size_t, POINTER(c_void_p), POINTER(c_uint64)]),
    "snerf_mc_create": (c_int, [_U, c_size_t, POINTER(c_uint64), POINTER(c_int)]),
    "snerf_mc_import": (c_int, [c_int, POINTER(c_uint64)]),
    "snerf_mc_add_device": (c_int, [c_uint64]),
    "snerf_mc_bind_and_map": (c_int, [c_uint64, c_uint64, c_size_t, c_size_t, POINTER(c_void_p)]),
    "snerf_mc_release": (c_int, [_P, _P, c_uint64, c_uint64, c_size_t]),
    "snerf_trunc_exp_forward": (c_int, [_P, _U, _P, _S]),
    "snerf_trunc_exp_backward": (c_int, [_P, _P, _U, _P, _S]),
}

# measurement / test hooks: exported by libsnerf_b200_dbg.so only (include/snerf.h, #ifdef SNERF_DEBUG_HOOKS)
DEBUG_SIGNATURES = {
    "snerf_tc_selftest": (c_int, [_P, _P, _P, _U, _U, c_int, c_int, _S]),
    "snerf_debug_set_march_warp_max_rays": (None, [_U]),
    "snerf_debug_set_field_stage_mask": (None, [_U]),
    "snerf_debug_set_side_reduce": (None, [_U]),
    "snerf_debug_set_dedupe_max_res": (None, [_U]),
    "snerf_debug_set_scatter_adaptive_scan": (None, [_U]),
    "snerf_debug_set_tail_prefetch": (None, [_U]),
    "snerf_tc_probe": (c_int, [_P, c_int, _S]),
    "snerf_debug_phase_buffer": (None, [_P, c_int]),
}

_lib = None
_dbg_lib = None
_use_debug = False


def _open(path, signatures):
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C stable_nerf_b200/csrc`. There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in signatures.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load libsnerf_b200.so once; raise loudly if it is not there (no fallback exists).  Inside ``debug_library()`` the
    debug build is returned instead."""
    global _lib
    if _use_debug:
        return load_debug()
    if _lib is None:
        _lib = _open(LIB_PATH, SIGNATURES)
    return _lib


def load_debug():
    """libsnerf_b200_dbg.so: the same sources compiled with -DSNERF_DEBUG_HOOKS (settable tunables, stage masks, phase
    marks, the tcgen05 self-test).  Tests, scripts and bench.py's per-kernel timings use it; the product path never does."""
    global _dbg_lib
    if _dbg_lib is None:
        _dbg_lib = _open(DBG_LIB_PATH, {**SIGNATURES, **DEBUG_SIGNATURES})
    return _dbg_lib


def use_debug_library(on=True):
    """Process-wide switch for measurement scripts: every later ``load()`` returns the debug build."""
    global _use_debug
    _use_debug = bool(on)


class debug_library:
    """``with debug_library() as lib:`` -- every call made through this package inside the block goes to the debug build
    (so that a tunable set through ``lib.snerf_debug_*`` is the one the kernels launched in the block see)."""

    def __enter__(self):
        global _use_debug
        self._prev = _use_debug
        _use_debug = True
        return load_debug()

    def __exit__(self, *exc):
        global _use_debug
        _use_debug = self._prev
        return False


def check(code, what):
    if code != 0:
        msg = load().snerf_error_string(int(code)).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stable_nerf_b200 runs on CUDA tensors only (sm_100a); got a CPU tensor")


def launch_count():
    return int(load().snerf_launch_count())


class Workspace:
    """Grow-only per-device scratch buffers keyed by name.

    The reference wrappers allocate fresh torch.zeros/empty tensors on every call and call
    torch.cuda.empty_cache() on the sync path (raymarching.py:206-231); the kernels here borrow scratch
    from this cache instead.
    """

    def __init__(self):
        self._bufs = {}

    def get(self, name, nbytes, device):
        key = (name, str(device))
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


workspace = Workspace()
