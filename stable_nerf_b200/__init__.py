"""stable_nerf_b200 -- B200-native (sm_100a) NeRF rendering hot path of earslan25/Stable-NeRF.

Layout mirrors the reference modules on the path:
  raymarching  <- submodules/raymarching/raymarching.py   (+ compact_rays)
  renderer     <- nerf/renderer.py
  network      <- nerf/network.py
  activation   <- nerf/activation.py
  config       <- nerf/config.py
All compute goes through libsnerf_b200.so (include/snerf.h); nothing here falls back to the CPU or to torch ops.
"""
from . import raymarching  # noqa: F401
from .activation import trunc_exp  # noqa: F401
from .config import BaseNeRFConfig  # noqa: F401
from .network import NeRFNetwork  # noqa: F401
from .renderer import NeRFRenderer  # noqa: F401

__all__ = ["raymarching", "trunc_exp", "BaseNeRFConfig", "NeRFNetwork", "NeRFRenderer"]
