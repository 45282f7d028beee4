"""NeRF hyper-parameter schema of the reference (nerf/config.py:5-72): dataclasses flattened to the dict that the
reference hands to tiny-cuda-nn.  Same keys and values; ``default_factory`` is used so the module imports on
Python >= 3.11 (the reference's instance defaults raise there)."""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class EncodingConfigHG:
    otype: str = "HashGrid"
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 16
    per_level_scale: float = float(np.exp2(np.log2(2048 / 16) / (16 - 1)))


@dataclass
class EncodingConfigSH:
    otype: str = "SphericalHarmonics"
    degree: int = 4


@dataclass
class NetworkConfig:
    otype: str = "FullyFusedMLP"
    activation: str = "ReLU"
    output_activation: str = "None"
    n_neurons: int = 128
    n_hidden_layers: int = 3


@dataclass
class NeRFConfig:
    encoding_sigma: EncodingConfigHG
    network_sigma: NetworkConfig
    encoding_dir: EncodingConfigSH
    network_color: NetworkConfig

    def as_dict(self):
        return {k: dict(v.__dict__) for k, v in self.__dict__.items()}


@dataclass
class BaseNeRFConfig(NeRFConfig):
    encoding_sigma: EncodingConfigHG = field(default_factory=EncodingConfigHG)
    network_sigma: NetworkConfig = field(default_factory=lambda: NetworkConfig(n_hidden_layers=3))
    encoding_dir: EncodingConfigSH = field(default_factory=EncodingConfigSH)
    network_color: NetworkConfig = field(default_factory=lambda: NetworkConfig(n_hidden_layers=4))
