"""Config family of MultiModalBridgeMatching (reference:
mp/config_classes/multimodal_bridge_matching_config.py:63-116)."""
from dataclasses import dataclass, field

from ._base import EncoderConfig, TrainingConfig, YamlConfig, make_jets_data_config, register_sections

JetsDataConfig = make_jets_data_config(max_num_particles=128, batch_size=1024)


@dataclass
class BridgeConfig:
    continuous: str = "LinearUniformBridge"
    discrete: str = "TelegraphBridge"
    sigma: float = 0.0001
    gamma: float = 0.125
    num_timesteps: int = 1000
    time_eps: float = 0.0001


@register_sections
@dataclass
class MultimodalBridgeMatchingConfig(YamlConfig):
    name_str: str = "ExampleModel"
    bridge: BridgeConfig = field(default_factory=BridgeConfig)
    data: JetsDataConfig = field(default_factory=JetsDataConfig)
    encoder: EncoderConfig = field(default_factory=EncoderConfig)
    train: TrainingConfig = field(default_factory=TrainingConfig)


__all__ = ["MultimodalBridgeMatchingConfig", "BridgeConfig", "JetsDataConfig", "EncoderConfig", "TrainingConfig"]
