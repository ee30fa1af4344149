"""Config family of AbsorbingFlow (reference: mp/config_classes/absorbing_flows_config.py:64-140)."""
from dataclasses import dataclass, field

from ._base import EncoderConfig, TrainingConfig, YamlConfig, make_jets_data_config, register_sections

JetsDataConfig = make_jets_data_config(max_num_particles=109, batch_size=28)


@dataclass
class BridgeConfig:
    continuous: str = "LinearUniformBridge"
    discrete: str = "TelegraphBridge"
    absorbing: str = "AbsorbingBridge"
    sigma: float = 0.0001
    gamma: float = 0.125
    gamma_absorb: float = 0.125
    num_timesteps: int = 1000
    time_eps: float = 0.0001


@dataclass
class GeneratorsHeadConfig:
    # absorbing-rate transformer head + discrete MLP head (absorbing_flows.py:41-88)
    rate_use_x0_pred: bool = True
    transformer_dim: int = 128
    temb_dim: int = 128
    n_heads: int = 2
    n_attn_blocks: int = 2
    detach_last_layer: bool = True
    augment_dim: int = 9
    discrete_head_hidden_dim: int = 56


@register_sections
@dataclass
class AbsorbingConfig(YamlConfig):
    name_str: str = "ExampleModel"
    experiment_name: str = "absorbing_flows"
    experiment_indentifier: str = None
    experiment_dir: str = None
    bridge: BridgeConfig = field(default_factory=BridgeConfig)
    data: JetsDataConfig = field(default_factory=JetsDataConfig)
    encoder: EncoderConfig = field(default_factory=EncoderConfig)
    generator: GeneratorsHeadConfig = field(default_factory=GeneratorsHeadConfig)
    train: TrainingConfig = field(default_factory=TrainingConfig)


__all__ = ["AbsorbingConfig", "BridgeConfig", "JetsDataConfig", "EncoderConfig", "GeneratorsHeadConfig",
           "TrainingConfig"]
