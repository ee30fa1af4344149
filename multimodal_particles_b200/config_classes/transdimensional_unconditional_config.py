"""Config family of the trans-dimensional jump diffusion
(reference: mp/config_classes/transdimensional_unconditional_config.py:5-154, 233-303).
Only the sections the generation path reads are given their own dataclass; the optimiser /
augmentation / grad-conditioner blocks are carried as plain dicts so YAML files round-trip."""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

from ._base import YamlConfig, _info_defaults, register_sections


@dataclass
class JetsDataConfig:
    target_name: str = "AspenOpenJets"
    target_path: List[str] = None
    target_preprocess_continuous: str = "standardize"
    target_preprocess_discrete: str = "tokens"
    target_info: Dict[str, Union[list, dict]] = field(default_factory=_info_defaults)
    source_name: str = "GaussNoise"
    source_path: List[str] = field(default_factory=lambda: None)
    source_preprocess_continuous: str = None
    source_preprocess_discrete: str = "tokens"
    source_info: Dict[str, Union[list, dict]] = field(default_factory=_info_defaults)
    source_masks_from_target_masks: bool = True
    min_num_particles: int = 0
    max_num_particles: int = 128
    num_jets: int = 100
    dim_features_continuous: int = 3
    dim_features_discrete: int = 1
    dim_context_continuous: int = 0
    dim_context_discrete: int = 0
    vocab_size_features: int = 8
    vocab_size_context: int = 0
    return_type: str = "namedtuple"
    graphical_structure: str = ""
    exist: List[int] = None
    observed: List[int] = None
    batch_size: int = 28
    data_split_frac: List[float] = field(default_factory=lambda: [0.8, 0.2, 0.0])


@dataclass
class LossKwargs:
    class_name: str = "training.loss.JumpLossFinalDim"
    score_loss_weight: float = 1.0
    rate_loss_weight: float = 1.0
    min_t: float = 0.001
    mean_or_sum_over_dim: str = "mean"
    nearest_atom_pred: bool = True
    rate_function_name: str = "step"
    noise_schedule_name: str = "vp_sde"
    auto_loss_weight: float = 1.0
    vp_sde_beta_max: float = 20.0
    nearest_atom_loss_weight: float = 1.0
    x0_logit_ce_loss_weight: float = 1.0
    vp_sde_beta_min: float = 0.1
    loss_type: str = "eps"
    rate_cut_t: float = 0.1


@dataclass
class SamplerKwargs:
    class_name: str = "training.sampler.JumpSampler"
    dt: float = 0.001
    do_jump_back: bool = False
    corrector_start_time: float = 0.1
    corrector_steps: int = 0
    corrector_finish_time: float = 0.003
    dt_schedule: str = "uniform"
    dt_schedule_h: float = 0.001
    condition_type: str = "sweep"
    do_jump_corrector: bool = False
    guidance_weight: float = 1.0
    dt_schedule_tc: float = 0.5
    condition_sweep_idx: int = 0
    sample_near_atom: bool = True
    do_conditioning: bool = False
    condition_sweep_path: Optional[str] = None
    dt_schedule_l: float = 0.001
    corrector_snr: float = 0.1
    jump_back_start_time: float = 0.5
    no_noise_final_step: bool = False


@dataclass
class EncoderConfig:
    name: str = "TransdimensionalEPiC"
    num_blocks: int = 2
    embedding_time: str = "SinusoidalPositionalEncoding"
    embedding_features_continuous: str = "Linear"
    embedding_features_discrete: str = "Embedding"
    embedding_context_continuous: Optional[str] = None
    embedding_context_discrete: Optional[str] = None
    dim_hidden_local: int = 16
    dim_hidden_glob: int = 19
    dim_emb_time: int = 16
    dim_emb_features_continuous: int = 16
    dim_emb_features_discrete: int = 16
    dim_emb_context_continuous: int = 0
    dim_emb_context_discrete: int = 0
    skip_connection: bool = True
    dropout: float = 0.1
    activation: str = "SELU"
    add_discrete_head: bool = True
    rate_use_x0_pred: bool = True
    transformer_dim: int = 128
    n_heads: int = 2
    n_attn_blocks: int = 2
    detach_last_layer: bool = True
    augment_dim: int = 9


def _structure_defaults():
    return {"exist": [1] * 9, "observed": [0, 0, 0, 1, 1, 1, 1, 1, 1]}


@register_sections
@dataclass
class TransdimensionalEpicConfig(YamlConfig):
    data: JetsDataConfig = field(default_factory=JetsDataConfig)
    encoder: EncoderConfig = field(default_factory=EncoderConfig)
    loss_kwargs: LossKwargs = field(default_factory=LossKwargs)
    optimizer_kwargs: dict = field(default_factory=lambda: {"class_name": "torch.optim.Adam", "lr": 3e-5,
                                                            "betas": [0.9, 0.999], "eps": 1e-8})
    structure_kwargs: dict = field(default_factory=_structure_defaults)
    sampler_kwargs: SamplerKwargs = field(default_factory=SamplerKwargs)
    grad_conditioner_kwargs: dict = field(default_factory=lambda: {"class_name": "training.grad_conditioning.MoleculeJump",
                                                                   "grad_norm_clip": 1.0, "lr_rampup_kimg": 320})
    augment_kwargs: dict = field(default_factory=lambda: {"class_name": "training.augment.AugmentPipe", "p": 0.12, "xflip": 1e8,
                                                          "yflip": 1, "scale": 1, "rotate_frac": 1, "aniso": 1,
                                                          "translate_frac": 1})
    just_visualize: bool = False
    distributed: bool = False
    device: str = "cuda"
    total_kimg: int = 200000
    ema_halflife_kimg: int = 500
    batch_size: int = 64
    batch_gpu: Optional[int] = None
    loss_scaling: float = 1.0
    cudnn_benchmark: bool = True
    kimg_per_tick: int = 50
    snapshot_ticks: int = 25
    state_dump_ticks: int = 25
    log_img_ticks: int = 50
    seed: int = 2047813205
    run_dir: str = ""


__all__ = ["TransdimensionalEpicConfig", "JetsDataConfig", "EncoderConfig", "LossKwargs", "SamplerKwargs"]
