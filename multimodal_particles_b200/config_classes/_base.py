"""Shared machinery for the YAML-backed config dataclasses.

The reference keeps three independent config families, each with hand-written
``from_yaml``/``to_yaml`` (mp/config_classes/*_config.py).  The generation path only *reads*
these objects, so here one mixin provides the YAML round trip for all of them and the
sections common to every family (training, jets data, encoder) are declared once.
Field names and defaults are the reference's: they are the drop-in contract.
"""
from dataclasses import asdict, dataclass, field, fields, is_dataclass
from typing import Dict, List, Optional, Union

import yaml


class YamlConfig:
    """``from_yaml`` builds nested section dataclasses from the mapping; unknown top-level keys
    (e.g. the ``experiment:`` block of config-mbm-test.yaml) are ignored as the reference does."""

    @classmethod
    def from_dict(cls, mapping: dict):
        kwargs = {}
        for f in fields(cls):
            if f.name not in mapping:
                continue
            value = mapping[f.name]
            section = _SECTION_TYPES.get((cls.__name__, f.name))
            if section is not None and isinstance(value, dict):
                value = section(**value)
            kwargs[f.name] = value
        return cls(**kwargs)

    @classmethod
    def from_yaml(cls, file_path: str):
        with open(file_path, "r") as handle:
            return cls.from_dict(yaml.safe_load(handle))

    def to_yaml(self, file_path: str):
        with open(file_path, "w") as handle:
            yaml.dump(asdict(self), handle, default_flow_style=False)


_SECTION_TYPES: Dict[tuple, type] = {}


def register_sections(cls):
    """Record which fields of a top-level config are nested dataclasses."""
    for f in fields(cls):
        default = f.default_factory if f.default_factory is not field().default_factory else None
        if isinstance(default, type) and is_dataclass(default):
            _SECTION_TYPES[(cls.__name__, f.name)] = default
    return cls


def _scheduler_defaults():
    return {"T_max": 1000, "eta_min": 5.0e-5, "last_epoch": -1}


def _info_defaults():
    return {"stats": None, "hist_num_particles": None}


@dataclass
class TrainingConfig:
    # mp/config_classes/multimodal_bridge_matching_config.py:6-21 (read only by the optimiser setup)
    epochs: int = 200
    gradient_clip_val: float = 1.0
    optimizer_name: str = "AdamW"
    lr: float = 0.001
    weight_decay: float = 5.0e-5
    betas: List[float] = field(default_factory=lambda: [0.9, 0.999])
    eps: float = 1.0e-8
    amsgrad: bool = False
    scheduler_name: str = "CosineAnnealingLR"
    scheduler_params: Dict[str, Union[float, int]] = field(default_factory=_scheduler_defaults)


def make_jets_data_config(max_num_particles: int, batch_size: int):
    """JetsDataConfig differs between families only in two defaults
    (mbm config :23-61 -> 128 / 1024; absorbing config :23-62 -> 109 / 28)."""

    @dataclass
    class JetsDataConfig:
        target_name: str = "AspenOpenJets"
        target_path: List[str] = field(default_factory=lambda: None)
        target_preprocess_continuous: str = "standardize"
        target_preprocess_discrete: str = "tokens"
        target_info: Dict[str, Union[list, dict]] = field(default_factory=_info_defaults)
        source_name: str = "GaussNoise"
        source_path: List[str] = field(default_factory=lambda: None)
        source_preprocess_continuous: str = None
        source_preprocess_discrete: str = "tokens"
        source_info: Dict[str, Union[list, dict]] = field(default_factory=_info_defaults)
        source_masks_from_target_masks: bool = True
        fill_target_with_noise: bool = True
        min_num_particles: int = 0
        num_jets: int = 1000
        dim_features_continuous: int = 3
        dim_features_discrete: int = 1
        dim_context_continuous: int = 0
        dim_context_discrete: int = 0
        vocab_size_features: int = 8
        vocab_size_context: int = 0
        return_type: str = "namedtuple"
        data_split_frac: List[float] = field(default_factory=lambda: [0.8, 0.2, 0.0])

    JetsDataConfig.__dataclass_fields__  # noqa: B018  (built)
    # the two family-specific defaults are appended as real dataclass fields
    JetsDataConfig = dataclass(type("JetsDataConfig", (JetsDataConfig,), {
        "__annotations__": {"max_num_particles": int, "batch_size": int},
        "max_num_particles": max_num_particles,
        "batch_size": batch_size,
    }))
    return JetsDataConfig


@dataclass
class EncoderConfig:
    # mp/config_classes/multimodal_bridge_matching_config.py:72-91; `dropout` and `activation`
    # are carried but never read by EPiC
    name: str = "MultiModalEPiC"
    num_blocks: int = 2
    embedding_time: str = "SinusoidalPositionalEncoding"
    embedding_features_continuous: str = "Linear"
    embedding_features_discrete: str = "Embedding"
    embedding_context_continuous: Optional[str] = None
    embedding_context_discrete: Optional[str] = None
    dim_hidden_local: int = 16
    dim_hidden_glob: int = 16
    dim_emb_time: int = 16
    dim_emb_features_continuous: int = 16
    dim_emb_features_discrete: int = 16
    dim_emb_context_continuous: int = 0
    dim_emb_context_discrete: int = 0
    skip_connection: bool = True
    dropout: float = 0.1
    activation: str = "SELU"
    add_discrete_head: bool = True
