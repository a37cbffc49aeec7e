"""Reference import path `face_replace.configs.train_config`: the configuration tree a checkpoint's `cfg` entry and
`config_files/*.yaml` decode into (reference configs/train_config.py; `pyrallis.decode(TrainConfig, ckpt['cfg'])` at
inference/test.py:43). pyrallis is not a dependency here: `TrainConfig.from_dict` / `from_yaml` / `decode` accept the
same nested mappings. The `model` and `data` sections — the only ones the inference path reads — are typed with the
reference's defaults (:94-147); the training-only sections (compute, optim, log, steps) are kept as attribute
namespaces holding whatever the file provides."""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from pathlib import Path
from types import SimpleNamespace
from typing import Any, Mapping, Optional


def _fill(cls, values: Optional[Mapping[str, Any]]):
    """Dataclass instance from a mapping: known fields are set, unknown keys are kept as plain attributes (newer or
    training-only options must not break loading)."""
    values = dict(values or {})
    known = {f.name for f in fields(cls)}
    obj = cls(**{k: v for k, v in values.items() if k in known})
    for k, v in values.items():
        if k not in known:
            setattr(obj, k, v)
    return obj


@dataclass
class ModelConfig:
    net_type: str = "pix2pix_turbo"
    use_pretrained: bool = True
    lora_rank_unet: int = 16
    lora_rank_vae: int = 16
    condition_on_face_embeds: bool = False
    concat_mask_and_landmarks: bool = False
    use_shared_attention: bool = True
    noise_timestep: int = 249
    train_vae: bool = True
    train_only_vae_encoder: bool = False
    checkpoint_path: Optional[Path] = None
    use_shortcuts: bool = False
    guidance_scale: float = 0.0
    train_reference_networks: bool = False
    use_adain: bool = False
    train_input: bool = True


@dataclass
class DataConfig:
    dataset_type: str = "debug"
    data_root: Any = None
    val_data_root: Any = None
    resolution: int = 512
    max_conditioning_images: int = 4
    store_landmarks: bool = False


@dataclass
class TrainConfig:
    model: ModelConfig = field(default_factory=ModelConfig)
    data: DataConfig = field(default_factory=DataConfig)
    compute: SimpleNamespace = field(default_factory=SimpleNamespace)
    optim: SimpleNamespace = field(default_factory=SimpleNamespace)
    log: SimpleNamespace = field(default_factory=SimpleNamespace)
    steps: SimpleNamespace = field(default_factory=SimpleNamespace)

    @classmethod
    def from_dict(cls, cfg: Optional[Mapping[str, Any]]) -> "TrainConfig":
        cfg = dict(cfg or {})
        loose = lambda name: SimpleNamespace(**dict(cfg.get(name) or {}))
        return cls(model=_fill(ModelConfig, cfg.get("model")), data=_fill(DataConfig, cfg.get("data")),
                   compute=loose("compute"), optim=loose("optim"), log=loose("log"), steps=loose("steps"))

    @classmethod
    def from_yaml(cls, path) -> "TrainConfig":
        import yaml
        with open(path) as fh:
            return cls.from_dict(yaml.safe_load(fh))


def decode(cls, cfg):
    """`pyrallis.decode(TrainConfig, cfg)` stand-in."""
    if isinstance(cfg, cls):
        return cfg
    return cls.from_dict(cfg)
