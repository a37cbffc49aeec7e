"""ORACLE (test infrastructure): seeded synthetic weights, configs and inputs.

No checkpoint or dataset is reachable offline, so every parity case runs on weights drawn here: variance-preserving
normal init (std = fan_in ** -0.5) so activations stay in fp16 range through ~60 layers, a gain on q/k projections so
the softmax is peaked rather than uniform, randomised norm affine parameters and biases so that a dropped bias or
gamma shows up in the output, and non-zero LoRA B factors so the load-time merge is exercised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from .diffusers024 import add_lora
from .unet import UNet2DConditionModel, UNetConfig

UNET_LORA_TARGETS = ["to_k", "to_q", "to_v", "to_out.0", "conv", "conv1", "conv2", "conv_shortcut", "conv_out",
                     "proj_in", "proj_out", "ff.net.2", "ff.net.0.proj"]  # reference pix2pix_turbo.py:171-174


@dataclass
class ModelFlags:
    """The ModelConfig fields read at inference (reference configs/train_config.py:118-147)."""
    use_shared_attention: bool = True
    use_adain: bool = False
    train_input: bool = True
    condition_on_face_embeds: bool = False
    lora_rank_unet: int = 32
    lora_rank_vae: int = 32
    use_shortcuts: bool = False
    train_reference_networks: bool = False


def seeded_init_(model: nn.Module, seed: int, qk_gain: float = 1.25) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "lora_" in name:
                continue
            if p.ndim >= 2:
                fan_in = p[0].numel()
                std = fan_in ** -0.5
                if name.endswith(("to_q.weight", "to_k.weight")) or ".to_q." in name or ".to_k." in name:
                    std *= qk_gain
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return model


def make_unet(cfg: Optional[UNetConfig] = None, seed: int = 0, lora_rank: int = 0, lora_b_std: float = 0.02,
              freeu: bool = True) -> UNet2DConditionModel:
    unet = UNet2DConditionModel(cfg)
    seeded_init_(unet, seed)
    if lora_rank > 0:
        g = torch.Generator().manual_seed(seed + 1000)
        add_lora(unet, UNET_LORA_TARGETS, r=lora_rank, alpha=lora_rank // 2, generator=g, b_std=lora_b_std)
    if freeu:
        unet.enable_freeu(0.9, 0.2, 1.4, 1.6)  # reference pix2pix_turbo.py:62-68
    return unet.eval().requires_grad_(False)


def caption_embedding(cross_dim: int = 1024, tokens: int = 77, seed: int = 42) -> torch.Tensor:
    """Stand-in for the constant CLIP caption encoding (reference pix2pix_turbo.py:100-106): unit-variance rows."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, tokens, cross_dim, generator=g)


def latents(batch: int, n_ref: int, size: int, seed: int = 1234):
    """Degraded-image latent, reference-image latents and the two DDPM noises, fp32 NCHW."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(batch, 4, size, size, generator=g) * 0.8
    refs = torch.randn(batch, n_ref, 4, size, size, generator=g) * 0.8
    noise_main = torch.randn(batch, 4, size, size, generator=g)
    noise_ref = torch.randn(batch * n_ref, 4, size, size, generator=g)
    return enc, refs, noise_main, noise_ref
