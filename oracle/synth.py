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


def make_vae(cfg=None, seed: int = 100, lora_rank: int = 0, lora_b_std: float = 0.02):
    """Seeded AutoencoderKL restatement; lora_rank > 0 wraps it with the 'vae_skip' adapter (pix2pix_turbo.py:150-162)."""
    from .vae import VAE_LORA_TARGETS, AutoencoderKL, VaeConfig
    cfg = cfg or VaeConfig()
    vae = AutoencoderKL(cfg)
    seeded_init_(vae, seed)
    with torch.no_grad():   # keep the posterior std moderate: logvar head (second half of conv_out/quant_conv) damped
        vae.quant_conv.weight[cfg.latent_channels:] *= 0.1
        vae.quant_conv.bias[cfg.latent_channels:] = -3.0
        vae.decoder.conv_out.weight *= 0.5     # keep most decoded pixels inside (-1, 1) so the final clamp hides little
    if lora_rank > 0:
        g = torch.Generator().manual_seed(seed + 1000)
        targets = list(VAE_LORA_TARGETS) + (["skip_conv_1", "skip_conv_2", "skip_conv_3", "skip_conv_4"] if cfg.use_shortcuts else [])
        add_lora(vae, targets, r=lora_rank, alpha=lora_rank // 2, adapter="vae_skip", generator=g, b_std=lora_b_std)
    return vae.eval().requires_grad_(False)


def images(batch: int, n_ref: int, size: int, latent: int, seed: int = 4321):
    """Degraded image, reference images in [-1, 1] (smooth random fields), and the four normal draws."""
    g = torch.Generator().manual_seed(seed)

    def img(*lead):
        low = torch.randn(*lead, 3, size // 8, size // 8, generator=g)
        x = torch.nn.functional.interpolate(low.flatten(0, -4), size=(size, size), mode="bilinear", align_corners=False)
        x = x + 0.1 * torch.randn(x.shape, generator=g)
        return (0.6 * x).clamp(-1, 1).reshape(*lead, 3, size, size)

    c_t = img(batch)
    cond = img(batch, n_ref)
    eps_main = torch.randn(batch, 4, latent, latent, generator=g)
    eps_ref = torch.randn(batch * n_ref, 4, latent, latent, generator=g)
    noise_main = torch.randn(batch, 4, latent, latent, generator=g)
    noise_ref = torch.randn(batch * n_ref, 4, latent, latent, generator=g)
    return c_t, cond, eps_main, eps_ref, noise_main, noise_ref


def seed_face_processors(unet, seed: int = 77):
    """Deterministic weights for the FaceIDAttnProcessor modules register_attention_processor creates (the reference
    leaves them at torch's default init; a trained checkpoint carries them as `...attn2.processor.*`)."""
    for i, (name, proc) in enumerate(sorted(unet.attn_processors.items())):
        if hasattr(proc, "face_projection"):
            seeded_init_(proc, seed + i)
    return unet


def face_embeddings(batch: int, n_faces: int = 4, dim: int = 512, seed: int = 5) -> torch.Tensor:
    """Stand-in for the normalised insightface embeddings of the reference images (test.py:117-126): unit-norm rows."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(batch, n_faces, dim, generator=g)
    return e / e.norm(dim=-1, keepdim=True)
