"""Seeded synthetic checkpoints in the reference's state_dict layout.

No released checkpoint is reachable offline, so bench.py / smoke() run on weights drawn here. The key names and
shapes are the ones `Pix2Pix_Turbo` saves for its UNets (reference pix2pix_turbo.py:56-60,171-179; coach.py:712-718):
diffusers-0.24 module paths, with the peft wrapping (`<mod>.base_layer.{weight,bias}`, `<mod>.lora_A.default.weight`,
`<mod>.lora_B.default.weight`) on the LoRA target modules when lora_rank > 0. Values: variance-preserving normal init
(std = fan_in ** -0.5) so activations stay inside fp16 range through the whole network, a gain on q/k so the softmax
is peaked, perturbed norm affines and biases.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .unet_engine import UNetSpec

LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0", "conv", "conv1", "conv2", "conv_shortcut", "conv_out",
                "proj_in", "proj_out", "ff.net.2", "ff.net.0.proj")   # reference pix2pix_turbo.py:171-174


def _is_lora_target(mod: str) -> bool:
    return any(mod == t or mod.endswith("." + t) for t in LORA_TARGETS)


def unet_parameter_shapes(spec: UNetSpec) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(module path, weight shape, kind) for every parameterised leaf of the SD-Turbo UNet, in module order.
    kind: 'conv' | 'linear' | 'linear_nobias' | 'norm'."""
    boc = spec.block_out_channels
    temb = boc[0] * 4
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def resnet(p, cin, cout):
        out.append((f"{p}.norm1", (cin,), "norm"))
        out.append((f"{p}.conv1", (cout, cin, 3, 3), "conv"))
        out.append((f"{p}.time_emb_proj", (cout, temb), "linear"))
        out.append((f"{p}.norm2", (cout,), "norm"))
        out.append((f"{p}.conv2", (cout, cout, 3, 3), "conv"))
        if cin != cout:
            out.append((f"{p}.conv_shortcut", (cout, cin, 1, 1), "conv"))

    def transformer(p, c):
        x = spec.cross_attention_dim
        out.append((f"{p}.norm", (c,), "norm"))
        out.append((f"{p}.proj_in", (c, c), "linear"))
        b = f"{p}.transformer_blocks.0"
        out.append((f"{b}.norm1", (c,), "norm"))
        for n in ("to_q", "to_k", "to_v"):
            out.append((f"{b}.attn1.{n}", (c, c), "linear_nobias"))
        out.append((f"{b}.attn1.to_out.0", (c, c), "linear"))
        out.append((f"{b}.norm2", (c,), "norm"))
        out.append((f"{b}.attn2.to_q", (c, c), "linear_nobias"))
        out.append((f"{b}.attn2.to_k", (c, x), "linear_nobias"))
        out.append((f"{b}.attn2.to_v", (c, x), "linear_nobias"))
        out.append((f"{b}.attn2.to_out.0", (c, c), "linear"))
        out.append((f"{b}.norm3", (c,), "norm"))
        out.append((f"{b}.ff.net.0.proj", (8 * c, c), "linear"))
        out.append((f"{b}.ff.net.2", (c, 4 * c), "linear"))
        out.append((f"{p}.proj_out", (c, c), "linear"))

    out.append(("conv_in", (boc[0], spec.in_channels, 3, 3), "conv"))
    out.append(("time_embedding.linear_1", (temb, boc[0]), "linear"))
    out.append(("time_embedding.linear_2", (temb, temb), "linear"))
    ch = boc[0]
    for i, c in enumerate(boc):
        for j in range(spec.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", ch if j == 0 else c, c)
        if spec.down_has_attn[i]:
            for j in range(spec.layers_per_block):
                transformer(f"down_blocks.{i}.attentions.{j}", c)
        if i != len(boc) - 1:
            out.append((f"down_blocks.{i}.downsamplers.0.conv", (c, c, 3, 3), "conv"))
        ch = c
    resnet("mid_block.resnets.0", boc[-1], boc[-1])
    transformer("mid_block.attentions.0", boc[-1])
    resnet("mid_block.resnets.1", boc[-1], boc[-1])
    rboc = list(reversed(boc))
    prev = rboc[0]
    n = spec.layers_per_block + 1
    for i, c in enumerate(rboc):
        cin = rboc[min(i + 1, len(boc) - 1)]
        for j in range(n):
            skip = cin if j == n - 1 else c
            rin = prev if j == 0 else c
            resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, c)
        if spec.up_has_attn[i]:
            for j in range(n):
                transformer(f"up_blocks.{i}.attentions.{j}", c)
        if i != len(boc) - 1:
            out.append((f"up_blocks.{i}.upsamplers.0.conv", (c, c, 3, 3), "conv"))
        prev = c
    out.append(("conv_norm_out", (boc[0],), "norm"))
    out.append(("conv_out", (spec.out_channels, boc[0], 3, 3), "conv"))
    return out


def synthetic_unet_state_dict(spec: UNetSpec | None = None, seed: int = 0, lora_rank: int = 0, qk_gain: float = 1.25,
                              lora_b_std: float = 0.02, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    spec = spec or UNetSpec()
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for mod, shape, kind in unet_parameter_shapes(spec):
        if kind == "norm":
            sd[f"{mod}.weight"] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
            sd[f"{mod}.bias"] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
            continue
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        std = fan_in ** -0.5
        if mod.endswith((".to_q", ".to_k")):
            std *= qk_gain
        w = (torch.randn(shape, generator=g) * std).to(dtype)
        b = None if kind == "linear_nobias" else (0.1 * torch.randn(shape[0], generator=g)).to(dtype)
        if lora_rank > 0 and _is_lora_target(mod):
            sd[f"{mod}.base_layer.weight"] = w
            if b is not None:
                sd[f"{mod}.base_layer.bias"] = b
            a_shape = (lora_rank,) + tuple(shape[1:])
            b_shape = (shape[0], lora_rank) + ((1, 1) if len(shape) == 4 else ())
            sd[f"{mod}.lora_A.default.weight"] = (torch.randn(a_shape, generator=g) / lora_rank).to(dtype)
            sd[f"{mod}.lora_B.default.weight"] = (torch.randn(b_shape, generator=g) * lora_b_std).to(dtype)
        else:
            sd[f"{mod}.weight"] = w
            if b is not None:
                sd[f"{mod}.bias"] = b
    return sd


def synthetic_caption(cross_dim: int = 1024, tokens: int = 77, seed: int = 42) -> torch.Tensor:
    """Stand-in for the constant CLIP encoding of the fixed prompt (reference pix2pix_turbo.py:100-106)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, tokens, cross_dim, generator=g)


def synthetic_latents(batch: int, n_ref: int, size: int = 64, seed: int = 1234):
    """Degraded-image latent, reference-image latents and the two DDPM noises (fp32 NCHW, CPU)."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(batch, 4, size, size, generator=g) * 0.8
    refs = torch.randn(batch, n_ref, 4, size, size, generator=g) * 0.8
    noise_main = torch.randn(batch, 4, size, size, generator=g)
    noise_ref = torch.randn(batch * n_ref, 4, size, size, generator=g)
    return enc, refs, noise_main, noise_ref


VAE_LORA_TARGETS = ("conv1", "conv2", "conv_in", "conv_shortcut", "conv", "conv_out", "to_k", "to_q", "to_v", "to_out.0")
# reference pix2pix_turbo.py:150-153


def vae_parameter_shapes(block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, latent_channels: int = 4,
                         use_shortcuts: bool = False):
    """(module path, weight shape, kind) for AutoencoderKL (sd-vae-ft-mse layout) incl. the reference's optional
    decoder.skip_conv_{1..4} (pix2pix_turbo.py:46-52)."""
    boc = tuple(block_out_channels)
    rboc = list(reversed(boc))
    out = []

    def resnet(p, cin, cout):
        out.append((f"{p}.norm1", (cin,), "norm"))
        out.append((f"{p}.conv1", (cout, cin, 3, 3), "conv"))
        out.append((f"{p}.norm2", (cout,), "norm"))
        out.append((f"{p}.conv2", (cout, cout, 3, 3), "conv"))
        if cin != cout:
            out.append((f"{p}.conv_shortcut", (cout, cin, 1, 1), "conv"))

    def mid(p, c):
        a = f"{p}.attentions.0"
        out.append((f"{a}.group_norm", (c,), "norm"))
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            out.append((f"{a}.{n}", (c, c), "linear"))
        resnet(f"{p}.resnets.0", c, c)
        resnet(f"{p}.resnets.1", c, c)

    out.append(("encoder.conv_in", (boc[0], 3, 3, 3), "conv"))
    ch = boc[0]
    for i, c in enumerate(boc):
        for j in range(layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", ch if j == 0 else c, c)
        if i != len(boc) - 1:
            out.append((f"encoder.down_blocks.{i}.downsamplers.0.conv", (c, c, 3, 3), "conv"))
        ch = c
    mid("encoder.mid_block", boc[-1])
    out.append(("encoder.conv_norm_out", (boc[-1],), "norm"))
    out.append(("encoder.conv_out", (2 * latent_channels, boc[-1], 3, 3), "conv"))
    out.append(("decoder.conv_in", (boc[-1], latent_channels, 3, 3), "conv"))
    mid("decoder.mid_block", boc[-1])
    prev = rboc[0]
    for i, c in enumerate(rboc):
        for j in range(layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i != len(boc) - 1:
            out.append((f"decoder.up_blocks.{i}.upsamplers.0.conv", (c, c, 3, 3), "conv"))
        prev = c
    out.append(("decoder.conv_norm_out", (boc[0],), "norm"))
    out.append(("decoder.conv_out", (3, boc[0], 3, 3), "conv"))
    if use_shortcuts:
        out.append(("decoder.skip_conv_1", (rboc[0], boc[2], 1, 1), "conv_nobias"))
        out.append(("decoder.skip_conv_2", (rboc[0], boc[1], 1, 1), "conv_nobias"))
        out.append(("decoder.skip_conv_3", (rboc[1], boc[0], 1, 1), "conv_nobias"))
        out.append(("decoder.skip_conv_4", (rboc[2], boc[0], 1, 1), "conv_nobias"))
    out.append(("quant_conv", (2 * latent_channels, 2 * latent_channels, 1, 1), "conv"))
    out.append(("post_quant_conv", (latent_channels, latent_channels, 1, 1), "conv"))
    return out


def synthetic_vae_state_dict(block_out_channels=(128, 256, 512, 512), seed: int = 100, lora_rank: int = 0,
                             use_shortcuts: bool = False, latent_channels: int = 4, lora_b_std: float = 0.02
                             ) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    targets = VAE_LORA_TARGETS + (("skip_conv_1", "skip_conv_2", "skip_conv_3", "skip_conv_4") if use_shortcuts else ())
    for mod, shape, kind in vae_parameter_shapes(block_out_channels, 2, latent_channels, use_shortcuts):
        if kind == "norm":
            sd[f"{mod}.weight"] = 1.0 + 0.1 * torch.randn(shape, generator=g)
            sd[f"{mod}.bias"] = 0.1 * torch.randn(shape, generator=g)
            continue
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        w = torch.randn(shape, generator=g) * fan_in ** -0.5
        b = None if kind == "conv_nobias" else 0.1 * torch.randn(shape[0], generator=g)
        if mod == "quant_conv":                      # moderate posterior std
            w[latent_channels:] *= 0.1
            b[latent_channels:] = -3.0
        if mod == "decoder.conv_out":
            w *= 0.5
        if mod.startswith("decoder.skip_conv"):
            w *= 0.1
        leaf = mod.rsplit(".", 1)[-1] if not mod.endswith("to_out.0") else "to_out.0"
        wrap = lora_rank > 0 and (leaf in targets) and mod not in ("quant_conv", "post_quant_conv")
        if wrap:
            sd[f"{mod}.base_layer.weight"] = w
            if b is not None:
                sd[f"{mod}.base_layer.bias"] = b
            a_shape = (lora_rank,) + tuple(shape[1:])
            b_shape = (shape[0], lora_rank) + ((1, 1) if len(shape) == 4 else ())
            sd[f"{mod}.lora_A.vae_skip.weight"] = torch.randn(a_shape, generator=g) / lora_rank
            sd[f"{mod}.lora_B.vae_skip.weight"] = torch.randn(b_shape, generator=g) * lora_b_std
        else:
            sd[f"{mod}.weight"] = w
            if b is not None:
                sd[f"{mod}.bias"] = b
    return sd


def synthetic_images(batch: int, n_ref: int, size: int = 512, seed: int = 4321):
    """Degraded image (B,3,S,S) and reference images (B,N,3,S,S) in [-1, 1]: smooth random fields + noise, fp16."""
    g = torch.Generator().manual_seed(seed)

    def img(*lead):
        low = torch.randn(*lead, 3, size // 8, size // 8, generator=g)
        x = torch.nn.functional.interpolate(low.flatten(0, -4), size=(size, size), mode="bilinear", align_corners=False)
        x = x + 0.1 * torch.randn(x.shape, generator=g)
        return (0.6 * x).clamp(-1, 1).reshape(*lead, 3, size, size).to(torch.float16)

    return img(batch), img(batch, n_ref)
