"""Checkpoint -> device weight packing for the UNet engine.

Accepts the reference state_dict layout: diffusers module paths, optionally peft-wrapped leaves
(`<mod>.base_layer.weight|bias`, `<mod>.lora_A.<adapter>.weight`, `<mod>.lora_B.<adapter>.weight`; reference
pix2pix_turbo.py:171-179) which are merged at load time as W' = W + (alpha/r) * B.A with alpha = r // 2.
Everything input-independent is folded here, once: conv weights go to [C_out, ky, kx, C_in] fp16 (the K-major B
operand the implicit-GEMM kernel TMA-loads), q/k/v of every self-attention are stacked into one [3C, C] GEMM, GEGLU
rows are interleaved in blocks of 64 so value and gate land in the same accumulator tile, conv_in channels are
zero-padded to 64.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

LORA_ADAPTERS = ("default", "vae_skip")


class StateDictView:
    """Prefix-scoped accessor that merges LoRA on the fly. All math in fp32 on the CPU."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = "", lora_alpha_over_r: Optional[float] = None):
        self.sd = sd
        self.prefix = prefix
        self.lora_scale = lora_alpha_over_r

    def sub(self, name: str) -> "StateDictView":
        return StateDictView(self.sd, f"{self.prefix}{name}.", self.lora_scale)

    def has(self, name: str) -> bool:
        k = self.prefix + name
        return k in self.sd or (self.prefix + name.replace(".weight", ".base_layer.weight").replace(".bias", ".base_layer.bias")) in self.sd

    def _raw(self, key: str) -> Optional[torch.Tensor]:
        t = self.sd.get(self.prefix + key)
        return None if t is None else t.detach().to(torch.float32).cpu()

    def weight(self, mod: str) -> torch.Tensor:
        """Merged fp32 weight of a Linear/Conv2d leaf."""
        w = self._raw(f"{mod}.weight")
        if w is not None:
            return w
        w = self._raw(f"{mod}.base_layer.weight")
        if w is None:
            raise KeyError(f"missing weight for {self.prefix}{mod}")
        for ad in LORA_ADAPTERS:
            a = self._raw(f"{mod}.lora_A.{ad}.weight")
            b = self._raw(f"{mod}.lora_B.{ad}.weight")
            if a is None or b is None:
                continue
            r = a.shape[0]
            scale = self.lora_scale if self.lora_scale is not None else (r // 2) / r
            if w.ndim == 2:
                w = w + scale * (b @ a)
            else:
                w = w + scale * torch.einsum("or,rikl->oikl", b[:, :, 0, 0], a)
        return w

    def bias(self, mod: str) -> Optional[torch.Tensor]:
        b = self._raw(f"{mod}.bias")
        if b is None:
            b = self._raw(f"{mod}.base_layer.bias")
        return b

    def param(self, name: str) -> torch.Tensor:
        t = self._raw(name)
        if t is None:
            raise KeyError(f"missing parameter {self.prefix}{name}")
        return t


def conv_weight_khwc(w: torch.Tensor, c_in_pad: int = 0) -> torch.Tensor:
    """[C_out, C_in, kh, kw] -> [C_out, kh*kw*C_in(_pad)] fp16, tap-major then channel."""
    co, ci, kh, kw = w.shape
    w = w.permute(0, 2, 3, 1)
    if c_in_pad and c_in_pad > ci:
        w = torch.nn.functional.pad(w, (0, c_in_pad - ci))
    return w.reshape(co, -1).contiguous().to(torch.float16)


def patch_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, 3, 3] with 9 * C_in <= 64 -> fp16 [C_out, 64]: (ky, kx, ch) order, zero padded — the weight of a 3x3
    convolution applied to the patches ir_image_in_patches3x3 writes (one 64-wide K block)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3 and 9 * ci <= 64, w.shape
    flat = w.permute(0, 2, 3, 1).reshape(co, 9 * ci)
    return torch.nn.functional.pad(flat, (0, 64 - 9 * ci)).contiguous().to(torch.float16)


def upsample_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """w: [C_out, C_in, 3, 3] of the 3x3 convolution that follows a nearest-2x upsampling (diffusers Upsample2D:
    F.interpolate(scale_factor=2.0, mode="nearest") + conv; reference block.py:2366,2476 and the VAE decoder's upsamplers).
    Returns the phase-folded fp16 matrix [4*C_out, 4*C_in] of ir_conv_gemm_params.upsample2x: for output sub-pixel phase
    (py, px), tap (a, b) reads low-resolution pixel (y + py - 1 + a, x + px - 1 + b) and carries the SUM of the 3x3 taps
    that land on that pixel — rows ky -> a: py=0: {0} -> 0, {1, 2} -> 1;  py=1: {0, 1} -> 0, {2} -> 1 (likewise kx -> b).
    Sums are taken in fp32 and rounded to fp16 once."""
    c_out, c_in = w.shape[:2]
    w = w.to(torch.float32).permute(0, 2, 3, 1)          # [C_out, ky, kx, C_in]
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = w.new_zeros((4, c_out, 2, 2, c_in))
    for py in (0, 1):
        for px in (0, 1):
            for a, kys in enumerate(groups[py]):
                for b, kxs in enumerate(groups[px]):
                    for ky in kys:
                        for kx in kxs:
                            out[py * 2 + px, :, a, b] += w[:, ky, kx]
    return out.reshape(4 * c_out, 4 * c_in).contiguous().to(torch.float16)


def geglu_interleave_index(n_total: int) -> torch.Tensor:
    """Row permutation for the GEGLU projection: per 128-row block, 64 value rows then their 64 gate rows."""
    half = n_total // 2
    assert half % 64 == 0, "GEGLU inner width must be a multiple of 64"
    idx = []
    for blk in range(half // 64):
        idx += list(range(blk * 64, blk * 64 + 64)) + list(range(half + blk * 64, half + blk * 64 + 64))
    return torch.tensor(idx, dtype=torch.long)


def _leaf_shape(sd: Dict[str, torch.Tensor], mod: str):
    t = sd.get(f"{mod}.weight")
    if t is None:
        t = sd.get(f"{mod}.base_layer.weight")
    return None if t is None else tuple(t.shape)


def infer_unet_geometry(sd: Dict[str, torch.Tensor]) -> dict:
    """block_out_channels / head counts / cross_attention_dim of a UNet state_dict in the reference layout (the
    released checkpoints are SD-Turbo: 320/640/1280/1280, head_dim 64, 1024-wide captions). head_dim is 64 in every
    supported model, so heads = channels // 64."""
    boc = []
    i = 0
    while True:
        s = _leaf_shape(sd, f"down_blocks.{i}.resnets.0.conv1")
        if s is None:
            break
        boc.append(int(s[0]))
        i += 1
    if not boc:
        raise KeyError("not a UNet state_dict: down_blocks.0.resnets.0.conv1 is missing")
    cross = None
    for i in range(len(boc)):
        s = _leaf_shape(sd, f"down_blocks.{i}.attentions.0.transformer_blocks.0.attn2.to_k")
        if s is not None:
            cross = int(s[1])
            break
    down_attn = tuple(_leaf_shape(sd, f"down_blocks.{i}.attentions.0.proj_in") is not None for i in range(len(boc)))
    up_attn = tuple(_leaf_shape(sd, f"up_blocks.{i}.attentions.0.proj_in") is not None for i in range(len(boc)))
    return dict(block_out_channels=tuple(boc), attention_head_dim=tuple(max(1, c // 64) for c in boc),
                cross_attention_dim=cross if cross is not None else 1024, down_has_attn=down_attn, up_has_attn=up_attn)


def infer_vae_channels(sd: Dict[str, torch.Tensor]):
    """encoder block_out_channels of an AutoencoderKL state_dict (sd-vae-ft-mse: 128/256/512/512)."""
    out = []
    i = 0
    while True:
        s = _leaf_shape(sd, f"encoder.down_blocks.{i}.resnets.0.conv1")
        if s is None:
            break
        out.append(int(s[0]))
        i += 1
    if not out:
        raise KeyError("not a VAE state_dict: encoder.down_blocks.0.resnets.0.conv1 is missing")
    return tuple(out)
