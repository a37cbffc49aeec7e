"""ORACLE (test infrastructure): CPU restatement of the reference UNet graph.

Follows reference face_replace/models/unet_2d_condition/unet.py:174-626 (construction), :628-686 (processor
registry), :772-794 (enable_freeu), :804-1179 (forward) and the live block classes of block.py:
UNetMidBlock2DCrossAttn :631-774, CrossAttnDownBlock2D :1024-1182, DownBlock2D :1185-1270,
CrossAttnUpBlock2D :2198-2368, UpBlock2D :2371-2478, apply_freeu :3495-3520 — restricted to the SD-Turbo
configuration the reference instantiates (pix2pix_turbo.py:17,56,60). Module/attribute names equal the diffusers
ones so reference-layout state_dicts load strict=True.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from .diffusers024 import (Downsample2D, ResnetBlock2D, Timesteps, TimestepEmbedding, Transformer2DModel, Upsample2D,
                           fourier_filter)


@dataclass
class UNetConfig:
    """SD-Turbo (SD-2.1 layout) values; `attention_head_dim` holds HEAD COUNTS (reference unet.py:239-245)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: Tuple[int, ...] = (5, 10, 20, 20)
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = True
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    sample_size: int = 64
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                         "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")

    @staticmethod
    def tiny(width: int = 64, cross_dim: int = 128, sample_size: int = 32) -> "UNetConfig":
        """Reduced widths (head_dim stays 64) for fast oracle/GPU parity cases."""
        return UNetConfig(block_out_channels=(width, 2 * width, 4 * width, 4 * width),
                          attention_head_dim=(width // 64, 2 * width // 64, 4 * width // 64, 4 * width // 64),
                          cross_attention_dim=cross_dim, sample_size=sample_size)


def apply_freeu(resolution_idx, hidden_states, res_hidden_states, s1, s2, b1, b2):
    """reference block.py:3495-3520."""
    dtype = res_hidden_states.dtype
    if resolution_idx == 0:
        half = hidden_states.shape[1] // 2
        hidden_states[:, :half] = hidden_states[:, :half] * b1
        res_hidden_states = fourier_filter(res_hidden_states.float(), threshold=1, scale=s1).to(dtype)
    if resolution_idx == 1:
        half = hidden_states.shape[1] // 2
        hidden_states[:, :half] = hidden_states[:, :half] * b2
        res_hidden_states = fourier_filter(res_hidden_states.float(), threshold=1, scale=s2).to(dtype)
    return hidden_states, res_hidden_states


class _FreeUMixin:
    s1 = s2 = b1 = b2 = None

    def _freeu(self, hidden_states, res_hidden_states):
        if self.s1 and self.s2 and self.b1 and self.b2:
            return apply_freeu(self.resolution_idx, hidden_states, res_hidden_states, self.s1, self.s2, self.b1, self.b2)
        return hidden_states, res_hidden_states


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, cfg: UNetConfig, in_ch, out_ch, temb_ch, heads, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList()
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(in_channels=in_ch if i == 0 else out_ch, out_channels=out_ch,
                                              temb_channels=temb_ch, eps=cfg.norm_eps, groups=cfg.norm_num_groups))
            self.attentions.append(Transformer2DModel(heads, out_ch // heads, in_channels=out_ch, num_layers=1,
                                                      cross_attention_dim=cfg.cross_attention_dim,
                                                      norm_num_groups=cfg.norm_num_groups,
                                                      use_linear_projection=cfg.use_linear_projection))
        self.downsamplers = (nn.ModuleList([Downsample2D(out_ch, use_conv=True, out_channels=out_ch, padding=1, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb, encoder_hidden_states, cross_attention_kwargs):
        output_states = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, cfg: UNetConfig, in_ch, out_ch, temb_ch, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_ch if i == 0 else out_ch, out_channels=out_ch, temb_channels=temb_ch,
                          eps=cfg.norm_eps, groups=cfg.norm_num_groups) for i in range(cfg.layers_per_block)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_ch, use_conv=True, out_channels=out_ch, padding=1, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, cfg: UNetConfig, ch, temb_ch, heads):
        super().__init__()
        mk = lambda: ResnetBlock2D(in_channels=ch, out_channels=ch, temb_channels=temb_ch, eps=cfg.norm_eps,
                                   groups=cfg.norm_num_groups)
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, in_channels=ch, num_layers=1,
                                                            cross_attention_dim=cfg.cross_attention_dim,
                                                            norm_num_groups=cfg.norm_num_groups,
                                                            use_linear_projection=cfg.use_linear_projection)])
        self.resnets = nn.ModuleList([mk(), mk()])

    def forward(self, hidden_states, temb, encoder_hidden_states, cross_attention_kwargs):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UpBlock2D(nn.Module, _FreeUMixin):
    has_cross_attention = False

    def __init__(self, cfg: UNetConfig, in_ch, prev_out_ch, out_ch, temb_ch, add_upsample, resolution_idx):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resolution_idx = resolution_idx
        self.resnets = nn.ModuleList()
        for i in range(n):
            res_skip = in_ch if i == n - 1 else out_ch
            res_in = prev_out_ch if i == 0 else out_ch
            self.resnets.append(ResnetBlock2D(in_channels=res_in + res_skip, out_channels=out_ch,
                                              temb_channels=temb_ch, eps=cfg.norm_eps, groups=cfg.norm_num_groups))
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch, use_conv=True, out_channels=out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states, res_hidden_states = self._freeu(hidden_states, res_hidden_states)
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module, _FreeUMixin):
    has_cross_attention = True

    def __init__(self, cfg: UNetConfig, in_ch, prev_out_ch, out_ch, temb_ch, heads, add_upsample, resolution_idx):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resolution_idx = resolution_idx
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList()
        for i in range(n):
            res_skip = in_ch if i == n - 1 else out_ch
            res_in = prev_out_ch if i == 0 else out_ch
            self.resnets.append(ResnetBlock2D(in_channels=res_in + res_skip, out_channels=out_ch,
                                              temb_channels=temb_ch, eps=cfg.norm_eps, groups=cfg.norm_num_groups))
            self.attentions.append(Transformer2DModel(heads, out_ch // heads, in_channels=out_ch, num_layers=1,
                                                      cross_attention_dim=cfg.cross_attention_dim,
                                                      norm_num_groups=cfg.norm_num_groups,
                                                      use_linear_projection=cfg.use_linear_projection))
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch, use_conv=True, out_channels=out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states, cross_attention_kwargs):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states, res_hidden_states = self._freeu(hidden_states, res_hidden_states)
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: Optional[UNetConfig] = None):
        super().__init__()
        cfg = cfg or UNetConfig()
        self.config = cfg
        boc = cfg.block_out_channels
        heads = cfg.attention_head_dim
        temb_ch = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], kernel_size=3, padding=1)
        self.time_proj = Timesteps(boc[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if t == "CrossAttnDownBlock2D":
                self.down_blocks.append(CrossAttnDownBlock2D(cfg, in_ch, out_ch, temb_ch, heads[i], not final))
            else:
                self.down_blocks.append(DownBlock2D(cfg, in_ch, out_ch, temb_ch, not final))
        self.mid_block = UNetMidBlock2DCrossAttn(cfg, boc[-1], temb_ch, heads[-1])
        self.up_blocks = nn.ModuleList()
        rboc, rheads = list(reversed(boc)), list(reversed(heads))
        out_ch = rboc[0]
        self.num_upsamplers = 0
        for i, t in enumerate(cfg.up_block_types):
            final = i == len(boc) - 1
            prev_out, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, len(boc) - 1)]
            if not final:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlock2D":
                self.up_blocks.append(CrossAttnUpBlock2D(cfg, in_ch, prev_out, out_ch, temb_ch, rheads[i], not final, i))
            else:
                self.up_blocks.append(UpBlock2D(cfg, in_ch, prev_out, out_ch, temb_ch, not final, i))
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=cfg.norm_num_groups, eps=cfg.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, kernel_size=3, padding=1)

    # ---- processor registry (reference unet.py:628-686)
    @property
    def attn_processors(self) -> Dict[str, object]:
        procs: Dict[str, object] = {}

        def rec(name, module):
            if hasattr(module, "get_processor"):
                procs[f"{name}.processor"] = module.get_processor(return_deprecated_lora=True)
            for sub, child in module.named_children():
                rec(f"{name}.{sub}", child)

        for name, module in self.named_children():
            rec(name, module)
        return procs

    def set_attn_processor(self, processor, _remove_lora=False):
        count = len(self.attn_processors.keys())
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not "
                             f"match the number of attention layers: {count}.")

        def rec(name, module):
            if hasattr(module, "set_processor"):
                module.set_processor(processor if not isinstance(processor, dict) else processor.pop(f"{name}.processor"))
            for sub, child in module.named_children():
                rec(f"{name}.{sub}", child)

        for name, module in self.named_children():
            rec(name, module)

    def enable_freeu(self, s1, s2, b1, b2):
        for blk in self.up_blocks:
            blk.s1, blk.s2, blk.b1, blk.b2 = s1, s2, b1, b2

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    # ---- forward (reference unet.py:804-1179, the branches SD-Turbo takes)
    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        sample = self.conv_in(sample)
        down_res = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(sample, emb, encoder_hidden_states, cross_attention_kwargs)
            else:
                sample, res = blk(sample, emb)
            down_res += res
        sample = self.mid_block(sample, emb, encoder_hidden_states, cross_attention_kwargs)
        for blk in self.up_blocks:
            res = down_res[-len(blk.resnets):]
            down_res = down_res[:-len(blk.resnets)]
            if blk.has_cross_attention:
                sample = blk(sample, res, emb, encoder_hidden_states, cross_attention_kwargs)
            else:
                sample = blk(sample, res, emb)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        return self.conv_out(sample)
