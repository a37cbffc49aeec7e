"""ORACLE (test infrastructure): CPU restatement of the VAE on either side of the UNet.

The reference builds `AutoencoderKL.from_pretrained("stabilityai/sd-vae-ft-mse")` twice (vae, original_vae;
face_replace/models/pix2pix_turbo.py:39,58), replaces the encoder/decoder forwards with
face_replace/models/model.py:15-31 (`my_vae_encoder_fwd`: records the input of every down block) and :34-63
(`my_vae_decoder_fwd`: optional 1x1 skip convolutions from those activations when cfg.use_shortcuts), LoRA-wraps the
`vae` with adapter "vae_skip" (pix2pix_turbo.py:150-162) and uses them at :245 (reference images), :291 (degraded
image) and :333 (decode, clamp to [-1, 1]).  diffusers 0.24.0 is not vendored by the reference, so AutoencoderKL /
Encoder / Decoder / UNetMidBlock2D / DownEncoderBlock2D / UpDecoderBlock2D / the deprecated-attention-block
`Attention(heads=1, dim_head=512, norm_num_groups=32, residual_connection=True, bias=True)` are restated here from
their published semantics with the sd-vae-ft-mse config: block_out_channels (128, 256, 512, 512),
layers_per_block 2, latent_channels 4, norm_num_groups 32, act silu, scaling_factor 0.18215.
Module names equal the diffusers ones so reference-layout state_dicts (`net.vae.*`, `net.original_vae.*`) load
strict=True.  "Parity unpinned" for these third-party leaves (see oracle/__init__.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from .diffusers024 import Attention, Downsample2D, ResnetBlock2D, Upsample2D


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    use_shortcuts: bool = False

    @staticmethod
    def tiny(width: int = 64) -> "VaeConfig":
        return VaeConfig(block_out_channels=(width, width, 2 * width, 2 * width))


class VaeAttnProcessor:
    """diffusers AttnProcessor on the (B, C, H, W) input of the VAE mid block: group norm, single head, softmax in
    fp32 (upcast_softmax), residual connection."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **_ignored):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = hidden_states.view(b, c, h * w).transpose(1, 2)
        x = attn.group_norm(x.transpose(1, 2)).transpose(1, 2)
        q = attn.head_to_batch_dim(attn.to_q(x))
        k = attn.head_to_batch_dim(attn.to_k(x))
        v = attn.head_to_batch_dim(attn.to_v(x))
        probs = attn.get_attention_scores(q, k, None)
        x = attn.batch_to_head_dim(torch.bmm(probs, v))
        x = attn.to_out[1](attn.to_out[0](x))
        x = x.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            x = x + residual
        return x / attn.rescale_output_factor


class UNetMidBlock2D(nn.Module):
    def __init__(self, ch: int, groups: int):
        super().__init__()
        mk = lambda: ResnetBlock2D(in_channels=ch, out_channels=ch, temb_channels=None, eps=1e-6, groups=groups)
        self.attentions = nn.ModuleList([Attention(ch, heads=1, dim_head=ch, rescale_output_factor=1.0, eps=1e-6,
                                                   norm_num_groups=groups, residual_connection=True, bias=True,
                                                   upcast_softmax=True, processor=VaeAttnProcessor())])
        self.resnets = nn.ModuleList([mk(), mk()])

    def forward(self, hidden_states, temb=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, layers, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_ch if i == 0 else out_ch, out_channels=out_ch,
                                                    temb_channels=None, eps=1e-6, groups=groups) for i in range(layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_ch, use_conv=True, out_channels=out_ch, padding=0, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states):
        for r in self.resnets:
            hidden_states = r(hidden_states, None)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
        return hidden_states


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_ch, out_ch, layers, groups, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_ch if i == 0 else out_ch, out_channels=out_ch,
                                                    temb_channels=None, eps=1e-6, groups=groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch, use_conv=True, out_channels=out_ch)]) if add_upsample else None

    def forward(self, hidden_states, temb=None):
        for r in self.resnets:
            hidden_states = r(hidden_states, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class Encoder(nn.Module):
    def __init__(self, cfg: VaeConfig):
        super().__init__()
        boc = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, c in enumerate(boc):
            inp, out = out, c
            self.down_blocks.append(DownEncoderBlock2D(inp, out, cfg.layers_per_block, cfg.norm_num_groups, i != len(boc) - 1))
        self.mid_block = UNetMidBlock2D(boc[-1], cfg.norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[-1], num_groups=cfg.norm_num_groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)
        self.current_down_blocks = None

    def forward(self, sample):
        """reference models/model.py:15-31 (my_vae_encoder_fwd)."""
        sample = self.conv_in(sample)
        l_blocks = []
        for down_block in self.down_blocks:
            l_blocks.append(sample)
            sample = down_block(sample)
        sample = self.mid_block(sample)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        sample = self.conv_out(sample)
        self.current_down_blocks = l_blocks
        return sample


class Decoder(nn.Module):
    def __init__(self, cfg: VaeConfig):
        super().__init__()
        boc = cfg.block_out_channels
        rboc = list(reversed(boc))
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(boc[-1], cfg.norm_num_groups)
        self.up_blocks = nn.ModuleList()
        out = rboc[0]
        for i, c in enumerate(rboc):
            prev, out = out, c
            self.up_blocks.append(UpDecoderBlock2D(prev, out, cfg.layers_per_block + 1, cfg.norm_num_groups, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=cfg.norm_num_groups, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)
        self.ignore_skip = not cfg.use_shortcuts
        self.gamma = 1
        self.incoming_skip_acts = None
        if cfg.use_shortcuts:
            # reference pix2pix_turbo.py:46-52: 512->512, 256->512, 128->512, 128->256 for sd-vae-ft-mse. In general the
            # reversed encoder activations have (boc[2], boc[1], boc[0], boc[0]) channels (model.py:19-23,45) and
            # are added to the INPUT of up block i, which has (rboc[0], rboc[0], rboc[1], rboc[2]) channels.
            self.skip_conv_1 = nn.Conv2d(boc[2], rboc[0], 1, bias=False)
            self.skip_conv_2 = nn.Conv2d(boc[1], rboc[0], 1, bias=False)
            self.skip_conv_3 = nn.Conv2d(boc[0], rboc[1], 1, bias=False)
            self.skip_conv_4 = nn.Conv2d(boc[0], rboc[2], 1, bias=False)

    def forward(self, sample, latent_embeds=None):
        """reference models/model.py:34-63 (my_vae_decoder_fwd)."""
        sample = self.conv_in(sample)
        sample = self.mid_block(sample, latent_embeds)
        if not self.ignore_skip:
            skip_convs = [self.skip_conv_1, self.skip_conv_2, self.skip_conv_3, self.skip_conv_4]
            for idx, up_block in enumerate(self.up_blocks):
                skip_in = skip_convs[idx](self.incoming_skip_acts[::-1][idx] * self.gamma)
                sample = sample + skip_in
                sample = up_block(sample, latent_embeds)
        else:
            for up_block in self.up_blocks:
                sample = up_block(sample, latent_embeds)
        sample = self.conv_norm_out(sample)
        sample = self.conv_act(sample)
        return self.conv_out(sample)


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: Optional[VaeConfig] = None):
        super().__init__()
        cfg = cfg or VaeConfig()
        self.config = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def encode_moments(self, x):
        """AutoencoderKL.encode(x).latent_dist parameters: (mean, std) with logvar clamped to [-30, 20]."""
        moments = self.quant_conv(self.encoder(x))
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        return mean, torch.exp(0.5 * logvar)

    def encode_sample(self, x, eps):
        """latent_dist.sample() with the normal draw injected: mean + std * eps."""
        mean, std = self.encode_moments(x)
        return mean + std * eps

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


VAE_LORA_TARGETS = ["conv1", "conv2", "conv_in", "conv_shortcut", "conv", "conv_out", "to_k", "to_q", "to_v", "to_out.0"]
# reference pix2pix_turbo.py:150-153 (+ skip_conv_1..4 when cfg.use_shortcuts, :154-155)
