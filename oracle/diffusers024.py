"""ORACLE (test infrastructure, never on the product path): CPU restatement of the diffusers==0.24.0 leaf modules
the reference hot path instantiates.

The reference (snap-research/InstantRestore) pins diffusers 0.24.0 (environment.yaml:59); that package is NOT vendored
under /root/reference and is not installed here, so its published module semantics are restated below in plain
PyTorch. Each class cites the reference call site that reaches it. Attribute names match diffusers so that
reference-layout state_dicts load strict=True and so that the reference's own unet.py/block.py/attn_processors.py can
run on top of these classes through oracle/shim (see oracle/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn


# ------------------------------------------------------------------------------------------------ embeddings
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    """diffusers.models.embeddings.get_timestep_embedding (reached from reference unet.py:305,932)."""
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                      downscale_freq_shift=self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    """reference unet.py:312-318."""

    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None):
        super().__init__()
        assert post_act_fn is None and cond_proj_dim is None
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = get_activation(act_fn)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


def get_activation(act_fn: str) -> nn.Module:
    act_fn = act_fn.lower()
    if act_fn in ("swish", "silu"):
        return nn.SiLU()
    if act_fn == "mish":
        return nn.Mish()
    if act_fn == "gelu":
        return nn.GELU()
    if act_fn == "relu":
        return nn.ReLU()
    raise ValueError(f"Unsupported activation function: {act_fn}")


# ------------------------------------------------------------------------------------------------ attention
class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention — the object API the reference processors touch
    (face_replace/models/attn_processors.py:44-95, 205-277)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, out_bias=True, scale_qk=True,
                 only_cross_attention=False, rescale_output_factor=1.0, residual_connection=False, processor=None,
                 norm_num_groups=None, eps=1e-5):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.dropout = dropout
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.only_cross_attention = only_cross_attention
        self.added_kv_proj_dim = None
        self.group_norm = (nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else DefaultAttnProcessor())

    def set_processor(self, processor, _remove_lora=False):
        if (hasattr(self, "processor") and isinstance(self.processor, nn.Module)
                and not isinstance(processor, nn.Module)):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self, return_deprecated_lora=False):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def batch_to_head_dim(self, tensor):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size // head_size, head_size, seq_len, dim)
        return tensor.permute(0, 2, 1, 3).reshape(batch_size // head_size, seq_len, dim * head_size)

    def head_to_batch_dim(self, tensor, out_dim=3):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size, seq_len, head_size, dim // head_size)
        tensor = tensor.permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(batch_size * head_size, seq_len, dim // head_size)
        return tensor

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query = query.float()
            key = key.float()
        if attention_mask is None:
            baddbmm_input = torch.zeros(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype,
                                        device=query.device)  # diffusers uses torch.empty with beta=0
            beta = 0
        else:
            baddbmm_input = attention_mask
            beta = 1
        attention_scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            attention_scores = attention_scores.float()
        attention_probs = attention_scores.softmax(dim=-1)
        return attention_probs.to(dtype)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return attention_mask
        raise NotImplementedError("attention masks are never passed on the reference hot path")


class DefaultAttnProcessor:
    """diffusers AttnProcessor (default; replaced on every layer of the main UNet by attn_processors.py:282-321 and
    kept on the non-captured layers of the reference-KV UNet, attn_processors.py:324-331)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 **_ignored):
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        query = attn.head_to_batch_dim(query)
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        attention_probs = attn.get_attention_scores(query, key, attention_mask)
        hidden_states = torch.bmm(attention_probs, value)
        hidden_states = attn.batch_to_head_dim(hidden_states)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states / attn.rescale_output_factor


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu"
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    """diffusers.models.attention.BasicTransformerBlock: cross_attention_kwargs are splatted into BOTH attn1 and
    attn2 — hence ref_keys/ref_values in every reference processor signature (attn_processors.py:124-125,200-201)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", attention_bias=False, only_cross_attention=False, upcast_attention=False,
                 norm_eps=1e-5, **_unused):
        super().__init__()
        self.only_cross_attention = only_cross_attention
        self.norm1 = nn.LayerNorm(dim, eps=norm_eps)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias,
                               cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention)
        self.norm2 = nn.LayerNorm(dim, eps=norm_eps)
        self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                               dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                               upcast_attention=upcast_attention)
        self.norm3 = nn.LayerNorm(dim, eps=norm_eps)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        norm_hidden_states = self.norm1(hidden_states)
        cross_attention_kwargs = cross_attention_kwargs.copy() if cross_attention_kwargs is not None else {}
        cross_attention_kwargs.pop("gligen", None)
        attn_output = self.attn1(norm_hidden_states,
                                 encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                                 attention_mask=attention_mask, **cross_attention_kwargs)
        hidden_states = attn_output + hidden_states
        norm_hidden_states = self.norm2(hidden_states)
        attn_output = self.attn2(norm_hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 attention_mask=encoder_attention_mask, **cross_attention_kwargs)
        hidden_states = attn_output + hidden_states
        norm_hidden_states = self.norm3(hidden_states)
        ff_output = self.ff(norm_hidden_states)
        return ff_output + hidden_states


class Transformer2DModel(nn.Module):
    """diffusers.models.transformer_2d.Transformer2DModel, continuous input (constructed at reference
    block.py:682,1076,2254)."""

    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 attention_type="default", **_unused):
        super().__init__()
        self.use_linear_projection = use_linear_projection
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention)
            for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None, return_dict=True):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask,
                                  encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        return (output,)


# ------------------------------------------------------------------------------------------------ resnet / sampling
class ResnetBlock2D(nn.Module):
    """diffusers.models.resnet.ResnetBlock2D, time_embedding_norm="default" (reference block.py:1061,2239,2397)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", output_scale_factor=1.0, use_in_shortcut=None, **_unused):
        super().__init__()
        assert time_embedding_norm == "default"
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        self.skip_time_act = skip_time_act
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = get_activation(non_linearity)
        self.use_in_shortcut = in_channels != out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = (nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
                              if self.use_in_shortcut else None)

    def forward(self, input_tensor, temb=None, scale: float = 1.0):
        hidden_states = self.norm1(input_tensor)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.conv1(hidden_states)
        if self.time_emb_proj is not None and temb is not None:
            if not self.skip_time_act:
                temb = self.nonlinearity(temb)
            temb = self.time_emb_proj(temb)[:, :, None, None]
            hidden_states = hidden_states + temb
        hidden_states = self.norm2(hidden_states)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.dropout(hidden_states)
        hidden_states = self.conv2(hidden_states)
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + hidden_states) / self.output_scale_factor


class Downsample2D(nn.Module):
    """3x3 stride-2 conv, padding 1 (reference block.py:1106)."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        assert use_conv
        self.channels = channels
        self.out_channels = out_channels or channels
        self.padding = padding
        self.conv = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, scale: float = 1.0):
        if self.padding == 0:
            hidden_states = F.pad(hidden_states, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(hidden_states)


class Upsample2D(nn.Module):
    """nearest x2 then 3x3 conv (reference block.py:2282,2414)."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        assert use_conv and not use_conv_transpose
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        return self.conv(hidden_states)


def fourier_filter(x_in, threshold, scale):
    """diffusers.utils.torch_utils.fourier_filter (imported at reference block.py:19, used at :3514,:3518)."""
    x = x_in
    B, C, H, W = x.shape
    if (W & (W - 1)) != 0 or (H & (H - 1)) != 0:
        x = x.to(dtype=torch.float32)
    x_freq = torch.fft.fftn(x, dim=(-2, -1))
    x_freq = torch.fft.fftshift(x_freq, dim=(-2, -1))
    B, C, H, W = x_freq.shape
    mask = torch.ones((B, C, H, W), device=x.device)
    crow, ccol = H // 2, W // 2
    mask[..., crow - threshold: crow + threshold, ccol - threshold: ccol + threshold] = scale
    x_freq = x_freq * mask
    x_freq = torch.fft.ifftshift(x_freq, dim=(-2, -1))
    x_filtered = torch.fft.ifftn(x_freq, dim=(-2, -1)).real
    return x_filtered.to(dtype=x_in.dtype)


# ------------------------------------------------------------------------------------------------ scheduler
class DDPMScheduler1Step:
    """The slice of diffusers.DDPMScheduler used by reference models/model.py:4-12 and pix2pix_turbo.py:250-251,
    277,310-311,331: scaled-linear betas 0.00085 -> 0.012 over 1000 steps (sd-turbo scheduler config), epsilon
    prediction. `clip_sample` defaults to False as in sd-turbo's scheduler_config.json (inherited from SD-2.1);
    when True the predicted x0 is clamped to [-1, 1] like DDPMScheduler.step does."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, clip_sample=False):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.clip_sample = clip_sample

    def coeffs(self, t: int):
        a = self.alphas_cumprod[int(t)]
        return float(a ** 0.5), float((1.0 - a) ** 0.5)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        while sa.ndim < original_samples.ndim:
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def scale_model_input(self, sample, timestep=None):
        return sample

    def pred_original_sample(self, model_output, t: int, sample):
        a = self.alphas_cumprod[int(t)].to(sample.dtype)
        x0 = (sample - (1 - a) ** 0.5 * model_output) / a ** 0.5
        if self.clip_sample:
            x0 = x0.clamp(-1.0, 1.0)
        return x0


# ------------------------------------------------------------------------------------------------ LoRA (peft 0.10.0)
class LoraLinear(nn.Module):
    """peft.tuners.lora.Linear: base(x) + lora_B(lora_A(x)) * (alpha / r); key layout <mod>.base_layer.*,
    <mod>.lora_A.<adapter>.weight, <mod>.lora_B.<adapter>.weight (reference pix2pix_turbo.py:158-188)."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, adapter: str = "default"):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({adapter: nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter: nn.Linear(r, base.out_features, bias=False)})
        self.scaling = alpha / r
        self.adapter = adapter

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x, *args, **kwargs):
        return self.base_layer(x) + self.lora_B[self.adapter](self.lora_A[self.adapter](x)) * self.scaling


class LoraConv2d(nn.Module):
    """peft.tuners.lora.Conv2d: A = conv(k x k, C_in -> r, same stride/padding), B = 1x1 (r -> C_out)."""

    def __init__(self, base: nn.Conv2d, r: int, alpha: float, adapter: str = "default"):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({adapter: nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride,
                                                        base.padding, bias=False)})
        self.lora_B = nn.ModuleDict({adapter: nn.Conv2d(r, base.out_channels, (1, 1), (1, 1), bias=False)})
        self.scaling = alpha / r
        self.adapter = adapter

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x, *args, **kwargs):
        return self.base_layer(x) + self.lora_B[self.adapter](self.lora_A[self.adapter](x)) * self.scaling


def add_lora(model: nn.Module, target_modules, r: int, alpha: float, adapter: str = "default",
             generator: Optional[torch.Generator] = None, b_std: float = 0.0) -> None:
    """peft get_peft_model/add_adapter restated: wraps every Linear/Conv2d whose dotted name ends with one of
    `target_modules` (peft suffix match). init_lora_weights="gaussian": A ~ N(0, 1/r), B = 0 (b_std > 0 gives a
    non-trivial synthetic B for parity tests)."""
    names = [n for n, m in model.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))
             and any(n == t or n.endswith("." + t) for t in target_modules)]
    for name in names:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        base = getattr(parent, leaf)
        wrapped = (LoraLinear if isinstance(base, nn.Linear) else LoraConv2d)(base, r, alpha, adapter)
        nn.init.normal_(wrapped.lora_A[adapter].weight, std=1.0 / r, generator=generator)
        if b_std > 0:
            nn.init.normal_(wrapped.lora_B[adapter].weight, std=b_std, generator=generator)
        else:
            nn.init.zeros_(wrapped.lora_B[adapter].weight)
        if isinstance(parent, (nn.ModuleList, nn.Sequential)):
            parent[int(leaf)] = wrapped
        else:
            setattr(parent, leaf, wrapped)
