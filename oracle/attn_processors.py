"""ORACLE (test infrastructure): CPU restatement of the reference attention processors.

Follows reference face_replace/models/attn_processors.py: adain :7-18, AttnProcessor :22-97 (plain attention that
records the projected keys/values of the reference-image UNet), SharedAttnProcessor :183-279 (queries of the degraded
image attend to keys/values concatenated from the reference images), register_attention_processor :282-321 and
register_attention_processor_kv_unet :324-331. FaceIDAttnProcessor (:100-180) is restated for API completeness; the
released configs keep condition_on_face_embeds False.

Semantics kept on purpose (see SURVEY.md 7.0 "quirks"): AdaIN statistics run over the token axis with the UNBIASED
std and eps added to the std; padded reference slots are zero tensors that still take softmax mass;
train_input=False drops the image's own keys/values from the extended set but still uses its values for the AdaIN style.
"""
from __future__ import annotations

import torch
from torch import nn

ADAIN_EPS = 1e-5
# bench.py's `gpu_eager_baseline` only: the same processors with torch's fused scaled_dot_product_attention instead of
# the reference's baddbmm + softmax + bmm (the reference never does this; it is the stronger eager-GPU bar of SURVEY 2b)
USE_SDPA = False


def _attend(attn, q, k, v):
    if USE_SDPA:
        return torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None], scale=attn.scale)[0]
    return torch.bmm(attn.get_attention_scores(q, k, None), v)


def token_stats(x):
    """mean / (unbiased std + eps) over the token axis of a (batch*heads, tokens, dim) tensor."""
    return x.mean(dim=1, keepdim=True), x.std(dim=1, keepdim=True) + ADAIN_EPS


def adain(content_features, style_mean, style_std):
    c_mean, c_std = token_stats(content_features)
    return (content_features - c_mean) / c_std * style_std + style_mean


def _prologue(attn, hidden_states, encoder_hidden_states):
    """Shared head of all three processors: (B,C,H,W) -> (B,HW,C) when needed, q/k/v source selection."""
    assert attn.spatial_norm is None and attn.group_norm is None and not attn.norm_cross
    shape4 = None
    if hidden_states.ndim == 4:
        shape4 = hidden_states.shape
        b, c, h, w = shape4
        hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
    is_self = encoder_hidden_states is None
    context = hidden_states if is_self else encoder_hidden_states
    return hidden_states, context, is_self, shape4


def _epilogue(attn, heads_out, residual, shape4):
    out = attn.batch_to_head_dim(heads_out)
    out = attn.to_out[1](attn.to_out[0](out))
    if shape4 is not None:
        out = out.transpose(-1, -2).reshape(*shape4)
    if attn.residual_connection:
        out = out + residual
    return out / attn.rescale_output_factor


class AttnProcessor(nn.Module):
    """Plain attention; keeps the un-split key/value projections for later sharing (reference :74)."""

    def __init__(self):
        super().__init__()
        self.reset()

    def reset(self):
        self.keys, self.values, self.is_self_attn = None, None, None

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        hidden_states, context, self.is_self_attn, shape4 = _prologue(attn, hidden_states, encoder_hidden_states)
        q = attn.to_q(hidden_states)
        k, v = attn.to_k(context), attn.to_v(context)
        self.keys, self.values = k, v
        q, k, v = (attn.head_to_batch_dim(t) for t in (q, k, v))
        return _epilogue(attn, _attend(attn, q, k, v), residual, shape4)


class FaceIDAttnProcessor(nn.Module):
    def __init__(self, hidden_size, self_attn_idx=None, cross_attention_dim=None, embed_dim: int = 512):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.self_attn_idx = hidden_size, cross_attention_dim, self_attn_idx
        width = cross_attention_dim or hidden_size
        self.face_projection = nn.Linear(embed_dim, width)
        self.to_k_face_embed = nn.Linear(width, hidden_size, bias=False)
        self.to_v_face_embed = nn.Linear(width, hidden_size, bias=False)
        self.reset()

    def reset(self):
        self.keys, self.values, self.is_self_attn = None, None, None

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, ref_keys=None,
                ref_values=None):
        residual = hidden_states
        hidden_states, context, self.is_self_attn, shape4 = _prologue(attn, hidden_states, encoder_hidden_states)
        q = attn.to_q(hidden_states)
        context = self.face_projection(context)
        k, v = self.to_k_face_embed(context), self.to_v_face_embed(context)
        q, k, v = (attn.head_to_batch_dim(t) for t in (q, k, v))
        probs = attn.get_attention_scores(q, k, None)
        return _epilogue(attn, torch.bmm(probs, v), residual, shape4)


class SharedAttnProcessor(nn.Module):
    def __init__(self, self_attn_idx: int = None, save_self_attentions: bool = False, use_adain: bool = False,
                 train_input: bool = True):
        super().__init__()
        self.self_attn_idx = self_attn_idx
        self.save_self_attentions = save_self_attentions
        self.use_adain = use_adain
        self.train_input = train_input

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, ref_keys=None,
                ref_values=None):
        residual = hidden_states
        hidden_states, context, _, shape4 = _prologue(attn, hidden_states, encoder_hidden_states)
        batch = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        k, v = attn.to_k(context), attn.to_v(context)
        q, k, v = (attn.head_to_batch_dim(t) for t in (q, k, v))
        if self.self_attn_idx is not None and ref_keys is not None and ref_values is not None:
            rk, rv = ref_keys[self.self_attn_idx], ref_values[self.self_attn_idx]      # (B, N, S, C)
            rks = [attn.head_to_batch_dim(rk[:, i]) for i in range(rk.shape[1])]
            rvs = [attn.head_to_batch_dim(rv[:, i]) for i in range(rv.shape[1])]
            if self.use_adain:
                style_mean, style_std = token_stats(v)
                rvs = [adain(x, style_mean, style_std) for x in rvs]
            own_k, own_v = ([k], [v]) if self.train_input else ([], [])
            k, v = torch.cat(own_k + rks, dim=1), torch.cat(own_v + rvs, dim=1)
        if not self.save_self_attentions:
            return _epilogue(attn, _attend(attn, q, k, v), residual, shape4)
        probs = attn.get_attention_scores(q, k, None)
        self.attention_probs = probs.reshape(batch, attn.heads, q.shape[1], k.shape[1])
        return _epilogue(attn, torch.bmm(probs, v), residual, shape4)


def _hidden_size(unet, name):
    boc = unet.config.block_out_channels
    if name.startswith("mid_block"):
        return boc[-1]
    if name.startswith("up_blocks"):
        return list(reversed(boc))[int(name[len("up_blocks.")])]
    return boc[int(name[len("down_blocks.")])]


def register_attention_processor(unet, cfg, save_self_attentions: bool = False):
    """Every layer gets a SharedAttnProcessor; only the up-block self-attentions are numbered (0..8, module order)
    and therefore consume reference keys/values."""
    procs, idx = {}, 0
    for name in unet.attn_processors.keys():
        is_cross = not name.endswith("attn1.processor")
        if is_cross and cfg.condition_on_face_embeds:
            procs[name] = FaceIDAttnProcessor(hidden_size=_hidden_size(unet, name), self_attn_idx=None,
                                              cross_attention_dim=unet.config.cross_attention_dim, embed_dim=512)
        elif is_cross:
            procs[name] = SharedAttnProcessor(self_attn_idx=None, use_adain=cfg.use_adain, train_input=cfg.train_input)
        elif name.startswith("up_blocks"):
            procs[name] = SharedAttnProcessor(self_attn_idx=idx, save_self_attentions=save_self_attentions,
                                              use_adain=cfg.use_adain, train_input=cfg.train_input)
            idx += 1
        else:
            procs[name] = SharedAttnProcessor(self_attn_idx=None, save_self_attentions=save_self_attentions,
                                              use_adain=cfg.use_adain, train_input=cfg.train_input)
        procs[name] = procs[name].to(unet.device, dtype=unet.dtype)
    unet.set_attn_processor(procs)


def register_attention_processor_kv_unet(unet):
    procs = {}
    for name, current in unet.attn_processors.items():
        capture = name.startswith("up_blocks") and "attn1" in name
        procs[name] = AttnProcessor().to(unet.device, dtype=unet.dtype) if capture else current
    unet.set_attn_processor(procs)
