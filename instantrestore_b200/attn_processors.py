"""Drop-in attention processors with the reference's API (face_replace/models/attn_processors.py), running on the
B200 kernels. Same class names, constructor arguments, attributes and call signature as the reference:

    SharedAttnProcessor(self_attn_idx=None, save_self_attentions=False, use_adain=False, train_input=True)   (:186)
    processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
              ref_keys=None, ref_values=None) -> Tensor shaped/typed like hidden_states                      (:193-279)
    AttnProcessor() with .keys / .values / .is_self_attn / .reset()                                           (:22-97)
    adain(content_features, style_mean, style_std)                                                            (:7-18)
    register_attention_processor(unet, cfg, save_self_attentions=False)                                       (:282-321)
    register_attention_processor_kv_unet(unet)                                                                (:324-331)

`attn` is the diffusers-style Attention module the reference UNet hands to its processors (to_q/to_k/to_v/to_out,
heads, scale ...). The q/k/v/out projections run on ir_conv_gemm, softmax(QK^T)V on ir_shared_attn_fwd with the
reference keys/values consumed in their native (B, N, S, C) layout (no head split, no torch.cat) and AdaIN folded
into the kernel through ir_adain_coeffs. Inputs must be CUDA tensors on a B200; there is no CPU path here (the CPU
restatement lives in oracle/ and is test infrastructure).

`save_self_attentions=True` keeps the reference behaviour of exposing `self.attention_probs` (B, H, S, S_k) after a
forward (:258-260); the fused kernel never builds that matrix, so it is recomputed by a GEMM + row softmax
(attn_probs.py) — diagnostic use only. `save_reference_mass=True` (new, cheap) stores `self.reference_mass`
(B, H, n_chunks): the softmax mass per KV chunk averaged over queries, which is what gradio_demo.py:118-133 derives
from the dense matrix.

`FaceIDAttnProcessor(hidden_size, self_attn_idx=None, cross_attention_dim=None, embed_dim=512)` (:100-180) is provided
as a drop-in too (face embeddings -> face_projection -> to_k/to_v_face_embed -> attention), although every released
config keeps condition_on_face_embeds False and the whole-step engine (pipeline.py) does not take face embeddings.

Not provided on this path: attention masks, attn.group_norm / spatial_norm / norm_cross (all None for the SD-Turbo
UNet).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L

ADAIN_EPS = 1e-5


def adain(content_features: torch.Tensor, style_mean: torch.Tensor, style_std: torch.Tensor) -> torch.Tensor:
    """Reference :7-18 on (batch*heads, tokens, 64) tensors, evaluated as the affine the kernels use:
    scale = style_std / (std(content) + eps), shift = style_mean - mean(content) * scale."""
    c_mean = content_features.float().mean(dim=1, keepdim=True)
    c_std = content_features.float().std(dim=1, keepdim=True) + ADAIN_EPS
    scale = style_std.float() / c_std
    return (content_features.float() * scale + (style_mean.float() - c_mean * scale)).to(content_features.dtype)


def _effective_weight(mod):
    """Weight a projection module actually applies, fp32, plus the parameters it was built from.

    In the reference's main UNet every to_q/to_k/to_v/to_out.0 is a peft LoRA layer (pix2pix_turbo.py:171-179) whose
    `.weight` property returns only `base_layer.weight`; the module computes base(x) + scaling * B(A(x)), so the weight
    to run is W + sum_adapters scaling * B @ A (adapters already merged into the base, or disabled, contribute nothing)."""
    base = getattr(mod, "base_layer", None)
    if base is None:
        return mod.weight.detach().float(), [mod.weight]
    params = [base.weight]
    w = base.weight.detach().float()
    lora_a, lora_b = getattr(mod, "lora_A", None), getattr(mod, "lora_B", None)
    if lora_a is None or lora_b is None:
        raise NotImplementedError(f"{type(mod).__name__}: adapter wrapper without lora_A / lora_B is not supported")
    if getattr(mod, "merged", False) or getattr(mod, "disable_adapters", False):
        return w, params
    names = getattr(mod, "active_adapters", None)
    if callable(names):
        names = names()
    if not names:
        one = getattr(mod, "adapter", None)
        names = [one] if one is not None else list(lora_a.keys())
    if isinstance(names, str):
        names = [names]
    for name in names:
        if name not in lora_a:
            continue
        a, b = lora_a[name].weight, lora_b[name].weight
        scaling = mod.scaling[name] if isinstance(mod.scaling, dict) else mod.scaling
        w = w + float(scaling) * (b.detach().float().flatten(1) @ a.detach().float().flatten(1))
        params += [a, b]
    return w, params


def _module_bias(mod):
    b = getattr(getattr(mod, "base_layer", mod), "bias", None)
    return b


class _ProjCache:
    """fp16 copies of an Attention module's projection weights (the reference keeps fp32 parameters and lets
    autocast cast them on every call, test.py:82-83), LoRA deltas merged; rebuilt when any parameter they were built
    from (base weight, LoRA A / B, bias) is replaced or modified in place."""

    def __init__(self):
        self.key = None
        self.w = {}

    @staticmethod
    def _sources(attn):
        mods = {"q": attn.to_q, "k": attn.to_k, "v": attn.to_v, "o": attn.to_out[0]}
        params = []
        for m in mods.values():
            base = getattr(m, "base_layer", m)
            params.append(base.weight)
            if getattr(base, "bias", None) is not None:
                params.append(base.bias)
            for d in (getattr(m, "lora_A", None), getattr(m, "lora_B", None)):
                if d is not None:
                    params += [l.weight for l in d.values()]
            params.append(bool(getattr(m, "merged", False)))
        return mods, params

    def get(self, attn, self_attention: bool):
        mods, params = self._sources(attn)
        key = tuple((p.data_ptr(), p._version) if isinstance(p, torch.Tensor) else p for p in params) + (self_attention,)
        if key != self.key:
            h = lambda t: t.to(torch.float16).contiguous()
            w = {name: h(_effective_weight(m)[0]) for name, m in mods.items()}
            for name in ("q", "k", "v", "o"):
                b = _module_bias(mods[name])
                w[name + "b"] = None if b is None else b.detach().float().contiguous()
            if self_attention:
                w["qkv"] = torch.cat([w["q"], w["k"], w["v"]], 0).contiguous()
                w["qkvb"] = None
                if any(w[n + "b"] is not None for n in ("q", "k", "v")):      # a bias on any of the three projections
                    zeros = lambda n: torch.zeros(w[n].shape[0], dtype=torch.float32, device=w[n].device)
                    w["qkvb"] = torch.cat([w[n + "b"] if w[n + "b"] is not None else zeros(n) for n in ("q", "k", "v")]).contiguous()
            self.w, self.key = w, key
        return self.w


def _check(attn, hidden_states, attention_mask):
    if not hidden_states.is_cuda:
        raise RuntimeError("instantrestore_b200 processors run on CUDA (B200) tensors only; there is no CPU fallback")
    if attention_mask is not None:
        raise NotImplementedError("attention masks are never passed on the reference hot path")
    if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "group_norm", None) is not None or getattr(attn, "norm_cross", None):
        raise NotImplementedError("spatial_norm / group_norm / norm_cross are None for the SD-Turbo UNet")
    if getattr(attn.to_q, "base_layer", attn.to_q).weight.shape[0] != attn.heads * 64:
        raise NotImplementedError("the B200 attention kernel is specialised for head_dim 64")


def _as_tokens(hidden_states):
    shape4 = None
    if hidden_states.ndim == 4:
        shape4 = hidden_states.shape
        b, c, h, w = shape4
        hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
    b, s, c = hidden_states.shape
    return hidden_states.to(torch.float16).reshape(b * s, c).contiguous(), b, s, c, shape4


def _finish(attn, out2d, w, b, s, residual, shape4, dtype):
    out = L.conv_gemm(out2d, w["o"], batch=1, h_in=1, w_in=out2d.shape[0], c_in=w["o"].shape[1], bias=w["ob"])
    out = out.view(b, s, -1)
    if shape4 is not None:
        out = out.transpose(-1, -2).reshape(*shape4)
    out = out.to(dtype)
    if attn.residual_connection:
        out = out + residual
    return out / attn.rescale_output_factor


def _project(x2d, w, bias):
    return L.conv_gemm(x2d, w, batch=1, h_in=1, w_in=x2d.shape[0], c_in=w.shape[1], bias=bias)


class AttnProcessor(nn.Module):
    def __init__(self):
        super().__init__()
        self._cache = _ProjCache()
        self.reset()

    def reset(self):
        self.keys, self.values, self.is_self_attn = None, None, None

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        _check(attn, hidden_states, attention_mask)
        residual, dtype = hidden_states, hidden_states.dtype
        x, b, s, c, shape4 = _as_tokens(hidden_states)
        self.is_self_attn = encoder_hidden_states is None
        w = self._cache.get(attn, self.is_self_attn)
        if self.is_self_attn:
            qkv = _project(x, w["qkv"], w["qkvb"])
            inner = w["q"].shape[0]
            q, k, v, s_kv = qkv, qkv[:, inner:2 * inner], qkv[:, 2 * inner:], s
        else:
            ctx = encoder_hidden_states.to(torch.float16)
            s_kv = ctx.shape[1]
            ctx = ctx.reshape(-1, ctx.shape[-1]).contiguous()
            q, k, v = _project(x, w["q"], w["qb"]), _project(ctx, w["k"], w["kb"]), _project(ctx, w["v"], w["vb"])
        self.keys, self.values = k.reshape(b, s_kv, -1), v.reshape(b, s_kv, -1)      # reference :74
        o = L.shared_attn(q, heads=attn.heads, scale=attn.scale, batch=b, s_q=s, k_own=k, v_own=v, s_own=s_kv)
        return _finish(attn, o, w, b, s, residual, shape4, dtype)


class FaceIDAttnProcessor(nn.Module):
    """Reference :100-180: keys/values come from face-recognition embeddings (B, n_faces, embed_dim) through
    face_projection -> to_k_face_embed / to_v_face_embed; the parameters live on the processor (state_dict keys
    `<attn>.processor.{face_projection,to_k_face_embed,to_v_face_embed}.*`)."""

    def __init__(self, hidden_size, self_attn_idx=None, cross_attention_dim=None, embed_dim: int = 512):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.self_attn_idx = hidden_size, cross_attention_dim, self_attn_idx
        width = cross_attention_dim or hidden_size
        self.face_projection = nn.Linear(embed_dim, width)
        self.to_k_face_embed = nn.Linear(width, hidden_size, bias=False)
        self.to_v_face_embed = nn.Linear(width, hidden_size, bias=False)
        self._cache = _ProjCache()
        self.reset()

    def reset(self):
        self.keys, self.values, self.is_self_attn = None, None, None

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, ref_keys=None,
                ref_values=None):
        _check(attn, hidden_states, attention_mask)
        residual, dtype = hidden_states, hidden_states.dtype
        x, b, s, c, shape4 = _as_tokens(hidden_states)
        self.is_self_attn = encoder_hidden_states is None
        w = self._cache.get(attn, False)
        if self.is_self_attn:        # reference :150-151: the (tokenised) hidden states are their own context
            ctx, s_kv = x, s
        else:
            ctx = encoder_hidden_states.to(torch.float16)
            s_kv = ctx.shape[1]
            ctx = ctx.reshape(-1, ctx.shape[-1]).contiguous()
        h = lambda t: t.detach().to(torch.float16).contiguous()
        f = lambda t: None if t is None else t.detach().float().contiguous()
        q = _project(x, w["q"], w["qb"])
        face = _project(ctx, h(self.face_projection.weight), f(self.face_projection.bias))
        k = _project(face, h(self.to_k_face_embed.weight), None)
        v = _project(face, h(self.to_v_face_embed.weight), None)
        o = L.shared_attn(q, heads=attn.heads, scale=attn.scale, batch=b, s_q=s, k_own=k, v_own=v, s_own=s_kv)
        return _finish(attn, o, w, b, s, residual, shape4, dtype)


class SharedAttnProcessor(nn.Module):
    def __init__(self, self_attn_idx: int = None, save_self_attentions: bool = False, use_adain: bool = False,
                 train_input: bool = True):
        super().__init__()
        self.self_attn_idx = self_attn_idx
        self.save_self_attentions = save_self_attentions
        self.use_adain = use_adain
        self.train_input = train_input
        self.save_reference_mass = False
        self.attention_probs = None
        self.reference_mass = None
        self._cache = _ProjCache()

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, ref_keys=None,
                ref_values=None):
        _check(attn, hidden_states, attention_mask)
        residual, dtype = hidden_states, hidden_states.dtype
        x, b, s, c, shape4 = _as_tokens(hidden_states)
        is_self = encoder_hidden_states is None
        w = self._cache.get(attn, is_self)
        inner = w["q"].shape[0]
        if is_self:
            qkv = _project(x, w["qkv"], w["qkvb"])
            q, k, v, s_kv = qkv, qkv[:, inner:2 * inner], qkv[:, 2 * inner:], s
        else:
            ctx = encoder_hidden_states.to(torch.float16)
            s_kv = ctx.shape[1]
            ctx = ctx.reshape(-1, ctx.shape[-1]).contiguous()
            q, k, v = _project(x, w["q"], w["qb"]), _project(ctx, w["k"], w["kb"]), _project(ctx, w["v"], w["vb"])
        kw = dict(k_own=k, v_own=v, s_own=s_kv)
        chunks = [(k, 0, s_kv, s_kv, 0)]
        has_refs = False
        if self.self_attn_idx is not None and ref_keys is not None and ref_values is not None:
            has_refs = True
            rk = ref_keys[self.self_attn_idx].to(torch.float16).contiguous()        # (B, N, S_ref, C)
            rv = ref_values[self.self_attn_idx].to(torch.float16).contiguous()
            n_ref, s_ref = rk.shape[1], rk.shape[2]
            rk2, rv2 = rk.view(-1, inner), rv.view(-1, inner)
            if not self.train_input:
                kw, chunks = {}, []
            kw.update(k_ref=rk2, v_ref=rv2, n_ref=n_ref, s_ref=s_ref)
            chunks += [(rk2, 0, s_ref, n_ref * s_ref, r * s_ref) for r in range(n_ref)]
            if self.use_adain:
                sc, sh = L.adain_coeffs(v, rv2, batch=b, s_own=s_kv, n_ref=n_ref, s_ref=s_ref, channels=inner, eps=ADAIN_EPS)
                kw.update(adain_scale=sc, adain_shift=sh)
        if self.save_reference_mass and has_refs:
            o, self.reference_mass = L.shared_attn(q, heads=attn.heads, scale=attn.scale, batch=b, s_q=s, chunk_mass=True, **kw)
        else:
            o = L.shared_attn(q, heads=attn.heads, scale=attn.scale, batch=b, s_q=s, **kw)
        if self.save_self_attentions:                                                  # reference :258-260
            from .attn_probs import dense_attention_probs
            self.attention_probs = dense_attention_probs(q, 0, chunks, batch=b, heads=attn.heads, s_q=s, scale=attn.scale)
        return _finish(attn, o, w, b, s, residual, shape4, dtype)


def _hidden_size(unet, name):
    boc = unet.config.block_out_channels
    if name.startswith("mid_block"):
        return boc[-1]
    if name.startswith("up_blocks"):
        return list(reversed(boc))[int(name[len("up_blocks.")])]
    return boc[int(name[len("down_blocks.")])]


def register_attention_processor(unet, cfg, save_self_attentions: bool = False):
    """Same numbering as the reference (:282-321): only `up_blocks.*.attn1` get a self_attn_idx (0..8, module order)."""
    procs, idx = {}, 0
    for name in unet.attn_processors.keys():
        shared_layer = name.endswith("attn1.processor") and name.startswith("up_blocks")
        if not name.endswith("attn1.processor") and getattr(cfg, "condition_on_face_embeds", False):   # :296-300
            procs[name] = FaceIDAttnProcessor(self_attn_idx=None, hidden_size=_hidden_size(unet, name),
                                              cross_attention_dim=unet.config.cross_attention_dim, embed_dim=512
                                              ).to(unet.device, dtype=unet.dtype)
            continue
        procs[name] = SharedAttnProcessor(self_attn_idx=idx if shared_layer else None,
                                          save_self_attentions=save_self_attentions and name.endswith("attn1.processor"),
                                          use_adain=cfg.use_adain, train_input=cfg.train_input)
        idx += int(shared_layer)
    unet.set_attn_processor(procs)


def register_attention_processor_kv_unet(unet):
    procs = {}
    for name, current in unet.attn_processors.items():
        procs[name] = AttnProcessor() if (name.startswith("up_blocks") and "attn1" in name) else current
    unet.set_attn_processor(procs)
