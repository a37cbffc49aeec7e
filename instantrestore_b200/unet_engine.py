"""B200 execution of the reference UNet forward (reference unet_2d_condition/unet.py:804-1179 and the live blocks of
block.py) on the C-ABI kernels. Activations stay channel-last fp16 end to end ([B*H*W, C] is both "NHWC" for the
convolutions and the token matrix of the transformer blocks, so the reference's permute/reshape pairs vanish).

Folded once per checkpoint (results unchanged): LoRA merged into base weights; the timestep is a
constant of the pipeline (249 main / 1 reference, reference test.py:62, pix2pix_turbo.py:247) so
time_embedding(...) and every time_emb_proj(silu(emb)) vector become part of conv1's bias; the caption embedding is
a constant (pix2pix_turbo.py:100-106) so the cross-attention K/V projections are precomputed.

No CPU fallback: everything here calls instantrestore_b200._lib, which raises if the CUDA library is missing.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .weights import StateDictView, conv_weight_khwc, geglu_interleave_index, upsample_conv_weight


FUSED_ADAIN_STATS = os.environ.get("IR_FUSED_ADAIN", "1") != "0"     # A-B measurement switch
FOLD_UPSAMPLE = os.environ.get("IR_FOLD_UPSAMPLE", "1") != "0"       # A-B measurement switch (also read by vae_engine)
UP_FOLD_MIN_ROWS = int(os.environ.get("IR_UP_FOLD_MIN_ROWS", "0"))   # measured ahead on every shape (tools/up_bench.py)


@dataclass
class UNetSpec:
    """Geometry of the SD-Turbo UNet (reference unet.py:182-199 defaults + sd-turbo config);
    attention_head_dim holds head COUNTS (unet.py:239-245), head_dim is 64 everywhere."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: Tuple[int, ...] = (5, 10, 20, 20)
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)


@dataclass
class RefKV:
    """Keys/values of the reference images for one shared-attention layer: a [B*N*S, row] fp16 buffer holding K at
    column k_off and V at column v_off (the reference-UNet's fused QKV projection output is used in place)."""
    buf: torch.Tensor
    k_off: int
    v_off: int
    n_ref: int
    s_ref: int
    v_partial: Optional[torch.Tensor] = None    # [B*N*S/32, C, 2] slab moments of V from the projection's epilogue (AdaIN)


class _Lin:
    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], dev):
        self.w = w.to(torch.float16).contiguous().to(dev)
        self.b = None if b is None else b.to(torch.float32).contiguous().to(dev)
        self.c_out, self.c_in = self.w.shape


class _Conv:
    def __init__(self, w4: torch.Tensor, b: Optional[torch.Tensor], dev, stride: int = 1, c_in_pad: int = 0, upsample: bool = False):
        self.ksize = w4.shape[-1]
        self.stride = stride
        self.c_in = max(w4.shape[1], c_in_pad)
        self.c_out = w4.shape[0]
        self.w = conv_weight_khwc(w4, c_in_pad).to(dev)
        # Upsample2D's convolution: also the phase-folded weights of the four 2x2 sub-pixel convolutions (ir_conv_gemm upsample2x)
        self.w_up = upsample_conv_weight(w4).to(dev) if upsample and FOLD_UPSAMPLE else None
        self.b = None if b is None else b.to(torch.float32).contiguous().to(dev)


class _Norm:
    def __init__(self, v: StateDictView, name: str, dev):
        self.g = v.param(f"{name}.weight").contiguous().to(dev)
        self.b = v.param(f"{name}.bias").contiguous().to(dev)


def timestep_embedding(t: int, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): [cos | sin] of t * exp(-ln(1e4) i / (dim/2))."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = float(t) * freqs
    return torch.cat([torch.cos(arg), torch.sin(arg)])[None]


class UNetEngine:
    def __init__(self, sd: StateDictView, spec: UNetSpec, timestep: int, caption_enc: torch.Tensor, device,
                 *, use_adain: bool = False, train_input: bool = True, consume_refs: bool = False,
                 capture_kv: bool = False, freeu: Optional[Tuple[float, float, float, float]] = (0.9, 0.2, 1.4, 1.6),
                 emit_v_stats: bool = False):
        L.load()
        self.spec, self.dev = spec, torch.device(device)
        self.use_adain, self.train_input = use_adain, train_input
        self.consume_refs, self.capture_kv = consume_refs, capture_kv
        # reference UNet of an AdaIN model: the captured QKV projections also emit the per-slab moments of V in their
        # epilogues (the statistics belong to K/V extraction time and travel with the K/V, RefCache included)
        self.emit_v_stats = emit_v_stats
        self.freeu = freeu
        self.captured: List[RefKV] = []
        # The reference UNet is only run for its 9 captured K/V projections (reference pix2pix_turbo.py:255-266; the
        # prediction itself feeds an image decode nobody reads, :277-278): stop right after the last projection and skip
        # that block's attention / cross-attention / feed-forward, conv_norm_out and conv_out.
        self.stop_after_last_kv = capture_kv
        self.debug: Optional[dict] = None     # set to {} to record per-module outputs (tools/debug_engine.py)
        # opt-in by-products of the 9 shared-attention layers (reference attn_processors.py:258-260; gradio_demo.py:118-133)
        self.save_attention_probs = False     # dense (B, H, S, S_k) per layer -> self.attention_probs
        self.save_reference_mass = False      # (B, H, n_chunks) per layer -> self.reference_mass
        self.attention_probs: List[torch.Tensor] = []
        self.reference_mass: List[torch.Tensor] = []
        dev = self.dev
        boc = spec.block_out_channels
        cap = caption_enc.detach().to(torch.float32).cpu()
        if cap.ndim == 3:
            cap = cap[0]
        self.n_ctx = cap.shape[0]

        # constant time embedding -> per-resnet bias vectors
        te = sd.sub("time_embedding")
        emb = timestep_embedding(timestep, boc[0])
        emb = torch.nn.functional.silu(emb @ te.weight("linear_1").T + te.bias("linear_1"))
        emb = emb @ te.weight("linear_2").T + te.bias("linear_2")
        self._silu_emb = torch.nn.functional.silu(emb)[0]

        # FaceIDAttnProcessor cross-attentions (reference attn_processors.py:100-180; cfg.condition_on_face_embeds):
        # K = to_k_face_embed(face_projection(e)), V = to_v_face_embed(face_projection(e)) depend only on the face
        # embeddings, so the two linears of every cross-attention fold into one [2C, 512] matrix and ALL of them run as
        # ONE GEMM per step ([B*N_f, 512] x [sum 2C, 512]^T); each layer then reads its K | V columns in place.
        self._face_parts: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self._face_cols = 0
        self.face_kv: Optional[_Lin] = None
        self._face: Optional[torch.Tensor] = None
        self._n_face = 0
        self.conv_in = _Conv(sd.weight("conv_in"), sd.bias("conv_in"), dev, c_in_pad=64)
        self.down = []
        for i, ch in enumerate(boc):
            blk = sd.sub(f"down_blocks.{i}")
            layers = []
            for j in range(spec.layers_per_block):
                res = self._load_resnet(blk.sub(f"resnets.{j}"))
                tr = self._load_transformer(blk.sub(f"attentions.{j}"), ch, spec.attention_head_dim[i], cap) if spec.down_has_attn[i] else None
                layers.append((res, tr))
            ds = None
            if i != len(boc) - 1:
                d = blk.sub("downsamplers.0")
                ds = _Conv(d.weight("conv"), d.bias("conv"), dev, stride=2)
            self.down.append((layers, ds))
        mid = sd.sub("mid_block")
        self.mid = (self._load_resnet(mid.sub("resnets.0")),
                    self._load_transformer(mid.sub("attentions.0"), boc[-1], spec.attention_head_dim[-1], cap),
                    self._load_resnet(mid.sub("resnets.1")))
        self.up = []
        rboc = list(reversed(boc))
        rheads = list(reversed(spec.attention_head_dim))
        for i, ch in enumerate(rboc):
            blk = sd.sub(f"up_blocks.{i}")
            layers = []
            for j in range(spec.layers_per_block + 1):
                res = self._load_resnet(blk.sub(f"resnets.{j}"))
                tr = self._load_transformer(blk.sub(f"attentions.{j}"), ch, rheads[i], cap) if spec.up_has_attn[i] else None
                layers.append((res, tr))
            us = None
            if i != len(boc) - 1:
                u = blk.sub("upsamplers.0")
                us = _Conv(u.weight("conv"), u.bias("conv"), dev, upsample=True)
            self.up.append((layers, us))
        self.norm_out = _Norm(sd, "conv_norm_out", dev)
        self.conv_out = _Conv(sd.weight("conv_out"), sd.bias("conv_out"), dev)
        if self._face_parts:
            self.face_kv = _Lin(torch.cat([w for w, _ in self._face_parts], 0), torch.cat([b for _, b in self._face_parts], 0), dev)
        self._face_parts = []

    # ------------------------------------------------------------------------------------------ loading
    def _load_resnet(self, v: StateDictView):
        dev = self.dev
        tvec = self._silu_emb @ v.weight("time_emb_proj").T + v.bias("time_emb_proj")
        conv1 = _Conv(v.weight("conv1"), v.bias("conv1") + tvec, dev)
        conv2 = _Conv(v.weight("conv2"), v.bias("conv2"), dev)
        sc = None
        if v.has("conv_shortcut.weight"):
            w = v.weight("conv_shortcut")
            sc = _Lin(w[:, :, 0, 0], v.bias("conv_shortcut"), dev)
        return dict(norm1=_Norm(v, "norm1", dev), conv1=conv1, norm2=_Norm(v, "norm2", dev), conv2=conv2, shortcut=sc)

    def _load_transformer(self, v: StateDictView, ch: int, heads: int, cap: torch.Tensor):
        dev = self.dev
        b = v.sub("transformer_blocks.0")
        a1, a2, ff = b.sub("attn1"), b.sub("attn2"), b.sub("ff")
        wqkv = torch.cat([a1.weight("to_q"), a1.weight("to_k"), a1.weight("to_v")], 0)
        k2 = (cap @ a2.weight("to_k").T).to(torch.float16).contiguous().to(dev)      # [n_ctx, C], constant
        v2 = (cap @ a2.weight("to_v").T).to(torch.float16).contiguous().to(dev)
        idx = geglu_interleave_index(ff.weight("net.0.proj").shape[0])
        face_off = None
        if a2.has("processor.face_projection.weight"):
            wp, bp = a2.weight("processor.face_projection"), a2.bias("processor.face_projection")     # [W, 512], [W]
            wk, wv = a2.weight("processor.to_k_face_embed"), a2.weight("processor.to_v_face_embed")   # [C, W]
            self._face_parts.append((torch.cat([wk @ wp, wv @ wp], 0), torch.cat([wk @ bp, wv @ bp], 0)))
            face_off = (self._face_cols, self._face_cols + ch)        # column offsets of this layer's K and V
            self._face_cols += 2 * ch
        return dict(
            heads=heads, ch=ch, face_off=face_off,
            norm=_Norm(v, "norm", dev), proj_in=_Lin(v.weight("proj_in"), v.bias("proj_in"), dev),
            ln1=_Norm(b, "norm1", dev), qkv=_Lin(wqkv, None, dev), out1=_Lin(a1.weight("to_out.0"), a1.bias("to_out.0"), dev),
            ln2=_Norm(b, "norm2", dev), q2=_Lin(a2.weight("to_q"), None, dev), k2=k2, v2=v2,
            out2=_Lin(a2.weight("to_out.0"), a2.bias("to_out.0"), dev),
            ln3=_Norm(b, "norm3", dev),
            ff1=_Lin(ff.weight("net.0.proj")[idx], ff.bias("net.0.proj")[idx], dev),
            ff2=_Lin(ff.weight("net.2"), ff.bias("net.2"), dev),
            proj_out=_Lin(v.weight("proj_out"), v.bias("proj_out"), dev),
        )

    # ------------------------------------------------------------------------------------------ ops
    def _lin(self, x, lin: _Lin, residual=None, act=L.IR_ACT_NONE, **kw):
        return L.conv_gemm(x, lin.w, batch=1, h_in=1, w_in=x.shape[0], c_in=lin.c_in, bias=lin.b, residual=residual, act=act, **kw)

    def _conv(self, x, cv: _Conv, B, H, W, residual=None):
        return L.conv_gemm(x, cv.w, batch=B, h_in=H, w_in=W, c_in=cv.c_in, ksize=cv.ksize, stride=cv.stride, bias=cv.b,
                           residual=residual)

    def _upsample_conv(self, x, cv: _Conv, B, H, W):
        """Upsample2D (nearest 2x + 3x3 conv; reference block.py:2366,2476). The upsampled tensor is never written: four 2x2
        sub-pixel convolutions on x (4/9 of the multiply-adds; 1.7-2.3x faster than upsample + conv on every shape of the
        step, tools/up_bench.py). IR_FOLD_UPSAMPLE=0 / IR_UP_FOLD_MIN_ROWS keep the materialised path for A/B runs."""
        if cv.w_up is not None and B * H * W >= UP_FOLD_MIN_ROWS:
            return L.conv_gemm(x, cv.w_up, batch=B, h_in=H, w_in=W, c_in=cv.c_in, ksize=3, bias=cv.b, upsample2x=True)
        return self._conv(L.upsample_nearest2x(x, batch=B, h=H, w=W), cv, B, 2 * H, 2 * W)

    def _gn(self, x, n: _Norm, B, HW, eps, silu):
        return L.groupnorm(x, n.g, n.b, batch=B, hw=HW, groups=self.spec.norm_num_groups, eps=eps, silu=silu)

    def _resnet(self, x, p, B, H, W):
        t = self._gn(x, p["norm1"], B, H * W, self.spec.norm_eps, True)
        h = self._conv(t, p["conv1"], B, H, W)
        t = self._gn(h, p["norm2"], B, H * W, self.spec.norm_eps, True)
        skip = x if p["shortcut"] is None else self._lin(x, p["shortcut"])
        return self._conv(t, p["conv2"], B, H, W, residual=skip)

    def _transformer(self, x, p, B, S, ref: Optional[RefKV] = None, capture: bool = False, stop_after_kv: bool = False):
        C, heads = p["ch"], p["heads"]
        scale = 0.125  # head_dim ** -0.5, head_dim == 64
        t = self._gn(x, p["norm"], B, S, 1e-6, False)
        h = self._lin(t, p["proj_in"])
        # --- attn1: self attention, shared with the reference images in the up blocks
        n = L.layernorm(h, p["ln1"].g, p["ln1"].b)
        shared_layer = self.consume_refs and ref is not None
        # AdaIN statistics of V (mean / M2 per 32-token slab and channel) ride in the epilogue of this projection
        want_stats = (capture and self.emit_v_stats) or (shared_layer and self.use_adain and ref.v_partial is not None)
        part = None
        if want_stats and FUSED_ADAIN_STATS and L.col_partial_supported(B * S, 3 * C, 2 * C):
            part = torch.empty((B * S // 32, C, 2), dtype=torch.float32, device=x.device)
        qkv = self._lin(n, p["qkv"], col_partial=part, col_begin=2 * C)    # [B*S, 3C]: q | k | v
        if capture:
            self.captured.append(RefKV(buf=qkv, k_off=C, v_off=2 * C, n_ref=0, s_ref=S, v_partial=part))
            if stop_after_kv:
                return None        # nothing downstream of the last captured K/V is ever read (see forward_up)
        kw = {}
        own = True
        if shared_layer:
            kw.update(k_ref=ref.buf[:, ref.k_off:], v_ref=ref.buf[:, ref.v_off:], n_ref=ref.n_ref, s_ref=ref.s_ref)
            if self.use_adain:
                if part is not None and ref.v_partial is not None:
                    sc, sh = L.adain_coeffs(None, None, batch=B, s_own=S, n_ref=ref.n_ref, s_ref=ref.s_ref, channels=C,
                                            own_partial=part, ref_partial=ref.v_partial)
                else:
                    sc, sh = L.adain_coeffs(qkv[:, 2 * C:], ref.buf[:, ref.v_off:], batch=B, s_own=S, n_ref=ref.n_ref,
                                            s_ref=ref.s_ref, channels=C)
                kw.update(adain_scale=sc, adain_shift=sh)
            own = self.train_input
        if own:
            kw.update(k_own=qkv[:, C:], v_own=qkv[:, 2 * C:], s_own=S)
        if shared_layer and self.save_reference_mass:
            a, mass = L.shared_attn(qkv, heads=heads, scale=scale, batch=B, s_q=S, chunk_mass=True, **kw)
            self.reference_mass.append(mass)
        else:
            a = L.shared_attn(qkv, heads=heads, scale=scale, batch=B, s_q=S, **kw)
        if shared_layer and self.save_attention_probs:
            from .attn_probs import dense_attention_probs
            chunks = [(qkv, C, S, S, 0)] if own else []
            chunks += [(ref.buf, ref.k_off, ref.s_ref, ref.n_ref * ref.s_ref, r * ref.s_ref) for r in range(ref.n_ref)]
            self.attention_probs.append(dense_attention_probs(qkv, 0, chunks, batch=B, heads=heads, s_q=S, scale=scale))
        h = self._lin(a, p["out1"], residual=h)
        # --- attn2: cross attention against the constant caption K/V
        n = L.layernorm(h, p["ln2"].g, p["ln2"].b)
        q = self._lin(n, p["q2"])
        if p["face_off"] is not None:       # keys / values from the face embeddings of this identity (N_f tokens)
            if self._face is None:
                raise ValueError("this checkpoint was trained with condition_on_face_embeds: forward(face_embeds=...) is required")
            a = L.shared_attn(q, heads=heads, scale=scale, batch=B, s_q=S, k_own=self._face, v_own=self._face,
                              k_own_col_off=p["face_off"][0], v_own_col_off=p["face_off"][1], s_own=self._n_face)
        else:
            a = L.shared_attn(q, heads=heads, scale=scale, batch=B, s_q=S, k_own=p["k2"], v_own=p["v2"], s_own=self.n_ctx,
                              own_shared=True)
        h = self._lin(a, p["out2"], residual=h)
        # --- feed-forward (GEGLU fused into the first GEMM's epilogue)
        n = L.layernorm(h, p["ln3"].g, p["ln3"].b)
        g = self._lin(n, p["ff1"], act=L.IR_ACT_GEGLU)
        h = self._lin(g, p["ff2"], residual=h)
        return self._lin(h, p["proj_out"], residual=x)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, B: int, H: int, W: int, ref_kv: Optional[Sequence[RefKV]] = None,
                face_embeds: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: fp16 channel-last latent [B*H*W, 64] (4 real channels). Returns the model output [B*H*W, 4] fp16."""
        return self.forward_up(self.forward_down_mid(x, B, H, W, face_embeds), ref_kv)

    def forward_down_mid(self, x: torch.Tensor, B: int, H: int, W: int, face_embeds: Optional[torch.Tensor] = None):
        """conv_in, down blocks and mid block (reference unet.py:1042-1121): independent of the reference K/V, so the
        pipeline runs it on a second stream while the reference UNet is still working.
        face_embeds: fp16 (B, N_f, 512) when the checkpoint's cross-attentions are FaceIDAttnProcessors."""
        self.captured = []
        self._face, self._n_face = None, 0
        if self.face_kv is not None and face_embeds is not None:
            fe = face_embeds.reshape(-1, face_embeds.shape[-1])
            self._n_face = face_embeds.shape[1]
            self._face = self._lin(fe, self.face_kv)        # [B*N_f, sum 2C]: every cross-attention's K | V
        dbg = self.debug
        h = self._conv(x, self.conv_in, B, H, W)
        if dbg is not None:
            dbg["conv_in"] = h
        skips = [(h, H, W)]
        for i, (layers, ds) in enumerate(self.down):
            for j, (res, tr) in enumerate(layers):
                h = self._resnet(h, res, B, H, W)
                if dbg is not None:
                    dbg[f"down_blocks.{i}.resnets.{j}"] = h
                if tr is not None:
                    h = self._transformer(h, tr, B, H * W)
                    if dbg is not None:
                        dbg[f"down_blocks.{i}.attentions.{j}"] = h
                skips.append((h, H, W))
            if ds is not None:
                h = self._conv(h, ds, B, H, W)
                H, W = H // 2, W // 2
                if dbg is not None:
                    dbg[f"down_blocks.{i}.downsamplers.0"] = h
                skips.append((h, H, W))
        r0, tr, r1 = self.mid
        h = self._resnet(h, r0, B, H, W)
        h = self._transformer(h, tr, B, H * W)
        h = self._resnet(h, r1, B, H, W)
        if dbg is not None:
            dbg["mid_block"] = h
        return h, skips, B, H, W

    def forward_up(self, state, ref_kv: Optional[Sequence[RefKV]] = None) -> torch.Tensor:
        """up blocks (shared attention against the reference K/V), conv_norm_out, conv_out (unet.py:1135-1170)."""
        h, skips, B, H, W = state
        skips = list(skips)
        shared_idx = 0
        self.attention_probs, self.reference_mass = [], []
        dbg = self.debug
        for i, (layers, us) in enumerate(self.up):
            bscale, sscale = 1.0, 1.0
            if self.freeu is not None and i < 2:
                s1, s2, b1, b2 = self.freeu
                bscale, sscale = (b1, s1) if i == 0 else (b2, s2)
            for j, (res, tr) in enumerate(layers):
                skip, sh, sw = skips.pop()
                assert (sh, sw) == (H, W)
                cat = L.concat_freeu(h, skip, batch=B, h=H, w=W, backbone_scale=bscale, skip_scale=sscale)
                h = self._resnet(cat, res, B, H, W)
                if dbg is not None:
                    dbg[f"up_blocks.{i}.resnets.{j}"] = h
                if tr is not None:
                    ref = ref_kv[shared_idx] if (self.consume_refs and ref_kv is not None) else None
                    shared_idx += 1
                    last_kv = self.capture_kv and self.stop_after_last_kv and i == len(self.up) - 1 and j == len(layers) - 1
                    h = self._transformer(h, tr, B, H * W, ref, capture=self.capture_kv, stop_after_kv=last_kv)
                    if last_kv:
                        return None
                    if dbg is not None:
                        dbg[f"up_blocks.{i}.attentions.{j}"] = h
            if us is not None:
                h = self._upsample_conv(h, us, B, H, W)
                H, W = 2 * H, 2 * W
                if dbg is not None:
                    dbg[f"up_blocks.{i}.upsamplers.0"] = h
        t = self._gn(h, self.norm_out, B, H * W, self.spec.norm_eps, True)
        return self._conv(t, self.conv_out, B, H, W)
