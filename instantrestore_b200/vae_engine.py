"""B200 execution of the VAE on either side of the UNet (reference pix2pix_turbo.py:245 reference-image encode,
:291 degraded-image encode, :333 decode; patched forwards models/model.py:15-63) on the C-ABI kernels.

AutoencoderKL "stabilityai/sd-vae-ft-mse" geometry (block_out_channels 128/256/512/512, 2 resnets per encoder block,
3 per decoder block, single-head 512-wide mid-block attention, GroupNorm eps 1e-6). Channel-last fp16 activations
throughout. Folded once at load: LoRA (adapter "vae_skip", pix2pix_turbo.py:150-162) into the base weights;
quant_conv (1x1) into encoder.conv_out; post_quant_conv (1x1) into decoder.conv_in — its bias becomes a constant
shift of the latent (W_pq^-1 b_pq) so the zero padding of conv_in stays exact; the value-projection bias of the
mid-block attention moves behind the softmax (rows sum to 1) into the output projection's bias.

Mid-block attention (head_dim 512): scores = (Q C^-1/2) K^T (the scale is folded into the Q projection at load, so
the fp16 scores have the magnitude of the reference's baddbmm(alpha=scale) output under autocast, diffusers
Attention.get_attention_scores) are materialised in fp16 by ir_conv_gemm, softmax-ed in place in fp32
(ir_softmax_rows), and P V runs as a GEMM against V^T, which a GEMM with swapped operands produces directly.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib as L
from .weights import StateDictView, conv_weight_khwc, patch_conv_weight, upsample_conv_weight

GN_EPS = 1e-6
GROUPS = 32
FUSED_GN_STATS = os.environ.get("IR_FUSED_GN", "1") != "0"     # A-B measurement switch
FOLD_UPSAMPLE = os.environ.get("IR_FOLD_UPSAMPLE", "1") != "0"  # A-B measurement switch
PATCH_CONV_IN = os.environ.get("IR_PATCH_CONV_IN", "1") != "0"  # A-B measurement switch
UP_FOLD_MIN_ROWS = int(os.environ.get("IR_UP_FOLD_MIN_ROWS", "0"))   # measured ahead on every shape (tools/up_bench.py)


class _Conv:
    def __init__(self, w4: torch.Tensor, b: Optional[torch.Tensor], dev, stride: int = 1, c_in_pad: int = 0,
                 pad_hi_only: bool = False, upsample: bool = False):
        self.ksize, self.stride, self.pad_hi_only = w4.shape[-1], stride, pad_hi_only
        self.c_in, self.c_out = max(w4.shape[1], c_in_pad), w4.shape[0]
        self.w = conv_weight_khwc(w4, c_in_pad).to(dev)
        # Upsample2D's convolution: also the phase-folded weights of the four 2x2 sub-pixel convolutions (ir_conv_gemm upsample2x)
        self.w_up = upsample_conv_weight(w4).to(dev) if upsample and FOLD_UPSAMPLE else None
        self.b = None if b is None else b.to(torch.float32).contiguous().to(dev)


class _Lin:
    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], dev):
        self.w = w.to(torch.float16).contiguous().to(dev)
        self.b = None if b is None else b.to(torch.float32).contiguous().to(dev)
        self.c_out, self.c_in = self.w.shape


class _Norm:
    def __init__(self, v: StateDictView, name: str, dev):
        self.g = v.param(f"{name}.weight").contiguous().to(dev)
        self.b = v.param(f"{name}.bias").contiguous().to(dev)


class VaeEngine:
    def __init__(self, sd, device, *, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2,
                 latent_channels: int = 4, scaling_factor: float = 0.18215, use_shortcuts: bool = False,
                 encoder: bool = True, decoder: bool = True):
        L.load()
        v = sd if isinstance(sd, StateDictView) else StateDictView(sd)
        self.dev = torch.device(device)
        self.boc = tuple(block_out_channels)
        self.lat = latent_channels
        self.sf = scaling_factor
        self.use_shortcuts = use_shortcuts
        dev = self.dev
        boc = self.boc
        if encoder:
            e = v.sub("encoder")
            self.e_conv_in = _Conv(e.weight("conv_in"), e.bias("conv_in"), dev, c_in_pad=64)
            # conv_in on 3x3 patches gathered by ir_image_in_patches3x3: a K = 64 GEMM (27 useful) instead of K = 576
            self.e_conv_in_patch = None
            if PATCH_CONV_IN and 9 * e.weight("conv_in").shape[1] <= 64:
                self.e_conv_in_patch = _Lin(patch_conv_weight(e.weight("conv_in")), e.bias("conv_in"), dev)
            self.e_down = []
            for i in range(len(boc)):
                blk = e.sub(f"down_blocks.{i}")
                res = [self._load_resnet(blk.sub(f"resnets.{j}")) for j in range(layers_per_block)]
                ds = None
                if i != len(boc) - 1:
                    d = blk.sub("downsamplers.0")
                    ds = _Conv(d.weight("conv"), d.bias("conv"), dev, stride=2, pad_hi_only=True)
                self.e_down.append((res, ds))
            self.e_mid = self._load_mid(e.sub("mid_block"))
            self.e_norm_out = _Norm(e, "conv_norm_out", dev)
            # quant_conv (1x1, 2*lat -> 2*lat) folded into conv_out (3x3, C -> 2*lat)
            wq = v.weight("quant_conv")[:, :, 0, 0]
            w_out = torch.einsum("om,mikl->oikl", wq, e.weight("conv_out"))
            b_out = wq @ e.bias("conv_out") + v.bias("quant_conv")
            self.e_conv_out = _Conv(w_out, b_out, dev)
        if decoder:
            d = v.sub("decoder")
            wpq = v.weight("post_quant_conv")[:, :, 0, 0]
            w_in = torch.einsum("omkl,mi->oikl", d.weight("conv_in"), wpq)
            self.d_conv_in = _Conv(w_in, d.bias("conv_in"), dev, c_in_pad=64)
            if torch.any(wpq):
                self.d_latent_shift = torch.linalg.solve(wpq.double(), v.bias("post_quant_conv").double()).float().to(dev)
            else:   # placeholder weights (dist.placeholder_state_dict): the prepared value arrives by dist.broadcast_engine
                self.d_latent_shift = torch.zeros(wpq.shape[0], dtype=torch.float32, device=dev)
            self.d_mid = self._load_mid(d.sub("mid_block"))
            self.d_up = []
            for i in range(len(boc)):
                blk = d.sub(f"up_blocks.{i}")
                res = [self._load_resnet(blk.sub(f"resnets.{j}")) for j in range(layers_per_block + 1)]
                us = None
                if i != len(boc) - 1:
                    u = blk.sub("upsamplers.0")
                    us = _Conv(u.weight("conv"), u.bias("conv"), dev, upsample=True)
                self.d_up.append((res, us))
            self.d_norm_out = _Norm(d, "conv_norm_out", dev)
            self.d_conv_out = _Conv(d.weight("conv_out"), d.bias("conv_out"), dev)
            self.d_skip = None
            if use_shortcuts:   # reference pix2pix_turbo.py:46-52, model.py:41-49
                self.d_skip = [_Lin(d.weight(f"skip_conv_{k}")[:, :, 0, 0], None, dev) for k in (1, 2, 3, 4)]
        self.skip_acts: Optional[List[torch.Tensor]] = None

    # ------------------------------------------------------------------------------------------ loading
    def _load_resnet(self, v: StateDictView):
        dev = self.dev
        sc = None
        if v.has("conv_shortcut.weight"):
            sc = _Lin(v.weight("conv_shortcut")[:, :, 0, 0], v.bias("conv_shortcut"), dev)
        return dict(norm1=_Norm(v, "norm1", dev), conv1=_Conv(v.weight("conv1"), v.bias("conv1"), dev),
                    norm2=_Norm(v, "norm2", dev), conv2=_Conv(v.weight("conv2"), v.bias("conv2"), dev), shortcut=sc)

    def _load_mid(self, v: StateDictView):
        dev = self.dev
        a = v.sub("attentions.0")
        wq, wk, wv, wo = a.weight("to_q"), a.weight("to_k"), a.weight("to_v"), a.weight("to_out.0")
        bq, bk, bv, bo = a.bias("to_q"), a.bias("to_k"), a.bias("to_v"), a.bias("to_out.0")
        return dict(
            res0=self._load_resnet(v.sub("resnets.0")), res1=self._load_resnet(v.sub("resnets.1")),
            norm=_Norm(a, "group_norm", dev),
            q=_Lin(wq * float(wq.shape[0]) ** -0.5, bq * float(wq.shape[0]) ** -0.5, dev), k=_Lin(wk, bk, dev),   # scale folded into Q
            v_t=wv.to(torch.float16).contiguous().to(dev),           # used as the A operand: V^T = W_v X^T
            out=_Lin(wo, bo + wo @ bv, dev),                            # value bias moved behind the softmax
            ch=wq.shape[0])

    # ------------------------------------------------------------------------------------------ ops
    # A producer whose output feeds a GroupNorm computes that norm's pass A (per-slab moments) in its epilogue
    # (`stats=True` -> returns (out, partial)); the partial travels with the tensor to `_gn`. Shapes the fused
    # statistics do not cover (tiny test geometries) return partial = None and take the three-kernel GroupNorm.
    def _partial(self, x, B, hw_out, c_out):
        if not FUSED_GN_STATS or not L.gn_partial_supported(hw_out, c_out, GROUPS):
            return None
        if L.gn_fused_supported(B, hw_out, c_out, GROUPS):    # the norm is ONE launch with one read: no pass A to move
            return None
        return torch.empty(L.gn_partial_numel(B, hw_out, GROUPS), dtype=torch.float32, device=x.device)

    def _conv(self, x, cv: _Conv, B, H, W, residual=None, stats=False):
        part = self._partial(x, B, (H // cv.stride) * (W // cv.stride), cv.c_out) if stats else None
        out = L.conv_gemm(x, cv.w, batch=B, h_in=H, w_in=W, c_in=cv.c_in, ksize=cv.ksize, stride=cv.stride, bias=cv.b,
                          residual=residual, pad_hi_only=cv.pad_hi_only, gn_partial=part, gn_groups=GROUPS)
        return (out, part) if stats else out

    def _upsample_conv(self, x, cv: _Conv, B, H, W):
        """Upsample2D of the decoder (nearest 2x + 3x3 conv): the upsampled tensor is never written — four 2x2 sub-pixel
        convolutions on the low-resolution x (4/9 of the multiply-adds), GroupNorm pass A of the result from the epilogue."""
        if cv.w_up is None or B * H * W < UP_FOLD_MIN_ROWS:
            return self._conv(L.upsample_nearest2x(x, batch=B, h=H, w=W), cv, B, 2 * H, 2 * W, stats=True)
        part = self._partial(x, B, 4 * H * W, cv.c_out) if (H * W) % 32 == 0 else None
        out = L.conv_gemm(x, cv.w_up, batch=B, h_in=H, w_in=W, c_in=cv.c_in, ksize=3, bias=cv.b, upsample2x=True,
                          gn_partial=part, gn_groups=GROUPS)
        return out, part

    def _lin(self, x, lin: _Lin, residual=None, stats_bhw=None):
        """stats_bhw = (B, hw): also emit the GroupNorm moments of the output, viewed as B images of hw pixels."""
        part = None
        if stats_bhw is not None:
            part = self._partial(x, stats_bhw[0], stats_bhw[1], lin.c_out)
        if part is None:
            out = L.conv_gemm(x, lin.w, batch=1, h_in=1, w_in=x.shape[0], c_in=lin.c_in, bias=lin.b, residual=residual)
        else:   # same GEMM, described as B "images" of 1 x hw tokens so the slabs are per image
            out = L.conv_gemm(x, lin.w, batch=stats_bhw[0], h_in=1, w_in=stats_bhw[1], c_in=lin.c_in, bias=lin.b,
                              residual=residual, gn_partial=part, gn_groups=GROUPS)
        return (out, part) if stats_bhw is not None else out

    def _gn(self, x, n: _Norm, B, HW, silu, part=None):
        return L.groupnorm(x, n.g, n.b, batch=B, hw=HW, groups=GROUPS, eps=GN_EPS, silu=silu, partial_in=part)

    def _resnet(self, x, p, B, H, W, xs=None, stats=True):
        """xs: pass-A moments of x (or None). Returns (out, moments of out) — the next consumer is a GroupNorm everywhere
        except in front of a down/upsampler (stats=False there)."""
        t = self._gn(x, p["norm1"], B, H * W, True, xs)
        h, hs = self._conv(t, p["conv1"], B, H, W, stats=True)
        t = self._gn(h, p["norm2"], B, H * W, True, hs)
        skip = x if p["shortcut"] is None else self._lin(x, p["shortcut"])
        if not stats:
            return self._conv(t, p["conv2"], B, H, W, residual=skip), None
        return self._conv(t, p["conv2"], B, H, W, residual=skip, stats=True)

    def _mid(self, x, p, B, H, W, xs=None):
        S, C = H * W, p["ch"]
        x, xs = self._resnet(x, p["res0"], B, H, W, xs)
        t = self._gn(x, p["norm"], B, S, False, xs)
        q_all, k_all = self._lin(t, p["q"]), self._lin(t, p["k"])           # [B*S, C] each
        attn = torch.empty((B * S, C), dtype=torch.float16, device=x.device)
        for b in range(B):
            rows = slice(b * S, (b + 1) * S)
            q, k, tb = q_all[rows], k_all[rows], t[rows]
            scores = L.conv_gemm(q, k, batch=1, h_in=1, w_in=S, c_in=C)     # [S, S] = (Q / sqrt(C)) K^T, fp16 like baddbmm(alpha)
            L.softmax_rows(scores, 1.0)
            v_t = L.conv_gemm(p["v_t"], tb, batch=1, h_in=1, w_in=C, c_in=C)   # [C, S] = W_v X^T
            L.conv_gemm(scores, v_t, batch=1, h_in=1, w_in=S, c_in=S, out=attn[rows])
        x, xs = self._lin(attn, p["out"], residual=x, stats_bhw=(B, S))
        return self._resnet(x, p["res1"], B, H, W, xs)

    # ------------------------------------------------------------------------------------------ public
    def encode(self, images: torch.Tensor, eps: Optional[torch.Tensor]) -> torch.Tensor:
        """images: (B,3,H,W) fp16/fp32 CUDA in [-1,1]; eps: (B,4,H/8,W/8) fp32 normal draw or None (posterior mode).
        Returns latent_dist.sample() * scaling_factor, fp32 NCHW. Records the skip activations (model.py:19-30)."""
        B, _, H, W = images.shape
        if self.e_conv_in_patch is not None:
            x, xs = self._lin(L.image_in_patches3x3(images.contiguous()), self.e_conv_in_patch, stats_bhw=(B, H * W))
        else:
            x, xs = self._conv(L.image_in(images.contiguous()), self.e_conv_in, B, H, W, stats=True)
        skips = []
        for res, ds in self.e_down:
            skips.append((x, H, W))
            for j, r in enumerate(res):
                x, xs = self._resnet(x, r, B, H, W, xs, stats=(ds is None or j + 1 < len(res)))
            if ds is not None:
                x, xs = self._conv(x, ds, B, H, W, stats=True)
                H, W = H // 2, W // 2
        x, xs = self._mid(x, self.e_mid, B, H, W, xs)
        t = self._gn(x, self.e_norm_out, B, H * W, True, xs)
        mom = self._conv(t, self.e_conv_out, B, H, W)                    # [B*hw, 2*lat] mean | logvar
        self.skip_acts = skips
        return L.vae_sample(mom, eps, self.sf, batch=B, c=self.lat, h=H, w=W)

    def decode(self, latents: torch.Tensor, skip_acts=None, dtype: torch.dtype = torch.float16) -> torch.Tensor:
        """latents: (B,4,h,w) fp32 CUDA (already scaled by scaling_factor, as the UNet works on them); returns
        vae.decode(latents / scaling_factor).sample.clamp(-1, 1) as NCHW `dtype` (pix2pix_turbo.py:333)."""
        B, _, H, W = latents.shape
        z = latents / self.sf + self.d_latent_shift.view(1, -1, 1, 1)
        x = L.latent_in(z.contiguous(), None, 1.0, 0.0)
        x, xs = self._conv(x, self.d_conv_in, B, H, W, stats=True)
        x, xs = self._mid(x, self.d_mid, B, H, W, xs)
        skips = skip_acts if skip_acts is not None else self.skip_acts
        for i, (res, us) in enumerate(self.d_up):
            if self.d_skip is not None:
                s_act, sh, sw = skips[::-1][i]
                assert (sh, sw) == (H, W)
                # sample + skip_conv(act * gamma), gamma = 1
                x, xs = self._lin(s_act, self.d_skip[i], residual=x, stats_bhw=(B, H * W))
            for j, r in enumerate(res):
                x, xs = self._resnet(x, r, B, H, W, xs, stats=(us is None or j + 1 < len(res)))
            if us is not None:
                x, xs = self._upsample_conv(x, us, B, H, W)
                H, W = 2 * H, 2 * W
        t = self._gn(x, self.d_norm_out, B, H * W, True, xs)
        y = self._conv(t, self.d_conv_out, B, H, W)                      # [B*HW, 3]
        return L.image_out(y, batch=B, c=y.shape[1], h=H, w=W, dtype=dtype)
