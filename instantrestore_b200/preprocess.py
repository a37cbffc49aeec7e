"""GPU pre-processing of the reference entry (face_replace/inference/test.py:54-59 + the fp16 cast of :92):

    Resize(512, LANCZOS) -> CenterCrop(512) -> ToTensor -> Normalize(0.5, 0.5) -> float16

on uint8 HWC images already on the device. The resize is Pillow's 8-bit resampling (what torchvision's Resize calls for
PIL images): two separable integer passes whose fixed-point weights are computed here on the host, in double precision,
the way Pillow's precompute_coeffs / normalize_coeffs_8bpc do (one table per (input size, output size), cached); the
passes themselves run in `ir_resample_u8_pass`, only over the crop window, and the last one writes the normalised fp16
NCHW tensor the network takes. Results are bit-identical to the PIL path (tests/test_preprocess.py).
There is no CPU fallback: the kernels come from the C-ABI library.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2      # Pillow Resample.c


def _lanczos3(x: float) -> float:
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a = x * math.pi
        b = x / 3.0 * math.pi
        return (math.sin(a) / a) * (math.sin(b) / b)
    return 0.0


def lanczos_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pillow's coefficient tables for one axis: bounds int32 [out, 2] (first tap, tap count), kk int32 [out, ksize]."""
    scale = in_size / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = [_lanczos3((x + lo - center + 0.5) * inv) for x in range(n)]
        total = 0.0
        for v in w:
            total += v
        for x in range(n):
            k = w[x] / total if total != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * one) if k < 0 else int(0.5 + k * one)
        bounds[xx, 0], bounds[xx, 1] = lo, n
    return bounds, kk


def resize_geometry(w: int, h: int, size: int = 512) -> Tuple[int, int, int, int]:
    """torchvision Resize(size) on the shorter side + CenterCrop(size): (new_w, new_h, crop_left, crop_top)."""
    if (w <= h and w == size) or (h <= w and h == size):
        nw, nh = w, h
    elif w <= h:
        nw, nh = size, int(size * h / w)
    else:
        nw, nh = int(size * w / h), size
    return nw, nh, int(round((nw - size) / 2.0)), int(round((nh - size) / 2.0))


class GpuPreprocessor:
    def __init__(self, device, size: int = 512):
        L.load()
        self.dev = torch.device(device)
        self.size = size
        self._tables: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self._tmp: Dict[int, torch.Tensor] = {}

    def tables(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            b, k = lanczos_tables(in_size, out_size)
            t = (torch.from_numpy(b).to(self.dev), torch.from_numpy(k).to(self.dev))
            self._tables[key] = t
        return t

    def _scratch(self, nbytes: int) -> torch.Tensor:
        """Per-stream scratch for the 8-bit intermediate image (kernels of one stream run in order)."""
        key = torch.cuda.current_stream(self.dev).cuda_stream
        t = self._tmp.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            self._tmp[key] = t
        return t

    def __call__(self, img_u8: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """img_u8: CUDA uint8 [H, W, 3] (contiguous); out: CUDA fp16 [3, size, size] (contiguous) <- the transformed image."""
        if img_u8.dtype != torch.uint8 or not img_u8.is_cuda or img_u8.ndim != 3 or img_u8.shape[2] != 3 or not img_u8.is_contiguous():
            raise TypeError("GpuPreprocessor: expected a contiguous CUDA uint8 [H, W, 3] image")
        h, w = int(img_u8.shape[0]), int(img_u8.shape[1])
        s = self.size
        if min(h, w) < 1:
            raise ValueError("empty image")
        nw, nh, left, top = resize_geometry(w, h, s)
        if nw < s or nh < s:
            raise ValueError(f"image {w}x{h} resizes to {nw}x{nh}: smaller than the {s}x{s} crop")
        row = 3 * w
        need_h, need_v = nw != w, nh != h
        hw = s * s
        if not need_h and not need_v:
            L.u8_to_f16(img_u8, top * row + left * 3, row, 3, s, s, out)
            return out
        if need_h and not need_v:
            bx, kx = self.tables(w, nw)
            # rows top .. top + s of the input, output columns left .. left + s; thread order follows the input rows
            L.resample_u8_pass(img_u8, top * row, 3, row, s, s, bx, kx, left, out, 1, s, hw, True)
            return out
        by, ky = self.tables(h, nh)
        if not need_h:
            L.resample_u8_pass(img_u8, left * 3, row, 3, s, s, by, ky, top, out, s, 1, hw, True)
            return out
        # both: horizontal pass into an 8-bit intermediate [rows, s, 3] covering the input rows the vertical pass reads
        bx, kx = self.tables(w, nw)
        key = ("v", h, nh, top)
        t = self._tables.get(key)
        if t is None:
            by_h = by[top:top + s].cpu()
            r0 = int(by_h[:, 0].min())
            r1 = int((by_h[:, 0] + by_h[:, 1]).max())
            b2 = by.clone()
            b2[:, 0] -= r0                      # tap indices relative to the first intermediate row
            t = (b2, ky, r0, r1 - r0)
            self._tables[key] = t
        b2, ky, r0, rows = t
        tmp = self._scratch(rows * s * 3)
        L.resample_u8_pass(img_u8, r0 * row, 3, row, s, rows, bx, kx, left, tmp, 3, 3 * s, 1, False)
        L.resample_u8_pass(tmp, 0, 3 * s, 3, s, s, b2, ky, top, out, s, 1, hw, True)
        return out
