"""ORACLE (test infrastructure): CPU restatement of the integer image path either side of the network.

Follows the reference's host pre/post-processing, face_replace/inference/test.py:54-59 (torchvision
Resize(512, LANCZOS) -> CenterCrop(512) -> ToTensor -> Normalize(0.5, 0.5), then `.to(device, float16)` :92) and
face_replace/training/utils/vis_utils.py:14-23 (tensor2im(unnorm=True) on the fp16 prediction). The resize itself lives
in a third-party dependency that is not vendored: Pillow (12.2.0 in this image), `Image.resize(size, LANCZOS)` =
src/libImaging/Resample.c `ImagingResample` for 8-bit images — restated here from its published algorithm:
precompute_coeffs (double precision Lanczos-3 weights, support scaled by the down-sampling factor, normalised),
normalize_coeffs_8bpc (fixed point, PRECISION_BITS = 32 - 8 - 2 = 22, round half away from zero),
ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc (int32 accumulators seeded with 1 << 21, `>> 22`,
clipped to [0, 255]; the intermediate image is 8-bit). Pinned bit-exactly against Pillow itself in
tests/test_preprocess.py. Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a, b = x * math.pi, x / 3.0 * math.pi
        return (math.sin(a) / a) * (1.0 if b == 0.0 else math.sin(b) / b)
    return 0.0


def coeffs_8bpc(in_size: int, out_size: int):
    """(bounds [out, 2] = (first input index, tap count), kk [out, ksize] int32 fixed-point weights)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One resampling pass along axis 0 of an (n, m, c) uint8 array."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        lo, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_lanczos_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), LANCZOS) of an (H, W, 3) uint8 image: horizontal pass, then vertical."""
    h, w = img.shape[:2]
    if w != out_w:
        bx, kx = coeffs_8bpc(w, out_w)
        img = _pass(img.transpose(1, 0, 2), bx, kx).transpose(1, 0, 2)
    if h != out_h:
        by, ky = coeffs_8bpc(h, out_h)
        img = _pass(img, by, ky)
    return np.ascontiguousarray(img)


def resize_geometry(w: int, h: int, size: int = 512):
    """torchvision Resize(size) (shorter side -> size, longer = int(size * long / short)) + CenterCrop(size):
    (new_w, new_h, left, top)."""
    if (w <= h and w == size) or (h <= w and h == size):
        nw, nh = w, h
    elif w <= h:
        nw, nh = size, int(size * h / w)
    else:
        nw, nh = int(size * w / h), size
    return nw, nh, int(round((nw - size) / 2.0)), int(round((nh - size) / 2.0))


def transform_u8(img: np.ndarray, size: int = 512) -> np.ndarray:
    """Resize + CenterCrop of test.py:54-57 on an (H, W, 3) uint8 image -> (size, size, 3) uint8."""
    h, w = img.shape[:2]
    nw, nh, left, top = resize_geometry(w, h, size)
    out = resize_lanczos_u8(img, nh, nw)
    return np.ascontiguousarray(out[top:top + size, left:left + size])


def normalize_f16(img_u8: np.ndarray) -> np.ndarray:
    """ToTensor + Normalize(0.5, 0.5) in fp32, then the fp16 cast of test.py:92 -> (3, H, W) float16."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float16)


def tensor2im_u8(pred_f16: np.ndarray) -> np.ndarray:
    """vis_utils.tensor2im(unnorm=True) on the fp16 (3, H, W) prediction: every step rounds to fp16 (in-place ops on a
    half tensor), clamp to [0, 1], * 255 in fp16, truncation to uint8 -> (H, W, 3)."""
    v = pred_f16.astype(np.float16)
    v = (v.astype(np.float32) * np.float32(0.5)).astype(np.float16)
    v = (v.astype(np.float32) + np.float32(0.5)).astype(np.float16)
    v = v.transpose(1, 2, 0).copy()
    v[v < 0] = 0
    v[v > 1] = 1
    v = (v.astype(np.float32) * np.float32(255.0)).astype(np.float16)
    return v.astype(np.uint8)
