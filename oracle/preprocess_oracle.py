"""TEST INFRASTRUCTURE (CPU oracle) -- numpy restatement of the reference's image preprocessing
(utils/image_utils.py:5-23): torchvision Resize((S, S)) on a PIL RGB image (= Pillow's two-pass antialiased bilinear
resampling on 8-bit pixels, fixed-point coefficients, uint8 intermediate), ToTensor (/255) and ImageNet Normalize.

The arithmetic lives in third-party code that is not under /root/reference: Pillow (libImaging/Resample.c;
requirements pin Pillow via torchvision 0.10) and torchvision.transforms.  The restated algorithm:
  * per output coordinate: centre = (xx + 0.5) * scale, support = max(scale, 1), taps xmin .. xmax around it, triangle
    weights normalised in double precision, then rounded to 22-bit fixed point (round half away from zero);
  * horizontal pass over every input row, then vertical pass, each  clip8((2^21 + sum pixel * k) >> 22);
  * float32: (u8 / 255 - mean) / std.
Pinned by tests/test_preprocess_oracle.py against Pillow + torchvision themselves (bit-exact) on synthetic images.
Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this module."""
from __future__ import annotations

import math
import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def precompute_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter, box = whole axis."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            if v < 0.0:
                v = -v
            wv = 1.0 - v if v < 1.0 else 0.0
            w[x] = wv
            ww += wv
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img (H, W, C) uint8 -> (out_h, out_w, C) uint8, Pillow Image.resize(..., BILINEAR)."""
    H, W, C = img.shape
    src = img.astype(np.int64)
    if out_w != W:
        bounds, kk = precompute_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (src[:, x0:x0 + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
    else:
        tmp = img
    if out_h != H:
        bounds, kk = precompute_coeffs(H, out_h)
        out = np.empty((out_h, src.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (src[y0:y0 + n, :, :] * kk[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        return out
    return tmp.astype(np.uint8)


def preprocess_rgb8(img: np.ndarray, img_size: int) -> np.ndarray:
    """(H, W, 3) uint8 RGB -> (3, S, S) float32, the tensor utils/image_utils.py:preprocess_image builds (before unsqueeze)."""
    r = resize_bilinear_u8(img, img_size, img_size)
    x = r.astype(np.float32) / np.float32(255.0)          # ToTensor: float32 division
    x = (x - MEAN[None, None, :]) / STD[None, None, :]    # Normalize: float32 sub, float32 div
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)
