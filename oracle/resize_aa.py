"""TEST INFRASTRUCTURE ONLY.  Integer restatement of torch's uint8 bicubic *antialias* resize (what transformers'
BitImageProcessorFast runs for the DINOv2 encoders: torchvision.transforms.v2.functional.resize(uint8 tensor, BICUBIC,
antialias=True) -> torch.nn.functional.interpolate(..., mode="bicubic", antialias=True) on uint8, ATen
UpSampleKernel.cpp separable_upsample_generic_Nd_kernel_impl / UpSampleKernelAVXAntialias.h).

Algorithm (Pillow's ImagingResample): per output index i, centre = scale (i + 0.5), support = 2 max(scale, 1),
taps [xmin, xmin + xsize) = [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the input, weights = Keys
cubic (a = -0.5) of (j + xmin - centre + 0.5) / max(scale, 1), normalised in float64, converted to int16 with the largest
precision p < 22 for which round(max_weight * 2^(p+1)) < 2^15; out = clamp((sum w_j src_j + 2^(p-1)) >> p, 0, 255).  The
horizontal pass runs first and its uint8 result feeds the vertical pass.

Pinned bit-exactly against torch itself (AVX-512 and ATEN_CPU_CAPABILITY=default builds give identical bytes) by
tests/test_oracle_dinov2.py.
"""
from __future__ import annotations

import numpy as np


def _cubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def _triangle(x: float) -> float:   # ATen HelperInterpLinear::aa_filter
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def aa_weights(n_in: int, n_out: int, mode: str = "bicubic"):
    """(xmin[n_out], xsize[n_out], list of int64 weight arrays, precision).  mode "bilinear": the triangle filter with
    interp_size 2 (support 1) through the same index / precision arithmetic (ViTImageProcessorFast with resample = 2: phikon.py)."""
    filt, half = (_cubic, 2.0) if mode == "bicubic" else (_triangle, 1.0)
    scale = n_in / n_out
    support = half * scale if scale >= 1.0 else half
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    xmins, sizes, ws = [], [], []
    for i in range(n_out):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), n_in) - xmin
        w = np.array([filt((j + xmin - center + 0.5) * invscale) for j in range(xsize)], dtype=np.float64)
        w = w / w.sum()
        xmins.append(xmin)
        sizes.append(xsize)
        ws.append(w)
    wt_max = max(float(w.max()) for w in ws)
    prec = 0
    for prec in range(22):
        if int(0.5 + wt_max * (1 << (prec + 1))) >= (1 << 15):
            break
    w16 = [np.array([int(v * (1 << prec) + (0.5 if v >= 0 else -0.5)) for v in w], dtype=np.int64) for w in ws]
    return np.asarray(xmins), np.asarray(sizes), w16, prec


def _resize_axis(a: np.ndarray, n_out: int, axis: int, mode: str = "bicubic") -> np.ndarray:
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    xmins, sizes, w16, prec = aa_weights(a.shape[0], n_out, mode)
    out = np.empty((n_out,) + a.shape[1:], dtype=np.int64)
    for i in range(n_out):
        acc = np.tensordot(w16[i], a[xmins[i]:xmins[i] + sizes[i]], axes=(0, 0)) + (1 << (prec - 1))
        out[i] = np.clip(acc >> prec, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_aa(img: np.ndarray, out_h: int, out_w: int, mode: str = "bicubic") -> np.ndarray:
    """HWC uint8 -> (out_h, out_w, C) uint8; an axis whose size does not change is not resampled."""
    t = _resize_axis(img, out_w, 1, mode) if img.shape[1] != out_w else img
    return _resize_axis(t, out_h, 0, mode) if img.shape[0] != out_h else t


def hf_vit_pixels(patch: np.ndarray, size: int = 224) -> np.ndarray:
    """uint8 (size, size, 3) that ViTImageProcessorFast (resize to size x size, resample 2 = bilinear, no crop) normalises."""
    return resize_aa(patch, size, size, "bilinear") if patch.shape[0] != size or patch.shape[1] != size else patch


def dinov2_pixels(patch: np.ndarray, resize_to: int = 256, crop: int = 224) -> np.ndarray:
    """uint8 (crop, crop, 3) that BitImageProcessorFast normalises for a square patch."""
    r = resize_aa(patch, resize_to, resize_to)
    o = (resize_to - crop) // 2
    return r[o:o + crop, o:o + crop]


# ---------------------------------------------------------------------------------------------------------------------
# Pillow's own 8-bit resampler (libImaging/Resample.c), BILINEAR: what torchvision's ImageClassification preset runs when the
# reference hands it a PIL image whose size differs from resize_size (models/patch/base.py:42-45 -> Image.fromarray ->
# weights.transforms(); [tv]transforms/_presets.py:58 -> F.resize -> _functional_pil.resize -> Image.resize(BILINEAR)).
# Same tap geometry as above with the triangle filter (support 1), but coefficients are int32 with a FIXED precision of
# 32 - 8 - 2 = 22 bits (normalize_coeffs_8bpc) and both passes accumulate in int32 from 1 << 21.  Pinned against Pillow itself
# in tests/test_oracle_vit.py.
# ---------------------------------------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 22


def pil_bilinear_weights(n_in: int, n_out: int, mode: str = "bilinear"):
    """precompute_coeffs + normalize_coeffs_8bpc: (xmin[n_out], xsize[n_out], list of int64 weight arrays).  mode "bicubic": Pillow's
    bicubic_filter (Keys cubic a = -0.5, support 2) through the same coefficient arithmetic (gigapath.py:17-26)."""
    filt, half = ((lambda t: max(0.0, 1.0 - abs(t))), 1.0) if mode == "bilinear" else (_cubic, 2.0)
    scale = n_in / n_out
    filterscale = max(scale, 1.0)
    support = half * filterscale
    ss = 1.0 / filterscale
    xmins, sizes, ws = [], [], []
    for i in range(n_out):
        center = (i + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), n_in) - xmin
        w = np.array([filt((x + xmin - center + 0.5) * ss) for x in range(xmax)], dtype=np.float64)
        tot = float(w.sum())
        if tot != 0.0:
            w = w / tot
        xmins.append(xmin)
        sizes.append(xmax)
        ws.append(np.array([int(v * (1 << PIL_PRECISION_BITS) + (0.5 if v >= 0 else -0.5)) for v in w], dtype=np.int64))
    return np.asarray(xmins), np.asarray(sizes), ws


def _resize_axis_pil(a: np.ndarray, n_out: int, axis: int, mode: str = "bilinear") -> np.ndarray:
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    xmins, sizes, ws = pil_bilinear_weights(a.shape[0], n_out, mode)
    out = np.empty((n_out,) + a.shape[1:], dtype=np.int64)
    for i in range(n_out):
        acc = np.tensordot(ws[i], a[xmins[i]:xmins[i] + sizes[i]], axes=(0, 0)) + (1 << (PIL_PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_pil_bilinear(img: np.ndarray, out_h: int, out_w: int, mode: str = "bilinear") -> np.ndarray:
    """PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR | BICUBIC) for HWC uint8: horizontal pass first, then vertical."""
    t = _resize_axis_pil(img, out_w, 1, mode) if img.shape[1] != out_w else img
    return _resize_axis_pil(t, out_h, 0, mode) if img.shape[0] != out_h else t


def vit_preset_pixels(patch: np.ndarray, resize_to: int = 256, crop: int = 224, mode: str = "bilinear") -> np.ndarray:
    """uint8 (crop, crop, 3) the torchvision preset normalises for a square PIL patch of any size."""
    r = resize_pil_bilinear(patch, resize_to, resize_to, mode) if patch.shape[0] != resize_to else patch
    o = int(round((resize_to - crop) / 2.0))
    return r[o:o + crop, o:o + crop]
