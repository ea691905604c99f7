"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the --no-fast-mode content filter.

Restates, in numpy integer arithmetic, what the reference does per candidate patch when ``fast_mode`` is off
(atlas_patch/services/extraction.py:105-119):

    patch = wsi.extract((x, y), lv, (read_w, read_h))              # extraction.py:105
    patch = cv2.resize(patch, (patch_size, patch_size))            # extraction.py:109-113 (bilinear) when read != patch
    drop if is_black_patch(patch, rgb_thresh=black_threshold)      # utils/image.py:7-18
    drop if is_white_patch(patch, sat_thresh=white_threshold)      # utils/image.py:21-41

The OpenCV pieces are restated from its published 8-bit fixed-point definitions and pinned EXHAUSTIVELY (all 2^24 colours)
against the cv2 build recorded in tests/golden/versions.json by tests/test_oracle_filter.py:

* COLOR_RGB2GRAY (8u):  gray = (9798 R + 19235 G + 3735 B + 2^14) >> 15
* COLOR_RGB2HSV  (8u):  v = max(R,G,B);  s = ((v - min) * sdiv[v] + 2^11) >> 12,  sdiv[v] = round((255 << 12) / v), sdiv[0] = 0
* cv2.resize INTER_LINEAR (8u), any ratio: two taps per axis with 11-bit coefficients (resize_linear_u8 below); at an integer
  ratio r it collapses to the rounded mean of each block's central 2 x 2 pixels (even r) or its centre pixel (odd r)

Pinned against the reference itself by tests/golden/filter_*.npz (make_golden.py --filter runs the reference's
_iter_patch_entries with fast_mode=False on the synthetic slides).
"""
from __future__ import annotations

import numpy as np

GRAY_R, GRAY_G, GRAY_B, GRAY_SHIFT = 9798, 19235, 3735, 15
HSV_SHIFT = 12
SDIV = np.zeros(256, dtype=np.int64)
SDIV[1:] = np.rint((255 << HSV_SHIFT) / np.arange(1, 256, dtype=np.float64)).astype(np.int64)


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    p = rgb.astype(np.int64)
    return (p[..., 0] * GRAY_R + p[..., 1] * GRAY_G + p[..., 2] * GRAY_B + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT


def rgb_to_sv(rgb: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    p = rgb.astype(np.int64)
    v = p.max(axis=-1)
    mn = p.min(axis=-1)
    s = ((v - mn) * SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    return s, v


def halve_bilinear(rgb: np.ndarray) -> np.ndarray:
    return shrink_bilinear(rgb, 2)


def shrink_bilinear(rgb: np.ndarray, r: int) -> np.ndarray:
    """cv2.resize(rgb, (w / r, h / r)) (INTER_LINEAR, uint8) for an integer ratio r: the sample point r d + (r - 1) / 2 is the
    midpoint of the block's central 2 x 2 pixels (even r: both taps weigh 1024/2048 -> (a + b + c + d + 2) >> 2) or exactly its
    centre pixel (odd r).  Pinned against cv2 for r = 2..8 in tests/test_oracle_filter.py."""
    a = rgb.astype(np.int64)
    k = (r - 1) // 2
    if r % 2:
        return a[k::r, k::r].astype(np.uint8)
    return ((a[k::r, k::r] + a[k::r, k + 1::r] + a[k + 1::r, k::r] + a[k + 1::r, k + 1::r] + 2) >> 2).astype(np.uint8)


def linear_taps(n_src: int, n_dst: int):
    """OpenCV's 8-bit INTER_LINEAR taps of one axis before any border rule: (first source index s[n_dst], weight of s, weight of s + 1),
    weights in units of 1 / 2048 (resize_linear_u8 below applies the same arithmetic and is pinned against cv2 itself)."""
    scale = 1.0 / (n_dst / n_src)
    f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64), np.rint(f * np.float32(2048)).astype(np.int64)


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h)) with the default INTER_LINEAR on uint8, any ratio (OpenCV resize.cpp, 8-bit path):
    per axis fx = float32((d + 0.5) * scale - 0.5), scale = 1 / (n_dst / n_src) in float64; s = floor(fx); fx -= s; on x the
    border zeroes fx and pins s, on y the two row indices are clamped; coefficients round-half-even((1 - fx) * 2048) and
    (fx * 2048); horizontal pass in int32; vertical pass (((b0 (h0 >> 4)) >> 16) + ((b1 (h1 >> 4)) >> 16) + 2) >> 2.
    Pinned bit-exactly against cv2 on ten size pairs in tests/test_oracle_filter.py."""
    h, w, _ = img.shape

    def axis(n_src, n_dst):
        scale = 1.0 / (n_dst / n_src)
        f = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        return s, (f - s.astype(np.float32)).astype(np.float32)

    def coef(f):
        return np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64), np.rint(f * np.float32(2048)).astype(np.int64)

    sx, fx = axis(w, out_w)
    lo = sx < 0
    fx, sx = np.where(lo, np.float32(0), fx), np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx, sx = np.where(hi, np.float32(0), fx), np.where(hi, w - 1, sx)
    a0, a1 = coef(fx)
    src = img.astype(np.int64)
    hp = src[:, sx, :] * a0[None, :, None] + src[:, np.minimum(sx + 1, w - 1), :] * a1[None, :, None]
    sy, fy = axis(h, out_h)
    b0, b1 = coef(fy)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    out = (((b0[:, None, None] * (hp[y0] >> 4)) >> 16) + ((b1[:, None, None] * (hp[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def patch_counts(patch: np.ndarray, black_thresh: int, white_thresh: int, value_thresh: int = 200) -> tuple[int, int]:
    """(# pixels with gray < black_thresh, # pixels with s < white_thresh and v >= value_thresh)."""
    gray = rgb_to_gray(patch)
    s, v = rgb_to_sv(patch)
    return int((gray < black_thresh).sum()), int(((s < white_thresh) & (v >= value_thresh)).sum())


def keep_patch(patch: np.ndarray, black_thresh: int, white_thresh: int, min_fraction: float = 0.7) -> bool:
    nb, nw = patch_counts(patch, black_thresh, white_thresh)
    n = patch.shape[0] * patch.shape[1]
    # numpy's bool mean is an exact integer sum divided once in float64 (utils/image.py:17,39)
    return not (nb / n >= float(min_fraction) or nw / n >= float(min_fraction))


def filter_rows(read_region, rows: np.ndarray, patch_size: int, black_thresh: int, white_thresh: int,
                min_fraction: float = 0.7) -> tuple[np.ndarray, np.ndarray]:
    """rows: int32 (N, 5) candidates (x, y, read_w, read_h, level); read_region(x, y, w, h) -> uint8 (h, w, 3).
    Returns (kept rows, counts (N, 2))."""
    keep = np.zeros(len(rows), dtype=bool)
    counts = np.zeros((len(rows), 2), dtype=np.int32)
    for i, (x, y, rw, rh, _lv) in enumerate(rows.tolist()):
        patch = read_region(x, y, rw, rh)
        if rw != patch_size:
            patch = resize_linear_u8(patch, patch_size, patch_size)
        counts[i] = patch_counts(patch, black_thresh, white_thresh)
        n = patch_size * patch_size
        keep[i] = not (counts[i, 0] / n >= float(min_fraction) or counts[i, 1] / n >= float(min_fraction))
    return rows[keep], counts
