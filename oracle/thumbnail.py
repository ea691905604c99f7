"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

a1: IWSI.get_thumbnail_at_power (reference core/wsi/iwsi.py:246-323) for a single-level slide whose
dimensions are multiples of the integer factor f = base_mag / power: the whole level is read and
`cv2.resize(..., INTER_AREA)` reduces it.  OpenCV's integer-factor area path on uint8 is
`saturate_cast<uchar>(sum * (1/f^2))` in float32 = round-half-to-even of the block mean
(SURVEY.md section 2.3 T1).  Pinned against the reference's own output in tests/golden/thumb_*.npz.
"""
from __future__ import annotations

import numpy as np


def thumbnail_factor(mag: int, power: float = 1.25) -> float:
    return max(1e-6, float(mag) / float(power))  # iwsi.py:283


def area_reduce(level0: np.ndarray, f: int) -> np.ndarray:
    """level0: (H, W, 3) uint8 with H % f == W % f == 0 -> (H/f, W/f, 3) uint8."""
    H, W, C = level0.shape
    assert H % f == 0 and W % f == 0
    s = level0.reshape(H // f, f, W // f, f, C).astype(np.uint32).sum(axis=(1, 3))
    scaled = s.astype(np.float32) * np.float32(1.0 / (f * f))
    return np.clip(np.rint(scaled), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# General case: round(W0/ds) x round(H0/ds) output for level sizes the factor does not divide (core/wsi/iwsi.py:302-321).
# cv2.resize(INTER_AREA) then leaves its integer-factor fast path (resize.cpp: is_area_fast needs BOTH scales integral) and
# runs computeResizeAreaTab + ResizeArea_<uchar, float>: fractional cell weights in float32, a horizontal pass per source row
# (buf += S * alpha, entries in table order), a vertical pass (sum += beta * buf), saturate_cast<uchar>(sum) = round half even.
# Restated below operation by operation (separate float32 multiply and add: the routine is compiled for the SSE baseline, no FMA).
# Pinned against cv2 4.13 itself in tests/test_oracle_thumbnail.py and by reference-generated goldens.
# ---------------------------------------------------------------------------------------------------------------------
def area_tab(ssize: int, dsize: int):
    """computeResizeAreaTab: per destination index the (source index, float32 weight) pairs, in the order OpenCV emits them.
    Returns (first[dsize + 1], si[k], alpha[k])."""
    import math

    scale = float(ssize) / float(dsize)
    first, si, alpha = [0], [], []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            si.append(sx1 - 1)
            alpha.append(np.float32((sx1 - fsx1) / cell))
        for sx in range(sx1, sx2):
            si.append(sx)
            alpha.append(np.float32(1.0 / cell))
        if fsx2 - sx2 > 1e-3:
            si.append(sx2)
            alpha.append(np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell))
        first.append(len(si))
    return np.asarray(first, np.int32), np.asarray(si, np.int32), np.asarray(alpha, np.float32)


def thumbnail_size(W0: int, H0: int, ds: float) -> tuple[int, int]:
    """iwsi.py:302-303 (Python round = half to even)."""
    return max(1, int(round(W0 / ds))), max(1, int(round(H0 / ds)))


def area_resize_general(src: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """cv2.resize(src, (out_w, out_h), interpolation=INTER_AREA) for a down-scale with non-integral scale factors."""
    H, W, C = src.shape
    xf, xs, xa = area_tab(W, out_w)
    yf, ys, ya = area_tab(H, out_h)
    out = np.empty((out_h, out_w, C), np.uint8)
    # horizontal pass of every source row that is used: buf[sy][dx] = sum_k S[sy][si_k] * alpha_k, k in table order (float32)
    kmax = int((xf[1:] - xf[:-1]).max())
    S = src.astype(np.float32)
    buf = np.zeros((H, out_w, C), np.float32)
    for k in range(kmax):
        idx = xf[:-1] + k
        valid = idx < xf[1:]
        idx = np.minimum(idx, len(xs) - 1)
        term = S[:, xs[idx], :] * xa[idx][None, :, None]          # float32 multiply
        buf = np.where(valid[None, :, None], buf + term, buf)     # float32 add (0 + x == x for the first entry)
    for dy in range(out_h):
        acc = None
        for j in range(yf[dy], yf[dy + 1]):
            t = ya[j] * buf[ys[j]]
            acc = t if acc is None else acc + t
        out[dy] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


def thumbnail_reference_rule(level0: np.ndarray, ds: float) -> np.ndarray:
    """What get_thumbnail_at_power returns for a single-level slide: integer-factor fast path when cv2 takes it, else the tables."""
    H, W, _ = level0.shape
    out_w, out_h = thumbnail_size(W, H, ds)
    if (out_w, out_h) == (W, H):
        return level0.copy()
    sx, sy = W / out_w, H / out_h
    if abs(sx - int(sx)) < 2.220446049250313e-16 and abs(sy - int(sy)) < 2.220446049250313e-16 and int(sx) == int(sy):
        return area_reduce(level0, int(sx))
    return area_resize_general(level0, out_w, out_h)
