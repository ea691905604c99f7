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
