"""Host-side slide geometry (scalars only), mirroring the reference so that coordinates are bit-exact.

reference: atlas_patch/services/extraction.py:44-64 (_prepare_geometry), core/wsi/iwsi.py:325-384
(optimal_level, _infer_mag), core/wsi/iwsi.py:283 (thumbnail downsample).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

MPP_MIN, MPP_MAX = 0.1, 10.0  # iwsi.py:13-14


def infer_mag(mpp: float) -> int:
    """iwsi.py:360-384."""
    for threshold, mag in ((0.16, 80), (0.2, 60), (0.3, 40), (0.6, 20), (1.2, 10), (2.4, 5)):
        if mpp < threshold:
            return mag
    raise ValueError(f"Cannot infer magnification from mpp {mpp}")


def validate_mpp(mpp: float, *, source: str = "metadata") -> float:
    """iwsi.py:126-154."""
    if not (MPP_MIN <= mpp <= MPP_MAX):
        raise ValueError(f"MPP value {mpp} from {source} is outside plausible range [{MPP_MIN}, {MPP_MAX}] um/pixel.")
    return mpp


def optimal_level(downsamples: Sequence[float], target_ds: float) -> tuple[int, float]:
    ds = list(downsamples) or [1.0]
    for i, d in enumerate(ds):
        if abs(d - target_ds) < 0.01:
            return i, 1.0
    if target_ds >= ds[0]:
        best_i, best_d = 0, ds[0]
        for i, d in enumerate(ds):
            if d <= target_ds:
                best_i, best_d = i, d
            else:
                break
        return best_i, target_ds / best_d
    for i, d in enumerate(ds):
        if d >= target_ds:
            return i, d / target_ds
    raise ValueError(f"No level for target downsample {target_ds}")


@dataclass(frozen=True)
class PatchGeometry:
    level: int
    read_w: int
    read_h: int
    patch_size_src: int
    step_src: int
    patch_size_level0: int


def prepare_geometry(*, src_mag: int | None, target_mag: int, patch_size: int, step_size: int | None,
                     downsamples: Sequence[float]) -> PatchGeometry:
    if src_mag is None:
        raise ValueError("WSI base magnification is required for patch extraction.")
    if int(target_mag) > int(src_mag):
        raise ValueError(f"Requested magnification {target_mag}x exceeds available {src_mag}x.")
    desired = float(src_mag) / float(target_mag)
    level, _ = optimal_level(downsamples, desired)
    level_ds = float((list(downsamples) or [1.0])[level])
    patch_src = int(round(patch_size * desired))                       # Python round: half-to-even
    step_src = int(round((step_size or patch_size) * desired))
    p0 = int(patch_size * int(src_mag) // int(target_mag))
    read = max(1, int(round(patch_src / level_ds)))
    return PatchGeometry(level, read, read, patch_src, step_src, p0)


def thumbnail_factor(mag: int, power: float = 1.25) -> float:
    if power <= 0:
        raise ValueError("thumbnail power must be positive")
    return max(1e-6, float(mag) / float(power))
