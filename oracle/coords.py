"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by atlaspatch_b200/).

numpy restatement of the reference's mask -> patch-coordinate path.  Parity
status: the reference ships no tests or golden vectors, so
this restatement is pinned against the reference's *own functions executed in the
build container* (`tests/golden/make_golden.py` -> `tests/golden/coords_*.npz`)
and against the OpenCV routines the reference calls (property tests in
`tests/test_oracle_coords.py`).

Every function cites the reference lines it follows (paths under /root/reference).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np


# ---------------------------------------------------------------------------------------
# a8  geometry  -- atlas_patch/services/extraction.py:44-64, core/wsi/iwsi.py:325-358
# ---------------------------------------------------------------------------------------
def optimal_level(downsamples: Sequence[float], target_ds: float) -> tuple[int, float]:
    """core/wsi/iwsi.py:325-358."""
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
class Geometry:
    level: int
    read_w: int
    read_h: int
    patch_size_src: int
    step_src: int
    patch_size_level0: int


def prepare_geometry(*, src_mag: int, target_mag: int, patch_size: int, step_size: int | None,
                     downsamples: Sequence[float]) -> Geometry:
    """services/extraction.py:44-64 (Python round() = half-to-even)."""
    if int(target_mag) > int(src_mag):
        raise ValueError(f"Requested magnification {target_mag}x exceeds available {src_mag}x.")
    desired = float(src_mag) / float(target_mag)
    level, _ = optimal_level(downsamples, desired)
    level_ds = float((list(downsamples) or [1.0])[level])
    patch_src = int(round(patch_size * desired))
    step_src = int(round((step_size or patch_size) * desired))
    p0 = int(patch_size * int(src_mag) // int(target_mag))
    read = max(1, int(round(patch_src / level_ds)))
    return Geometry(level, read, read, patch_src, step_src, p0)


# ---------------------------------------------------------------------------------------
# a6  mask -> contours  -- atlas_patch/utils/contours.py:41-116
# ---------------------------------------------------------------------------------------
def mask_to_contours(mask: np.ndarray, *, tissue_area_thresh: float = 0.01, a_h: int = 16,
                     max_n_holes: int = 10):
    """utils/contours.py:41-116.  cv2.findContours IS the algorithm the reference calls
    (utils/contours.py:58-59); the filtering/ordering logic around it is restated."""
    import cv2

    m8 = (mask > 0.5).astype(np.uint8) * 255
    contours, hierarchy = cv2.findContours(m8, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    if hierarchy is None or len(contours) == 0:
        return [], []
    hierarchy = np.asarray(hierarchy).reshape(-1, 4)
    H, W = mask.shape[:2]
    min_area = tissue_area_thresh * float(H * W)
    tissue_idx: list[int] = []
    holes_by_parent: dict[int, list[np.ndarray]] = {}
    for i, c in enumerate(contours):
        area = cv2.contourArea(c)
        parent = int(hierarchy[i][3])
        if parent == -1:
            if area >= min_area:
                tissue_idx.append(i)
        elif area >= float(a_h):
            holes_by_parent.setdefault(parent, []).append(c)
    all_holes = [h for hs in holes_by_parent.values() for h in hs]
    if max_n_holes > 0 and len(all_holes) > max_n_holes:
        keep = sorted(all_holes, key=cv2.contourArea, reverse=True)[:max_n_holes]  # stable
        allowed = set(map(id, keep))
        for p in list(holes_by_parent):
            holes_by_parent[p] = [h for h in holes_by_parent[p] if id(h) in allowed]
    tissue = [contours[i] for i in tissue_idx]
    holes = [list(holes_by_parent.get(i, [])) for i in tissue_idx]
    return tissue, holes


# ---------------------------------------------------------------------------------------
# a7  scale  -- atlas_patch/utils/contours.py:119-131, services/extraction.py:35-41
# ---------------------------------------------------------------------------------------
def scale_contour(c: np.ndarray, sx: float, sy: float) -> np.ndarray:
    """float32 in-place multiply by a Python float, then truncate to int32."""
    f = c.astype(np.float32)
    f[:, :, 0] *= sx
    f[:, :, 1] *= sy
    return f.astype(np.int32)


def prepare_contours(mask: np.ndarray, level0_wh: tuple[int, int], *, tissue_area_thresh: float):
    """services/extraction.py:30-42."""
    tissue, holes = mask_to_contours(mask, tissue_area_thresh=tissue_area_thresh)
    W, H = level0_wh
    mh, mw = mask.shape[:2]
    sx, sy = W / float(mw), H / float(mh)
    return ([scale_contour(c, sx, sy) for c in tissue],
            [[scale_contour(h, sx, sy) for h in hs] for hs in holes])


# ---------------------------------------------------------------------------------------
# a9  containment  -- cv2.pointPolygonTest(measureDist=False) on CV_32S contours with
#     integral query points (called at utils/contours.py:37 and services/extraction.py:79).
#     OpenCV 4.13 integer branch restated; verified against cv2.
# ---------------------------------------------------------------------------------------
def point_polygon_test(contour: np.ndarray, px: np.ndarray, py: np.ndarray) -> np.ndarray:
    """Vectorised over query points.  Returns int8 array of +1 (inside) / 0 (edge) / -1."""
    pts = np.asarray(contour, dtype=np.int64).reshape(-1, 2)
    px = np.asarray(px, dtype=np.int64)
    py = np.asarray(py, dtype=np.int64)
    counter = np.zeros(px.shape, dtype=np.int64)
    on_edge = np.zeros(px.shape, dtype=bool)
    K = pts.shape[0]
    if K == 0:
        return np.full(px.shape, -1, dtype=np.int8)
    v = pts[K - 1]
    for i in range(K):
        v0, v = v, pts[i]
        skip = ((v0[1] <= py) & (v[1] <= py)) | ((v0[1] > py) & (v[1] > py)) | ((v0[0] < px) & (v[0] < px))
        on_v = (py == v[1]) & ((px == v[0]) | ((py == v0[1]) & (((v0[0] <= px) & (px <= v[0])) |
                                                                ((v[0] <= px) & (px <= v0[0])))))
        on_edge |= skip & on_v
        dist = (py - v0[1]) * (v[0] - v0[0]) - (px - v0[0]) * (v[1] - v0[1])
        on_edge |= (~skip) & (dist == 0)
        if v[1] < v0[1]:
            dist = -dist
        counter += ((~skip) & (dist > 0)).astype(np.int64)
    res = np.where(counter % 2 == 0, -1, 1).astype(np.int8)
    res[on_edge] = 0
    return res


def bounding_rect(contour: np.ndarray) -> tuple[int, int, int, int]:
    """cv2.boundingRect on an int32 point set (services/extraction.py:94)."""
    pts = np.asarray(contour).reshape(-1, 2)
    x0, y0 = int(pts[:, 0].min()), int(pts[:, 1].min())
    return x0, y0, int(pts[:, 0].max()) - x0 + 1, int(pts[:, 1].max()) - y0 + 1


def extract_coords(tissue: Sequence[np.ndarray], holes: Sequence[Sequence[np.ndarray]],
                   geo: Geometry) -> np.ndarray:
    """services/extraction.py:83-103 (fast mode) + :67-81 + utils/contours.py:22-38.

    Returns int32 (N, 5) rows (x, y, read_w, read_h, level) in the reference's order:
    contour-major, then y, then x; no clipping, no de-duplication.
    """
    P, step = geo.patch_size_src, geo.step_src
    half = P // 2
    shift = int(half * 0.5)
    rows = []
    for contour, hs in zip(tissue, holes):
        x0, y0, ww, hh = bounding_rect(contour)
        xs = np.arange(x0, x0 + ww, step, dtype=np.int64)
        ys = np.arange(y0, y0 + hh, step, dtype=np.int64)
        if xs.size == 0 or ys.size == 0:
            continue
        X, Y = np.meshgrid(xs, ys)  # row-major: y outer, x inner
        X, Y = X.ravel(), Y.ravel()
        cx, cy = X + half, Y + half
        in_hole = np.zeros(X.shape, dtype=bool)
        for h in hs:
            in_hole |= point_polygon_test(h, cx, cy) > 0
        if shift > 0:
            probes = [(cx - shift, cy - shift), (cx + shift, cy + shift),
                      (cx + shift, cy - shift), (cx - shift, cy + shift)]
        else:
            probes = [(cx, cy)]
        inside = np.zeros(X.shape, dtype=bool)
        for (qx, qy) in probes:
            inside |= point_polygon_test(contour, qx, qy) >= 0
        keep = inside & ~in_hole
        n = int(keep.sum())
        if n:
            r = np.empty((n, 5), dtype=np.int32)
            r[:, 0], r[:, 1] = X[keep], Y[keep]
            r[:, 2], r[:, 3], r[:, 4] = geo.read_w, geo.read_h, geo.level
            rows.append(r)
    if not rows:
        return np.empty((0, 5), dtype=np.int32)
    return np.concatenate(rows, axis=0)


def coords_from_mask(mask: np.ndarray, *, level0_wh: tuple[int, int], src_mag: int, target_mag: int,
                     patch_size: int, step_size: int | None = None, tissue_thresh: float = 0.0,
                     downsamples: Sequence[float] = (1.0,)) -> np.ndarray:
    """End to end a6-a9: what PatchExtractionService.extract writes into `coords`."""
    geo = prepare_geometry(src_mag=src_mag, target_mag=target_mag, patch_size=patch_size,
                           step_size=step_size, downsamples=downsamples)
    tissue, holes = prepare_contours(mask, level0_wh, tissue_area_thresh=tissue_thresh)
    return extract_coords(tissue, holes, geo)
