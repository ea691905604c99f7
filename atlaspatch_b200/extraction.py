"""Mask -> patch coordinates through the CUDA coordinate kernel (C ABI: ap_extract_coords).

Host part (<= 1 MPx mask, a few thousand vertices): threshold + cv2.findContours + hierarchy/area filters +
float32 scaling, exactly as the reference does with the same library calls (atlas_patch/utils/contours.py:41-131,
services/extraction.py:30-42) -- SURVEY.md section 2.3 "Placement guidance".  Device part: the candidate grid,
the 4-probe cv2.pointPolygonTest containment, hole rejection and ordered compaction
(services/extraction.py:67-103, utils/contours.py:22-38).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from atlaspatch_b200._lib import Context, current_stream_ptr
from atlaspatch_b200.geometry import PatchGeometry, prepare_geometry


def mask_to_contours(mask: np.ndarray, *, tissue_area_thresh: float = 0.01, a_h: int = 16, max_n_holes: int = 10):
    """Same calls, same order as utils/contours.py:41-116."""
    import cv2

    mask_uint8 = (mask > 0.5).astype(np.uint8) * 255
    contours, hierarchy = cv2.findContours(mask_uint8, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_NONE)
    if hierarchy is None or len(contours) == 0:
        return [], []
    hierarchy = np.asarray(hierarchy).reshape(-1, 4)
    H, W = mask.shape[:2]
    min_area = tissue_area_thresh * float(H * W)
    tissue_idx, holes_by_parent = [], {}
    for i, cont in enumerate(contours):
        area = cv2.contourArea(cont)
        parent = int(hierarchy[i][3])
        if parent == -1:
            if area >= min_area:
                tissue_idx.append(i)
        elif area >= float(a_h):
            holes_by_parent.setdefault(parent, []).append(cont)
    all_holes = [h for hs in holes_by_parent.values() for h in hs]
    if max_n_holes > 0 and len(all_holes) > max_n_holes:
        allowed = set(map(id, sorted(all_holes, key=cv2.contourArea, reverse=True)[:max_n_holes]))
        for p in list(holes_by_parent):
            holes_by_parent[p] = [h for h in holes_by_parent[p] if id(h) in allowed]
    return [contours[i] for i in tissue_idx], [list(holes_by_parent.get(i, [])) for i in tissue_idx]


def scale_contours(contours: Sequence[np.ndarray], sx: float, sy: float) -> list[np.ndarray]:
    """utils/contours.py:119-131: float32 multiply by a Python float, truncate to int32."""
    out = []
    for c in contours:
        f = c.astype(np.float32)
        f[:, :, 0] *= sx
        f[:, :, 1] *= sy
        out.append(f.astype(np.int32))
    return out


@dataclass
class FlatContours:
    """The flattened arrays the C ABI takes (include/atlaspatch_b200.h: ap_extract_coords)."""
    contour_xy: np.ndarray
    contour_offsets: np.ndarray
    hole_xy: np.ndarray
    hole_offsets: np.ndarray
    hole_first: np.ndarray

    @property
    def n_contours(self) -> int:
        return int(self.contour_offsets.shape[0] - 1)


def flatten_contours(tissue: Sequence[np.ndarray], holes: Sequence[Sequence[np.ndarray]]) -> FlatContours:
    cxy = [np.asarray(c, dtype=np.int32).reshape(-1, 2) for c in tissue]
    coff = np.zeros(len(cxy) + 1, dtype=np.int32)
    if cxy:
        coff[1:] = np.cumsum([c.shape[0] for c in cxy])
    hxy, hoff, hfirst = [], [0], [0]
    for hs in holes:
        for h in hs:
            a = np.asarray(h, dtype=np.int32).reshape(-1, 2)
            hxy.append(a)
            hoff.append(hoff[-1] + a.shape[0])
        hfirst.append(len(hoff) - 1)
    return FlatContours(
        np.ascontiguousarray(np.concatenate(cxy) if cxy else np.zeros((0, 2), np.int32)),
        coff,
        np.ascontiguousarray(np.concatenate(hxy) if hxy else np.zeros((0, 2), np.int32)),
        np.asarray(hoff, dtype=np.int32),
        np.asarray(hfirst, dtype=np.int32),
    )


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def extract_coords_from_contours(flat: FlatContours, geo: PatchGeometry, *, ctx: Context | None = None,
                                 return_device: bool = False):
    """Run the coordinate kernels.  Returns int32 (N, 5) numpy rows, plus a CUDA tensor copy if asked."""
    import torch

    ctx = ctx or Context.get(torch.cuda.current_device() if torch.cuda.is_available() else 0)
    n_c = flat.n_contours
    if n_c == 0:
        empty = np.empty((0, 5), dtype=np.int32)
        return (empty, torch.empty((0, 5), dtype=torch.int32, device="cuda")) if return_device else empty
    cap = int(ctx.lib.ap_coords_capacity(_ptr(flat.contour_xy), _ptr(flat.contour_offsets), n_c, geo.step_src))
    if cap < 0:
        raise ValueError("invalid contour arrays")
    rows_dev = torch.empty((max(cap, 1), 5), dtype=torch.int32, device="cuda")
    rows_host = np.empty((max(cap, 1), 5), dtype=np.int32)
    count = C.c_int64(0)
    ctx.check(ctx.lib.ap_extract_coords(
        ctx.handle, _ptr(flat.contour_xy), _ptr(flat.contour_offsets), n_c, _ptr(flat.hole_xy), _ptr(flat.hole_offsets),
        _ptr(flat.hole_first), geo.patch_size_src, geo.step_src, geo.read_w, geo.read_h, geo.level,
        C.c_void_p(rows_dev.data_ptr()), _ptr(rows_host), cap, C.byref(count), C.c_void_p(current_stream_ptr())))
    n = int(count.value)
    coords = rows_host[:n].copy()
    return (coords, rows_dev[:n]) if return_device else coords


def extract_coords(mask: np.ndarray, *, level0_wh: tuple[int, int], src_mag: int, target_mag: int, patch_size: int,
                   step_size: int | None = None, tissue_thresh: float = 0.0, downsamples: Sequence[float] = (1.0,),
                   ctx: Context | None = None, return_device: bool = False):
    """a6-a9 end to end: what the reference writes into the H5 `coords` dataset."""
    geo = prepare_geometry(src_mag=src_mag, target_mag=target_mag, patch_size=patch_size, step_size=step_size,
                           downsamples=downsamples)
    tissue_t, holes_t = mask_to_contours(mask, tissue_area_thresh=tissue_thresh)
    W, H = level0_wh
    mh, mw = mask.shape[:2]
    sx, sy = W / float(mw), H / float(mh)
    tissue = scale_contours(tissue_t, sx, sy)
    holes = [scale_contours(hs, sx, sy) for hs in holes_t]
    return extract_coords_from_contours(flatten_contours(tissue, holes), geo, ctx=ctx, return_device=return_device)


def filter_patches(slide_dev, W: int, H: int, pitch: int, rows_dev, *, patch_size: int, black_threshold: int = 50,
                   white_threshold: int = 15, min_fraction: float = 0.7, ctx: Context | None = None, return_counts: bool = False):
    """The --no-fast-mode content filter (services/extraction.py:105-119, utils/image.py:7-41) on a slide resident in HBM.

    rows_dev: int32 (N, 5) CUDA tensor of candidates from extract_coords_from_contours.  Returns (kept rows as numpy,
    kept rows as a CUDA tensor[, per-candidate (black, white) pixel counts as numpy])."""
    import torch

    ctx = ctx or Context.get(torch.cuda.current_device())
    n = int(rows_dev.shape[0])
    out_dev = torch.empty((max(n, 1), 5), dtype=torch.int32, device="cuda")
    out_host = np.empty((max(n, 1), 5), dtype=np.int32)
    counts = torch.zeros((max(n, 1), 2), dtype=torch.int32, device="cuda") if return_counts else None
    count = C.c_int64(0)
    if n:
        rows_dev = rows_dev.contiguous()
        read_size = int(rows_dev[0, 2].item())
        ctx.check(ctx.lib.ap_filter_patches(
            ctx.handle, C.c_void_p(slide_dev.data_ptr()), W, H, pitch, C.c_void_p(rows_dev.data_ptr()), n, read_size,
            int(patch_size), int(black_threshold), int(white_threshold), float(min_fraction), C.c_void_p(out_dev.data_ptr()),
            _ptr(out_host), C.byref(count), C.c_void_p(counts.data_ptr()) if counts is not None else None,
            C.c_void_p(current_stream_ptr())))
    k = int(count.value)
    res = (out_host[:k].copy(), out_dev[:k])
    return res + (counts[:n].cpu().numpy(),) if return_counts else res
