"""a10 / a14: the reference's per-slide HDF5 container, written from in-memory results.

reference: atlas_patch/services/storage.py (H5PatchWriter.write_coords :106-161, append_features :250-337, _passport :387-392),
atlas_patch/utils/h5.py (H5AppendWriter: tmp file + os.replace, chunked resizable datasets), file attributes from
storage.py:55-70 and services/extraction.py:146-164.

Layout produced (identical names, dtypes, shapes, chunking and attributes):

    /coords               int32 (N, 5)  rows (x, y, read_w, read_h, level), maxshape (None, 5), chunks (write_batch, 5)
    /passports            S160  (N,)    "{stem}__x{X}_y{Y}_rw{RW}_rh{RH}_lv{LV}_mag{MAG}_tmag{TMAG}_total{TOTAL}"
    /features/<encoder>   float32 (N, D), maxshape (None, D), chunks (feature_batch, D); written as __tmp_<encoder>, then moved
    file attrs            patch_size, patch_size_level0, level0_magnification, target_magnification, overlap, level0_width,
                          level0_height, wsi_path, passport_format, passport_version, creation_date, filename, num_patches

The HDF5 library is the reference's own dependency (h5py) and is used when it is installed.  This build image and the GPU
boxes have neither h5py nor libhdf5: there the same calls go to `atlaspatch_b200.h5lite`, a self-contained writer / reader of
the HDF5 subset this container needs (see its header for what is and is not pinned: the reader is checked against a file
written by the real library, the writer by round trips; acceptance of h5lite's bytes by libhdf5 is UNPINNED until
tests/test_h5lite.py::test_h5py_reads_h5lite_file_and_back runs where h5py exists).
"""
from __future__ import annotations

import json
import os
import uuid
from datetime import datetime, timezone
from pathlib import Path
from typing import Any, Mapping

import numpy as np

PASSPORT_FORMAT = "{stem}__x{X}_y{Y}_rw{RW}_rh{RH}_lv{LV}_mag{MAG}_tmag{TMAG}_total{TOTAL}"
PASSPORT_DTYPE = np.dtype("S160")


def _h5py():
    """h5py where it exists, else the in-tree HDF5 subset implementation with the same calls."""
    try:
        import h5py

        if getattr(h5py, "File", None) is not None:   # tests may have a stub module installed under this name
            return h5py
    except ImportError:
        pass
    from atlaspatch_b200 import h5lite

    return h5lite


def passports(stem: str, coords: np.ndarray, level0_mag: int, target_mag: int) -> np.ndarray:
    """storage.py:387-392, one S160 string per coordinate row."""
    total = int(coords.shape[0])
    mag = level0_mag if level0_mag else "na"
    tmag = target_mag if target_mag else "na"
    out = [f"{stem}__x{x}_y{y}_rw{rw}_rh{rh}_lv{lv}_mag{mag}_tmag{tmag}_total{total}" for x, y, rw, rh, lv in coords.tolist()]
    return np.asarray(out, dtype=PASSPORT_DTYPE).reshape(total)


def write_coords(path: str | os.PathLike, coords: np.ndarray, *, slide_stem: str, wsi_path: str, patch_size: int,
                 patch_size_level0: int, level0_mag: int, target_mag: int, level0_wh: tuple[int, int], step_size: int | None = None,
                 write_batch: int = 8192, extra_file_attrs: Mapping[str, Any] | None = None, h5=None) -> int:
    """a10: coords + passports + file attributes, atomically (tmp file in the target directory, then os.replace)."""
    h5 = h5 or _h5py()
    coords = np.ascontiguousarray(coords, dtype=np.int32).reshape(-1, 5)
    n = int(coords.shape[0])
    target = os.path.abspath(os.fspath(path))
    tmp = os.path.join(os.path.dirname(target) or ".", f".{os.path.basename(target)}.tmp.{uuid.uuid4().hex}")
    rows = max(1, int(write_batch))
    step = int(step_size or patch_size)
    attrs: dict[str, Any] = {
        "patch_size": int(patch_size), "patch_size_level0": int(patch_size_level0), "level0_magnification": int(level0_mag),
        "target_magnification": int(target_mag), "overlap": max(0, int(patch_size) - step), "level0_width": int(level0_wh[0]),
        "level0_height": int(level0_wh[1]), "wsi_path": str(wsi_path), "passport_format": PASSPORT_FORMAT, "passport_version": 2,
        "creation_date": datetime.now(timezone.utc).isoformat(), "filename": Path(wsi_path).name,
    }
    attrs.update(dict(extra_file_attrs or {}))
    attrs = {k: (json.dumps(v) if isinstance(v, dict) else v) for k, v in attrs.items()}   # utils/h5.py:69-73
    f = h5.File(tmp, "w")
    try:
        dc = f.create_dataset("coords", shape=(0, 5), maxshape=(None, 5), chunks=(rows, 5), dtype=np.int32)
        dp = f.create_dataset("passports", shape=(0,), maxshape=(None,), chunks=(rows,), dtype=PASSPORT_DTYPE)
        pp = passports(slide_stem, coords, level0_mag, target_mag)
        for s in range(0, n, rows):                          # the reference appends write_batch rows at a time
            e = min(n, s + rows)
            dc.resize(e, axis=0)
            dc[s:e] = coords[s:e]
            dp.resize(e, axis=0)
            dp[s:e] = pp[s:e]
        for k, v in attrs.items():
            f.attrs[k] = "None" if v is None else v
        f.attrs["num_patches"] = n
        f.close()
        os.replace(tmp, target)
    except Exception:
        try:
            f.close()
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
        raise
    return n


def append_features(path: str | os.PathLike, name: str, feats: np.ndarray, *, feature_batch: int = 32,
                    expected_total: int | None = None, h5=None) -> int:
    """a14: features/<name> float32 (N, D): written under __tmp_<name> and moved into place once the row count is verified."""
    h5 = h5 or _h5py()
    arr = np.asarray(feats, dtype=np.float32)
    if arr.ndim != 2:
        raise ValueError(f"Feature extractor '{name}' must return a 2D array, got shape {arr.shape}")
    if expected_total is not None and arr.shape[0] != int(expected_total):
        raise ValueError(f"Feature rows written ({arr.shape[0]}) do not match expected coords ({expected_total})")
    batch = max(1, int(feature_batch))
    tmp = f"__tmp_{name}"
    with h5.File(os.fspath(path), "a") as f:
        grp = f.require_group("features")
        if name in grp:
            raise ValueError(f"Feature dataset '{name}' already exists in {path}.")
        if tmp in grp:
            del grp[tmp]
        try:
            ds = grp.create_dataset(tmp, shape=(0, arr.shape[1]), maxshape=(None, arr.shape[1]), chunks=(batch, arr.shape[1]),
                                    dtype=np.float32)
            for s in range(0, arr.shape[0], batch):
                e = min(arr.shape[0], s + batch)
                ds.resize((e, arr.shape[1]))
                ds[s:e, :] = arr[s:e]
            grp.move(tmp, name)
        except Exception:
            if tmp in grp:
                del grp[tmp]
            raise
    return int(arr.shape[0])


def save_patch_images(wsi, coords: np.ndarray, image_dir: str | os.PathLike, stem: str, *, patch_size: int) -> int:
    """--save-images (services/storage.py:163-248, services/extraction.py:105-113): one PNG per coordinate row,
    `<image_dir>/<stem>_x{X}_y{Y}.png`, the read resized to the patch size with cv2.resize when it differs, written by a pool of
    max(2, min(8, cpu_count)) threads with at most 4 x workers saves pending."""
    import concurrent.futures as fut
    from collections import deque

    from PIL import Image

    image_dir = Path(image_dir)
    image_dir.mkdir(parents=True, exist_ok=True)
    workers = max(2, min(8, os.cpu_count() or 4))
    pending: deque = deque()
    n = 0
    with fut.ThreadPoolExecutor(max_workers=workers, thread_name_prefix="patch-img") as pool:
        for x, y, rw, rh, lv in np.asarray(coords).reshape(-1, 5).tolist():
            patch = wsi.extract((int(x), int(y)), int(lv), (int(rw), int(rh)), mode="array")
            if patch.shape[0] != patch_size or patch.shape[1] != patch_size:
                import cv2

                patch = cv2.resize(patch, (patch_size, patch_size))
            out = image_dir / f"{stem}_x{int(x)}_y{int(y)}.png"
            pending.append(pool.submit(lambda a, o: Image.fromarray(a).save(str(o)), np.ascontiguousarray(patch), out))
            n += 1
            if len(pending) >= workers * 4:
                pending.popleft().result()
        while pending:
            pending.popleft().result()
    return n


def write_result(path: str | os.PathLike, result, *, wsi, cfg, write_batch: int = 8192, feature_batch: int = 32, h5=None) -> Path:
    """ExtractionResult (services.py) -> the reference's H5: coords first, then one dataset per embedded feature set.
    File attributes as services/extraction.py:146-164 passes them: filename = the slide's file name, then wsi.metadata_attrs()
    (mpp, magnification, vendor ...)."""
    extra = {"filename": Path(result.slide.path).name}
    extra.update(wsi.metadata_attrs() if hasattr(wsi, "metadata_attrs") else {})
    write_coords(path, result.coords, slide_stem=result.slide.stem, wsi_path=str(wsi.path), patch_size=cfg.patch_size,
                 patch_size_level0=int(result.patch_size_level0), level0_mag=int(wsi.mag or 0), target_mag=cfg.target_magnification,
                 level0_wh=tuple(int(v) for v in wsi.get_size(lv=0)), step_size=cfg.step_size, write_batch=write_batch,
                 extra_file_attrs=extra, h5=h5)
    for name, feats in result.features.items():
        append_features(path, name, feats, feature_batch=feature_batch, expected_total=result.num_patches, h5=h5)
    result.h5_path = Path(path)
    return result.h5_path
