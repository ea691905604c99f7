"""Service adapters with the reference's seam-#1 signatures (atlas_patch/services/interfaces.py:12-32).

`B200PatchExtractionService.extract(wsi, mask, *, slide)` replaces PatchExtractionService.extract
(services/extraction.py:131-197) for the coordinate path (fast mode, and the black/white content filter of --no-fast-mode on an HBM-resident slide):
same contours / geometry / order, coordinates computed by the CUDA kernels.  `B200FeatureEmbeddingService.embed_features(result, *, wsi)` replaces
PatchFeatureEmbeddingService._embed_with_extractor (services/feature_embedding.py:179-249) with the zero-copy path
when the slide is resident in HBM.  Results are returned in memory; `storage.write_result` writes them into the reference's H5
container through h5py (services/storage.py layout) where that library is installed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

from atlaspatch_b200.extraction import (extract_coords_from_contours, filter_patches, flatten_contours, mask_to_contours,
                                        scale_contours)
from atlaspatch_b200.geometry import prepare_geometry


@dataclass(frozen=True)
class Slide:  # core/models.py:10-18
    path: Path
    mpp: float | None = None
    backend: str | None = None

    @property
    def stem(self) -> str:
        return Path(self.path).stem


@dataclass
class ExtractionResult:  # core/models.py:27-36 (h5_path is set by storage.write_result)
    slide: Slide
    h5_path: Path | None
    num_patches: int
    image_dir: Path | None = None
    visualizations: dict[str, Path] = field(default_factory=dict)
    metadata: dict[str, Any] = field(default_factory=dict)
    coords: np.ndarray | None = None
    patch_size_level0: int | None = None
    coords_device: Any = None          # int32 (N, 5) CUDA tensor for the embed step
    features: dict[str, np.ndarray] = field(default_factory=dict)


@dataclass
class ExtractionConfig:  # core/config.py:62-89 (fields used on the path; CLI default tissue_threshold is 0.0, cli.py:85-91)
    patch_size: int
    target_magnification: int
    step_size: int | None = None
    tissue_threshold: float = 0.0
    white_threshold: int = 15
    black_threshold: int = 50
    fast_mode: bool = True

    def validated(self) -> "ExtractionConfig":
        if self.patch_size <= 0 or self.target_magnification <= 0:
            raise ValueError("patch_size and target_magnification must be > 0")
        if self.step_size is None:
            self.step_size = self.patch_size
        if self.step_size <= 0:
            raise ValueError("step_size must be > 0")
        if not (0 <= self.tissue_threshold <= 1):
            raise ValueError("tissue_threshold must be between 0 and 1")
        if self.white_threshold <= 0 or self.black_threshold <= 0:
            raise ValueError("white_threshold and black_threshold must be > 0")
        return self


class B200PatchExtractionService:
    def __init__(self, extraction_cfg: ExtractionConfig):
        self.cfg = extraction_cfg.validated()

    def _prepare_contours(self, mask: np.ndarray, wsi):  # services/extraction.py:30-42
        tissue_t, holes_t = mask_to_contours(mask, tissue_area_thresh=self.cfg.tissue_threshold)
        W, H = wsi.get_size(lv=0)
        mh, mw = mask.shape[:2]
        sx, sy = W / float(mw), H / float(mh)
        return scale_contours(tissue_t, sx, sy), [scale_contours(hs, sx, sy) for hs in holes_t]

    def _prepare_geometry(self, wsi):  # services/extraction.py:44-64
        return prepare_geometry(src_mag=wsi.mag, target_mag=self.cfg.target_magnification, patch_size=self.cfg.patch_size,
                                step_size=self.cfg.step_size, downsamples=wsi.ds or [1.0])

    def extract(self, wsi, mask: np.ndarray, *, slide: Slide) -> ExtractionResult:
        tissue, holes = self._prepare_contours(np.asarray(mask), wsi)
        geo = self._prepare_geometry(wsi)
        coords, coords_dev = extract_coords_from_contours(flatten_contours(tissue, holes), geo, return_device=True)
        if not self.cfg.fast_mode and coords.shape[0]:  # services/extraction.py:105-119
            if not hasattr(wsi, "device_image"):
                raise RuntimeError("fast_mode=False needs the slide resident in device memory (wsi.device_image); "
                                   "there is no host fallback")
            _require_level0(coords, "extract(fast_mode=False)")
            coords, coords_dev = filter_patches(wsi.device_image, wsi.w, wsi.h, wsi.pitch, coords_dev,
                                                patch_size=self.cfg.patch_size, black_threshold=self.cfg.black_threshold,
                                                white_threshold=self.cfg.white_threshold)
        return ExtractionResult(slide=slide, h5_path=None, num_patches=int(coords.shape[0]), coords=coords,
                                patch_size_level0=geo.patch_size_level0, coords_device=coords_dev)


def _require_level0(coords: np.ndarray, what: str) -> None:
    """The device paths read level-0 pixels of `wsi.device_image`; rows that name another pyramid level would silently get the
    wrong field of view (services/extraction.py:44-64 picks level > 0 for multi-level slides)."""
    if coords.shape[0] and not (np.asarray(coords)[:, 4] == 0).all():
        raise NotImplementedError(f"{what}: coordinate rows on pyramid level > 0 need that level resident in device memory; "
                                  "only single-level (level 0) device slides are supported")


class B200FeatureEmbeddingService:
    """One extractor at a time, like embed_all (services/feature_embedding.py:251-316).  `extraction_cfg` carries the patch size
    the reference resizes every read to (feature_embedding.py:93-95); without it the extractor's own input size is used."""

    def __init__(self, extractor, extraction_cfg: ExtractionConfig | None = None):
        self.extractor = extractor
        self.cfg = extraction_cfg.validated() if extraction_cfg is not None else None

    @property
    def patch_size(self) -> int:
        return int(self.cfg.patch_size) if self.cfg is not None else int(getattr(self.extractor, "input_patch", 0) or 0)

    def embed_features(self, result: ExtractionResult, *, wsi, shard_group=None, sharded: bool = False) -> ExtractionResult:
        """sharded=True (intra-slide mode, BASELINE.json configs[4]): every rank of `shard_group` holds the same slide and
        coordinate rows, embeds its contiguous row range and all-gathers the (N, D) matrix (sharding.gather_rows)."""
        name = self.extractor.name
        if result.num_patches == 0:
            result.features[name] = np.empty((0, self.extractor.embedding_dim), dtype=np.float32)
        elif (hasattr(wsi, "device_image") and result.coords_device is not None and
              (int(result.coords[0, 2]) == int(getattr(self.extractor, "input_patch", result.coords[0, 2]))
               or getattr(self.extractor, "supports_large_reads", True))):
            # (an extractor whose CUDA preprocess resizes -- DINOv2 / hub families -- reads exactly its patch size on the device; a larger
            # read goes through the reference-style host path below, cv2.resize included)
            _require_level0(result.coords, "embed_features")
            rows = result.coords_device.contiguous()
            if sharded:
                import torch.distributed as dist

                from atlaspatch_b200.sharding import gather_rows, row_range

                b, e = row_range(result.num_patches, dist.get_rank(shard_group), dist.get_world_size(shard_group))
                local = self.extractor.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows[b:e].contiguous(),
                                                    read_size=int(result.coords[0, 2]))
                feats = gather_rows(local, result.num_patches, group=shard_group)
            else:
                feats = self.extractor.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows, read_size=int(result.coords[0, 2]))
            result.features[name] = feats.cpu().numpy()
        else:  # reference-style host reads (feature_embedding.py:81-96): read, cv2.resize to the patch size if the read differs
            P = self.patch_size
            patches = []
            for x, y, rw, rh, lv in result.coords.tolist():
                patch = wsi.extract((int(x), int(y)), int(lv), (int(rw), int(rh)))
                if P and (patch.shape[0] != P or patch.shape[1] != P):
                    import cv2

                    patch = cv2.resize(patch, (P, P))
                patches.append(patch)
            result.features[name] = self.extractor.extract_batch(patches, batch_size=32)
        result.metadata.setdefault("feature_sets", []).append(name)
        return result
