"""Segmentation seam around a pluggable SAM2 forward (row a4).

The reference's SAM2SegmentationService (atlas_patch/services/segmentation.py:195-236) does, per slide:
  a1  wsi.get_thumbnail_at_power(1.25)                       -> device kernel here (slide.py / ap_thumbnail_area)
  a2  thumb.thumbnail((1024, 1024))                          -> Pillow, same call (host; <= 3 MB)
  a3  PIL resize to 1024 x 1024, BILINEAR                     -> Pillow, same call (host)
  a4  SAM2ImagePredictor.set_image + predict(box=full image)  -> `predict_logits` callable = sam2.py: B200Sam2Predictor
      (ap_sam2_* : Hiera trunk + FPN + prompt encoder + mask decoder in CUDA)
  a5  mask > threshold, *255, PIL NEAREST back to thumbnail   -> Pillow, same call (host)
Pillow's fixed-point resamplers define the exact result of a2/a3/a5 and the images are tiny, so these steps deliberately stay
on the host with the very same library calls (SURVEY.md section 2.3 "Placement guidance").
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np
from PIL import Image

SAM_INPUT = 1024


@dataclass
class Mask:  # core/models.py:21-24
    data: np.ndarray
    source_shape: tuple[int, int]


def cap_thumbnail(thumb: Image.Image, thumbnail_max: int = 1024) -> Image.Image:
    """a2: segmentation.py:202-206 (`Image.thumbnail` is in place; a copy is returned here)."""
    t = thumb.copy()
    if thumbnail_max:
        t.thumbnail((thumbnail_max, thumbnail_max))
    return t


def resize_for_sam(image: np.ndarray) -> tuple[np.ndarray, tuple[int, int]]:
    """a3: segmentation.py:104-110."""
    shape = (int(image.shape[0]), int(image.shape[1]))
    if shape == (SAM_INPUT, SAM_INPUT):
        return image, shape
    resized = Image.fromarray(image).resize((SAM_INPUT, SAM_INPUT), Image.Resampling.BILINEAR)
    return np.array(resized, copy=True), shape


def resize_mask(mask: np.ndarray, target_shape: tuple[int, int]) -> np.ndarray:
    """a5: segmentation.py:112-118."""
    m8 = (mask * 255).astype(np.uint8)
    out = Image.fromarray(m8, mode="L").resize((target_shape[1], target_shape[0]), resample=Image.Resampling.NEAREST)
    return np.asarray(out, dtype=np.float32) / 255.0


class B200SegmentationService:
    """SegmentationService contract (services/interfaces.py:12-18): segment_thumbnail(wsi) / segment_batch(wsis) -> Mask."""

    def __init__(self, predict_logits: Callable[[np.ndarray], np.ndarray] | None = None, *, thumbnail_power: float = 1.25,
                 thumbnail_max: int = 1024, mask_threshold: float = 0.0):
        self.predict_logits = predict_logits
        self.thumbnail_power, self.thumbnail_max, self.mask_threshold = thumbnail_power, thumbnail_max, mask_threshold

    def _prepare_thumbnail(self, wsi) -> Image.Image:
        return cap_thumbnail(wsi.get_thumbnail_at_power(power=self.thumbnail_power, interpolation="optimise"), self.thumbnail_max)

    def segment_thumbnail(self, wsi) -> Mask:
        if self.predict_logits is None:
            raise NotImplementedError("the SAM2 forward (SURVEY.md section 8 row a4) is not built in this round; pass predict_logits=...")
        thumb = np.array(self._prepare_thumbnail(wsi).convert("RGB"), copy=True)
        sam_in, orig = resize_for_sam(thumb)
        logits = np.asarray(self.predict_logits(sam_in), dtype=np.float32)        # (1024, 1024) mask logits
        mask = (logits > self.mask_threshold).astype(np.float32)                  # SAM2ImagePredictor(return_logits=False)
        if mask.shape[:2] != orig:
            mask = resize_mask(mask, orig)
        return Mask(data=mask.astype(np.float32), source_shape=(int(mask.shape[0]), int(mask.shape[1])))

    def segment_batch(self, wsis: Sequence) -> list[Mask]:
        """services/segmentation.py:216-229: thumbnails are prepared by up to 8 threads (device thumbnail kernel + Pillow cap), then
        the batch goes through the predictor's batch call (`predict_logits_batch` of a bound B200Sam2Predictor method when the
        callable has one, else image by image)."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        if self.predict_logits is None:
            raise NotImplementedError("no SAM2 predictor bound; pass predict_logits=B200Sam2Predictor(...).predict_logits")
        wsis = list(wsis)
        if not wsis:
            return []
        workers = max(1, min(8, len(wsis), os.cpu_count() or 8))
        with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="thumb") as ex:
            thumbs = list(ex.map(lambda w: np.array(self._prepare_thumbnail(w).convert("RGB"), copy=True), wsis))
        ins, shapes = zip(*(resize_for_sam(t) for t in thumbs))
        owner = getattr(self.predict_logits, "__self__", None)
        if owner is not None and hasattr(owner, "predict_logits_batch"):
            logits = owner.predict_logits_batch(np.stack(ins))
        else:
            logits = [np.asarray(self.predict_logits(i), dtype=np.float32) for i in ins]
        out = []
        for lg, orig in zip(logits, shapes):
            mask = (np.asarray(lg) > self.mask_threshold).astype(np.float32)
            if mask.shape[:2] != orig:
                mask = resize_mask(mask, orig)
            out.append(Mask(data=mask.astype(np.float32), source_shape=(int(mask.shape[0]), int(mask.shape[1]))))
        return out
