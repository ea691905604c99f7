"""Per-GPU slide scheduler with the reference runner's semantics.

reference: atlas_patch/orchestration/runner.py:106-181 (skip / reuse of existing outputs, O_CREAT|O_EXCL lock files),
:202-306 (per-slide try / except -> failures list, the run continues), orchestration/parallel.py:149-160 (lock released when
the slide's extraction is done), services/feature_embedding.py:98-126,179-316 (feature lock, "feature already present" check,
one failure entry per (slide, extractor)), utils/features.py:37-71 (a feature set counts only if its row count equals
num_patches), core/paths.py (output layout).

The reference scales out by starting one OS process per GPU and letting the lock files arbitrate (README.md:527,628).  Here
the ranks of a `torch.distributed` group take the slides the longest-processing-time-first assignment gives them
(sharding.assign_slides) -- no data-path collective -- and the lock files still guard against a SECOND run (or the reference
itself) working in the same output directory.  Every H5 is written by the rank that owns the slide; the (results, failures)
lists are gathered to every rank at the end.

Per slide, on its rank:  skip / reuse check -> lock -> open -> segment -> extract -> write coords H5 (atomic) [-> PNG export]
-> unlock;  then per requested encoder: feature lock -> present? -> embed -> append features/<name> (tmp dataset, then move)
-> unlock.  Encoders are built once per rank on first use and kept until the run ends (180 GB of HBM hold every registered
encoder beside a slide; the reference rebuilds the model per encoder and re-opens every slide instead).
"""
from __future__ import annotations

import logging
import os
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Mapping, Sequence

import numpy as np

from atlaspatch_b200 import storage
from atlaspatch_b200.services import ExtractionConfig, ExtractionResult, Slide
from atlaspatch_b200.sharding import assign_slides

logger = logging.getLogger("atlaspatch_b200.runner")


@dataclass
class RunConfig:
    output_root: Path
    extraction: ExtractionConfig
    feature_extractors: list[str] = field(default_factory=list)
    feature_batch: int = 32
    write_batch: int = 8192
    skip_existing: bool = True          # cli.py:140-142: --skip-existing is the default, --force turns it off
    save_images: bool = False

    def validated(self) -> "RunConfig":
        self.output_root = Path(self.output_root)
        self.extraction = self.extraction.validated()
        self.feature_extractors = [n.lower() for n in self.feature_extractors]
        if len(set(self.feature_extractors)) != len(self.feature_extractors):
            raise ValueError("Duplicate feature extractor names")
        if self.feature_batch <= 0 or self.write_batch <= 0:
            raise ValueError("feature_batch and write_batch must be > 0")
        return self


# ---- output layout (core/paths.py:9-42) -----------------------------------------------------------------------------
def patch_h5_path(slide: Slide, cfg: RunConfig) -> Path:
    return Path(cfg.output_root) / "patches" / f"{slide.stem}.h5"


def patch_lock_path(slide: Slide, cfg: RunConfig) -> Path:
    return Path(cfg.output_root) / "patches" / f"{slide.stem}.lock"


def images_dir(slide: Slide, cfg: RunConfig) -> Path:
    return Path(cfg.output_root) / "images" / slide.stem


# ---- lock files (runner.py:154-181, feature_embedding.py:98-126) -----------------------------------------------------
def acquire_lock(path: Path, slide: Slide, phase: str | None = None) -> int | None:
    """fd when the lock file could be created exclusively, None when somebody else holds it."""
    path.parent.mkdir(parents=True, exist_ok=True)
    payload = f"pid={os.getpid()},time={int(time.time())},slide={slide.path}" + (f",phase={phase}" if phase else "")
    try:
        fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
    except FileExistsError:
        return None
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(f"Failed to create lock {path}: {e}") from e
    os.write(fd, payload.encode())
    os.fsync(fd)
    return fd


def release_lock(fd: int | None, path: Path) -> None:
    if fd is not None:
        try:
            os.close(fd)
        except OSError:
            pass
    try:
        path.unlink()
    except OSError:
        pass


# ---- existing outputs (runner.py:71-152, utils/features.py:37-71) -----------------------------------------------------
def existing_features(h5_path: Path, expected_total: int | None) -> set[str]:
    h5 = storage._h5py()
    try:
        with h5.File(str(h5_path), "r") as f:
            if "features" not in f:
                return set()
            out = set()
            for name, ds in f["features"].items():
                if expected_total is not None and int(ds.shape[0]) != int(expected_total):
                    continue                                  # a partial embedding does not count
                out.add(str(name).lower())
            return out
    except Exception:  # noqa: BLE001  unreadable file: treat as missing so that it is regenerated
        return set()


def missing_features(h5_path: Path, required: Sequence[str], expected_total: int | None) -> list[str]:
    have = existing_features(h5_path, expected_total)
    return [n.lower() for n in required if n.lower() not in have]


def load_existing_result(slide: Slide, h5_path: Path) -> ExtractionResult | None:
    """Lightweight result from an existing H5 (no re-segmentation); None when the file is unreadable or holds no patches."""
    h5 = storage._h5py()
    try:
        with h5.File(str(h5_path), "r") as f:
            n = f.attrs.get("num_patches")
            coords = np.asarray(f["coords"][...], dtype=np.int32) if "coords" in f else None
            n = int(n) if n is not None else (int(coords.shape[0]) if coords is not None else None)
            p0 = f.attrs.get("patch_size_level0")
    except Exception as e:  # noqa: BLE001
        logger.warning("Failed to read existing output for %s; will reprocess. Error: %s", slide.path.name, e)
        return None
    if n is None or n <= 0 or coords is None or coords.shape[0] != n:
        return None
    return ExtractionResult(slide=slide, h5_path=h5_path, num_patches=n, coords=coords,
                            patch_size_level0=int(p0) if p0 is not None else None)


def default_slide_cost(slide: Slide) -> int:
    """Relative processing cost for the assignment: level-0 pixels for .synth descriptors, else the file size."""
    p = Path(slide.path)
    try:
        if p.suffix.lower() == ".synth":
            from atlaspatch_b200.ref_backend import read_synth_descriptor

            d = read_synth_descriptor(p)
            return d["width"] * d["height"]
        return max(1, p.stat().st_size)
    except Exception:  # noqa: BLE001
        return 1


class B200Runner:
    """run(slides) -> (results, failures), like ProcessingRunner.run + PatchFeatureEmbeddingService.embed_all.

    segmentation:  object with segment_thumbnail(wsi) -> Mask                     (services/interfaces.py:12-18)
    extraction:    object with extract(wsi, mask_array, *, slide) -> result       (:20-24; B200PatchExtractionService)
    wsi_loader:    object with open(slide) -> wsi                                 (:34-40)
    extractor_builders: {name: () -> FeatureExtractor}                            (the registry's builders, registry.py:11-44)
    embed:         (extractor, result, wsi) -> (N, D) float32 features; default = B200FeatureEmbeddingService (device path when
                   the slide is resident in HBM, reference-style host reads otherwise)
    group:         torch.distributed process group (None: the default group when initialised, else a single process)
    """

    def __init__(self, cfg: RunConfig, *, segmentation, extraction, wsi_loader, extractor_builders: Mapping[str, Callable[[], Any]] | None = None,
                 embed: Callable | None = None, mpp_resolver=None, slide_cost: Callable[[Slide], int] = default_slide_cost, group=None):
        self.cfg = cfg.validated()
        self.segmentation, self.extraction, self.wsi_loader = segmentation, extraction, wsi_loader
        self.builders = {k.lower(): v for k, v in (extractor_builders or {}).items()}
        unknown = [n for n in self.cfg.feature_extractors if n not in self.builders]
        if unknown:
            raise KeyError(f"Unknown feature extractor(s): {', '.join(unknown)}. Available: {', '.join(sorted(self.builders))}")
        self.embed = embed or self._default_embed
        self.mpp_resolver = mpp_resolver
        self.slide_cost = slide_cost
        self.group = group
        self._extractors: dict[str, Any] = {}
        self._extractor_errors: dict[str, Exception] = {}

    # ---- distributed plumbing ----
    def _dist(self):
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                return dist
        except ImportError:
            pass
        return None

    def _rank_world(self) -> tuple[int, int]:
        d = self._dist()
        return (d.get_rank(self.group), d.get_world_size(self.group)) if d else (0, 1)

    # ---- encoders (one instance per rank and name, built on first use) ----
    def _extractor(self, name: str):
        if name in self._extractor_errors:
            raise self._extractor_errors[name]
        if name not in self._extractors:
            try:
                self._extractors[name] = self.builders[name]()
            except Exception as e:  # noqa: BLE001
                self._extractor_errors[name] = e
                raise
        return self._extractors[name]

    def _default_embed(self, extractor, result: ExtractionResult, wsi) -> np.ndarray:
        from atlaspatch_b200.services import B200FeatureEmbeddingService

        if result.coords_device is None and hasattr(wsi, "device_image") and result.coords is not None and result.num_patches:
            import torch

            result.coords_device = torch.from_numpy(np.ascontiguousarray(result.coords, dtype=np.int32)).cuda()
        svc = B200FeatureEmbeddingService(extractor, self.cfg.extraction)
        return svc.embed_features(result, wsi=wsi).features[extractor.name]

    # ---- one slide ----
    def _existing(self, slide: Slide) -> tuple[str, ExtractionResult | None]:
        """runner.py:106-152 -> ("process" | "skip" | "reuse", result)."""
        if not self.cfg.skip_existing:
            return "process", None
        h5 = patch_h5_path(slide, self.cfg)
        if not h5.exists():
            return "process", None
        if not self.cfg.feature_extractors:
            logger.info("Skipping %s (already processed).", slide.path.name)
            return "skip", None
        res = load_existing_result(slide, h5)
        if res is None:
            logger.info("Existing output invalid for %s; reprocessing.", slide.path.name)
            return "process", None
        missing = missing_features(h5, self.cfg.feature_extractors, res.num_patches)
        if not missing:
            logger.info("Skipping %s (features complete).", slide.path.name)
            return "skip", res
        logger.info("Reusing existing patches for %s; missing features: %s", slide.path.name, ", ".join(missing))
        return "reuse", res

    def _segment_and_extract(self, slide: Slide, wsi) -> ExtractionResult:
        mask = self.segmentation.segment_thumbnail(wsi)
        res = self.extraction.extract(wsi, np.asarray(getattr(mask, "data", mask)), slide=slide)
        h5 = patch_h5_path(slide, self.cfg)
        h5.parent.mkdir(parents=True, exist_ok=True)
        c = self.cfg.extraction
        extra = {"filename": Path(slide.path).name}
        extra.update(wsi.metadata_attrs() if hasattr(wsi, "metadata_attrs") else {})
        storage.write_coords(h5, res.coords, slide_stem=slide.stem, wsi_path=str(wsi.path), patch_size=c.patch_size,
                             patch_size_level0=int(res.patch_size_level0), level0_mag=int(wsi.mag or 0),
                             target_mag=c.target_magnification, level0_wh=tuple(int(v) for v in wsi.get_size(lv=0)),
                             step_size=c.step_size, write_batch=self.cfg.write_batch, extra_file_attrs=extra)
        res.h5_path = h5
        if self.cfg.save_images:
            res.image_dir = images_dir(slide, self.cfg)
            storage.save_patch_images(wsi, res.coords, res.image_dir, slide.stem, patch_size=c.patch_size)
        return res

    def _embed_slide(self, res: ExtractionResult, wsi, failures: list) -> None:
        """services/feature_embedding.py:179-249 for every requested encoder that is still missing in the slide's H5."""
        for name in self.cfg.feature_extractors:
            if name in existing_features(res.h5_path, res.num_patches):
                continue
            lock = patch_lock_path(res.slide, self.cfg)
            fd = None
            try:
                extractor = self._extractor(name)
                fd = acquire_lock(lock, res.slide, phase="features")
                if fd is None:
                    logger.info("Skipping feature embedding for %s (locked by another process).", res.slide.path.name)
                    continue
                if extractor.name.lower() in existing_features(res.h5_path, res.num_patches):
                    continue
                feats = self.embed(extractor, res, wsi)
                storage.append_features(res.h5_path, extractor.name, feats, feature_batch=self.cfg.feature_batch,
                                        expected_total=res.num_patches)
            except Exception as e:  # noqa: BLE001
                failures.append((res.slide, e))
                logger.error("Feature embedding '%s' failed for %s: %s", name, res.slide.path.name, e)
            finally:
                if fd is not None:
                    release_lock(fd, lock)
        have = sorted(existing_features(res.h5_path, res.num_patches))
        if have:
            res.metadata["feature_sets"] = have

    def _run_slide(self, slide: Slide, results: list, failures: list) -> None:
        action, existing = self._existing(slide)
        if action == "skip":
            return
        wsi, res = None, existing
        try:
            if action == "process":
                lock = patch_lock_path(slide, self.cfg)
                fd = acquire_lock(lock, slide)
                if fd is None:
                    logger.info("Skipping %s (locked by another process).", slide.path.name)
                    return
                try:
                    try:
                        wsi = self.wsi_loader.open(slide)
                    except Exception as e:  # noqa: BLE001
                        failures.append((slide, e))
                        logger.error("Failed to open %s: %s", slide.path.name, e)
                        return
                    try:
                        res = self._segment_and_extract(slide, wsi)
                    except Exception as e:  # noqa: BLE001
                        failures.append((slide, e))
                        logger.error("Segmentation / extraction failed for %s: %s", slide.path.name, e)
                        return
                finally:
                    release_lock(fd, lock)
            results.append(res)
            if self.cfg.feature_extractors and res.num_patches > 0:
                if wsi is None:
                    try:
                        wsi = self.wsi_loader.open(slide)
                    except Exception as e:  # noqa: BLE001
                        failures.append((slide, e))
                        return
                self._embed_slide(res, wsi, failures)
        finally:
            if wsi is not None:
                try:
                    wsi.cleanup()
                except Exception:  # noqa: BLE001
                    pass

    # ---- the run ----
    def run(self, slides: Sequence[Slide]) -> tuple[list[ExtractionResult], list[tuple[Slide, Exception | str]]]:
        slides = list(slides)
        if self.mpp_resolver is not None:
            slides = [Slide(path=s.path, mpp=self.mpp_resolver.resolve(s), backend=s.backend) for s in slides]
        if not slides:
            logger.warning("No slides found to process.")
            return [], []
        rank, world = self._rank_world()
        mine = assign_slides([self.slide_cost(s) for s in slides], world)[rank]
        results: list[ExtractionResult] = []
        failures: list[tuple[Slide, Exception | str]] = []
        try:
            for i in mine:
                self._run_slide(slides[i], results, failures)
        finally:
            for ext in self._extractors.values():
                try:
                    ext.cleanup()
                except Exception:  # noqa: BLE001
                    pass
            self._extractors.clear()
        d = self._dist()
        if d and world > 1:
            # the lists travel as plain data (paths, counts, messages); device tensors stay on their rank
            light = [dict(path=str(r.slide.path), mpp=r.slide.mpp, h5=str(r.h5_path), n=r.num_patches, p0=r.patch_size_level0,
                          feats=r.metadata.get("feature_sets", [])) for r in results]
            fails = [(str(s.path), f"{type(e).__name__}: {e}" if isinstance(e, Exception) else str(e)) for s, e in failures]
            gathered: list = [None] * world
            d.all_gather_object(gathered, (light, fails), group=self.group)
            results, failures = [], []
            for lg, fl in gathered:
                for r in lg:
                    results.append(ExtractionResult(slide=Slide(Path(r["path"]), mpp=r["mpp"]), h5_path=Path(r["h5"]), num_patches=r["n"],
                                                    patch_size_level0=r["p0"], metadata={"feature_sets": r["feats"]} if r["feats"] else {}))
                failures += [(Slide(Path(p)), msg) for p, msg in fl]
        return results, failures
