"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference's own services (imported through oracle/refimport.py: from
/root/reference in the build container, from baseline/_ref -- installed by oracle/make_ref.sh -- on the GPU box).

Users: bench.py --impl reference (the reference's CPU pipeline as the timed baseline, cpu_baseline.kind "reference"), and the
reference-in-the-loop tests (tests/test_ref_in_loop.py on the CPU, tests/test_gpu_ref_in_loop.py on the GPU), which run

    PatchExtractionService.extract                      services/extraction.py:131-197
    PatchFeatureEmbeddingService._embed_with_extractor   services/feature_embedding.py:179-249
    H5PatchWriter / H5AppendWriter                      services/storage.py, utils/h5.py   (through h5py, or atlaspatch_b200.h5lite
                                                                                               where no HDF5 library exists)

with either the reference's own CPU extractor or the B200 plug-in / backend swapped in at the reference's seams.
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Mapping

import numpy as np

from oracle.refimport import import_reference, reference_available  # noqa: F401


def host_synthetic_wsi_class():
    """Synthetic slide behind the reference's IWSI ABC, rendered on the HOST region by region (what a CPU-only user has)."""
    import_reference()
    from atlas_patch.core.wsi.iwsi import IWSI
    from PIL import Image

    from atlaspatch_b200.synthetic import render_region_host

    class HostSyntheticWSI(IWSI):
        def __init__(self, spec, path=None):
            super().__init__(path=path or f"synthetic_{spec.width}x{spec.height}_s{spec.seed}.synth", mpp=spec.mpp)
            self.spec = spec
            self._ensure_loaded()

        def _setup(self):
            self.w, self.h = self.spec.width, self.spec.height
            self.nlvl, self.ds, self.dims = 1, [1.0], [(self.w, self.h)]
            self.meta = {}
            self.mpp = self._extract_mpp()
            self.mag = self._extract_mag()

        def _extract_mpp(self):
            return self.validate_mpp(float(self._mpp_manual), source="manual")

        def _extract_mag(self):
            return self._infer_mag(self.mpp)

        def extract(self, xy, lv, wh, *, mode="array"):
            arr = render_region_host(self.spec, int(xy[0]), int(xy[1]), int(wh[0]), int(wh[1]))
            return arr if mode == "array" else Image.fromarray(arr)

        def get_size(self, lv=0):
            return self.w, self.h

        def get_thumb(self, max_hw):
            t = self.get_thumbnail_at_power(power=1.25)
            t.thumbnail(max_hw)
            return t

        def cleanup(self):
            pass

    return HostSyntheticWSI


def reference_vit_builder(state_dict: Mapping, *, name: str = "vit_b_16", num_workers: int = 4, device: str = "cpu"):
    """Builder for the reference's registry: its own PatchFeatureExtractor (models/patch/base.py:46-107) around torchvision's
    ViT with heads -> Identity and the ImageClassification preset, exactly as build_torchvision_extractor assembles it
    (base.py:148-180) -- except that the weights are the seeded state_dict (no checkpoint can be downloaded here)."""
    import_reference()
    import torch
    from atlas_patch.models.patch.base import PatchFeatureExtractor
    from torchvision import models

    enum = {"vit_b_16": "ViT_B_16_Weights", "vit_l_16": "ViT_L_16_Weights", "vit_b_32": "ViT_B_32_Weights",
            "vit_l_32": "ViT_L_32_Weights"}[name]

    def build():
        model = getattr(models, name)(weights=None)
        dim = model.heads.head.in_features
        model.heads = torch.nn.Identity()
        model.load_state_dict({k: (v if hasattr(v, "detach") else torch.from_numpy(np.asarray(v))) for k, v in state_dict.items()},
                              strict=True)
        preprocess = getattr(models, enum).IMAGENET1K_V1.transforms()
        return PatchFeatureExtractor(name=name, model=model, embedding_dim=dim, preprocess=preprocess, device=torch.device(device),
                                     dtype=torch.float32, num_workers=num_workers)

    return build


def reference_services(out_root, *, patch_size: int, target_mag: int, step_size: int | None = None, tissue_threshold: float = 0.0,
                       fast_mode: bool = True, extractors: Mapping[str, Callable] | None = None, feature_batch: int = 32,
                       num_workers: int = 4, device: str = "cpu"):
    """(PatchExtractionService, PatchFeatureEmbeddingService | None) of the reference, configured like cli.py:238-288 does."""
    import_reference()
    from atlas_patch.core.config import ExtractionConfig, FeatureExtractionConfig, OutputConfig
    from atlas_patch.models.patch.registry import PatchFeatureExtractorRegistry
    from atlas_patch.services.extraction import PatchExtractionService
    from atlas_patch.services.feature_embedding import PatchFeatureEmbeddingService

    ecfg = ExtractionConfig(patch_size=patch_size, target_magnification=target_mag, step_size=step_size,
                            tissue_threshold=tissue_threshold, fast_mode=fast_mode)
    ocfg = OutputConfig(output_root=Path(out_root))
    extraction = PatchExtractionService(ecfg, ocfg)
    embedding = None
    if extractors:
        reg = PatchFeatureExtractorRegistry()
        for n, b in extractors.items():
            reg.register(n, b)
        fcfg = FeatureExtractionConfig(extractors=list(extractors), batch_size=feature_batch, device=device, num_workers=num_workers,
                                       precision="float32")
        embedding = PatchFeatureEmbeddingService(ecfg, ocfg, fcfg, registry=reg)
    return extraction, embedding


def reference_slide(path, mpp=None):
    import_reference()
    from atlas_patch.core.models import Slide

    return Slide(path=Path(path), mpp=mpp)


def write_reference_coords(embedding_or_extraction, wsi, slide, coords: np.ndarray, *, patch_size_level0: int):
    """An H5 holding exactly `coords`, written by the reference's own H5PatchWriter.write_coords (services/storage.py:106-161);
    returns the reference ExtractionResult that _embed_with_extractor takes."""
    import_reference()
    from atlas_patch.core.models import ExtractionResult
    from atlas_patch.core.paths import patch_h5_path
    from atlas_patch.services.storage import H5PatchWriter

    svc = embedding_or_extraction
    cfg, ocfg = svc.cfg, svc.output_cfg
    out = patch_h5_path(slide, ocfg, cfg)
    out.parent.mkdir(parents=True, exist_ok=True)
    step = cfg.step_size or cfg.patch_size
    w = H5PatchWriter(chunk_rows=cfg.write_batch, patch_size=cfg.patch_size, patch_size_level0=patch_size_level0,
                      level0_mag=int(wsi.mag), target_mag=cfg.target_magnification, level0_wh=wsi.get_size(lv=0),
                      overlap=max(0, int(cfg.patch_size) - int(step)), slide_stem=slide.stem, wsi_path=str(wsi.path),
                      extra_file_attrs={"filename": slide.path.name, **wsi.metadata_attrs()})
    total, _ = w.write_coords(out, ((int(x), int(y), int(rw), int(rh), int(lv), None) for x, y, rw, rh, lv in coords.tolist()),
                              batch=cfg.write_batch)
    return ExtractionResult(slide=slide, h5_path=out, num_patches=int(total), patch_size_level0=patch_size_level0)


def read_h5(path):
    """{'coords', 'passports', 'attrs', 'features': {name: array}} of a container file, through the same HDF5 layer the reference used."""
    import h5py   # the real library, or h5lite installed under this name by refimport.install_stubs

    with h5py.File(str(path), "r") as f:
        out = {"coords": np.asarray(f["coords"][...]), "passports": np.asarray(f["passports"][...]), "attrs": dict(f.attrs.items()),
               "features": {}}
        if "features" in f:
            for k, ds in f["features"].items():
                out["features"][k] = np.asarray(ds[...])
    return out
