"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through oracle/refimport.py (stubbed optional deps), plugs a
synthetic-slide backend into the reference's own IWSI ABC (core/wsi/iwsi.py:9-124), and
records what the reference's functions return:

* coords_<case>.npz  -- PatchExtractionService._prepare_contours + _iter_patch_entries
                        (services/extraction.py:30-42,83-128) -> int32 (N,5) rows
* thumb_<case>.npz   -- IWSI.get_thumbnail_at_power(1.25) (core/wsi/iwsi.py:246-323)
* vit_b_16_feats.npz -- PatchFeatureExtractor.extract_batch (models/patch/base.py:76-107)
                        on torchvision vit_b_16 + its ImageClassification preset, fp32 CPU,
                        seeded weights from oracle/weights.py
* library versions the vectors were made with (versions.json)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.refimport import import_reference  # noqa: E402

import_reference()

import cv2  # noqa: E402
import PIL  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402
from PIL import Image  # noqa: E402

from atlas_patch.core.config import ExtractionConfig, OutputConfig  # noqa: E402
from atlas_patch.core.wsi.iwsi import IWSI  # noqa: E402
from atlas_patch.services.extraction import PatchExtractionService  # noqa: E402

from atlaspatch_b200.synthetic import make_spec, render_region_host, truth_mask  # noqa: E402
from tests.cases import COORD_CASES, THUMB_CASES, build_mask  # noqa: E402

OUT = Path(__file__).resolve().parent


class RefSyntheticWSI(IWSI):
    """Synthetic slide behind the reference's own IWSI ABC (single level, ds=[1.0])."""

    def __init__(self, spec):
        super().__init__(path=f"synthetic_{spec.width}x{spec.height}_s{spec.seed}.synth", mpp=spec.mpp)
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


def reference_coords(case) -> np.ndarray:
    spec = make_spec(case["width"], case["height"], case["seed"], mpp=case["mpp"])
    wsi = RefSyntheticWSI(spec)
    mask = build_mask(case, spec)
    with tempfile.TemporaryDirectory() as td:
        svc = PatchExtractionService(
            ExtractionConfig(patch_size=case["patch"], target_magnification=case["target_mag"],
                             step_size=case["step"], tissue_threshold=case["tissue_thresh"]),
            OutputConfig(output_root=Path(td)),
        )
        tissue, holes = svc._prepare_contours(mask, wsi)
        rows = [e[:5] for e in svc._iter_patch_entries(wsi, tissue, holes, include_patch=False)]
    return np.asarray(rows, dtype=np.int32).reshape(-1, 5)


def main() -> None:
    versions = {"cv2": cv2.__version__, "numpy": np.__version__, "PIL": PIL.__version__,
                "torch": torch.__version__, "torchvision": torchvision.__version__}
    for case in COORD_CASES:
        coords = reference_coords(case)
        np.savez_compressed(OUT / f"coords_{case['name']}.npz", coords=coords)
        print(f"coords_{case['name']}: N={coords.shape[0]} first={coords[0].tolist() if len(coords) else None}")

    for case in THUMB_CASES:
        spec = make_spec(case["width"], case["height"], case["seed"], mpp=case["mpp"])
        thumb = np.asarray(RefSyntheticWSI(spec).get_thumbnail_at_power(power=1.25))
        np.savez_compressed(OUT / f"thumb_{case['name']}.npz", thumb=thumb)
        print(f"thumb_{case['name']}: {thumb.shape}")

    # ---- encoder features through the reference's own extractor class -------------------
    from atlas_patch.models.patch.base import PatchFeatureExtractor
    from torchvision import models as tvm

    from oracle.weights import vit_state_dict
    from tests.cases import FEATURE_CASE, feature_patches

    model = tvm.vit_b_16(weights=None)
    model.heads = torch.nn.Identity()
    model.load_state_dict(vit_state_dict("vit_b_16", seed=FEATURE_CASE["weight_seed"]), strict=True)
    preprocess = tvm.ViT_B_16_Weights.IMAGENET1K_V1.transforms()
    ext = PatchFeatureExtractor(name="vit_b_16", model=model, embedding_dim=768, preprocess=preprocess,
                                device=torch.device("cpu"), dtype=torch.float32, num_workers=0)
    patches = feature_patches()
    feats = ext.extract_batch(patches, batch_size=8)
    np.savez_compressed(OUT / "vit_b_16_feats.npz", feats=feats.astype(np.float32))
    print("vit_b_16_feats:", feats.shape, float(np.abs(feats).mean()))

    (OUT / "versions.json").write_text(json.dumps(versions, indent=1) + "\n")


if __name__ == "__main__" and not ({"--sam2", "--sam2-large", "--filter", "--dinov2", "--thumb-general", "--vit-resize", "--hub"} & set(sys.argv)):
    os.environ.setdefault("OMP_NUM_THREADS", "8")
    main()


def make_sam2_golden() -> None:
    """a4 has no runnable reference here (no `sam2` package / checkpoint): the golden logits come from transformers' Sam2Model
    (an independent restatement of the same architecture) with the seeded weights of oracle/sam2_hf.py."""
    from oracle import sam2_hf
    from tests.cases import sam2_input_image

    model = sam2_hf.build_model(sam2_hf.sam2_state_dict(0))
    up, low = sam2_hf.predict_logits(model, sam2_input_image())
    np.savez_compressed(OUT / "sam2_hiera_t_lowres.npz", low=low.astype(np.float16), positives=np.int64((up > 0).sum()))
    print("sam2 golden:", low.shape, float(low.std()), int((up > 0).sum()))


def make_sam2_large_golden() -> None:
    """BASELINE.json configs[2] (Hiera-L): low-res logits (fp16) and the packed 1024 x 1024 mask of the same restatement, weights
    seed 1 as in tests/test_gpu_sam2.py::test_hiera_large_matches_live_hf_model (bench.py's aux.c2 IoU check reads this file)."""
    from oracle import sam2_hf
    from tests.cases import sam2_input_image

    model = sam2_hf.build_model(sam2_hf.sam2_state_dict(1, "large"), "large")
    up, low = sam2_hf.predict_logits(model, sam2_input_image())
    np.savez_compressed(OUT / "sam2_hiera_l_mask.npz", low=low.astype(np.float16), mask_bits=np.packbits(up > 0),
                        positives=np.int64((up > 0).sum()))
    print("sam2 large golden:", low.shape, float(low.std()), int((up > 0).sum()))


if __name__ == "__main__" and "--sam2" in sys.argv:
    make_sam2_golden()
if __name__ == "__main__" and "--sam2-large" in sys.argv:
    make_sam2_large_golden()


def make_filter_golden() -> None:
    """filter_<case>.npz: rows the reference keeps with fast_mode=False (services/extraction.py:105-119; every candidate is
    read through IWSI.extract on the host, resized by cv2.resize when read != patch, and tested by utils/image.py)."""
    from tests.cases import FILTER_CASES

    by_name = {c["name"]: c for c in COORD_CASES}
    for fc in FILTER_CASES:
        case = by_name[fc["coords"]]
        spec = make_spec(case["width"], case["height"], case["seed"], mpp=case["mpp"])
        wsi = RefSyntheticWSI(spec)
        mask = build_mask(case, spec)
        with tempfile.TemporaryDirectory() as td:
            svc = PatchExtractionService(
                ExtractionConfig(patch_size=case["patch"], target_magnification=case["target_mag"], step_size=case["step"],
                                 tissue_threshold=case["tissue_thresh"], fast_mode=False,
                                 white_threshold=fc["white"], black_threshold=fc["black"]),
                OutputConfig(output_root=Path(td)),
            )
            tissue, holes = svc._prepare_contours(mask, wsi)
            rows = [e[:5] for e in svc._iter_patch_entries(wsi, tissue, holes, include_patch=False)]
        rows = np.asarray(rows, dtype=np.int32).reshape(-1, 5)
        cand = np.load(OUT / f"coords_{case['name']}.npz")["coords"]
        np.savez_compressed(OUT / f"filter_{fc['name']}.npz", coords=rows)
        print(f"filter_{fc['name']}: kept {rows.shape[0]} of {cand.shape[0]} candidates")


if __name__ == "__main__" and "--filter" in sys.argv:
    make_filter_golden()


def make_dinov2_golden() -> None:
    """dinov2_<name>.npz: features of transformers' Dinov2Model + BitImageProcessorFast (what the reference's DinoV2Encoder calls,
    models/patch/dinov2.py:49-62) with the seeded weights of oracle/dinov2_hf.py, plus the uint8 pixels the processor normalises."""
    from oracle import dinov2_hf
    from tests.cases import DINOV2_CASES, dinov2_patches

    for name, case in DINOV2_CASES.items():
        patches = dinov2_patches(name)
        sd = dinov2_hf.dinov2_state_dict(name, seed=case["weight_seed"])
        feats = dinov2_hf.extract_features(patches, sd, name, batch_size=4)
        x = dinov2_hf.preprocess(patches[-2:])
        mean, std = torch.tensor(dinov2_hf.MEAN).view(1, 3, 1, 1), torch.tensor(dinov2_hf.STD).view(1, 3, 1, 1)
        pix = torch.round((x * std + mean) * 255.0).to(torch.uint8).permute(0, 2, 3, 1).numpy()
        np.savez_compressed(OUT / f"{name}.npz", feats=feats.astype(np.float32), pixels=pix)
        print(name, feats.shape, float(np.abs(feats).mean()), pix.shape)


if __name__ == "__main__" and "--dinov2" in sys.argv:
    os.environ.setdefault("OMP_NUM_THREADS", "8")
    make_dinov2_golden()


def make_thumb_general_golden() -> None:
    """IWSI.get_thumbnail_at_power (core/wsi/iwsi.py:246-323), unmodified, on slides whose level size the factor does not divide
    and on a 60x slide (factor 48)."""
    from tests.cases import THUMB_GENERAL_CASES

    for case in THUMB_GENERAL_CASES:
        spec = make_spec(case["width"], case["height"], case["seed"], mpp=case["mpp"])
        thumb = np.asarray(RefSyntheticWSI(spec).get_thumbnail_at_power(power=1.25))
        np.savez_compressed(OUT / f"thumb_{case['name']}.npz", thumb=thumb)
        print(f"thumb_{case['name']}: {thumb.shape}")


if __name__ == "__main__" and "--thumb-general" in sys.argv:
    make_thumb_general_golden()


def make_vit_resize_golden() -> None:
    """--patch-size 224 / 512 with vit_b_16: the reference's own PatchFeatureExtractor (models/patch/base.py:76-107) on the torchvision
    preset, which resizes the PIL patch to 256 (Pillow BILINEAR) before the 224 crop.  Seeded weights, fp32 CPU."""
    import torch
    from torchvision import models

    from atlas_patch.models.patch.base import PatchFeatureExtractor
    from oracle.weights import vit_state_dict
    from tests.cases import VIT_RESIZE_CASES, vit_resize_patches

    sd = vit_state_dict("vit_b_16", seed=VIT_RESIZE_CASES["weight_seed"])
    model = models.vit_b_16(weights=None)
    model.heads = torch.nn.Identity()
    model.load_state_dict(sd, strict=True)
    ext = PatchFeatureExtractor(name="vit_b_16", model=model, embedding_dim=768, preprocess=models.ViT_B_16_Weights.IMAGENET1K_V1.transforms(),
                                device=torch.device("cpu"), dtype=torch.float32, num_workers=0)
    out = {}
    for P in VIT_RESIZE_CASES["sizes"]:
        patches = vit_resize_patches(P)
        out[f"feats_{P}"] = ext.extract_batch(patches, batch_size=4)
        print(f"vit_b_16 patch {P}: {out[f'feats_{P}'].shape}")
    np.savez_compressed(OUT / "vit_b_16_resize_feats.npz", **out)


if __name__ == "__main__" and "--vit-resize" in sys.argv:
    make_vit_resize_golden()


def make_hub_golden() -> None:
    """Features of the hub encoder families on three seeded 256 px patches, tiny seeded configs, fp32 CPU.  Where the reference's own
    extractor class can run with only its hub call replaced (tests/test_ref_hub_families.py: Midnight, Phikon, PhikonV2, HibouEncoder,
    PLIPExtractor, QuiltNet, PathOrchestraEncoder, HOptimus0, ProvGigaPathExtractor) the rows come from THAT class; openmidnight
    (torch.hub + checkpoint file) and the open_clip CLIP towers come from the oracle (oracle/hub_families.py), and `source_<name>` says which."""
    import importlib
    from unittest import mock

    import timm
    import torch

    from oracle import hub_families as hf
    from tests import test_ref_hub_families as T

    cpu = T.CPU
    patches = T._patches()
    out = {}

    def token_model(model):
        class _Token(torch.nn.Module):
            def forward(self, x):
                return model(pixel_values=x).last_hidden_state[:, 0]
        return _Token()

    via_class = {
        "midnight_test_tiny": ("midnight", "Midnight", {}, "hf"),
        "phikon_v1_test_tiny": ("phikon", "Phikon", {}, "hf+proc"),
        "phikon_v2_test_tiny": ("phikon", "PhikonV2", {}, "hf+proc"),
        "hibou_test_tiny": ("hibou", "HibouEncoder", dict(name="hibou_b", model_id="histai/hibou-B", embedding_dim=256), "hf+proc"),
        "plip_test_tiny": ("plip", "PLIPExtractor", {}, "clip"),
        "quilt_b_16_test_tiny": ("quilt", "QuiltNet", dict(name="quilt_b_16", model_id="wisdomik/QuiltNet-B-16"), "clip"),
        "pathorchestra_test_tiny": ("pathorchestra", "PathOrchestraEncoder", {}, "timm"),
        "h_optimus_test_tiny": ("hoptimus", "HOptimus0", {}, "timm"),
        "prov_gigapath_test_tiny": ("gigapath", "ProvGigaPathExtractor", {}, "timm"),
    }
    for name in ["midnight_test_tiny", "phikon_v1_test_tiny", "phikon_v2_test_tiny", "hibou_test_tiny", "openmidnight_test_tiny",
                 "h_optimus_test_tiny", "pathorchestra_test_tiny", "prov_gigapath_test_tiny", "plip_test_tiny", "quilt_b_16_test_tiny",
                 "clip_vit_b_32_test_tiny", "clip_vit_l_14_test_tiny"]:
        sd = hf.state_dict(name, seed=21)
        if name in via_class:
            module, cls, kw, how = via_class[name]
            klass = getattr(importlib.import_module(f"atlas_patch.models.patch.{module}"), cls)
            model = hf.build_model(name, sd)
            if how == "timm":
                with mock.patch.object(timm, "create_model", lambda *a, m=model, **k: token_model(m), create=True):
                    ext = klass(**kw, **cpu)
            else:
                with T._hub(T._Clip4x(model) if how == "clip" else model, None if how == "hf" else T._Processor(name)):
                    ext = klass(**kw, **cpu)
            feats, src = ext.extract_batch(patches, batch_size=2), f"reference class {module}.{cls}"
        else:
            feats, src = hf.extract_features(patches, sd, name), "oracle/hub_families.py (reference loader not runnable offline)"
        out[f"feats_{name}"] = feats.astype(np.float32)
        out[f"source_{name}"] = np.array(src)
        print(f"{name}: {feats.shape} <- {src}")
    np.savez_compressed(OUT / "hub_families.npz", **out)


if __name__ == "__main__" and "--hub" in sys.argv:
    make_hub_golden()
