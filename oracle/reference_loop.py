"""CPU ORACLE -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py `cpu_baseline` and `--impl reference`).

Port of the reference's CPU embedding pipeline for one slide, used as the timed CPU baseline on the GPU box
(where /root/reference does not exist).  It keeps the reference's structure, including its inefficiencies,
because that is what a user of the reference runs:

* per-row patch read + optional cv2.resize            services/feature_embedding.py:81-96
* a NEW DataLoader for every `feature_batch` patches    services/storage.py:279-294 -> models/patch/base.py:83-91
* PatchDataset: np -> PIL -> torchvision preset         models/patch/base.py:32-45,170
* torchvision vit_b_16 (heads -> Identity), fp32, CPU   models/patch/vit.py:9-38, base.py:66,95-106
* feats.cpu().float().numpy() per batch                 models/patch/base.py:105-106

`kind` for bench.py: "port".  Validated against the reference's own PatchFeatureExtractor output in
tests/golden/vit_b_16_feats.npz (tests/test_oracle_vit.py::test_reference_loop_matches_golden).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
from PIL import Image
from torch.utils.data import DataLoader, Dataset


class _PatchDataset(Dataset):
    def __init__(self, patches, transform):
        self._patches, self.transform = patches, transform

    def __len__(self):
        return len(self._patches)

    def __getitem__(self, idx):
        return self.transform(Image.fromarray(self._patches[idx]))


def build_vit_b_16(state_dict):
    from torchvision import models

    model = models.vit_b_16(weights=None)
    model.heads = torch.nn.Identity()
    model.load_state_dict(state_dict, strict=True)
    preprocess = models.ViT_B_16_Weights.IMAGENET1K_V1.transforms()
    return model.to(device="cpu", dtype=torch.float32).eval(), preprocess


@torch.inference_mode()
def extract_batch(model, preprocess, patches: Sequence[np.ndarray], *, batch_size: int, num_workers: int) -> np.ndarray:
    if not patches:
        return np.empty((0, 768), dtype=np.float32)
    bs = min(len(patches), batch_size or len(patches))
    loader = DataLoader(_PatchDataset(patches, preprocess), batch_size=bs, shuffle=False, num_workers=num_workers,
                        pin_memory=False)
    outs = []
    for batch in loader:
        out = model(batch.to(device="cpu", dtype=torch.float32))
        if out.ndim > 2:
            out = torch.flatten(out, start_dim=1)
        outs.append(out.detach())
    feats = torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]
    return feats.cpu().to(dtype=torch.float32).numpy()


def embed_slide_rows(model, preprocess, read_patch: Callable[[int, int, int, int], np.ndarray], coords: np.ndarray, *,
                     patch_size: int = 256, feature_batch: int = 32, num_workers: int = 4) -> np.ndarray:
    """The append_features loop (services/storage.py:250-337): rows of `coords` in order, batches of feature_batch."""
    import cv2

    feats, batch = [], []
    for x, y, rw, rh, _lv in coords.tolist():
        patch = read_patch(int(x), int(y), int(rw), int(rh))
        if patch.shape[0] != patch_size or patch.shape[1] != patch_size:
            patch = cv2.resize(patch, (patch_size, patch_size))
        batch.append(patch)
        if len(batch) == feature_batch:
            feats.append(extract_batch(model, preprocess, batch, batch_size=feature_batch, num_workers=num_workers))
            batch = []
    if batch:
        feats.append(extract_batch(model, preprocess, batch, batch_size=feature_batch, num_workers=num_workers))
    return np.concatenate(feats, axis=0) if feats else np.empty((0, 768), dtype=np.float32)
