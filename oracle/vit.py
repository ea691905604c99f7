"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

a12 + a13: encoder preprocessing and forward of the reference's `vit_*` extractors, in plain fp32
torch ops on the CPU.  The reference builds torchvision's VisionTransformer with `heads -> Identity`
and its `weights.transforms()` preset (atlas_patch/models/patch/vit.py:9-38,
models/patch/base.py:148-180); the arithmetic lives in torchvision 0.26.0
([tv]models/vision_transformer.py:86-117,154-157,268-305; [tv]transforms/_presets.py:39-65), restated
here.  Pinned by tests/golden/vit_b_16_feats.npz, which was produced by the reference's own
PatchFeatureExtractor.extract_batch on the torchvision model (tests/golden/make_golden.py), and by
tests/test_oracle_vit.py against torchvision itself.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle.weights import VIT_SPECS

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess(patches: Sequence[np.ndarray], *, crop: int = 224, resize: int = 256) -> torch.Tensor:
    """ImageClassification(crop_size=224, resize_size=256, bilinear) preset on the PIL image the reference's PatchDataset builds
    (models/patch/base.py:42-45; [tv]_presets.py:39-65).

    For (256,256,3) patches the resize is a no-op ([tv]transforms/functional.py:470-471), so this is centre-crop [16:240) ->
    /255 -> (x-mean)/std.  Other sizes are resized by Pillow (Image.resize(BILINEAR): antialiased triangle filter, 8-bit fixed
    point) -- NOT by torch's tensor resize, whose int16 coefficients round differently.
    """
    from PIL import Image

    out = []
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    for p in patches:
        a = np.ascontiguousarray(p)
        h, w = a.shape[:2]
        if min(h, w) != resize:
            if h <= w:
                nh, nw = resize, int(resize * w / h)
            else:
                nh, nw = int(resize * h / w), resize
            a = np.asarray(Image.fromarray(a).resize((nw, nh), Image.Resampling.BILINEAR))
            h, w = nh, nw
        t = torch.from_numpy(np.ascontiguousarray(a)).permute(2, 0, 1)  # uint8 CHW
        top, left = int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))
        t = t[:, top:top + crop, left:left + crop]
        t = t.to(torch.float32) / 255.0
        out.append((t - mean) / std)
    return torch.stack(out) if out else torch.empty(0, 3, crop, crop)


@torch.inference_mode()
def forward(x: torch.Tensor, sd: dict[str, torch.Tensor], name: str = "vit_b_16",
            return_hidden: bool = False):
    """VisionTransformer.forward with heads=Identity -> (B, D) CLS features."""
    patch, layers, heads, d, _mlp = VIT_SPECS[name]
    B = x.shape[0]
    x = F.conv2d(x, sd["conv_proj.weight"], sd["conv_proj.bias"], stride=patch)   # (B, d, g, g)
    x = x.reshape(B, d, -1).permute(0, 2, 1)                                      # (B, g*g, d)
    x = torch.cat([sd["class_token"].expand(B, -1, -1), x], dim=1)
    x = x + sd["encoder.pos_embedding"]
    hd = d // heads
    hidden = []
    for i in range(layers):
        p = f"encoder.layers.encoder_layer_{i}."
        y = F.layer_norm(x, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps=1e-6)
        qkv = y @ sd[p + "self_attention.in_proj_weight"].T + sd[p + "self_attention.in_proj_bias"]
        q, k, v = qkv.split(d, dim=-1)
        S = x.shape[1]
        q = q.view(B, S, heads, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
        k = k.view(B, S, heads, hd).transpose(1, 2)
        v = v.view(B, S, heads, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, d)
        x = x + (a @ sd[p + "self_attention.out_proj.weight"].T + sd[p + "self_attention.out_proj.bias"])
        y = F.layer_norm(x, (d,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps=1e-6)
        y = F.gelu(y @ sd[p + "mlp.0.weight"].T + sd[p + "mlp.0.bias"])
        x = x + (y @ sd[p + "mlp.3.weight"].T + sd[p + "mlp.3.bias"])
        if return_hidden:
            hidden.append(x.clone())
    x = F.layer_norm(x, (d,), sd["encoder.ln.weight"], sd["encoder.ln.bias"], eps=1e-6)
    feats = x[:, 0]
    return (feats, hidden) if return_hidden else feats


def extract_features(patches: Sequence[np.ndarray], sd: dict[str, torch.Tensor], name: str = "vit_b_16",
                     batch_size: int = 32) -> np.ndarray:
    """FeatureExtractor.extract_batch contract (models/patch/base.py:76-107): (n, D) float32."""
    _, _, _, d, _ = VIT_SPECS[name]
    if len(patches) == 0:
        return np.empty((0, d), dtype=np.float32)
    outs = []
    for i in range(0, len(patches), batch_size):
        outs.append(forward(preprocess(patches[i:i + batch_size]), sd, name))
    return torch.cat(outs).to(torch.float32).numpy()
