"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the DINOv2 encoders of BASELINE.json configs[3..4] (a12 + a13).

The reference builds these extractors from the HuggingFace hub (atlas_patch/models/patch/dinov2.py:12-17,49-50):

    processor = AutoImageProcessor.from_pretrained("facebook/dinov2-<size>", use_fast=True)    # dinov2.py:49
    model     = AutoModel.from_pretrained("facebook/dinov2-<size>")                             # dinov2.py:50
    feature   = model(pixel_values=x).last_hidden_state[:, 0, :]                                # dinov2.py:60-62

There is no network here, so the oracle constructs the same two classes the hub files resolve to, with the published
contents of those files and seeded weights:

* `preprocessor_config.json` of facebook/dinov2-{small,base,large,giant} (identical for all four): BitImageProcessor,
  resize shortest_edge 256 with resample 3 (bicubic), center crop 224, rescale 1/255, ImageNet mean/std.  With use_fast=True
  this is transformers' BitImageProcessorFast = torch uint8 bicubic *antialias* interpolation (see oracle/resize_aa.py for
  the integer restatement, pinned bit-exactly against torch) -> crop -> fused rescale/normalise.
* `config.json`: Dinov2Config(hidden 1024 / 1536, layers 24 / 40, heads 16 / 24, patch 14, image_size 518, layerscale,
  use_swiglu_ffn for giant, qkv_bias, layer_norm_eps 1e-6).  image_size 518 means the 37 x 37 position grid is bicubically
  interpolated to 16 x 16 for 224 px inputs (modeling_dinov2.py: interpolate_pos_encoding) -- the oracle keeps that.

Parity of the CUDA path is therefore pinned against transformers' implementation run in this container (golden features in
tests/golden/dinov2_*.npz, made by tests/golden/make_golden.py --dinov2), which is the code the reference itself calls.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from atlaspatch_b200.weights import DINOV2_SPECS, PATCH, dinov2_state_dict, swiglu_hidden  # noqa: E402,F401  (seeded input data)

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def make_config(name: str, image_size: int = 518):
    from transformers import Dinov2Config

    layers, heads, d, swiglu = DINOV2_SPECS[name]
    return Dinov2Config(hidden_size=d, num_hidden_layers=layers, num_attention_heads=heads, mlp_ratio=4, patch_size=PATCH,
                        image_size=image_size, use_swiglu_ffn=swiglu, layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0)


def build_model(name: str, sd: dict[str, torch.Tensor], image_size: int = 518):
    from transformers import Dinov2Model

    model = Dinov2Model(make_config(name, image_size)).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return model


def make_processor():
    """The fast (torchvision-backend) BitImageProcessor with the contents of facebook/dinov2-*'s preprocessor_config.json;
    transformers >= 5 names it BitImageProcessor, earlier releases BitImageProcessorFast (same bytes, checked here)."""
    import transformers

    cls = transformers.BitImageProcessor if int(transformers.__version__.split(".")[0]) >= 5 else transformers.BitImageProcessorFast
    return cls(do_resize=True, size={"shortest_edge": 256}, resample=3, do_center_crop=True,
                                 crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                                 do_normalize=True, image_mean=list(MEAN), image_std=list(STD), do_convert_rgb=True)


def preprocess(patches: Sequence[np.ndarray], processor=None) -> torch.Tensor:
    """The reference's per-patch preprocess (dinov2.py:20-25 via base.py:42-45): PIL image -> processor -> (3, 224, 224)."""
    from PIL import Image

    processor = processor or make_processor()
    out = [processor(images=Image.fromarray(np.asarray(p)), return_tensors="pt")["pixel_values"].squeeze(0) for p in patches]
    return torch.stack(out) if out else torch.empty(0, 3, 224, 224)


@torch.inference_mode()
def extract_features(patches: Sequence[np.ndarray], sd: dict[str, torch.Tensor], name: str, *, image_size: int = 518,
                     batch_size: int = 8, model=None) -> np.ndarray:
    d = DINOV2_SPECS[name][2]
    if len(patches) == 0:
        return np.empty((0, d), dtype=np.float32)
    model = model or build_model(name, sd, image_size)
    proc = make_processor()
    outs = []
    for i in range(0, len(patches), batch_size):
        x = preprocess(patches[i:i + batch_size], proc)
        outs.append(model(pixel_values=x).last_hidden_state[:, 0, :])
    return torch.cat(outs).to(torch.float32).numpy()
