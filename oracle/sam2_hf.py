"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

a4: SAM2 box-prompted tissue mask.  The reference calls the un-vendored `sam2` package
(atlas_patch/services/segmentation.py:60-71,120-140: SAM2ImagePredictor.set_image + predict(box=[0,0,w,h],
multimask_output=False, return_logits=False)); neither `sam2` nor the AtlasPatch checkpoint exist offline, so **parity for this
row is unpinned by the reference**.  The independent restatement available in this image is `transformers.Sam2Model`
(transformers 5.5.0, models/sam2/modeling_sam2.py; defaults = the Hiera-T config the reference ships,
atlas_patch/configs/sam2.1_hiera_t.yaml), configured like the reference's YAML (dynamic_multimask_via_stability off) and
driven exactly as upstream's predictor drives the model: pixels/255 -> ImageNet normalise -> image encoder ->
box as two corner points (labels 2, 3) + one padding point -> mask decoder (mask token 0) -> bilinear x4 -> > threshold.
Weights are seeded random values keyed by HF parameter names (numpy PCG64, like oracle/weights.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


from atlaspatch_b200.weights import sam2_config as make_config, sam2_state_dict  # noqa: E402,F401  (seeded input data)


def build_model(sd, variant: str = "tiny"):
    from transformers import Sam2Model

    m = Sam2Model(make_config(variant))
    m.load_state_dict(sd, strict=True)
    return m.eval()


@torch.inference_mode()
def predict_logits(model, image_u8: np.ndarray, return_intermediates: bool = False):
    """image_u8: (1024, 1024, 3) uint8 -> (1024, 1024) float32 mask logits (box prompt = whole image, mask token 0)."""
    assert image_u8.shape == (1024, 1024, 3) and image_u8.dtype == np.uint8
    x = torch.from_numpy(image_u8).permute(2, 0, 1).float() / 255.0
    x = (x - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)
    pts = torch.tensor([[[[0.0, 0.0], [1024.0, 1024.0]]]])
    labels = torch.tensor([[[2, 3]]], dtype=torch.int32)
    out = model(pixel_values=x[None], input_points=pts, input_labels=labels, multimask_output=False,
                output_hidden_states=return_intermediates)
    low = out.pred_masks[0, 0, 0].float()                                                      # (256, 256)
    up = torch.nn.functional.interpolate(low[None, None], size=(1024, 1024), mode="bilinear", align_corners=False)[0, 0]
    if return_intermediates:
        return up.numpy(), low.numpy(), out
    return up.numpy(), low.numpy()
