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


def make_config(variant: str = "tiny"):
    """'tiny' = the model the reference ships (configs/sam2.1_hiera_t.yaml); 'large' = BASELINE.json configs[2] (Hiera-L)."""
    from transformers import Sam2Config, Sam2HieraDetConfig, Sam2VisionConfig

    if variant == "tiny":
        cfg = Sam2Config()
    elif variant == "large":
        bb = Sam2HieraDetConfig(hidden_size=144, num_attention_heads=2, blocks_per_stage=[2, 6, 36, 4],
                                embed_dim_per_stage=[144, 288, 576, 1152], num_attention_heads_per_stage=[2, 4, 8, 16],
                                window_size_per_stage=[8, 4, 16, 8], global_attention_blocks=[23, 33, 43])
        cfg = Sam2Config(vision_config=Sam2VisionConfig(backbone_config=bb, backbone_channel_list=[1152, 576, 288, 144]))
    else:
        raise ValueError(variant)
    cfg.mask_decoder_config.dynamic_multimask_via_stability = False
    return cfg


def sam2_state_dict(seed: int = 0, variant: str = "tiny") -> dict[str, torch.Tensor]:
    """Seeded random parameters in transformers' Sam2Model naming."""
    from transformers import Sam2Model

    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in Sam2Model(make_config(variant)).state_dict().items()}
    rng = np.random.default_rng(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("positional_embedding"):                     # gaussian Fourier matrices (scale 1)
            a = rng.standard_normal(shp)
        elif "layer_norm" in k or ".norm" in k:
            a = 1.0 + 0.1 * rng.standard_normal(shp) if k.endswith("weight") else 0.05 * rng.standard_normal(shp)
        elif k.endswith("bias"):
            a = 0.02 * rng.standard_normal(shp)
        elif len(shp) >= 2 and k.endswith("weight") and not any(s in k for s in ("token", "embed")):
            if "upscale_conv" in k:                                # ConvTranspose2d weight is (in, out, kh, kw)
                fan_in = shp[0]
            else:
                fan_in = int(np.prod(shp[1:]))
            a = rng.standard_normal(shp) / math.sqrt(fan_in)
        elif "patch_embed.projection.weight" in k:
            a = rng.standard_normal(shp) / math.sqrt(int(np.prod(shp[1:])))
        else:                                                      # tokens, embeddings, pos_embed, no_memory_embedding
            a = 0.5 * rng.standard_normal(shp)
        sd[k] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


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
