"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the hub encoders of SURVEY.md section 8(f) rank 4 that run on the DINOv2 / ViT kernels:

* `midnight`  (atlas_patch/models/patch/midnight.py:12-25,44,55-61): `AutoModel.from_pretrained("kaiko-ai/midnight")` = a
  transformers Dinov2Model (ViT-g/14, SwiGLU), torchvision `Resize(224) -> CenterCrop(224) -> ToTensor -> Normalize(0.5, 0.5)` on
  the PIL patch, feature = cat(last_hidden_state[:, 0], last_hidden_state[:, 1:].mean(1))  -> 2 x 1536.
* `phikon_v2` (phikon.py:90-93,103-105): `AutoModel.from_pretrained("owkin/phikon-v2")` = Dinov2Model (ViT-L/16, 224 px) with the
  repo's BitImageProcessor (fast: shortest_edge 224 bicubic, crop 224, ImageNet mean / std), feature = last_hidden_state[:, 0].
* `phikon_v1` (phikon.py:41-46,54-56): `ViTModel.from_pretrained("owkin/phikon", add_pooling_layer=False)` with the repo's
  ViTImageProcessor (fast: 224 x 224, resample 2 = bilinear, ImageNet mean / std), feature = last_hidden_state[:, 0].

* `hibou_b`, `hibou_l` (hibou.py:12-15,51-54,67-69): `AutoModel.from_pretrained("histai/hibou-*", trust_remote_code=True)` = DINOv2
  with 4 register tokens (the architecture transformers ships as Dinov2WithRegistersModel), the repo's BitImageProcessor (fast:
  shortest_edge 224 bicubic, crop 224, mean (0.7068, 0.5755, 0.722), std (0.195, 0.2316, 0.1816)), feature = pooler_output.
* `openmidnight` (openmidnight.py:17-30,49-63): facebookresearch `dinov2_vitg14_reg` (ViT-g/14, 4 registers, SwiGLU) with the
  checkpoint's 224 px position grid, torchvision `Resize((224, 224)) -> ToTensor -> Normalize(ImageNet)`, feature = model(x) = the
  class token after the final LayerNorm.  facebookresearch/dinov2 is not importable offline; Dinov2WithRegistersModel is the same
  architecture under transformers' key names (atlaspatch_b200/dinov2.py: fb_to_hf_dinov2_names maps one layout onto the other).
* `h_optimus_0`, `h_optimus_1` (hoptimus.py:14-31,53-58): timm `vit_giant_patch14_reg4_dinov2` (ViT-g/14, 4 registers, packed SwiGLU,
  no_embed_class) with torchvision `Resize((224, 224)) -> ToTensor -> Normalize(own mean / std)`, feature = the class token.  timm is
  not in this image; the architecture is Dinov2WithRegistersModel's, and fb_to_hf_dinov2_names accepts timm's key names.
* `pathorchestra` (pathorchestra.py:38-58) and `prov_gigapath` (gigapath.py:17-26,46): timm ViT-L/16 / ViT-g/16 (packed SwiGLU) with
  LayerScale = Dinov2Model's architecture at patch 16; torchvision `Resize(224)` resp. `Resize(256, BICUBIC) -> CenterCrop(224)` on the
  PIL patch, ImageNet mean / std, feature = the class token (timm global_pool "token").
* `plip`, `quilt_b_32`, `quilt_b_16` (plip.py:34-35,56; quilt.py:12-16,56-60): transformers `CLIPModel` + `CLIPProcessor` (fast image
  processor: shortest_edge 224 bicubic, crop 224, OpenAI CLIP mean / std), feature = `get_image_features(pixel_values=x)` =
  visual_projection(post_layernorm(class token)) -> 512.  transformers 4.x returns that tensor (what the reference's forward_fn
  hands on); transformers >= 5 returns the vision ModelOutput with the projected features in `pooler_output` -- same numbers.
* `clip_vit_b_32`, `clip_vit_b_16`, `clip_vit_l_14` (clip.py:15-17,36-40,58): OpenAI CLIP through open_clip; the eval transform
  open_clip builds is torchvision `Resize(224, BICUBIC) -> CenterCrop(224) -> ToTensor -> Normalize(CLIP mean / std)` on the PIL patch,
  feature = `encode_image` = ln_post(class token) @ visual.proj.  open_clip is not in this image; transformers' CLIPModel is the same
  network (the OpenAI checkpoints converted), and atlaspatch_b200/dinov2.py: openclip_to_hf_clip_names maps open_clip's keys onto it.

There is no network: the oracle builds the classes those hub files resolve to, from the published contents of their config.json /
preprocessor_config.json (restated from memory of the public repos -- parity of the *settings* is unpinned; parity of the
*arithmetic* is pinned against transformers / torchvision run in this container on the same settings), with seeded weights.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from atlaspatch_b200.weights import (DINOV2_PATCH, DINOV2_REGISTERS, DINOV2_SPECS, HF_CLIP_SPECS, HF_VIT_SPECS,  # noqa: F401
                                     dinov2_state_dict, hf_clip_state_dict, hf_vit_state_dict)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
HIBOU_MEAN, HIBOU_STD = (0.7068, 0.5755, 0.722), (0.195, 0.2316, 0.1816)
HOPT_MEAN, HOPT_STD = (0.707223, 0.578729, 0.703617), (0.211883, 0.230117, 0.177517)
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def _family(name: str) -> str:
    for fam in ("midnight", "phikon_v2", "phikon_v1", "hibou", "openmidnight", "h_optimus", "pathorchestra", "prov_gigapath", "plip", "quilt", "clip_vit"):
        if name.startswith(fam):
            return fam
    raise KeyError(name)


def state_dict(name: str, seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded weights in the key layout of the model class the reference loads (224 px position grids)."""
    if _family(name) == "phikon_v1":
        return hf_vit_state_dict(name, seed=seed, image_size=224)
    if _family(name) in ("plip", "quilt", "clip_vit"):
        return hf_clip_state_dict(name, seed=seed, image_size=224)
    return dinov2_state_dict(name, seed=seed, image_size=224)


def build_model(name: str, sd: dict[str, torch.Tensor]):
    fam = _family(name)
    if fam == "phikon_v1":
        from transformers import ViTConfig, ViTModel

        patch, layers, heads, d, mlp = HF_VIT_SPECS[name]
        cfg = ViTConfig(hidden_size=d, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=mlp, patch_size=patch,
                        image_size=224, hidden_act="gelu", layer_norm_eps=1e-12, qkv_bias=True)
        model = ViTModel(cfg, add_pooling_layer=False).eval()
    elif fam in ("plip", "quilt", "clip_vit"):
        from transformers import CLIPConfig, CLIPModel, CLIPTextConfig, CLIPVisionConfig

        patch, layers, heads, d, mlp, proj = HF_CLIP_SPECS[name]
        vc = CLIPVisionConfig(hidden_size=d, intermediate_size=mlp, num_hidden_layers=layers, num_attention_heads=heads, image_size=224,
                              patch_size=patch, hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=proj)
        tc = CLIPTextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, vocab_size=100,
                            max_position_embeddings=16, projection_dim=proj)         # the text tower is not on the path: kept minimal
        model = CLIPModel(CLIPConfig(text_config=tc.to_dict(), vision_config=vc.to_dict(), projection_dim=proj)).eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(("text_model.", "text_projection.", "logit_scale")) for k in missing), (missing, unexpected)
        return model
    else:
        from transformers import Dinov2Config, Dinov2Model, Dinov2WithRegistersConfig, Dinov2WithRegistersModel

        layers, heads, d, swiglu = DINOV2_SPECS[name]
        kw = dict(hidden_size=d, num_hidden_layers=layers, num_attention_heads=heads, mlp_ratio=4, patch_size=DINOV2_PATCH.get(name, 14),
                  image_size=224, use_swiglu_ffn=swiglu, layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0)
        if DINOV2_REGISTERS.get(name, 0):
            model = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(num_register_tokens=DINOV2_REGISTERS[name], **kw)).eval()
        else:
            model = Dinov2Model(Dinov2Config(**kw)).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return model


def make_preprocess(name: str):
    """PIL image -> (3, 224, 224) float32, built from the same classes the reference builds."""
    fam = _family(name)
    if fam == "midnight":                                   # midnight.py:15-25
        from torchvision import transforms

        return transforms.Compose([transforms.Resize(224), transforms.CenterCrop(224), transforms.ToTensor(),
                                   transforms.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])
    if fam == "clip_vit":                                   # open_clip.transform.image_transform(is_train=False) for "openai" weights
        from torchvision import transforms

        return transforms.Compose([transforms.Resize(224, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(224),
                                   transforms.ToTensor(), transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])
    if fam in ("pathorchestra", "prov_gigapath"):           # pathorchestra.py:52-58, gigapath.py:17-26
        from torchvision import transforms

        head = ([transforms.Resize(224)] if fam == "pathorchestra" else
                [transforms.Resize(256, interpolation=transforms.InterpolationMode.BICUBIC), transforms.CenterCrop(224)])
        return transforms.Compose(head + [transforms.ToTensor(), transforms.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD)])
    if fam in ("openmidnight", "h_optimus"):                # openmidnight.py:17-30, hoptimus.py:14-31
        from torchvision import transforms

        mean, std = (IMAGENET_MEAN, IMAGENET_STD) if fam == "openmidnight" else (HOPT_MEAN, HOPT_STD)
        return transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(), transforms.Normalize(mean=list(mean), std=list(std))])
    import transformers

    if fam in ("plip", "quilt"):
        proc = transformers.CLIPImageProcessor(do_resize=True, size={"shortest_edge": 224}, resample=3, do_center_crop=True,
                                               crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                                               do_normalize=True, image_mean=list(CLIP_MEAN), image_std=list(CLIP_STD), do_convert_rgb=True)
    elif fam == "hibou":
        proc = transformers.BitImageProcessor(do_resize=True, size={"shortest_edge": 224}, resample=3, do_center_crop=True,
                                              crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                                              do_normalize=True, image_mean=list(HIBOU_MEAN), image_std=list(HIBOU_STD), do_convert_rgb=True)
    elif fam == "phikon_v2":
        proc = transformers.BitImageProcessor(do_resize=True, size={"shortest_edge": 224}, resample=3, do_center_crop=True,
                                              crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                                              do_normalize=True, image_mean=list(IMAGENET_MEAN), image_std=list(IMAGENET_STD), do_convert_rgb=True)
    else:
        proc = transformers.ViTImageProcessor(do_resize=True, size={"height": 224, "width": 224}, resample=2, do_rescale=True,
                                              rescale_factor=1 / 255, do_normalize=True, image_mean=list(IMAGENET_MEAN),
                                              image_std=list(IMAGENET_STD))
    return lambda pil: proc(images=pil, return_tensors="pt")["pixel_values"].squeeze(0)       # phikon.py:15-21


def pixels(name: str, patch: np.ndarray) -> np.ndarray:
    """The uint8 (224, 224, 3) pixels the family's preprocess normalises, by the integer restatements of oracle/resize_aa.py."""
    from oracle import resize_aa as ra

    fam = _family(name)
    if fam == "midnight":
        return ra.vit_preset_pixels(patch, resize_to=224, crop=224)
    if fam in ("openmidnight", "h_optimus", "pathorchestra"):
        return ra.resize_pil_bilinear(patch, 224, 224)
    if fam == "prov_gigapath":
        return ra.vit_preset_pixels(patch, resize_to=256, crop=224, mode="bicubic")
    if fam == "clip_vit":
        return ra.vit_preset_pixels(patch, resize_to=224, crop=224, mode="bicubic")
    if fam in ("phikon_v2", "hibou", "plip", "quilt"):
        return ra.dinov2_pixels(patch, resize_to=224, crop=224) if patch.shape[0] != 224 else patch
    return ra.hf_vit_pixels(patch, 224)


@torch.inference_mode()
def extract_features(patches: Sequence[np.ndarray], sd: dict[str, torch.Tensor], name: str, batch_size: int = 8) -> np.ndarray:
    """The reference's extract_batch for these families (base.py:76-107 with each class's forward_fn), fp32 on the CPU."""
    from PIL import Image

    model = build_model(name, sd)
    pre = make_preprocess(name)
    cls_mean = _family(name) == "midnight"
    outs = []
    for i in range(0, len(patches), batch_size):
        x = torch.stack([pre(Image.fromarray(np.asarray(p))) for p in patches[i:i + batch_size]])
        if _family(name) in ("plip", "quilt", "clip_vit"):
            out = model.get_image_features(pixel_values=x)                     # plip.py:56, quilt.py:60
            outs.append(out if isinstance(out, torch.Tensor) else out.pooler_output)
            continue
        out = model(pixel_values=x)
        h = out.last_hidden_state
        if _family(name) == "hibou":
            outs.append(out.pooler_output)                                     # hibou.py:67-69
        else:
            outs.append(torch.cat([h[:, 0], h[:, 1:].mean(1)], dim=-1) if cls_mean else h[:, 0])
    d = getattr(model.config, "projection_dim", None) or model.config.hidden_size * (2 if cls_mean else 1)
    return torch.cat(outs).to(torch.float32).numpy() if outs else np.empty((0, d), dtype=np.float32)
