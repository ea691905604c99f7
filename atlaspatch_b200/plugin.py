"""`--feature-plugin` entry point for the reference CLI (atlas_patch/cli.py:182-191,621-628).

    atlaspatch process slide.svs -o out --feature-plugin /path/to/atlaspatch_b200/plugin.py \
        --feature-extractors b200_vit_b_16

The hook signature is the reference's CustomRegistryHook (atlas_patch/models/patch/custom.py:92-146).  Weights come
from torchvision exactly as the reference's own `vit_b_16` builder resolves them (models/patch/base.py:126-180); the
forward runs on the sm_100a kernels.  `b200_dinov2_*` load transformers' Dinov2Model like the reference's DinoV2Encoder
(models/patch/dinov2.py); `b200_midnight`, `b200_phikon_v1`, `b200_phikon_v2` load the hub checkpoints like models/patch/midnight.py and
phikon.py.  A CUDA device is mandatory: there is no CPU fallback.
"""
from __future__ import annotations

import sys
from pathlib import Path

# The reference loads a plug-in BY FILE PATH (models/patch/custom.py:104-111: spec_from_file_location + exec_module), so
# nothing has put this package's parent directory on sys.path: do it here, before the first package import.
_ROOT = str(Path(__file__).resolve().parent.parent)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402

_TORCHVISION = {"vit_b_16": ("vit_b_16", "ViT_B_16_Weights"), "vit_l_16": ("vit_l_16", "ViT_L_16_Weights"),
                "vit_b_32": ("vit_b_32", "ViT_B_32_Weights"), "vit_l_32": ("vit_l_32", "ViT_L_32_Weights")}   # models/patch/vit.py:9-15


def _build(name: str, device) -> B200FeatureExtractor:
    import torch
    from torchvision import models

    if torch.device(device).type != "cuda":
        raise RuntimeError("atlaspatch_b200 encoders need a CUDA device (B200); no CPU fallback exists")
    ctor_name, enum_name = _TORCHVISION[name]
    weights_enum = getattr(models, enum_name)
    weights = getattr(weights_enum, "IMAGENET1K_V1", None) or weights_enum.DEFAULT
    model = getattr(models, ctor_name)(weights=weights)
    idx = torch.device(device).index or 0
    return B200FeatureExtractor(name, model.state_dict(), device=idx, registry_name=f"b200_{name}")


_DINOV2 = {"dinov2_small": "facebook/dinov2-small", "dinov2_base": "facebook/dinov2-base",
           "dinov2_large": "facebook/dinov2-large", "dinov2_giant": "facebook/dinov2-giant"}  # models/patch/dinov2.py:12-17


def _build_dinov2(name: str, device, patch_size: int | None) -> B200FeatureExtractor:
    """Weights exactly as the reference's DinoV2Encoder loads them (models/patch/dinov2.py:50); the processor's resize / crop /
    normalise (dinov2.py:49) runs in the CUDA preprocess kernel.  The engine is built for one input patch size: pass
    ATLASPATCH_B200_PATCH_SIZE (the --patch-size of the run; default 224)."""
    import os

    import torch
    from transformers import AutoModel

    if torch.device(device).type != "cuda":
        raise RuntimeError("atlaspatch_b200 encoders need a CUDA device (B200); no CPU fallback exists")
    model = AutoModel.from_pretrained(_DINOV2[name])
    idx = torch.device(device).index or 0
    patch = int(patch_size or os.environ.get("ATLASPATCH_B200_PATCH_SIZE", 224))
    return B200FeatureExtractor(name, model.state_dict(), input_patch=patch, device=idx, registry_name=f"b200_{name}")


# Hub encoders on the same kernels: name -> (hub id, loader), loaded exactly as the reference's classes do
_HUB = {"midnight": "kaiko-ai/midnight",        # models/patch/midnight.py:12,44  AutoModel (Dinov2Model ViT-g/14), [class || mean] -> 3072
        "phikon_v1": "owkin/phikon",            # models/patch/phikon.py:41-44    ViTModel(add_pooling_layer=False) -> 768
        "phikon_v2": "owkin/phikon-v2",         # models/patch/phikon.py:90-92    AutoModel (Dinov2Model ViT-L/16) -> 1024
        "hibou_b": "histai/hibou-B",            # models/patch/hibou.py:12-15,54  AutoModel(trust_remote_code): DINOv2 + 4 registers -> 768
        "hibou_l": "histai/hibou-L",            #                                 -> 1024
        "openmidnight": "SophontAI/OpenMidnight",   # models/patch/openmidnight.py:49-63  torch.hub dinov2_vitg14_reg + checkpoint -> 1536
        "plip": "vinid/plip",                   # models/patch/plip.py:34         CLIPModel (ViT-B/32), get_image_features -> 512
        "quilt_b_32": "wisdomik/QuiltNet-B-32",  # models/patch/quilt.py:12-16,56  CLIPModel (ViT-B/32 / ViT-B/16) -> 512
        "quilt_b_16": "wisdomik/QuiltNet-B-16",
        "h_optimus_0": "hf-hub:bioptimus/H-optimus-0",   # models/patch/hoptimus.py:53-58,98-132  timm ViT-g/14 reg4 -> 1536
        "h_optimus_1": "hf-hub:bioptimus/H-optimus-1",
        "pathorchestra": "hf-hub:AI4Pathology/PathOrchestra",   # models/patch/pathorchestra.py:38-43  timm ViT-L/16 -> 1024
        "prov_gigapath": "hf_hub:prov-gigapath/prov-gigapath",  # models/patch/gigapath.py:12,46       timm ViT-g/16 -> 1536
        "clip_vit_b_32": "ViT-B-32", "clip_vit_b_16": "ViT-B-16", "clip_vit_l_14": "ViT-L-14"}   # models/patch/clip.py:15-17  open_clip, "openai"


def _build_hub(name: str, device, patch_size: int | None) -> B200FeatureExtractor:
    """Weights as the reference loads them; each family's preprocess (torchvision Resize / CenterCrop / Normalize(0.5) for midnight,
    the checkpoints' fast image processors for phikon) and head run in the CUDA path (encoder.py: FAMILY_RECIPES)."""
    import os

    import torch

    if torch.device(device).type != "cuda":
        raise RuntimeError("atlaspatch_b200 encoders need a CUDA device (B200); no CPU fallback exists")
    if name == "phikon_v1":
        from transformers import ViTModel

        model = ViTModel.from_pretrained(_HUB[name], add_pooling_layer=False)
    elif name == "openmidnight":            # openmidnight.py:49-63, step for step; the state_dict stays in facebookresearch's key layout
        from huggingface_hub import hf_hub_download

        model = torch.hub.load("facebookresearch/dinov2", "dinov2_vitg14_reg", weights=None)
        checkpoint = torch.load(hf_hub_download(repo_id=_HUB[name], filename="teacher_checkpoint_load.pt"), map_location="cpu")
        model.pos_embed = torch.nn.parameter.Parameter(checkpoint["pos_embed"])
        model.load_state_dict(checkpoint)
    elif name.startswith("h_optimus"):      # hoptimus.py:53-58; the state_dict stays in timm's key layout
        import timm

        model = timm.create_model(_HUB[name], pretrained=True, init_values=1e-5, dynamic_img_size=False)
    elif name.startswith("clip_vit"):       # clip.py:36-40; the state_dict stays in open_clip's key layout
        import open_clip

        model, _, _ = open_clip.create_model_and_transforms(_HUB[name], pretrained="openai")
    elif name == "pathorchestra":           # pathorchestra.py:38-43
        import timm

        model = timm.create_model(_HUB[name], pretrained=True, init_values=1e-5, dynamic_img_size=True)
    elif name == "prov_gigapath":           # gigapath.py:46
        import timm

        model = timm.create_model(_HUB[name], pretrained=True)
    elif name == "plip" or name.startswith("quilt"):
        from transformers import CLIPModel

        model = CLIPModel.from_pretrained(_HUB[name])
    elif name.startswith("hibou"):
        from transformers import AutoModel

        model = AutoModel.from_pretrained(_HUB[name], trust_remote_code=True)
    else:
        from transformers import AutoModel

        model = AutoModel.from_pretrained(_HUB[name])
    idx = torch.device(device).index or 0
    patch = int(patch_size or os.environ.get("ATLASPATCH_B200_PATCH_SIZE", 224))
    return B200FeatureExtractor(name, model.state_dict(), input_patch=patch, device=idx, registry_name=f"b200_{name}")


# ---- timm-loaded ViTs whose evaluation transform comes from the checkpoint's data config ---------------------------------------------
# (models/patch/uni.py:30-45,80-110, hoptimus.py:141-161 H0-mini, lunit.py:46-60): nothing about them is restated here -- the
# architecture is read off the timm model object and the preprocess off timm's resolved data config at load time.
_TIMM = {"uni_v1": ("hf-hub:MahmoodLab/uni", 0), "uni_v2": ("hf-hub:MahmoodLab/UNI2-h", 0), "h0_mini": ("hf-hub:bioptimus/H0-mini", 1),
         "lunit_vit_small_patch16_dino": ("hf-hub:1aurent/vit_small_patch16_224.lunit_dino", 0)}   # name -> (hub id, pool)


def recipe_from_timm_data_config(cfg, *, pool: int = 0) -> tuple[dict, int]:
    """timm.data.create_transform(**cfg) for evaluation (transforms_factory.transforms_imagenet_eval) applied to a SQUARE PIL patch:
    Resize(floor(img / crop_pct), interpolation) -> CenterCrop(img) -> ToTensor -> Normalize(mean, std); for a square image the
    'center', 'squash' and 'border' crop modes coincide.  Returns (recipe for B200FeatureExtractor, image size).  PIL's BILINEAR /
    BICUBIC are the two filters the CUDA preprocess restates bit for bit (ap_vit_desc.preprocess 2 / 4)."""
    import math

    size = tuple(cfg["input_size"])
    img = int(size[-1])
    if int(size[-2]) != img:
        raise ValueError(f"timm data config: input_size {size} is not square")
    kind = {"bilinear": 2, "bicubic": 4}.get(cfg.get("interpolation", "bilinear"))
    if kind is None:
        raise ValueError(f"timm data config: interpolation '{cfg.get('interpolation')}' has no CUDA restatement (bilinear, bicubic)")
    if cfg.get("crop_border_pixels"):
        raise ValueError("timm data config: crop_border_pixels is not supported")
    scale = int(math.floor(img / (cfg.get("crop_pct") or 0.875)))         # timm's DEFAULT_CROP_PCT
    if scale < img:
        raise ValueError(f"timm data config: crop_pct {cfg.get('crop_pct')} > 1 pads instead of cropping; not supported")
    return dict(preprocess=kind, resize_to=scale, mean=tuple(float(m) for m in cfg["mean"]), std=tuple(float(v) for v in cfg["std"]),
                pool=int(pool), ln_eps=1e-6, default_patch=img), img


def arch_from_timm_vit(model) -> tuple:
    """(patch, layers, heads, hidden, mlp, swiglu, registers) of a timm VisionTransformer, read off the module tree."""
    blk = model.blocks[0]
    patch = model.patch_embed.patch_size
    patch = int(patch[0] if isinstance(patch, (tuple, list)) else patch)
    fc1_out, fc2_in = int(blk.mlp.fc1.out_features), int(blk.mlp.fc2.in_features)
    return (patch, len(model.blocks), int(blk.attn.num_heads), int(model.embed_dim), fc2_in, fc1_out == 2 * fc2_in,
            int(getattr(model, "num_reg_tokens", 0) or 0))


def _build_timm(name: str, device, patch_size: int | None) -> B200FeatureExtractor:
    import os

    import torch

    if torch.device(device).type != "cuda":
        raise RuntimeError("atlaspatch_b200 encoders need a CUDA device (B200); no CPU fallback exists")
    import timm

    hub_id, pool = _TIMM[name]
    if name == "uni_v1":                    # uni.py:30-36
        model = timm.create_model(hub_id, pretrained=True, init_values=1e-5, dynamic_img_size=True, num_classes=0)
    elif name == "uni_v2":                  # uni.py:80-99
        model = timm.create_model(hub_id, pretrained=True, img_size=224, patch_size=14, depth=24, num_heads=24, init_values=1e-5, embed_dim=1536,
                                  mlp_ratio=2.66667 * 2, num_classes=0, no_embed_class=True, mlp_layer=timm.layers.SwiGLUPacked,
                                  act_layer=torch.nn.SiLU, reg_tokens=8, dynamic_img_size=True)
    elif name == "h0_mini":                 # hoptimus.py:141-146
        model = timm.create_model(hub_id, pretrained=True, mlp_layer=timm.layers.SwiGLUPacked, act_layer=torch.nn.SiLU)
    else:                                   # lunit.py:46-49
        model = timm.create_model(hub_id, pretrained=True)
    cfg = timm.data.resolve_data_config(model.pretrained_cfg, model=model)      # uni.py:42, hoptimus.py:155, lunit.py:58
    recipe, image_size = recipe_from_timm_data_config(cfg, pool=pool)
    idx = torch.device(device).index or 0
    patch = int(patch_size or os.environ.get("ATLASPATCH_B200_PATCH_SIZE", image_size))
    return B200FeatureExtractor(name, model.state_dict(), input_patch=patch, image_size=image_size, device=idx, registry_name=f"b200_{name}",
                                arch=arch_from_timm_vit(model), recipe=recipe)


def resolve_feature_dtype(device, precision: str):
    """services/feature_embedding.py:28-39: the reference's dtype policy (float16 is downgraded to float32 on CPU devices)."""
    import torch

    dtype = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}.get(precision, torch.float32)
    if torch.device(device).type == "cpu" and dtype == torch.float16:
        dtype = torch.float32
    return dtype


def register_feature_extractors(registry, device, dtype, num_workers) -> None:
    """The reference's CustomRegistryHook (models/patch/custom.py:92-103).  `dtype` (cli.py --feature-precision through
    resolve_feature_dtype) selects the precision the REFERENCE runs its torch model in; the features it returns are float32 numpy
    either way (models/patch/base.py:105-106).  The B200 engine has one arithmetic (fp16 operands, fp32 accumulate / residual /
    LayerNorm / softmax) that meets the fp32 tolerance, so every dtype maps onto it and the returned array is float32 as well;
    `num_workers` configures the reference's DataLoader, which this path does not have."""
    for name in _TORCHVISION:
        registry.register(f"b200_{name}", lambda n=name: _build(n, device))
    for name in _DINOV2:
        registry.register(f"b200_{name}", lambda n=name: _build_dinov2(n, device, None))
    for name in _HUB:
        registry.register(f"b200_{name}", lambda n=name: _build_hub(n, device, None))
    for name in _TIMM:
        registry.register(f"b200_{name}", lambda n=name: _build_timm(n, device, None))
