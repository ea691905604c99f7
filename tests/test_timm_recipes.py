"""CPU: the load-time translation of timm-loaded ViTs (plugin.py: uni_v1, uni_v2, h0_mini, lunit_vit_small_patch16_dino) -- the data config ->
preprocess recipe, the module tree -> architecture tuple, and the timm key layout -> engine tensors -- without timm (absent in this image):
the transform timm builds is a torchvision Compose, restated here from transforms_factory.transforms_imagenet_eval, and the model side is
transformers' Dinov2(WithRegisters)Model, whose architecture these timm ViTs share."""
import math
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from atlaspatch_b200.plugin import arch_from_timm_vit, recipe_from_timm_data_config
from oracle import resize_aa
from tests.test_oracle_hub_families import _engine_forward, _patch, hf_to_fb_names

IMNET = dict(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))


def _timm_eval_transform(cfg):
    """timm.data.transforms_factory.transforms_imagenet_eval for a square input size and the 'center' crop mode."""
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode as IM

    img = cfg["input_size"][-1]
    scale = math.floor(img / (cfg.get("crop_pct") or 0.875))
    interp = {"bilinear": IM.BILINEAR, "bicubic": IM.BICUBIC}[cfg.get("interpolation", "bilinear")]
    return transforms.Compose([transforms.Resize(scale, interpolation=interp), transforms.CenterCrop(img), transforms.ToTensor(),
                               transforms.Normalize(mean=cfg["mean"], std=cfg["std"])])


@pytest.mark.parametrize("cfg", [dict(input_size=(3, 224, 224), interpolation="bicubic", crop_pct=0.875, crop_mode="center", **IMNET),
                                 dict(input_size=(3, 224, 224), interpolation="bilinear", crop_pct=1.0, crop_mode="center", **IMNET),
                                 dict(input_size=(3, 224, 224), interpolation="bicubic", crop_pct=0.95, crop_mode="squash",      # odd margin: 235 -> 224
                                      mean=(0.7032, 0.5361, 0.661), std=(0.2172, 0.2608, 0.2072)),
                                 dict(input_size=(3, 224, 224), interpolation="bicubic", crop_pct=None, **IMNET)])
@pytest.mark.parametrize("P", [224, 256, 300])
def test_recipe_reproduces_the_timm_transform(cfg, P):
    from PIL import Image

    recipe, img = recipe_from_timm_data_config(cfg)
    assert img == 224 and recipe["preprocess"] == (4 if cfg["interpolation"] == "bicubic" else 2)
    assert recipe["resize_to"] == math.floor(224 / (cfg.get("crop_pct") or 0.875))
    patch = _patch(P, seed=P)
    want = _timm_eval_transform(cfg)(Image.fromarray(patch)).permute(1, 2, 0).numpy()
    mode = "bicubic" if recipe["preprocess"] == 4 else "bilinear"
    pix = resize_aa.vit_preset_pixels(patch, resize_to=recipe["resize_to"], crop=224, mode=mode)      # what preprocess 2 / 4 implement
    got = (pix.astype(np.float32) / 255.0 - np.asarray(recipe["mean"], np.float32)) / np.asarray(recipe["std"], np.float32)
    assert np.abs(got - want).max() < 2e-6


def test_recipe_refuses_what_the_kernels_do_not_restate():
    base = dict(input_size=(3, 224, 224), interpolation="bicubic", crop_pct=0.9, **IMNET)
    for bad in (dict(interpolation="lanczos"), dict(crop_pct=1.1), dict(crop_border_pixels=8), dict(input_size=(3, 224, 256))):
        with pytest.raises(ValueError):
            recipe_from_timm_data_config({**base, **bad})
    assert recipe_from_timm_data_config(base, pool=1)[0]["pool"] == 1


def _fake_timm_vit(patch, layers, heads, d, fc1_out, fc2_in, regs):
    blk = NS(attn=NS(num_heads=heads), mlp=NS(fc1=NS(out_features=fc1_out), fc2=NS(in_features=fc2_in)))
    return NS(blocks=[blk] * layers, patch_embed=NS(patch_size=(patch, patch)), embed_dim=d, num_reg_tokens=regs)


def test_arch_is_read_off_the_module_tree():
    assert arch_from_timm_vit(_fake_timm_vit(16, 24, 16, 1024, 4096, 4096, 0)) == (16, 24, 16, 1024, 4096, False, 0)         # uni_v1
    assert arch_from_timm_vit(_fake_timm_vit(14, 24, 24, 1536, 8192, 4096, 8)) == (14, 24, 24, 1536, 4096, True, 8)          # uni_v2 (packed SwiGLU)
    assert arch_from_timm_vit(_fake_timm_vit(14, 12, 12, 768, 4096, 2048, 4)) == (14, 12, 12, 768, 2048, True, 4)            # h0_mini
    assert arch_from_timm_vit(_fake_timm_vit(16, 12, 6, 384, 1536, 1536, 0)) == (16, 12, 6, 384, 1536, False, 0)             # lunit ViT-S/16


@pytest.mark.parametrize("regs,swiglu,layerscale,pool", [(8, True, True, 0), (4, True, True, 1), (0, False, False, 0)])
def test_timm_state_dict_through_arch_and_recipe(regs, swiglu, layerscale, pool):
    """uni_v2-like (8 registers, packed SwiGLU, no_embed_class), h0_mini-like ([class || mean] over a 4-register sequence) and lunit-like
    (no LayerScale) tiny models in timm's key layout: converted tensors evaluated with plain torch ops against transformers."""
    from transformers import Dinov2Config, Dinov2Model, Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    from atlaspatch_b200 import weights as wt
    from atlaspatch_b200.dinov2 import convert_dinov2_state_dict

    name = "timm_test_tiny"
    d, heads, layers, patch = (384, 6, 2, 14) if swiglu else (256, 4, 2, 16)
    wt.DINOV2_SPECS[name], wt.DINOV2_PATCH[name] = (layers, heads, d, swiglu), patch
    if regs:
        wt.DINOV2_REGISTERS[name] = regs
    try:
        sd = wt.dinov2_state_dict(name, seed=3, image_size=224)
    finally:
        for m in (wt.DINOV2_SPECS, wt.DINOV2_PATCH, wt.DINOV2_REGISTERS):
            m.pop(name, None)
    if regs:
        sd["embeddings.position_embeddings"][:, 0] = 0.0                   # no_embed_class
    if not layerscale:
        for i in range(layers):
            sd[f"encoder.layer.{i}.layer_scale1.lambda1"].fill_(1.0)
            sd[f"encoder.layer.{i}.layer_scale2.lambda1"].fill_(1.0)
    kw = dict(hidden_size=d, num_hidden_layers=layers, num_attention_heads=heads, mlp_ratio=4, patch_size=patch, image_size=224,
              use_swiglu_ffn=swiglu, layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0)
    model = (Dinov2WithRegistersModel(Dinov2WithRegistersConfig(num_register_tokens=regs, **kw)) if regs else Dinov2Model(Dinov2Config(**kw))).eval()
    model.load_state_dict(sd, strict=True)
    x = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 3, 224, 224)).astype(np.float32))
    with torch.inference_mode():
        h = model(pixel_values=x).last_hidden_state
        want = (torch.cat([h[:, 0], h[:, 1 + regs:].mean(1)], dim=-1) if pool else h[:, 0]).numpy()
    # the same weights as timm holds them
    t = hf_to_fb_names(sd, layers, swiglu)
    t.pop("mask_token")
    if regs:
        t["reg_token"] = t.pop("register_tokens")
        t["pos_embed"] = t["pos_embed"][:, 1:]
    for i in range(layers):
        for k in ("weight", "bias"):
            if swiglu:
                t[f"blocks.{i}.mlp.fc1.{k}"], t[f"blocks.{i}.mlp.fc2.{k}"] = t.pop(f"blocks.{i}.mlp.w12.{k}"), t.pop(f"blocks.{i}.mlp.w3.{k}")
        if not layerscale:
            t.pop(f"blocks.{i}.ls1.gamma"), t.pop(f"blocks.{i}.ls2.gamma")
    mlp = 1024 if swiglu else 4 * d
    arch = arch_from_timm_vit(_fake_timm_vit(patch, layers, heads, d, 2 * mlp if swiglu else mlp, mlp, regs))
    assert arch == (patch, layers, heads, d, mlp, swiglu, regs)
    w = convert_dinov2_state_dict(t, layers=layers, swiglu=swiglu, image_size=224, patch=patch, registers=regs)
    w = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}
    with torch.inference_mode():
        got = _engine_forward(x, w, patch=patch, layers=layers, heads=heads, d=d, mlp=mlp, swiglu=swiglu, eps=1e-6, pool=pool).numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert got.shape == want.shape and rel.max() < 2e-5, rel
    # and the extractor accepts the pair up to the device init
    if not torch.cuda.is_available():
        from atlaspatch_b200._lib import AtlasB200Error
        from atlaspatch_b200.encoder import B200FeatureExtractor

        recipe, img = recipe_from_timm_data_config(dict(input_size=(3, 224, 224), interpolation="bicubic", crop_pct=1.0, **IMNET), pool=pool)
        ext = B200FeatureExtractor.__new__(B200FeatureExtractor)
        with pytest.raises(AtlasB200Error, match="no CUDA device"):
            ext.__init__("uni_like", t, input_patch=224, image_size=img, max_batch=4, arch=arch, recipe=recipe)
        ext._h = None
        assert ext.embedding_dim == want.shape[1] and ext.supports_large_reads is False
