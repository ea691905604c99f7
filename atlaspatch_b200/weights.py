"""Seeded random-init encoder weights (benchmark / test input data, no arithmetic of the path).

BASELINE.json's benchmark configs are quoted on a random-init encoder; bench.py's GPU arm, its CPU arms and the parity tests
all draw the same weights from here (oracle/weights.py re-exports this module for the test side).

No pretrained checkpoint is reachable offline (torchvision downloads at
[tv]models/vision_transformer.py:353; reference call: atlas_patch/models/patch/base.py:165),
so parity is checked with seeded weights in torchvision's own state_dict key layout
(`encoder.layers.encoder_layer_i.*`), which is also what a real checkpoint would provide.
Distributions follow torchvision's init, except that biases / LayerNorm affine parameters /
class token are perturbed away from 0/1 so that a kernel that drops one of them fails parity.
numpy's PCG64 is used (not torch's generator) so the values do not depend on the torch build.
"""
from __future__ import annotations

import math

import numpy as np
import torch

VIT_SPECS = {
    # name: (patch, layers, heads, hidden, mlp)   -- [tv]models/vision_transformer.py:635-760
    "vit_b_16": (16, 12, 12, 768, 3072),
    "vit_b_32": (32, 12, 12, 768, 3072),
    "vit_l_16": (16, 24, 16, 1024, 4096),
    "vit_l_32": (32, 24, 16, 1024, 4096),
    "vit_h_14": (14, 32, 16, 1280, 5120),
    # tiny config used only by fast unit tests
    "vit_test_tiny": (16, 2, 4, 256, 512),
}


def vit_state_dict(name: str, seed: int = 0, image_size: int = 224) -> dict[str, torch.Tensor]:
    patch, layers, heads, d, mlp = VIT_SPECS[name]
    rng = np.random.default_rng(seed)
    n_tok = (image_size // patch) ** 2 + 1

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))

    def uniform(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))

    sd: dict[str, torch.Tensor] = {}
    fan_in = 3 * patch * patch
    sd["conv_proj.weight"] = normal((d, 3, patch, patch), math.sqrt(1.0 / fan_in))
    sd["conv_proj.bias"] = normal((d,), 0.02)
    sd["class_token"] = normal((1, 1, d), 0.02)
    sd["encoder.pos_embedding"] = normal((1, n_tok, d), 0.02)
    for i in range(layers):
        p = f"encoder.layers.encoder_layer_{i}."
        sd[p + "ln_1.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "ln_1.bias"] = normal((d,), 0.05)
        sd[p + "self_attention.in_proj_weight"] = uniform((3 * d, d), math.sqrt(6.0 / (d + 3 * d)))
        sd[p + "self_attention.in_proj_bias"] = normal((3 * d,), 0.02)
        sd[p + "self_attention.out_proj.weight"] = uniform((d, d), math.sqrt(1.0 / d))
        sd[p + "self_attention.out_proj.bias"] = normal((d,), 0.02)
        sd[p + "ln_2.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "ln_2.bias"] = normal((d,), 0.05)
        sd[p + "mlp.0.weight"] = uniform((mlp, d), math.sqrt(6.0 / (d + mlp)))
        sd[p + "mlp.0.bias"] = normal((mlp,), 0.02)
        sd[p + "mlp.3.weight"] = uniform((d, mlp), math.sqrt(6.0 / (d + mlp)))
        sd[p + "mlp.3.bias"] = normal((d,), 0.02)
    sd["encoder.ln.weight"] = 1.0 + normal((d,), 0.1)
    sd["encoder.ln.bias"] = normal((d,), 0.05)
    return sd


# ---- DINOv2 (transformers Dinov2Model key layout; reference: atlas_patch/models/patch/dinov2.py:12-17 + the hub config.json files) ----
DINOV2_SPECS = {
    # name: (layers, heads, hidden, swiglu)
    "dinov2_small": (12, 6, 384, False),
    "dinov2_base": (12, 12, 768, False),
    "dinov2_large": (24, 16, 1024, False),
    "dinov2_giant": (40, 24, 1536, True),
    # tiny configs used only by fast unit tests
    "dinov2_test_tiny": (2, 4, 256, False),
    "dinov2_test_tiny_swiglu": (2, 6, 384, True),
    # other hub encoders of the reference with the Dinov2Model architecture: kaiko-ai/midnight (ViT-g/14, atlas_patch/models/patch/
    # midnight.py:12,44) and owkin/phikon-v2 (ViT-L/16, phikon.py:90-93)
    "midnight": (40, 24, 1536, True),
    "phikon_v2": (24, 16, 1024, False),
    "midnight_test_tiny": (2, 6, 384, True),
    "phikon_v2_test_tiny": (2, 4, 256, False),
    # DINOv2 with 4 register tokens: histai/hibou-B / -L (hibou.py:12-15) and the ViT-g/14 reg4 of openmidnight.py:49
    "hibou_b": (12, 12, 768, False),
    "hibou_l": (24, 16, 1024, False),
    "openmidnight": (40, 24, 1536, True),
    "hibou_test_tiny": (2, 4, 256, False),
    "openmidnight_test_tiny": (2, 6, 384, True),
    "h_optimus_0": (40, 24, 1536, True),        # hoptimus.py: timm vit_giant_patch14_reg4_dinov2
    "h_optimus_1": (40, 24, 1536, True),
    "h_optimus_test_tiny": (2, 6, 384, True),
    "pathorchestra": (24, 16, 1024, False),     # pathorchestra.py: timm ViT-L/16 with LayerScale
    "prov_gigapath": (40, 24, 1536, True),      # gigapath.py: timm ViT-g/16, packed SwiGLU
    "pathorchestra_test_tiny": (2, 4, 256, False),
    "prov_gigapath_test_tiny": (2, 6, 384, True),
}
PATCH = 14
DINOV2_PATCH = {"phikon_v2": 16, "phikon_v2_test_tiny": 16, "pathorchestra": 16, "prov_gigapath": 16, "pathorchestra_test_tiny": 16,
                "prov_gigapath_test_tiny": 16}   # conv patch where it is not 14
DINOV2_REGISTERS = {"hibou_b": 4, "hibou_l": 4, "openmidnight": 4, "hibou_test_tiny": 4, "openmidnight_test_tiny": 4,
                    "h_optimus_0": 4, "h_optimus_1": 4, "h_optimus_test_tiny": 4}


def swiglu_hidden(d: int) -> int:
    """modeling_dinov2.py Dinov2SwiGLUFFN: hidden = (int(4 d * 2 / 3) + 7) // 8 * 8."""
    return (int(int(d * 4) * 2 / 3) + 7) // 8 * 8


def dinov2_state_dict(name: str, seed: int = 0, image_size: int = 518, patch: int | None = None) -> dict[str, torch.Tensor]:
    """Seeded weights in transformers' Dinov2Model key layout (numpy PCG64, independent of the torch build).  Biases, LayerNorm
    affine parameters, LayerScale and the class token are perturbed away from their init so that dropping one fails parity."""
    layers, heads, d, swiglu = DINOV2_SPECS[name]
    rng = np.random.default_rng(seed)
    PATCH = int(patch or DINOV2_PATCH.get(name, 14))
    g = image_size // PATCH

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))

    def uniform(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))

    sd: dict[str, torch.Tensor] = {}
    sd["embeddings.cls_token"] = normal((1, 1, d), 0.02)
    sd["embeddings.mask_token"] = normal((1, d), 0.02)
    if DINOV2_REGISTERS.get(name, 0):
        sd["embeddings.register_tokens"] = normal((1, DINOV2_REGISTERS[name], d), 0.5)   # large against the 0.02 tokens: a dropped row shows
    sd["embeddings.position_embeddings"] = normal((1, g * g + 1, d), 0.02)
    sd["embeddings.patch_embeddings.projection.weight"] = normal((d, 3, PATCH, PATCH), math.sqrt(1.0 / (3 * PATCH * PATCH)))
    sd["embeddings.patch_embeddings.projection.bias"] = normal((d,), 0.02)
    hs = swiglu_hidden(d) if swiglu else 4 * d
    for i in range(layers):
        p = f"encoder.layer.{i}."
        sd[p + "norm1.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "norm1.bias"] = normal((d,), 0.05)
        for nm in ("query", "key", "value"):
            sd[p + f"attention.attention.{nm}.weight"] = uniform((d, d), math.sqrt(6.0 / (d + 3 * d)))
            sd[p + f"attention.attention.{nm}.bias"] = normal((d,), 0.02)
        sd[p + "attention.output.dense.weight"] = uniform((d, d), math.sqrt(1.0 / d))
        sd[p + "attention.output.dense.bias"] = normal((d,), 0.02)
        sd[p + "layer_scale1.lambda1"] = uniform((d,), 0.4) + 0.6        # 0.2 .. 1.0
        sd[p + "norm2.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "norm2.bias"] = normal((d,), 0.05)
        if swiglu:
            sd[p + "mlp.weights_in.weight"] = uniform((2 * hs, d), math.sqrt(6.0 / (d + hs)))
            sd[p + "mlp.weights_in.bias"] = normal((2 * hs,), 0.02)
            sd[p + "mlp.weights_out.weight"] = uniform((d, hs), math.sqrt(6.0 / (d + hs)))
            sd[p + "mlp.weights_out.bias"] = normal((d,), 0.02)
        else:
            sd[p + "mlp.fc1.weight"] = uniform((hs, d), math.sqrt(6.0 / (d + hs)))
            sd[p + "mlp.fc1.bias"] = normal((hs,), 0.02)
            sd[p + "mlp.fc2.weight"] = uniform((d, hs), math.sqrt(6.0 / (d + hs)))
            sd[p + "mlp.fc2.bias"] = normal((d,), 0.02)
        sd[p + "layer_scale2.lambda1"] = uniform((d,), 0.4) + 0.6
    sd["layernorm.weight"] = 1.0 + normal((d,), 0.1)
    sd["layernorm.bias"] = normal((d,), 0.05)
    return sd



# ---- transformers ViTModel key layout (owkin/phikon: atlas_patch/models/patch/phikon.py:41-46, add_pooling_layer=False) ----
HF_VIT_SPECS = {
    # name: (patch, layers, heads, hidden, mlp)
    "phikon_v1": (16, 12, 12, 768, 3072),
    "phikon_v1_test_tiny": (16, 2, 4, 256, 512),
}


def hf_vit_state_dict(name: str, seed: int = 0, image_size: int = 224) -> dict[str, torch.Tensor]:
    """Seeded weights in transformers' ViTModel key layout; every bias / LayerNorm parameter away from its init value."""
    patch, layers, heads, d, mlp = HF_VIT_SPECS[name]
    rng = np.random.default_rng(seed)
    g = image_size // patch

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))

    def uniform(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))

    sd: dict[str, torch.Tensor] = {}
    sd["embeddings.cls_token"] = normal((1, 1, d), 0.02)
    sd["embeddings.position_embeddings"] = normal((1, g * g + 1, d), 0.02)
    sd["embeddings.patch_embeddings.projection.weight"] = normal((d, 3, patch, patch), math.sqrt(1.0 / (3 * patch * patch)))
    sd["embeddings.patch_embeddings.projection.bias"] = normal((d,), 0.02)
    for i in range(layers):
        p = f"encoder.layer.{i}."
        sd[p + "layernorm_before.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "layernorm_before.bias"] = normal((d,), 0.05)
        for nm in ("query", "key", "value"):
            sd[p + f"attention.attention.{nm}.weight"] = uniform((d, d), math.sqrt(6.0 / (d + 3 * d)))
            sd[p + f"attention.attention.{nm}.bias"] = normal((d,), 0.02)
        sd[p + "attention.output.dense.weight"] = uniform((d, d), math.sqrt(1.0 / d))
        sd[p + "attention.output.dense.bias"] = normal((d,), 0.02)
        sd[p + "layernorm_after.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "layernorm_after.bias"] = normal((d,), 0.05)
        sd[p + "intermediate.dense.weight"] = uniform((mlp, d), math.sqrt(6.0 / (d + mlp)))
        sd[p + "intermediate.dense.bias"] = normal((mlp,), 0.02)
        sd[p + "output.dense.weight"] = uniform((d, mlp), math.sqrt(6.0 / (d + mlp)))
        sd[p + "output.dense.bias"] = normal((d,), 0.02)
    sd["layernorm.weight"] = 1.0 + normal((d,), 0.1)
    sd["layernorm.bias"] = normal((d,), 0.05)
    return sd


# ---- transformers CLIPModel key layout, image tower + visual projection (vinid/plip: atlas_patch/models/patch/plip.py:34;
#      wisdomik/QuiltNet-B-32 / -B-16: quilt.py:12-16,56) ----
HF_CLIP_SPECS = {
    # name: (patch, layers, heads, hidden, mlp, projection)
    "plip": (32, 12, 12, 768, 3072, 512),
    "quilt_b_32": (32, 12, 12, 768, 3072, 512),
    "quilt_b_16": (16, 12, 12, 768, 3072, 512),
    "plip_test_tiny": (32, 2, 4, 256, 512, 128),
    "quilt_b_16_test_tiny": (16, 2, 4, 256, 512, 128),
    "clip_vit_b_32": (32, 12, 12, 768, 3072, 512),      # clip.py:15-17 (open_clip, pretrained "openai")
    "clip_vit_b_16": (16, 12, 12, 768, 3072, 512),
    "clip_vit_l_14": (14, 24, 16, 1024, 4096, 768),
    "clip_vit_b_32_test_tiny": (32, 2, 4, 256, 512, 128),
    "clip_vit_l_14_test_tiny": (14, 2, 4, 256, 512, 128),
}


def hf_clip_state_dict(name: str, seed: int = 0, image_size: int = 224) -> dict[str, torch.Tensor]:
    """Seeded weights of the image side of transformers' CLIPModel (vision_model.* + visual_projection.weight)."""
    patch, layers, heads, d, mlp, proj = HF_CLIP_SPECS[name]
    rng = np.random.default_rng(seed)
    g = image_size // patch

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))

    def uniform(shape, bound):
        return torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))

    v = "vision_model."
    sd: dict[str, torch.Tensor] = {}
    sd[v + "embeddings.class_embedding"] = normal((d,), 0.02)
    sd[v + "embeddings.patch_embedding.weight"] = normal((d, 3, patch, patch), math.sqrt(1.0 / (3 * patch * patch)))
    sd[v + "embeddings.position_embedding.weight"] = normal((g * g + 1, d), 0.02)
    sd[v + "pre_layrnorm.weight"] = 1.0 + normal((d,), 0.1)
    sd[v + "pre_layrnorm.bias"] = normal((d,), 0.05)
    for i in range(layers):
        p = v + f"encoder.layers.{i}."
        sd[p + "layer_norm1.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "layer_norm1.bias"] = normal((d,), 0.05)
        for nm in ("q_proj", "k_proj", "v_proj"):
            sd[p + f"self_attn.{nm}.weight"] = uniform((d, d), math.sqrt(6.0 / (d + 3 * d)))
            sd[p + f"self_attn.{nm}.bias"] = normal((d,), 0.02)
        sd[p + "self_attn.out_proj.weight"] = uniform((d, d), math.sqrt(1.0 / d))
        sd[p + "self_attn.out_proj.bias"] = normal((d,), 0.02)
        sd[p + "layer_norm2.weight"] = 1.0 + normal((d,), 0.1)
        sd[p + "layer_norm2.bias"] = normal((d,), 0.05)
        sd[p + "mlp.fc1.weight"] = uniform((mlp, d), math.sqrt(6.0 / (d + mlp)))
        sd[p + "mlp.fc1.bias"] = normal((mlp,), 0.02)
        sd[p + "mlp.fc2.weight"] = uniform((d, mlp), math.sqrt(6.0 / (d + mlp)))
        sd[p + "mlp.fc2.bias"] = normal((d,), 0.02)
    sd[v + "post_layernorm.weight"] = 1.0 + normal((d,), 0.1)
    sd[v + "post_layernorm.bias"] = normal((d,), 0.05)
    sd["visual_projection.weight"] = uniform((proj, d), math.sqrt(3.0 / d))
    return sd


# ---- SAM2 (transformers Sam2Model key layout; 'tiny' = the model the reference ships, 'large' = BASELINE.json configs[2]) ----
def sam2_config(variant: str = "tiny"):
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
        shapes = {k: tuple(v.shape) for k, v in Sam2Model(sam2_config(variant)).state_dict().items()}
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


