"""DINOv2 encoders (`dinov2_large`, `dinov2_giant`; reference: atlas_patch/models/patch/dinov2.py) on the B200 engine, and the key maps of
the other hub families that share its kernels (bottom of the file: transformers ViTModel and CLIPModel, facebookresearch / timm DINOv2-style
ViTs, open_clip's CLIP.visual).

Host-side weight preparation only: transformers' Dinov2Model state_dict (what `AutoModel.from_pretrained` gives the
reference, dinov2.py:50) is renamed to the engine's tensor names, with
* query / key / value stacked into one in_proj (one GEMM),
* LayerScale folded into the rows of attention.output.dense and mlp.fc2 / mlp.weights_out (x + ls * (W y + b) = x + (ls W) y + ls b),
* the 37 x 37 position grid of the 518 px checkpoints bicubically interpolated to the 16 x 16 grid of 224 px inputs, as
  Dinov2Embeddings.interpolate_pos_encoding does at run time (torch bicubic, A = -0.75, align_corners False),
* SwiGLU's weights_in rows interleaved in blocks of 16 (gate block, value block) so that one 32-column accumulator chunk of the
  GEMM epilogue holds both operands of silu(x1) * x2.
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

# name -> (patch, layers, heads, hidden, mlp hidden features, swiglu)
DINOV2_CONFIGS = {
    "dinov2_small": (14, 12, 6, 384, 1536, False),     # models/patch/dinov2.py:12-17
    "dinov2_base": (14, 12, 12, 768, 3072, False),
    "dinov2_large": (14, 24, 16, 1024, 4096, False),
    "dinov2_giant": (14, 40, 24, 1536, 4096, True),
    "dinov2_test_tiny": (14, 2, 4, 256, 1024, False),
    "dinov2_test_tiny_swiglu": (14, 2, 6, 384, 1024, True),
    # hub encoders with the same Dinov2Model architecture: kaiko-ai/midnight (models/patch/midnight.py:12,44: ViT-g/14, feature =
    # [class || mean of patch tokens]) and owkin/phikon-v2 (models/patch/phikon.py:90-93: ViT-L/16)
    "midnight": (14, 40, 24, 1536, 4096, True),
    "phikon_v2": (16, 24, 16, 1024, 4096, False),
    "midnight_test_tiny": (14, 2, 6, 384, 1024, True),
    "phikon_v2_test_tiny": (16, 2, 4, 256, 1024, False),
    # DINOv2 with 4 register tokens: histai/hibou-B / -L (models/patch/hibou.py:12-15, pooler_output = class token) and the
    # facebookresearch dinov2_vitg14_reg that models/patch/openmidnight.py:49 fills with the OpenMidnight weights
    "hibou_b": (14, 12, 12, 768, 3072, False),
    "hibou_l": (14, 24, 16, 1024, 4096, False),
    "openmidnight": (14, 40, 24, 1536, 4096, True),
    "hibou_test_tiny": (14, 2, 4, 256, 1024, False),
    "openmidnight_test_tiny": (14, 2, 6, 384, 1024, True),
    # bioptimus/H-optimus-0 / -1 (models/patch/hoptimus.py:53-58,98-132): timm vit_giant_patch14_reg4_dinov2 = the same ViT-g/14 with
    # 4 registers and a packed SwiGLU, class-token feature (1 536)
    "h_optimus_0": (14, 40, 24, 1536, 4096, True),
    "h_optimus_1": (14, 40, 24, 1536, 4096, True),
    "h_optimus_test_tiny": (14, 2, 6, 384, 1024, True),
    # timm ViTs with LayerScale whose preprocess the reference spells out: AI4Pathology/PathOrchestra (models/patch/pathorchestra.py:38-58,
    # ViT-L/16) and prov-gigapath/prov-gigapath (gigapath.py:17-26,46: vit_giant_patch14_dinov2 with patch 16, packed SwiGLU)
    "pathorchestra": (16, 24, 16, 1024, 4096, False),
    "prov_gigapath": (16, 40, 24, 1536, 4096, True),
    "pathorchestra_test_tiny": (16, 2, 4, 256, 1024, False),
    "prov_gigapath_test_tiny": (16, 2, 6, 384, 1024, True),
}
# name -> (patch, layers, heads, hidden, mlp, projection): image towers of transformers CLIPModel checkpoints (models/patch/plip.py:34,
# quilt.py:12-16,56): pre-LayerNorm after the embeddings, QuickGELU MLP, bias-free visual projection of the class token
HF_CLIP_CONFIGS = {
    "plip": (32, 12, 12, 768, 3072, 512),
    "quilt_b_32": (32, 12, 12, 768, 3072, 512),
    "quilt_b_16": (16, 12, 12, 768, 3072, 512),
    "plip_test_tiny": (32, 2, 4, 256, 512, 128),
    "quilt_b_16_test_tiny": (16, 2, 4, 256, 512, 128),
    # OpenAI CLIP ViTs through open_clip (models/patch/clip.py:15-17: ViT-B-32 / ViT-B-16 / ViT-L-14, pretrained "openai"); ViT-L/14 is
    # the other reading of BASELINE.json configs[3]'s "ViT-L/14".  ViT-L-14-336 (577 tokens) and the ResNets are not on these kernels
    "clip_vit_b_32": (32, 12, 12, 768, 3072, 512),
    "clip_vit_b_16": (16, 12, 12, 768, 3072, 512),
    "clip_vit_l_14": (14, 24, 16, 1024, 4096, 768),
    "clip_vit_b_32_test_tiny": (32, 2, 4, 256, 512, 128),
    "clip_vit_l_14_test_tiny": (14, 2, 4, 256, 512, 128),
}
DINOV2_REGISTERS = {"hibou_b": 4, "hibou_l": 4, "openmidnight": 4, "hibou_test_tiny": 4, "openmidnight_test_tiny": 4,
                    "h_optimus_0": 4, "h_optimus_1": 4, "h_optimus_test_tiny": 4}
# name -> (patch, layers, heads, hidden, mlp): transformers ViTModel checkpoints (owkin/phikon, models/patch/phikon.py:41-44)
HF_VIT_CONFIGS = {
    "phikon_v1": (16, 12, 12, 768, 3072),
    "phikon_v1_test_tiny": (16, 2, 4, 256, 512),
}
SWIGLU_BLOCK = 16


def _np(t) -> np.ndarray:
    return t.detach().to("cpu").float().numpy() if hasattr(t, "detach") else np.asarray(t, dtype=np.float32)


def _cubic_coeffs(t: np.ndarray, a: float = -0.75) -> np.ndarray:
    """torch's get_cubic_upsample_coefficients."""
    x1, x2 = t + 1.0, 1.0 - t
    return np.stack([((a * x1 - 5 * a) * x1 + 8 * a) * x1 - 4 * a,
                     ((a + 2) * t - (a + 3)) * t * t + 1,
                     ((a + 2) * x2 - (a + 3)) * x2 * x2 + 1,
                     ((a * (x2 + 1) - 5 * a) * (x2 + 1) + 8 * a) * (x2 + 1) - 4 * a], axis=-1)


def _bicubic_matrix(n_in: int, n_out: int) -> np.ndarray:
    """(n_out, n_in) matrix of torch.nn.functional.interpolate(mode="bicubic", align_corners=False) along one axis."""
    src = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5
    i0 = np.floor(src)
    w = _cubic_coeffs(src - i0)
    m = np.zeros((n_out, n_in), dtype=np.float64)
    for k in range(4):
        idx = np.clip(i0.astype(np.int64) - 1 + k, 0, n_in - 1)
        np.add.at(m, (np.arange(n_out), idx), w[:, k])
    return m


def _bicubic_aa_matrix(n_in: int, n_out: int) -> np.ndarray:
    """(n_out, n_in) matrix of torch.nn.functional.interpolate(mode="bicubic", antialias=True, align_corners=False) along one axis for
    float tensors (ATen _compute_indices_weights_aa: Keys cubic a = -0.5, support 2 max(scale, 1), weights normalised)."""
    scale = n_in / n_out
    support = 2.0 * scale if scale >= 1.0 else 2.0
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    a = -0.5
    m = np.zeros((n_out, n_in), dtype=np.float64)
    for i in range(n_out):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(int(center + support + 0.5), n_in) - xmin
        t = np.abs((np.arange(xsize) + xmin - center + 0.5) * invscale)
        w = np.where(t < 1.0, ((a + 2.0) * t - (a + 3.0)) * t * t + 1.0, np.where(t < 2.0, (((t - 5.0) * t + 8.0) * t - 4.0) * a, 0.0))
        m[i, xmin:xmin + xsize] = w / w.sum()
    return m


def interpolate_pos_embedding(pos: np.ndarray, grid_out: int, antialias: bool = False) -> np.ndarray:
    """(1 + g*g, D) -> (1 + grid_out^2, D); identity when the grids match (modeling_dinov2.py interpolate_pos_encoding).  antialias:
    the register-token models' variant (modeling_dinov2_with_registers.py, facebookresearch dinov2 `interpolate_antialias=True`)."""
    n = pos.shape[0] - 1
    g = int(round(n ** 0.5))
    assert g * g == n, "position embedding is not a square grid"
    if g == grid_out:
        return pos.astype(np.float32)
    m = _bicubic_aa_matrix(g, grid_out) if antialias else _bicubic_matrix(g, grid_out)
    grid = pos[1:].astype(np.float64).reshape(g, g, -1)
    out = np.einsum("yi,xj,ijd->yxd", m, m, grid)
    return np.concatenate([pos[:1].astype(np.float32), out.reshape(grid_out * grid_out, -1).astype(np.float32)], axis=0)


def swiglu_interleave(hidden_features: int) -> np.ndarray:
    """Row permutation of weights_in: new row r reads old row perm[r]; blocks of 16 gate rows then their 16 value rows."""
    assert hidden_features % SWIGLU_BLOCK == 0
    b = np.arange(hidden_features).reshape(-1, SWIGLU_BLOCK)
    return np.concatenate([b, b + hidden_features], axis=1).reshape(-1)


def convert_dinov2_state_dict(sd: Mapping[str, object], *, layers: int, swiglu: bool, image_size: int = 224,
                              patch: int = 14, registers: int = 0) -> dict[str, np.ndarray]:
    """transformers Dinov2Model / Dinov2WithRegistersModel names -> engine names (the torchvision layout of encoder.py:
    vit_state_dict_names).  Facebook's own key layout (torch.hub dinov2_*: cls_token, blocks.i.attn.qkv ...) is accepted too."""
    if "cls_token" in sd and "embeddings.cls_token" not in sd:
        sd = fb_to_hf_dinov2_names(sd, layers=layers, swiglu=swiglu)
    sd = {k[len("dinov2."):] if k.startswith("dinov2.") else k: v for k, v in sd.items()}
    out: dict[str, np.ndarray] = {}
    if registers:
        # Dinov2WithRegistersEmbeddings: [class + pos_0 ; registers (no position) ; patches + pos]; a position grid of another
        # resolution (OpenMidnight's checkpoint carries its training grid: openmidnight.py:58-61) is interpolated with antialiasing
        out["register_tokens"] = _np(sd["embeddings.register_tokens"]).reshape(registers, -1)
    out["conv_proj.weight"] = _np(sd["embeddings.patch_embeddings.projection.weight"])
    out["conv_proj.bias"] = _np(sd["embeddings.patch_embeddings.projection.bias"])
    out["class_token"] = _np(sd["embeddings.cls_token"]).reshape(1, 1, -1)
    pos = _np(sd["embeddings.position_embeddings"])
    out["encoder.pos_embedding"] = interpolate_pos_embedding(pos.reshape(pos.shape[-2], pos.shape[-1]), image_size // patch,
                                                             antialias=registers > 0)[None]
    out["encoder.ln.weight"] = _np(sd["layernorm.weight"])
    out["encoder.ln.bias"] = _np(sd["layernorm.bias"])
    for i in range(layers):
        s, d = f"encoder.layer.{i}.", f"encoder.layers.encoder_layer_{i}."
        out[d + "ln_1.weight"], out[d + "ln_1.bias"] = _np(sd[s + "norm1.weight"]), _np(sd[s + "norm1.bias"])
        out[d + "ln_2.weight"], out[d + "ln_2.bias"] = _np(sd[s + "norm2.weight"]), _np(sd[s + "norm2.bias"])
        out[d + "self_attention.in_proj_weight"] = np.concatenate(
            [_np(sd[s + f"attention.attention.{n}.weight"]) for n in ("query", "key", "value")], axis=0)
        out[d + "self_attention.in_proj_bias"] = np.concatenate(
            [_np(sd[s + f"attention.attention.{n}.bias"]) for n in ("query", "key", "value")], axis=0)
        ls1, ls2 = _np(sd[s + "layer_scale1.lambda1"]), _np(sd[s + "layer_scale2.lambda1"])
        out[d + "self_attention.out_proj.weight"] = _np(sd[s + "attention.output.dense.weight"]) * ls1[:, None]
        out[d + "self_attention.out_proj.bias"] = _np(sd[s + "attention.output.dense.bias"]) * ls1
        if swiglu:
            w_in, b_in = _np(sd[s + "mlp.weights_in.weight"]), _np(sd[s + "mlp.weights_in.bias"])
            perm = swiglu_interleave(w_in.shape[0] // 2)
            out[d + "mlp.0.weight"], out[d + "mlp.0.bias"] = np.ascontiguousarray(w_in[perm]), np.ascontiguousarray(b_in[perm])
            w_out, b_out = _np(sd[s + "mlp.weights_out.weight"]), _np(sd[s + "mlp.weights_out.bias"])
        else:
            out[d + "mlp.0.weight"], out[d + "mlp.0.bias"] = _np(sd[s + "mlp.fc1.weight"]), _np(sd[s + "mlp.fc1.bias"])
            w_out, b_out = _np(sd[s + "mlp.fc2.weight"]), _np(sd[s + "mlp.fc2.bias"])
        out[d + "mlp.3.weight"] = w_out * ls2[:, None]
        out[d + "mlp.3.bias"] = b_out * ls2
    return out


def convert_hf_vit_state_dict(sd: Mapping[str, object], *, layers: int) -> dict[str, np.ndarray]:
    """transformers ViTModel names (what `ViTModel.from_pretrained("owkin/phikon", add_pooling_layer=False)` gives the reference,
    models/patch/phikon.py:41-44) -> engine names.  Same pre-LayerNorm block as torchvision's (layernorm_before / attention /
    layernorm_after / intermediate GELU / output), query / key / value stacked into one in_proj; no LayerScale, learned position
    embedding used as stored (224 px checkpoints, 14 x 14 grid)."""
    sd = {k[len("vit."):] if k.startswith("vit.") else k: v for k, v in sd.items()}
    out: dict[str, np.ndarray] = {}
    out["conv_proj.weight"] = _np(sd["embeddings.patch_embeddings.projection.weight"])
    out["conv_proj.bias"] = _np(sd["embeddings.patch_embeddings.projection.bias"])
    out["class_token"] = _np(sd["embeddings.cls_token"]).reshape(1, 1, -1)
    pos = _np(sd["embeddings.position_embeddings"])
    out["encoder.pos_embedding"] = pos.reshape(1, pos.shape[-2], pos.shape[-1])
    out["encoder.ln.weight"], out["encoder.ln.bias"] = _np(sd["layernorm.weight"]), _np(sd["layernorm.bias"])
    for i in range(layers):
        s, d = f"encoder.layer.{i}.", f"encoder.layers.encoder_layer_{i}."
        out[d + "ln_1.weight"], out[d + "ln_1.bias"] = _np(sd[s + "layernorm_before.weight"]), _np(sd[s + "layernorm_before.bias"])
        out[d + "ln_2.weight"], out[d + "ln_2.bias"] = _np(sd[s + "layernorm_after.weight"]), _np(sd[s + "layernorm_after.bias"])
        out[d + "self_attention.in_proj_weight"] = np.concatenate(
            [_np(sd[s + f"attention.attention.{n}.weight"]) for n in ("query", "key", "value")], axis=0)
        out[d + "self_attention.in_proj_bias"] = np.concatenate(
            [_np(sd[s + f"attention.attention.{n}.bias"]) for n in ("query", "key", "value")], axis=0)
        out[d + "self_attention.out_proj.weight"] = _np(sd[s + "attention.output.dense.weight"])
        out[d + "self_attention.out_proj.bias"] = _np(sd[s + "attention.output.dense.bias"])
        out[d + "mlp.0.weight"], out[d + "mlp.0.bias"] = _np(sd[s + "intermediate.dense.weight"]), _np(sd[s + "intermediate.dense.bias"])
        out[d + "mlp.3.weight"], out[d + "mlp.3.bias"] = _np(sd[s + "output.dense.weight"]), _np(sd[s + "output.dense.bias"])
    return out


def fb_to_hf_dinov2_names(sd: Mapping[str, object], *, layers: int, swiglu: bool) -> dict[str, object]:
    """facebookresearch/dinov2 DinoVisionTransformer keys (what torch.hub's dinov2_vitg14_reg holds after models/patch/
    openmidnight.py:49-63 loads the checkpoint into it; block_chunks = 0 as the hub entry points build it) -> transformers names.
    Same tensors, same arithmetic: qkv rows are [q ; k ; v], SwiGLUFFNFused.w12 = weights_in (gate half first), w3 = weights_out.
    timm's VisionTransformer (`vit_giant_patch14_reg4_dinov2` of bioptimus/H-optimus-0 / -1, models/patch/hoptimus.py:53-58) uses the
    same module tree with three differences, accepted here: `reg_token` for the registers, SwiGLUPacked's `mlp.fc1` / `mlp.fc2`
    for w12 / w3 (x1, x2 = fc1(x).chunk(2); fc2(silu(x1) * x2)), and -- with no_embed_class -- a `pos_embed` without the class row
    (the class and register tokens get no position: a zero row is put in front)."""
    pos = _np(sd["pos_embed"])
    pos = pos.reshape(1, pos.shape[-2], pos.shape[-1])
    if int(round((pos.shape[1] - 1) ** 0.5)) ** 2 != pos.shape[1] - 1:      # timm no_embed_class: rows = patches only
        pos = np.concatenate([np.zeros((1, 1, pos.shape[-1]), dtype=np.float32), pos], axis=1)
    out: dict[str, object] = {"embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": pos,
                              "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
                              "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
                              "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for key in ("register_tokens", "reg_token"):
        if key in sd:
            out["embeddings.register_tokens"] = sd[key]
    for i in range(layers):
        s, d = f"blocks.{i}.", f"encoder.layer.{i}."
        for a, b in (("norm1", "norm1"), ("norm2", "norm2"), ("attn.proj", "attention.output.dense")):
            out[d + b + ".weight"], out[d + b + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
        qkv_w, qkv_b = _np(sd[s + "attn.qkv.weight"]), _np(sd[s + "attn.qkv.bias"])
        D = qkv_w.shape[1]
        for j, n in enumerate(("query", "key", "value")):
            out[d + f"attention.attention.{n}.weight"] = qkv_w[j * D:(j + 1) * D]
            out[d + f"attention.attention.{n}.bias"] = qkv_b[j * D:(j + 1) * D]
        ones = np.ones((D,), dtype=np.float32)                          # timm ViTs built without init_values carry no LayerScale
        out[d + "layer_scale1.lambda1"], out[d + "layer_scale2.lambda1"] = sd.get(s + "ls1.gamma", ones), sd.get(s + "ls2.gamma", ones)
        if swiglu:
            packed = (s + "mlp.w12.weight") not in sd                  # timm SwiGLUPacked
            pairs = (("mlp.fc1" if packed else "mlp.w12", "mlp.weights_in"), ("mlp.fc2" if packed else "mlp.w3", "mlp.weights_out"))
        else:
            pairs = (("mlp.fc1", "mlp.fc1"), ("mlp.fc2", "mlp.fc2"))
        for a, b in pairs:
            out[d + b + ".weight"], out[d + b + ".bias"] = sd[s + a + ".weight"], sd[s + a + ".bias"]
    return out


def convert_hf_clip_state_dict(sd: Mapping[str, object], *, layers: int) -> dict[str, np.ndarray]:
    """transformers CLIPModel names (`CLIPModel.from_pretrained("vinid/plip")`, models/patch/plip.py:34; text tower ignored) -> engine
    names.  CLIPVisionEmbeddings: bias-free patch convolution, class_embedding + position_embedding; pre_layrnorm -> "encoder.pre_ln";
    q / k / v projections stacked into in_proj (the 1 / sqrt(head_dim) scale is the attention kernel's); post_layernorm ->
    "encoder.ln"; visual_projection -> "head.proj.weight".  open_clip's key layout (`visual.conv1`, `visual.transformer.resblocks.i` ...,
    what models/patch/clip.py:36-40 holds) is accepted too."""
    if "visual.conv1.weight" in sd:
        sd = openclip_to_hf_clip_names(sd, layers=layers)
    v = "vision_model."
    out: dict[str, np.ndarray] = {}
    w = _np(sd[v + "embeddings.patch_embedding.weight"])
    out["conv_proj.weight"] = w
    out["conv_proj.bias"] = np.zeros((w.shape[0],), dtype=np.float32)
    out["class_token"] = _np(sd[v + "embeddings.class_embedding"]).reshape(1, 1, -1)
    pos = _np(sd[v + "embeddings.position_embedding.weight"])
    out["encoder.pos_embedding"] = pos.reshape(1, pos.shape[-2], pos.shape[-1])
    out["encoder.pre_ln.weight"], out["encoder.pre_ln.bias"] = _np(sd[v + "pre_layrnorm.weight"]), _np(sd[v + "pre_layrnorm.bias"])
    out["encoder.ln.weight"], out["encoder.ln.bias"] = _np(sd[v + "post_layernorm.weight"]), _np(sd[v + "post_layernorm.bias"])
    out["head.proj.weight"] = _np(sd["visual_projection.weight"])
    for i in range(layers):
        s, d = v + f"encoder.layers.{i}.", f"encoder.layers.encoder_layer_{i}."
        out[d + "ln_1.weight"], out[d + "ln_1.bias"] = _np(sd[s + "layer_norm1.weight"]), _np(sd[s + "layer_norm1.bias"])
        out[d + "ln_2.weight"], out[d + "ln_2.bias"] = _np(sd[s + "layer_norm2.weight"]), _np(sd[s + "layer_norm2.bias"])
        out[d + "self_attention.in_proj_weight"] = np.concatenate([_np(sd[s + f"self_attn.{n}.weight"]) for n in ("q_proj", "k_proj", "v_proj")], axis=0)
        out[d + "self_attention.in_proj_bias"] = np.concatenate([_np(sd[s + f"self_attn.{n}.bias"]) for n in ("q_proj", "k_proj", "v_proj")], axis=0)
        out[d + "self_attention.out_proj.weight"], out[d + "self_attention.out_proj.bias"] = _np(sd[s + "self_attn.out_proj.weight"]), _np(sd[s + "self_attn.out_proj.bias"])
        out[d + "mlp.0.weight"], out[d + "mlp.0.bias"] = _np(sd[s + "mlp.fc1.weight"]), _np(sd[s + "mlp.fc1.bias"])
        out[d + "mlp.3.weight"], out[d + "mlp.3.bias"] = _np(sd[s + "mlp.fc2.weight"]), _np(sd[s + "mlp.fc2.bias"])
    return out


def openclip_to_hf_clip_names(sd: Mapping[str, object], *, layers: int) -> dict[str, object]:
    """open_clip CLIP.visual (VisionTransformer) keys -> transformers CLIPModel names.  Same tensors: nn.MultiheadAttention's
    in_proj rows are [q ; k ; v]; `visual.proj` is stored [width, output_dim] and applied as `pooled @ proj`, i.e. the transpose of
    visual_projection.weight."""
    v = "vision_model."
    out: dict[str, object] = {v + "embeddings.class_embedding": sd["visual.class_embedding"],
                              v + "embeddings.patch_embedding.weight": sd["visual.conv1.weight"],
                              v + "embeddings.position_embedding.weight": sd["visual.positional_embedding"],
                              "visual_projection.weight": np.ascontiguousarray(_np(sd["visual.proj"]).T)}
    for a, b in (("visual.ln_pre", "pre_layrnorm"), ("visual.ln_post", "post_layernorm")):
        out[v + b + ".weight"], out[v + b + ".bias"] = sd[a + ".weight"], sd[a + ".bias"]
    for i in range(layers):
        s, d = f"visual.transformer.resblocks.{i}.", v + f"encoder.layers.{i}."
        w, b = _np(sd[s + "attn.in_proj_weight"]), _np(sd[s + "attn.in_proj_bias"])
        D = w.shape[1]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            out[d + f"self_attn.{n}.weight"], out[d + f"self_attn.{n}.bias"] = w[j * D:(j + 1) * D], b[j * D:(j + 1) * D]
        for a2, b2 in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("attn.out_proj", "self_attn.out_proj"), ("mlp.c_fc", "mlp.fc1"),
                       ("mlp.c_proj", "mlp.fc2")):
            out[d + b2 + ".weight"], out[d + b2 + ".bias"] = sd[s + a2 + ".weight"], sd[s + a2 + ".bias"]
    return out
