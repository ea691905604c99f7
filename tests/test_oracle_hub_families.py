"""CPU: the hub encoder families of SURVEY.md section 8(f) rank 4 (midnight, phikon_v1, phikon_v2) -- the integer pixel restatements
against the preprocess objects the reference builds, and the host-side recipe (weight conversion, LayerNorm eps, mean / std,
[class || mean] head) evaluated with plain torch ops against transformers' own models on tiny configs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hub_families as hf
from oracle import resize_aa

FAMILIES = ["midnight_test_tiny", "phikon_v2_test_tiny", "phikon_v1_test_tiny", "hibou_test_tiny", "openmidnight_test_tiny",
            "plip_test_tiny", "quilt_b_16_test_tiny", "h_optimus_test_tiny", "pathorchestra_test_tiny", "prov_gigapath_test_tiny",
            "clip_vit_b_32_test_tiny", "clip_vit_l_14_test_tiny"]


def _patch(P, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (P // 8 + 1, P // 8 + 1, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))[:P, :P]                       # blocky structure + noise
    return (img.astype(np.int32) + rng.integers(-20, 20, img.shape)).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("shape", [(256, 256, 224, 224), (512, 512, 224, 224), (224, 224, 224, 224), (100, 130, 224, 224), (300, 257, 224, 200)])
def test_bilinear_restatement_is_bit_exact_vs_torch(shape):
    h, w, oh, ow = shape
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1)[None]
    ref = F.interpolate(t, size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    assert np.array_equal(resize_aa.resize_aa(img, oh, ow, "bilinear"), ref)


@pytest.mark.parametrize("name", FAMILIES)
@pytest.mark.parametrize("P", [224, 256, 512])
def test_pixels_match_the_reference_preprocess(name, P):
    from PIL import Image

    from atlaspatch_b200.encoder import FAMILY_RECIPES, IMAGENET_MEAN, IMAGENET_STD

    patch = _patch(P, seed=P)
    got = hf.make_preprocess(name)(Image.fromarray(patch)).permute(1, 2, 0).numpy()
    r = FAMILY_RECIPES[name]
    mean, std = np.asarray(r.get("mean", IMAGENET_MEAN), np.float32), np.asarray(r.get("std", IMAGENET_STD), np.float32)
    want = (hf.pixels(name, patch).astype(np.float32) / 255.0 - mean) / std
    assert got.shape == want.shape == (224, 224, 3)
    assert np.abs(got - want).max() < 2e-6


def _engine_forward(x, w, *, patch, layers, heads, d, mlp, swiglu, eps, pool, quick_gelu=False):
    """The engine's tensor layout (torchvision names) evaluated with plain torch ops."""
    from atlaspatch_b200.dinov2 import SWIGLU_BLOCK

    B = x.shape[0]
    t = F.conv2d(x, w["conv_proj.weight"], w["conv_proj.bias"], stride=patch).reshape(B, d, -1).permute(0, 2, 1)
    t = torch.cat([w["class_token"].expand(B, -1, -1), t], dim=1) + w["encoder.pos_embedding"]
    lead = 1
    if "register_tokens" in w:      # engine layout: [class + pos_0 ; registers ; patches + pos]
        lead += w["register_tokens"].shape[0]
        t = torch.cat([t[:, :1], w["register_tokens"][None].expand(B, -1, -1), t[:, 1:]], dim=1)
    if "encoder.pre_ln.weight" in w:
        t = F.layer_norm(t, (d,), w["encoder.pre_ln.weight"], w["encoder.pre_ln.bias"], eps=eps)
    for i in range(layers):
        p = f"encoder.layers.encoder_layer_{i}."
        y = F.layer_norm(t, (d,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=eps)
        q, k, v = (y @ w[p + "self_attention.in_proj_weight"].T + w[p + "self_attention.in_proj_bias"]).split(d, dim=-1)
        sh = lambda z: z.view(B, -1, heads, d // heads).transpose(1, 2)  # noqa: E731
        a = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / (d // heads) ** 0.5, dim=-1) @ sh(v)
        t = t + a.transpose(1, 2).reshape(B, -1, d) @ w[p + "self_attention.out_proj.weight"].T + w[p + "self_attention.out_proj.bias"]
        y = F.layer_norm(t, (d,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=eps)
        h = y @ w[p + "mlp.0.weight"].T + w[p + "mlp.0.bias"]
        if swiglu:
            h = h.reshape(B, -1, mlp // SWIGLU_BLOCK, 2, SWIGLU_BLOCK)
            h = (F.silu(h[..., 0, :]) * h[..., 1, :]).reshape(B, -1, mlp)
        elif quick_gelu:
            h = h * torch.sigmoid(1.702 * h)
        else:
            h = F.gelu(h)
        t = t + h @ w[p + "mlp.3.weight"].T + w[p + "mlp.3.bias"]
    t = F.layer_norm(t, (d,), w["encoder.ln.weight"], w["encoder.ln.bias"], eps=eps)
    if "head.proj.weight" in w:
        return t[:, 0] @ w["head.proj.weight"].T
    return torch.cat([t[:, 0], t[:, lead:].mean(1)], dim=-1) if pool == 1 else t[:, 0]


@pytest.mark.parametrize("name", FAMILIES)
def test_recipe_and_converted_weights_reproduce_transformers(name):
    from atlaspatch_b200.dinov2 import (DINOV2_CONFIGS, DINOV2_REGISTERS, HF_CLIP_CONFIGS, HF_VIT_CONFIGS, convert_dinov2_state_dict,
                                        convert_hf_clip_state_dict, convert_hf_vit_state_dict)
    from atlaspatch_b200.encoder import FAMILY_RECIPES, IMAGENET_MEAN, IMAGENET_STD

    r = FAMILY_RECIPES[name]
    sd = hf.state_dict(name, seed=4)
    out_dim = None
    patches = [_patch(256, seed=s) for s in range(3)]
    want = hf.extract_features(patches, sd, name)
    if name in HF_CLIP_CONFIGS:
        (patch, layers, heads, d, mlp, out_dim), swiglu = HF_CLIP_CONFIGS[name], False
        w = convert_hf_clip_state_dict(sd, layers=layers)
    elif name in DINOV2_CONFIGS:
        patch, layers, heads, d, mlp, swiglu = DINOV2_CONFIGS[name]
        w = convert_dinov2_state_dict(sd, layers=layers, swiglu=swiglu, image_size=224, patch=patch, registers=DINOV2_REGISTERS.get(name, 0))
    else:
        (patch, layers, heads, d, mlp), swiglu = HF_VIT_CONFIGS[name], False
        w = convert_hf_vit_state_dict(sd, layers=layers)
    w = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}
    mean, std = np.asarray(r.get("mean", IMAGENET_MEAN), np.float32), np.asarray(r.get("std", IMAGENET_STD), np.float32)
    x = np.stack([(hf.pixels(name, p).astype(np.float32) / 255.0 - mean) / std for p in patches]).transpose(0, 3, 1, 2)
    with torch.inference_mode():
        got = _engine_forward(torch.from_numpy(np.ascontiguousarray(x)), w, patch=patch, layers=layers, heads=heads, d=d, mlp=mlp,
                              swiglu=swiglu, eps=r["ln_eps"], pool=r["pool"], quick_gelu=name in HF_CLIP_CONFIGS).numpy()
    assert got.shape == want.shape == (3, out_dim or d * (2 if r["pool"] == 1 else 1))
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 2e-5, rel


def test_full_size_configs_have_the_published_shapes():
    """midnight = ViT-g/14 -> 3072 features (midnight.py:53), phikon_v1 768, phikon_v2 1024 (phikon.py:11-12)."""
    from atlaspatch_b200.dinov2 import DINOV2_CONFIGS, HF_VIT_CONFIGS
    from atlaspatch_b200.encoder import FAMILY_RECIPES

    assert DINOV2_CONFIGS["midnight"][3] * (2 if FAMILY_RECIPES["midnight"]["pool"] == 1 else 1) == 3072
    assert HF_VIT_CONFIGS["phikon_v1"][3] == 768 and DINOV2_CONFIGS["phikon_v2"][3] == 1024
    for name in ("midnight", "phikon_v1", "phikon_v2"):
        cfg = DINOV2_CONFIGS.get(name) or HF_VIT_CONFIGS[name]
        assert cfg[3] // cfg[2] == 64 and 224 % cfg[0] == 0 and (224 // cfg[0]) ** 2 + 1 <= 257      # what the attention kernels cover


def hf_to_fb_names(sd, layers, swiglu):
    """Inverse of atlaspatch_b200.dinov2.fb_to_hf_dinov2_names, written independently from the facebookresearch/dinov2 module tree
    (DinoVisionTransformer: cls_token, pos_embed, register_tokens, patch_embed.proj, blocks.i.{norm1, attn.qkv, attn.proj, ls1, norm2,
    mlp.{fc1, fc2 | w12, w3}, ls2}, norm): the key layout torch.hub's dinov2_vitg14_reg holds in openmidnight.py:49-63."""
    out = {"cls_token": sd["embeddings.cls_token"], "pos_embed": sd["embeddings.position_embeddings"],
           "mask_token": sd["embeddings.mask_token"], "patch_embed.proj.weight": sd["embeddings.patch_embeddings.projection.weight"],
           "patch_embed.proj.bias": sd["embeddings.patch_embeddings.projection.bias"], "norm.weight": sd["layernorm.weight"],
           "norm.bias": sd["layernorm.bias"]}
    if "embeddings.register_tokens" in sd:
        out["register_tokens"] = sd["embeddings.register_tokens"]
    for i in range(layers):
        a, b = f"encoder.layer.{i}.", f"blocks.{i}."
        for k in ("weight", "bias"):
            out[b + f"norm1.{k}"], out[b + f"norm2.{k}"] = sd[a + f"norm1.{k}"], sd[a + f"norm2.{k}"]
            out[b + f"attn.qkv.{k}"] = torch.cat([sd[a + f"attention.attention.{n}.{k}"] for n in ("query", "key", "value")], dim=0)
            out[b + f"attn.proj.{k}"] = sd[a + f"attention.output.dense.{k}"]
            if swiglu:
                out[b + f"mlp.w12.{k}"], out[b + f"mlp.w3.{k}"] = sd[a + f"mlp.weights_in.{k}"], sd[a + f"mlp.weights_out.{k}"]
            else:
                out[b + f"mlp.fc1.{k}"], out[b + f"mlp.fc2.{k}"] = sd[a + f"mlp.fc1.{k}"], sd[a + f"mlp.fc2.{k}"]
        out[b + "ls1.gamma"], out[b + "ls2.gamma"] = sd[a + "layer_scale1.lambda1"], sd[a + "layer_scale2.lambda1"]
    return out


def hf_to_timm_names(sd, layers, swiglu):
    """timm VisionTransformer keys of vit_giant_patch14_reg4_dinov2 (hoptimus.py:53-58), written from timm's module tree: like
    facebookresearch's, but `reg_token`, SwiGLUPacked `mlp.fc1 / fc2`, and (no_embed_class) a pos_embed without the class row -- the
    class row of the transformers layout must therefore be zero for the two to mean the same thing."""
    out = hf_to_fb_names(sd, layers, swiglu)
    out.pop("mask_token")
    if "register_tokens" in out:
        out["reg_token"] = out.pop("register_tokens")
    assert float(out["pos_embed"][:, 0].abs().max()) == 0.0
    out["pos_embed"] = out["pos_embed"][:, 1:]
    for i in range(layers):
        for k in ("weight", "bias"):
            if swiglu:
                out[f"blocks.{i}.mlp.fc1.{k}"], out[f"blocks.{i}.mlp.fc2.{k}"] = out.pop(f"blocks.{i}.mlp.w12.{k}"), out.pop(f"blocks.{i}.mlp.w3.{k}")
    return out


def test_timm_key_layout_converts_to_the_same_tensors():
    from atlaspatch_b200.dinov2 import DINOV2_CONFIGS, DINOV2_REGISTERS, convert_dinov2_state_dict

    name = "h_optimus_test_tiny"
    patch, layers, heads, d, mlp, swiglu = DINOV2_CONFIGS[name]
    sd = hf.state_dict(name, seed=8)
    sd["embeddings.position_embeddings"][:, 0] = 0.0                      # no_embed_class: the class token carries no position
    kw = dict(layers=layers, swiglu=swiglu, image_size=224, patch=patch, registers=DINOV2_REGISTERS[name])
    a = convert_dinov2_state_dict(sd, **kw)
    b = convert_dinov2_state_dict(hf_to_timm_names(sd, layers, swiglu), **kw)
    assert a.keys() == b.keys() and b["encoder.pos_embedding"].shape == (1, 257, d) and not b["encoder.pos_embedding"][0, 0].any()
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name", ["openmidnight_test_tiny", "hibou_test_tiny"])
def test_facebook_key_layout_converts_to_the_same_tensors(name):
    from atlaspatch_b200.dinov2 import DINOV2_CONFIGS, DINOV2_REGISTERS, convert_dinov2_state_dict

    patch, layers, heads, d, mlp, swiglu = DINOV2_CONFIGS[name]
    sd = hf.state_dict(name, seed=8)
    kw = dict(layers=layers, swiglu=swiglu, image_size=224, patch=patch, registers=DINOV2_REGISTERS[name])
    a = convert_dinov2_state_dict(sd, **kw)
    b = convert_dinov2_state_dict(hf_to_fb_names(sd, layers, swiglu), **kw)
    assert a.keys() == b.keys() and "register_tokens" in a and a["register_tokens"].shape == (4, d)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_register_checkpoints_interpolate_their_position_grid_with_antialiasing():
    """OpenMidnight's checkpoint carries the position grid of its training resolution (openmidnight.py:58-61: "from 392 to 224"); the
    register-token models interpolate it with antialias=True (modeling_dinov2_with_registers.py, facebookresearch dinov2
    interpolate_antialias).  The converted 16 x 16 grid must reproduce transformers' forward on a 224 px input."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    from atlaspatch_b200.dinov2 import _bicubic_aa_matrix, convert_dinov2_state_dict
    from atlaspatch_b200.weights import dinov2_state_dict

    for n_in, n_out in ((28, 16), (37, 16), (10, 16)):
        x = torch.from_numpy(np.random.default_rng(n_in).standard_normal((1, 3, n_in, n_in)).astype(np.float32))
        ref = F.interpolate(x, size=(n_out, n_out), mode="bicubic", antialias=True, align_corners=False)[0].numpy()
        m = _bicubic_aa_matrix(n_in, n_out)
        assert np.abs(np.einsum("yi,xj,cij->cyx", m, m, x[0].numpy().astype(np.float64)) - ref).max() < 1e-5
    name = "openmidnight_test_tiny"
    sd = dinov2_state_dict(name, seed=0, image_size=392)                   # 28 x 28 grid
    cfg = Dinov2WithRegistersConfig(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4, patch_size=14, image_size=392,
                                    use_swiglu_ffn=True, layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0, num_register_tokens=4)
    model = Dinov2WithRegistersModel(cfg).eval()
    model.load_state_dict(sd, strict=True)
    x = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 3, 224, 224)).astype(np.float32))
    with torch.inference_mode():
        want = model(pixel_values=x).last_hidden_state[:, 0].numpy()
        w = convert_dinov2_state_dict(sd, layers=2, swiglu=True, image_size=224, patch=14, registers=4)
        assert w["encoder.pos_embedding"].shape == (1, 257, 384)
        w = {k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}
        got = _engine_forward(x, w, patch=14, layers=2, heads=6, d=384, mlp=1024, swiglu=True, eps=1e-6, pool=0).numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 2e-5, rel


def test_oracle_reproduces_the_committed_golden_features():
    """tests/golden/hub_families.npz (make_golden.py --hub): rows produced by the reference's own extractor classes where they can run
    offline (hub call replaced), by the oracle otherwise.  The live oracle must reproduce every row."""
    from pathlib import Path

    from tests.test_ref_hub_families import _patches

    g = np.load(Path(__file__).parent / "golden" / "hub_families.npz")
    names = [k[len("feats_"):] for k in g.files if k.startswith("feats_")]
    assert len(names) == 12 and sum("reference class" in str(g[f"source_{n}"]) for n in names) == 9
    patches = _patches()
    for name in names:
        want = g[f"feats_{name}"]
        got = hf.extract_features(patches, hf.state_dict(name, seed=21), name)
        assert got.shape == want.shape, name
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max(), (name, np.abs(got - want).max())


@pytest.mark.parametrize("name", FAMILIES)
def test_extractor_host_side_runs_up_to_the_device_init(name):
    """B200FeatureExtractor's host side (recipe lookup, key map, derived attributes) for every family, up to the point where the library
    is asked for a device: without a GPU that must fail loudly (no CPU fallback), with the attributes of the family already in place."""
    import torch

    from atlaspatch_b200._lib import AtlasB200Error
    from atlaspatch_b200.encoder import FAMILY_RECIPES, B200FeatureExtractor

    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    ext = B200FeatureExtractor.__new__(B200FeatureExtractor)
    with pytest.raises(AtlasB200Error, match="no CUDA device"):
        ext.__init__(name, hf.state_dict(name, seed=1), input_patch=256, max_batch=4)
    ext._h = None
    want = hf.extract_features([_patch(256)], hf.state_dict(name, seed=1), name).shape[1]
    assert ext.embedding_dim == want and ext.input_patch == 256
    assert ext.supports_large_reads is (FAMILY_RECIPES[name]["preprocess"] == 0)


def test_clip_recipe_equals_the_processor_class_defaults():
    """The CLIP checkpoints' preprocessor_config.json (vinid/plip, wisdomik/QuiltNet-*) cannot be fetched here; they are the defaults of
    transformers' CLIPImageProcessor (the OpenAI CLIP values), which this pins the recipe and the oracle to."""
    import transformers

    from atlaspatch_b200.encoder import FAMILY_RECIPES

    p = transformers.CLIPImageProcessor()
    r = FAMILY_RECIPES["plip"]
    assert dict(p.size) == {"shortest_edge": 224} and dict(p.crop_size) == {"height": 224, "width": 224} and int(p.resample) == 3
    assert p.do_resize and p.do_center_crop and p.do_rescale and p.do_normalize and abs(p.rescale_factor - 1 / 255) < 1e-12
    assert np.allclose(p.image_mean, r["mean"], atol=0, rtol=0) and np.allclose(p.image_std, r["std"], atol=0, rtol=0)
    assert r["resize_to"] == 224 and r["preprocess"] == 1
    assert np.allclose(p.image_mean, hf.CLIP_MEAN, atol=0, rtol=0) and np.allclose(p.image_std, hf.CLIP_STD, atol=0, rtol=0)
