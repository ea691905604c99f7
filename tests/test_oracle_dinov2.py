"""CPU: DINOv2 oracle pieces -- the integer restatement of torch's uint8 antialias bicubic resize against torch itself, the
preprocess against transformers' BitImageProcessorFast, the host-side weight conversion (position-embedding interpolation,
LayerScale folding, SwiGLU interleave) against transformers' Dinov2Model on a tiny config, and the committed golden vectors."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import dinov2_hf, resize_aa

GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("shape", [(224, 224, 256, 256), (512, 512, 256, 256), (448, 300, 256, 171), (37, 53, 256, 300), (256, 256, 256, 256)])
def test_resize_restatement_is_bit_exact_vs_torch(shape):
    h, w, oh, ow = shape
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).unsqueeze(0)
    ref = torch.nn.functional.interpolate(t, size=(oh, ow), mode="bicubic", antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    assert np.array_equal(resize_aa.resize_aa(img, oh, ow), ref)


@pytest.mark.parametrize("P", [224, 512, 256])
def test_pixels_match_bit_image_processor(P):
    from tests.cases import dinov2_patches

    patch = dinov2_patches("dinov2_large" if P != 512 else "dinov2_giant")[-1]
    if patch.shape[0] != P:
        patch = np.random.default_rng(P).integers(0, 256, (P, P, 3), dtype=np.uint8)
    x = dinov2_hf.preprocess([patch])[0]
    pix = resize_aa.dinov2_pixels(patch).astype(np.float32)
    want = (torch.from_numpy(pix).permute(2, 0, 1) / 255.0 - torch.tensor(dinov2_hf.MEAN).view(3, 1, 1)) / torch.tensor(dinov2_hf.STD).view(3, 1, 1)
    assert float((x - want).abs().max()) < 1e-6


def test_golden_pixels_match_restatement():
    from tests.cases import DINOV2_CASES, dinov2_patches

    for name in DINOV2_CASES:
        g = np.load(GOLDEN / f"{name}.npz")
        patches = dinov2_patches(name)[-2:]
        for i, p in enumerate(patches):
            assert np.array_equal(resize_aa.dinov2_pixels(p), g["pixels"][i])


@pytest.mark.parametrize("name", ["dinov2_test_tiny", "dinov2_test_tiny_swiglu"])
def test_converted_weights_reproduce_hf_forward(name):
    """convert_dinov2_state_dict (in_proj stacking, LayerScale folding, pos-embedding interpolation, SwiGLU interleave) evaluated
    with plain torch ops must equal transformers' Dinov2Model."""
    import torch.nn.functional as F

    from atlaspatch_b200.dinov2 import DINOV2_CONFIGS, SWIGLU_BLOCK, convert_dinov2_state_dict

    patch, layers, heads, d, mlp, swiglu = DINOV2_CONFIGS[name]
    sd = dinov2_hf.dinov2_state_dict(name, seed=3)
    model = dinov2_hf.build_model(name, sd)
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 3, 224, 224)).astype(np.float32))
    with torch.inference_mode():
        want = model(pixel_values=x).last_hidden_state[:, 0, :].numpy()
    w = {k: torch.from_numpy(np.asarray(v)) for k, v in convert_dinov2_state_dict(sd, layers=layers, swiglu=swiglu).items()}
    with torch.inference_mode():
        t = F.conv2d(x, w["conv_proj.weight"], w["conv_proj.bias"], stride=patch).reshape(2, d, -1).permute(0, 2, 1)
        t = torch.cat([w["class_token"].expand(2, -1, -1), t], dim=1) + w["encoder.pos_embedding"]
        for i in range(layers):
            p = f"encoder.layers.encoder_layer_{i}."
            y = F.layer_norm(t, (d,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=1e-6)
            q, k, v = (y @ w[p + "self_attention.in_proj_weight"].T + w[p + "self_attention.in_proj_bias"]).split(d, dim=-1)
            sh = lambda z: z.view(2, -1, heads, d // heads).transpose(1, 2)
            a = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / (d // heads) ** 0.5, dim=-1) @ sh(v)
            t = t + a.transpose(1, 2).reshape(2, -1, d) @ w[p + "self_attention.out_proj.weight"].T + w[p + "self_attention.out_proj.bias"]
            y = F.layer_norm(t, (d,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=1e-6)
            h = y @ w[p + "mlp.0.weight"].T + w[p + "mlp.0.bias"]
            if swiglu:
                h = h.reshape(2, -1, mlp // SWIGLU_BLOCK, 2, SWIGLU_BLOCK)
                h = (F.silu(h[..., 0, :]) * h[..., 1, :]).reshape(2, -1, mlp)
            else:
                h = F.gelu(h)
            t = t + h @ w[p + "mlp.3.weight"].T + w[p + "mlp.3.bias"]
        got = F.layer_norm(t, (d,), w["encoder.ln.weight"], w["encoder.ln.bias"], eps=1e-6)[:, 0].numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 2e-5, rel
