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
