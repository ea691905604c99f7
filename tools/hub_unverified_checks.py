"""GPU checks of the hub-family paths that were built after the round's GPU budget was spent (DESIGN.md section 1 "Encoder registry
coverage", section 7.5).  A tool, not a test: nothing here has been run on a GPU yet, so it must not be able to turn the driver's
`pytest -m gpu` red.  Run on a B200 first thing next round:

    python tools/hub_unverified_checks.py [regs8] [depth40] [aa_pos]

regs8    an 8-register, packed-SwiGLU, class-less-pos_embed tiny model in timm's key layout through B200FeatureExtractor(arch=, recipe=)
         (the uni_v2 path of plugin.py: 265-token sequence) against transformers' Dinov2WithRegistersModel;
depth40  feature error of the full-depth 40-layer register model (openmidnight / h_optimus shape, 261 tokens) on 4 patches: the A-operand
         splits that keep dinov2_giant at 7.9e-4 are unavailable above 257 tokens, so this is the number that says whether they are needed;
aa_pos   a register-token checkpoint carrying a 28 x 28 position grid (OpenMidnight's training resolution) interpolated on the host.
Each check prints the per-row relative error; the bar is 1e-3.
Run so far: regs8 (B200, end of round 2): 4.6e-4 .. 5.8e-4 on 9 rows, extract_batch identical to embed_coords.
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200 import weights as wt  # noqa: E402
from atlaspatch_b200.encoder import FAMILY_RECIPES, B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, render_region_host  # noqa: E402
from oracle import hub_families as hf  # noqa: E402


def _inputs(n, P, seed):
    wsi = SyntheticWSI(make_spec(4096, 3072, 11, mpp=0.5))
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
    rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
    return wsi, torch.from_numpy(rows).cuda(), [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]


def _rel(got, want):
    return np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)


def regs8():
    from tests.test_oracle_hub_families import hf_to_timm_names
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    name, regs, layers, heads, d, mlp = "uni2_like_tiny", 8, 2, 6, 384, 1024
    wt.DINOV2_SPECS[name], wt.DINOV2_REGISTERS[name] = (layers, heads, d, True), regs
    sd = wt.dinov2_state_dict(name, seed=5, image_size=224)
    sd["embeddings.position_embeddings"][:, 0] = 0.0
    model = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(hidden_size=d, num_hidden_layers=layers, num_attention_heads=heads, mlp_ratio=4,
                                                               patch_size=14, image_size=224, use_swiglu_ffn=True, layer_norm_eps=1e-6,
                                                               qkv_bias=True, layerscale_value=1.0, num_register_tokens=regs)).eval()
    model.load_state_dict(sd, strict=True)
    recipe = dict(FAMILY_RECIPES["pathorchestra"])                      # Pillow BILINEAR to 224, ImageNet mean / std
    wsi, rows, patches = _inputs(9, 256, 1)
    pre = hf.make_preprocess("pathorchestra_test_tiny")
    from PIL import Image

    with torch.inference_mode():
        want = model(pixel_values=torch.stack([pre(Image.fromarray(p)) for p in patches])).last_hidden_state[:, 0].numpy()
    ext = B200FeatureExtractor(name, hf_to_timm_names(sd, layers, True), input_patch=256, max_batch=4, arch=(14, layers, heads, d, mlp, True, regs),
                               recipe=recipe)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    print("regs8: 265-token sequence, rel err per row", _rel(got, want))
    assert np.abs(ext.extract_batch(patches) - got).max() < 1e-5
    ext.cleanup()


def depth40():
    name = "openmidnight"
    sd = hf.state_dict(name, seed=3)
    wsi, rows, patches = _inputs(4, 224, 2)
    want = hf.extract_features(patches, sd, name, batch_size=2)          # 40-layer ViT-g on the CPU: about a minute
    ext = B200FeatureExtractor(name, sd, input_patch=224, max_batch=4)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    print("depth40: openmidnight shape (40 layers, 261 tokens, weight-split precise layers only), rel err per row", _rel(got, want))
    ext.cleanup()


def aa_pos():
    name = "openmidnight_test_tiny"
    sd = wt.dinov2_state_dict(name, seed=0, image_size=392)              # 28 x 28 grid
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel

    model = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4, patch_size=14,
                                                               image_size=392, use_swiglu_ffn=True, layer_norm_eps=1e-6, qkv_bias=True,
                                                               layerscale_value=1.0, num_register_tokens=4)).eval()
    model.load_state_dict(sd, strict=True)
    wsi, rows, patches = _inputs(5, 224, 3)
    pre = hf.make_preprocess(name)
    from PIL import Image

    with torch.inference_mode():
        want = model(pixel_values=torch.stack([pre(Image.fromarray(p)) for p in patches])).last_hidden_state[:, 0].numpy()
    ext = B200FeatureExtractor(name, sd, input_patch=224, max_batch=4)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    print("aa_pos: 28 x 28 -> 16 x 16 antialiased position grid, rel err per row", _rel(got, want))
    ext.cleanup()


if __name__ == "__main__":
    todo = sys.argv[1:] or ["regs8", "aa_pos", "depth40"]
    for t in todo:
        {"regs8": regs8, "depth40": depth40, "aa_pos": aa_pos}[t]()
