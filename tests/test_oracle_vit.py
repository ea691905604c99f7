"""CPU tests: oracle/vit.py and oracle/thumbnail.py against the golden vectors produced by the
unmodified reference, and against the libraries the reference calls."""
import numpy as np
import pytest
import torch

from oracle import thumbnail as ot
from oracle import vit as ov
from oracle.weights import vit_state_dict
from tests.cases import FEATURE_CASE, THUMB_CASES, case_spec, feature_patches
from atlaspatch_b200.synthetic import render_region_host


def test_vit_oracle_matches_reference_golden(golden_dir):
    gold = np.load(golden_dir / "vit_b_16_feats.npz")["feats"]
    sd = vit_state_dict("vit_b_16", seed=FEATURE_CASE["weight_seed"])
    got = ov.extract_features(feature_patches(), sd, "vit_b_16", batch_size=8)
    assert got.shape == gold.shape == (16, 768)
    rel = np.linalg.norm(got - gold, axis=1) / np.linalg.norm(gold, axis=1)
    assert rel.max() < 2e-5, rel.max()   # same fp32 arithmetic, different op order only


def test_vit_oracle_matches_torchvision_tiny():
    from torchvision.models.vision_transformer import VisionTransformer

    sd = vit_state_dict("vit_test_tiny", seed=3)
    m = VisionTransformer(image_size=224, patch_size=16, num_layers=2, num_heads=4, hidden_dim=256, mlp_dim=512)
    m.heads = torch.nn.Identity()
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = ov.preprocess(feature_patches()[:4])
    with torch.inference_mode():
        want = m(x)
    got = ov.forward(x, sd, "vit_test_tiny")
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


def test_preprocess_is_crop_scale_normalize():
    p = feature_patches()[:2]
    x = ov.preprocess(p)
    from torchvision.models import ViT_B_16_Weights
    from PIL import Image

    tf = ViT_B_16_Weights.IMAGENET1K_V1.transforms()
    want = torch.stack([tf(Image.fromarray(a)) for a in p])
    assert torch.equal(x, want)
    assert ov.extract_features([], vit_state_dict("vit_test_tiny", 0), "vit_test_tiny").shape == (0, 256)


@pytest.mark.parametrize("case", THUMB_CASES, ids=[c["name"] for c in THUMB_CASES])
def test_thumbnail_oracle_matches_reference_golden(case, golden_dir):
    gold = np.load(golden_dir / f"thumb_{case['name']}.npz")["thumb"]
    spec = case_spec(case)
    mag = {0.5: 20, 0.25: 40, 1.0: 10}[case["mpp"]]
    f = ot.thumbnail_factor(mag)
    assert f == int(f)
    lvl0 = render_region_host(spec, 0, 0, spec.width, spec.height)
    got = ot.area_reduce(lvl0, int(f))
    assert np.array_equal(got, gold)


def test_reference_loop_port_matches_golden(golden_dir):
    """oracle/reference_loop.py (the timed CPU baseline) reproduces the reference's extractor output."""
    from oracle import reference_loop as rl

    gold = np.load(golden_dir / "vit_b_16_feats.npz")["feats"]
    model, preprocess = rl.build_vit_b_16(vit_state_dict("vit_b_16", seed=FEATURE_CASE["weight_seed"]))
    got = rl.extract_batch(model, preprocess, feature_patches()[:8], batch_size=8, num_workers=0)
    assert np.allclose(got, gold[:8], rtol=1e-4, atol=1e-5)


def test_cv2_exact_2x_resize_is_box_mean():
    """T9: cv2.resize(patch, (P, P)) on a 2P x 2P uint8 read = (a+b+c+d+2)>>2, what the CUDA preprocess implements."""
    import cv2

    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    got = cv2.resize(a, (256, 256))
    s = a.astype(np.uint32)
    want = ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [224, 512, 300, 100, 257, 768])
def test_pillow_bilinear_restatement_is_bit_exact(n):
    """oracle/resize_aa.py: resize_pil_bilinear (what the CUDA preprocess implements for --patch-size != 256) against Pillow itself."""
    from PIL import Image

    from oracle import resize_aa as ra

    img = np.random.default_rng(n).integers(0, 256, (n, n, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((256, 256), Image.Resampling.BILINEAR))
    assert np.array_equal(ra.resize_pil_bilinear(img, 256, 256), want)


@pytest.mark.parametrize("P", [224, 512])
def test_oracle_matches_reference_golden_for_other_patch_sizes(P, golden_dir):
    """tests/golden/vit_b_16_resize_feats.npz: produced by the reference's PatchFeatureExtractor + torchvision preset on PIL patches."""
    from tests.cases import VIT_RESIZE_CASES, vit_resize_patches

    g = np.load(golden_dir / "vit_b_16_resize_feats.npz")[f"feats_{P}"]
    sd = vit_state_dict("vit_b_16", seed=VIT_RESIZE_CASES["weight_seed"])
    patches = vit_resize_patches(P)
    got = ov.extract_features(patches, sd, "vit_b_16")
    rel = np.linalg.norm(got - g, axis=1) / np.linalg.norm(g, axis=1)
    assert rel.max() < 2e-5, rel
    # and the integer restatement of the preset's pixels equals what torchvision normalises
    from oracle import resize_aa as ra

    x = ov.preprocess(patches[:2])
    mean, std = np.asarray(ov.IMAGENET_MEAN, np.float32), np.asarray(ov.IMAGENET_STD, np.float32)
    for i in range(2):
        px = ra.vit_preset_pixels(patches[i]).astype(np.float32) / 255.0
        assert np.allclose(((px - mean) / std).transpose(2, 0, 1), x[i].numpy(), atol=1e-6)
