"""GPU: DINOv2 encoders (BASELINE.json configs[3..4]) through the C ABI -- preprocess pixels bit-exact vs transformers'
BitImageProcessorFast (golden + integer oracle), features within 1e-3 relative of transformers' Dinov2Model (golden)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import dinov2_hf, resize_aa
from tests.cases import DINOV2_CASES, dinov2_coords, dinov2_patches

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _slide(name):
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    s = DINOV2_CASES[name]["slide"]
    return SyntheticWSI(make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"]))


@pytest.mark.parametrize("name,P", [("dinov2_test_tiny", 224), ("dinov2_test_tiny_swiglu", 512), ("dinov2_test_tiny", 256), ("dinov2_test_tiny", 300)])
def test_tiny_preprocess_bit_exact_and_features(name, P):
    import torch

    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.synthetic import render_region_host

    wsi = _slide("dinov2_large")
    rng = np.random.default_rng(P)
    n = 9
    xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
    xy[-1] = (wsi.w - P // 2, wsi.h - P // 3)
    rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
    patches = [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]
    sd = dinov2_hf.dinov2_state_dict(name, seed=5)
    ext = B200FeatureExtractor(name, sd, input_patch=P, max_batch=4)   # 9 patches -> three forward chunks
    rows_dev = torch.from_numpy(rows).cuda()
    pix = ext.preprocess_pixels(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev[:4])
    for i in range(4):
        assert np.array_equal(pix[i], resize_aa.dinov2_pixels(patches[i])), i
    want = dinov2_hf.extract_features(patches, sd, name)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev).cpu().numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 1e-3, rel
    got_host = ext.extract_batch(patches, batch_size=4)                 # FeatureExtractor contract, host patches
    assert np.abs(got_host - got).max() < 1e-5
    ext.cleanup()


@pytest.mark.parametrize("name", ["dinov2_small", "dinov2_base"])
def test_small_base_match_transformers(name):
    """models/patch/dinov2.py:12-13: the two smaller checkpoints (hidden 384 / 768) against transformers' Dinov2Model on the CPU."""
    import torch

    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.synthetic import render_region_host

    wsi = _slide("dinov2_large")
    rng = np.random.default_rng(3)
    n, P = 5, 224
    xy = np.stack([rng.integers(0, wsi.w - P, n), rng.integers(0, wsi.h - P, n)], 1)
    rows = np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
    patches = [render_region_host(wsi.spec, int(x), int(y), P, P) for x, y in xy]
    sd = dinov2_hf.dinov2_state_dict(name, seed=9)
    want = dinov2_hf.extract_features(patches, sd, name)
    ext = B200FeatureExtractor(name, sd, input_patch=P, max_batch=4)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(rows).cuda()).cpu().numpy()
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    print(name, "rel err per row:", rel)
    assert got.shape == want.shape and rel.max() < 1e-3, rel
    ext.cleanup()


@pytest.mark.parametrize("name", ["dinov2_large", "dinov2_giant"])
def test_matches_transformers_golden(name):
    import torch

    from atlaspatch_b200.encoder import B200FeatureExtractor

    case = DINOV2_CASES[name]
    g = np.load(GOLDEN / f"{name}.npz")
    wsi = _slide(name)
    sd = dinov2_hf.dinov2_state_dict(name, seed=case["weight_seed"])
    ext = B200FeatureExtractor(name, sd, input_patch=case["patch"], max_batch=32)
    rows_dev = torch.from_numpy(dinov2_coords(name)).cuda()
    pix = ext.preprocess_pixels(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev[-2:])
    assert np.array_equal(pix, g["pixels"])
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows_dev).cpu().numpy()
    rel = np.linalg.norm(got - g["feats"], axis=1) / np.linalg.norm(g["feats"], axis=1)
    print(name, "default precise_layers: max rel", rel.max(), "mean", rel.mean())
    ext.cleanup()
    # every golden row, at the library's default setting, under the stated tolerance.  For the 40-layer giant the default is 8
    # leading layers with hi/lo split weights AND split A operands: the last case row is 83 % black overhang (hundreds of
    # near-identical tokens -> coherent rounding errors) and sits at 7.9e-4 there.
    assert rel.max() < 1e-3, rel
