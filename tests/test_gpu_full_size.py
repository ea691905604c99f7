"""GPU: BASELINE.json configs[1] at its full size -- an 80000 x 60000 slide (14.4 GB) resident in HBM -- through size-independent
checks: sampled thumbnail pixels, sampled device-vs-host pixels beyond the 4 GB offset, the reference's golden coordinate list,
filter counts and encoder inputs / features for candidates in the far corner (64-bit addressing everywhere)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import patch_filter as pf
from oracle import vit as ov
from oracle.weights import vit_state_dict
from tests.cases import COORD_CASES, build_mask

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
W, H = 80000, 60000


@pytest.fixture(scope="module")
def big():
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    free, _ = torch.cuda.mem_get_info()
    if free < 20 << 30:
        pytest.skip("needs 20 GB of free device memory")
    wsi = SyntheticWSI(make_spec(W, H, 0))
    wsi.device_image
    yield wsi
    wsi._image = None
    torch.cuda.empty_cache()


def test_device_pixels_equal_host_generator_beyond_4gb(big):
    from atlaspatch_b200.synthetic import render_region_host

    img = big.device_image
    assert img.shape[0] == H and big.pitch >= 3 * W and H * big.pitch > (1 << 33)
    rng = np.random.default_rng(0)
    spots = [(0, 0), (W - 64, H - 48), (W - 64, 0), (0, H - 48), (40000, 30000)] + \
            [(int(rng.integers(0, W - 64)), int(rng.integers(0, H - 48))) for _ in range(20)]
    for x, y in spots:
        got = img[y:y + 48, 3 * x:3 * (x + 64)].reshape(48, 64, 3).cpu().numpy()
        assert np.array_equal(got, render_region_host(big.spec, x, y, 64, 48)), (x, y)


def test_thumbnail_sampled_pixels_are_exact(big):
    from atlaspatch_b200.synthetic import render_region_host

    thumb = big.thumbnail_at_power_device(1.25).cpu().numpy()
    assert thumb.shape == (H // 16, W // 16, 3)
    rng = np.random.default_rng(1)
    pts = [(0, 0), (W // 16 - 1, H // 16 - 1), (W // 16 - 1, 0), (0, H // 16 - 1)] + \
          [(int(rng.integers(0, W // 16)), int(rng.integers(0, H // 16))) for _ in range(300)]
    for tx, ty in pts:
        blk = render_region_host(big.spec, tx * 16, ty * 16, 16, 16).astype(np.uint32).sum(axis=(0, 1))
        want = np.clip(np.rint(blk.astype(np.float32) * np.float32(1 / 256.0)), 0, 255).astype(np.uint8)   # exact INTER_AREA mean
        assert np.array_equal(thumb[ty, tx], want), (tx, ty)


def test_coords_filter_and_encoder_inputs_at_full_size(big):
    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.extraction import filter_patches
    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.synthetic import render_region_host

    case = {c["name"]: c for c in COORD_CASES}["c1_80000x60000_p256"]
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=256))
    res = svc.extract(big, build_mask(case, big.spec), slide=Slide(Path(big.path), mpp=0.5))
    want = np.load(GOLDEN / "coords_c1_80000x60000_p256.npz")["coords"]
    assert np.array_equal(res.coords, want) and res.num_patches > 20000            # the reference's list, bit for bit

    # content filter over every candidate: order-preserving subset, counts exact on the candidates farthest into the slide
    kept, _, counts = filter_patches(big.device_image, W, H, big.pitch, res.coords_device, patch_size=256, black_threshold=142,
                                     white_threshold=6, return_counts=True)
    pos = {tuple(r): i for i, r in enumerate(res.coords.tolist())}
    idx = [pos[tuple(r)] for r in kept.tolist()]
    assert idx == sorted(idx) and 0 < len(idx) < res.num_patches
    far = np.argsort(res.coords[:, 1].astype(np.int64) * W + res.coords[:, 0])[-12:]
    _, ocounts = pf.filter_rows(lambda x, y, w, h: render_region_host(big.spec, x, y, w, h), res.coords[far], 256, 142, 6)
    assert np.array_equal(counts[far], ocounts)

    # encoder inputs (bit-exact pixels) and features for the same far candidates: byte offsets up to 14.4 GB
    sd = vit_state_dict("vit_test_tiny", seed=4)
    ext = B200FeatureExtractor("vit_test_tiny", sd, max_batch=16)
    rows = torch.from_numpy(res.coords[far]).cuda()
    pix = ext.preprocess_pixels(big.device_image, W, H, big.pitch, rows)
    patches = [render_region_host(big.spec, int(x), int(y), 256, 256) for x, y in res.coords[far, :2]]
    for i, p in enumerate(patches):
        assert np.array_equal(pix[i], p[16:240, 16:240]), i
    got = ext.embed_coords(big.device_image, W, H, big.pitch, rows).cpu().numpy()
    ref = ov.extract_features(patches, sd, "vit_test_tiny")
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-3, rel
    ext.cleanup()


def test_config3_dinov2_large_on_a_40000_square_slide():
    """BASELINE.json configs[3] (one rank's share): 40000 x 40000 slide, 224 px patches, dinov2_large.  The coordinate list must be
    the CPU oracle's (the reference's algorithm), the encoder inputs bit-exact and the features within 1e-3 of transformers' for the
    candidates farthest into the 4.8 GB image."""
    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec, render_region_host, truth_mask
    from oracle import coords as oc
    from oracle import dinov2_hf, resize_aa
    from PIL import Image

    spec = make_spec(40000, 40000, 3)
    wsi = SyntheticWSI(spec)
    m = (truth_mask(spec) * 255).astype(np.uint8)
    mask = np.asarray(Image.fromarray(m).resize((1000, 1000), Image.Resampling.NEAREST), dtype=np.float32) / 255.0
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=224, target_magnification=20, step_size=224))
    res = svc.extract(wsi, mask, slide=Slide(Path(wsi.path), mpp=0.5))
    want = oc.coords_from_mask(mask, level0_wh=(40000, 40000), src_mag=20, target_mag=20, patch_size=224, step_size=224, tissue_thresh=0.0)
    assert res.num_patches > 8000 and np.array_equal(res.coords, want)

    sd = dinov2_hf.dinov2_state_dict("dinov2_large", seed=4321)
    ext = B200FeatureExtractor("dinov2_large", sd, input_patch=224, max_batch=16)
    far = np.argsort(res.coords[:, 1].astype(np.int64) * 40000 + res.coords[:, 0])[-3:]
    rows = torch.from_numpy(res.coords[far]).cuda()
    patches = [render_region_host(spec, int(x), int(y), 224, 224) for x, y in res.coords[far, :2]]
    pix = ext.preprocess_pixels(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows)
    for i, p in enumerate(patches):
        assert np.array_equal(pix[i], resize_aa.dinov2_pixels(p)), i
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    ref = dinov2_hf.extract_features(patches, sd, "dinov2_large")
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-3, rel
    ext.cleanup()
