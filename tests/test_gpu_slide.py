"""GPU parity tests: synthetic generator, thumbnail (a1), coordinate extraction (a6-a9)."""
import numpy as np
import pytest
import torch

from oracle import coords as oc
from oracle import thumbnail as ot
from tests.cases import THUMB_GENERAL_CASES, COORD_CASES, THUMB_CASES, build_mask, case_spec

pytestmark = pytest.mark.gpu


def _mag(mpp):
    from atlaspatch_b200.geometry import infer_mag

    return infer_mag(mpp)


def test_synth_device_matches_host():
    from atlaspatch_b200.slide import render_device
    from atlaspatch_b200.synthetic import make_spec, render_region_host

    spec = make_spec(6000, 4000, seed=5)
    for (x, y, w, h) in [(0, 0, 512, 300), (1777, 901, 333, 257), (-40, -30, 200, 100), (5900, 3900, 300, 300), (3000, 1990, 16, 16)]:
        buf, pitch = render_device(spec, x, y, w, h)
        got = buf[:, : w * 3].reshape(h, w, 3).cpu().numpy()
        assert np.array_equal(got, render_region_host(spec, x, y, w, h)), (x, y, w, h)


@pytest.mark.parametrize("case", THUMB_CASES, ids=[c["name"] for c in THUMB_CASES])
def test_thumbnail_matches_reference_golden(case, golden_dir):
    from atlaspatch_b200.slide import SyntheticWSI

    gold = np.load(golden_dir / f"thumb_{case['name']}.npz")["thumb"]
    wsi = SyntheticWSI(case_spec(case))
    got = wsi.thumbnail_at_power_device(1.25).cpu().numpy()
    assert got.shape == gold.shape and np.array_equal(got, gold)


@pytest.mark.parametrize("f", [2, 5, 16, 32, 64])
def test_thumbnail_matches_oracle_random(f):
    from atlaspatch_b200.slide import thumbnail_area

    rng = np.random.default_rng(f)
    H, W = 4 * f * 3, 16 * f * 5
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: f, : f] = 255
    img[f: 2 * f, : f] = 0
    pitch = (W * 3 + 15) // 16 * 16
    dev = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda")
    dev[:, : W * 3] = torch.from_numpy(img.reshape(H, W * 3)).cuda()
    got = thumbnail_area(dev, W, H, pitch, f).cpu().numpy()
    assert np.array_equal(got, ot.area_reduce(img, f))


def test_thumbnail_rejects_non_dividing_factor():
    from atlaspatch_b200.slide import thumbnail_area

    dev = torch.zeros((100, 304), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        thumbnail_area(dev, 100, 100, 304, 16)


@pytest.mark.parametrize("case", THUMB_GENERAL_CASES, ids=[c["name"] for c in THUMB_GENERAL_CASES])
def test_thumbnail_general_matches_reference_golden(case, golden_dir):
    """Level sizes the factor does not divide (cv2's fractional INTER_AREA weights) and the 60x factor 48, against thumbnails the
    unmodified reference produced (IWSI.get_thumbnail_at_power)."""
    from atlaspatch_b200.slide import SyntheticWSI

    gold = np.load(golden_dir / f"thumb_{case['name']}.npz")["thumb"]
    got = SyntheticWSI(case_spec(case)).thumbnail_at_power_device(1.25).cpu().numpy()
    assert got.shape == gold.shape and np.array_equal(got, gold)


@pytest.mark.parametrize("W,H,ow,oh", [(1000, 777, 63, 49), (515, 300, 32, 19), (999, 1001, 125, 63), (640, 480, 40, 60), (20000, 9000, 1237, 561)])
def test_thumbnail_resize_matches_cv2(W, H, ow, oh):
    """ap_thumbnail_resize against cv2.resize(INTER_AREA) itself: fractional scales, unequal integer scales (16 x 8), a large level."""
    import cv2

    from atlaspatch_b200.slide import thumbnail_resize

    src = np.random.default_rng(W + H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    pitch = (W * 3 + 15) // 16 * 16
    dev = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda")
    dev[:, :W * 3] = torch.from_numpy(src.reshape(H, W * 3)).cuda()
    got = thumbnail_resize(dev, W, H, pitch, ow, oh).cpu().numpy()
    assert np.array_equal(got, cv2.resize(src, (ow, oh), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize("case", COORD_CASES, ids=[c["name"] for c in COORD_CASES])
def test_coords_match_reference_golden(case, golden_dir):
    from atlaspatch_b200.extraction import extract_coords

    gold = np.load(golden_dir / f"coords_{case['name']}.npz")["coords"]
    spec = case_spec(case)
    mask = build_mask(case, spec)
    got, got_dev = extract_coords(mask, level0_wh=(spec.width, spec.height), src_mag=_mag(spec.mpp),
                                  target_mag=case["target_mag"], patch_size=case["patch"], step_size=case["step"],
                                  tissue_thresh=case["tissue_thresh"], return_device=True)
    assert got.dtype == np.int32 and got.shape == gold.shape
    assert np.array_equal(got, gold)                       # bit-exact, including order
    assert np.array_equal(got_dev.cpu().numpy(), gold)


@pytest.mark.parametrize("seed", range(6))
def test_coords_match_oracle_random_masks(seed):
    from atlaspatch_b200.extraction import extract_coords

    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 300)), int(rng.integers(40, 300))
    mask = build_mask(dict(mask="noisy", mask_hw=(h, w), seed=100 + seed), None)
    W, H = int(w * rng.uniform(3, 40)), int(h * rng.uniform(3, 40))
    patch = int(rng.choice([32, 64, 100, 224, 256]))
    step = int(rng.choice([patch, patch // 2, patch + 7]))
    kw = dict(level0_wh=(W, H), src_mag=20, target_mag=20, patch_size=patch, step_size=step, tissue_thresh=0.0)
    want = oc.coords_from_mask(mask, **kw)
    got = extract_coords(mask, **kw)
    assert np.array_equal(got, want)


def test_coords_properties_at_full_size():
    """BASELINE configs[1] size: every row lies on its contour's grid, rows are unique per contour order,
    and a second run is identical (idempotence)."""
    from atlaspatch_b200.extraction import extract_coords

    case = next(c for c in COORD_CASES if c["name"].startswith("c1_"))
    spec = case_spec(case)
    mask = build_mask(case, spec)
    kw = dict(level0_wh=(spec.width, spec.height), src_mag=20, target_mag=20, patch_size=256, step_size=256)
    a = extract_coords(mask, **kw)
    b = extract_coords(mask, **kw)
    assert np.array_equal(a, b) and a.shape[0] > 10000
    assert (a[:, 2] == 256).all() and (a[:, 3] == 256).all() and (a[:, 4] == 0).all()
    assert a[:, 0].min() >= 0 and a[:, 0].max() < spec.width and a[:, 1].max() < spec.height
