"""CPU tests: the oracle restatement (oracle/coords.py) against (a) golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) and (b) the OpenCV routines the reference calls."""
import cv2
import numpy as np
import pytest

from oracle import coords as oc
from tests.cases import COORD_CASES, build_mask, case_spec


def _mag_from_mpp(mpp):  # core/wsi/iwsi.py:360-384
    for thr, mag in [(0.16, 80), (0.2, 60), (0.3, 40), (0.6, 20), (1.2, 10), (2.4, 5)]:
        if mpp < thr:
            return mag
    raise ValueError(mpp)


@pytest.mark.parametrize("case", COORD_CASES, ids=[c["name"] for c in COORD_CASES])
def test_oracle_matches_reference_golden(case, golden_dir):
    gold = np.load(golden_dir / f"coords_{case['name']}.npz")["coords"]
    spec = case_spec(case)
    mask = build_mask(case, spec)
    got = oc.coords_from_mask(mask, level0_wh=(spec.width, spec.height), src_mag=_mag_from_mpp(spec.mpp),
                              target_mag=case["target_mag"], patch_size=case["patch"], step_size=case["step"],
                              tissue_thresh=case["tissue_thresh"])
    assert got.dtype == np.int32 and got.shape == gold.shape
    assert np.array_equal(got, gold)


def test_point_polygon_test_matches_cv2():
    rng = np.random.default_rng(0)
    n_checked = 0
    for seed in range(6):
        m = build_mask(dict(mask="noisy", mask_hw=(96, 128), seed=seed), None)
        tissue, holes = oc.mask_to_contours(m, tissue_area_thresh=0.0)
        for scale in (1.0, 7.3, 16.0, 39.0625):
            for c in (tissue + [h for hs in holes for h in hs])[:12]:
                cs = oc.scale_contour(c, scale, scale)
                pts = cs.reshape(-1, 2)
                x0, y0, w, h = oc.bounding_rect(cs)
                assert (x0, y0, w, h) == cv2.boundingRect(cs)
                qx = np.concatenate([pts[:, 0], pts[:, 0] + 1, pts[:, 0] - 1, rng.integers(x0 - 3, x0 + w + 3, 64)])
                qy = np.concatenate([pts[:, 1], pts[:, 1], pts[:, 1] + 1, rng.integers(y0 - 3, y0 + h + 3, 64)])
                got = oc.point_polygon_test(cs, qx, qy)
                want = np.array([cv2.pointPolygonTest(cs, (int(a), int(b)), False) for a, b in zip(qx, qy)])
                assert np.array_equal(got.astype(np.int64), want.astype(np.int64))
                n_checked += qx.size
    assert n_checked > 10000


def test_scale_contour_matches_float32_rule():
    rng = np.random.default_rng(1)
    c = rng.integers(0, 1024, (500, 1, 2)).astype(np.int32)
    for s in rng.uniform(1, 100, 50):
        s = float(s)  # the reference passes a Python float (services/extraction.py:37-38)
        f = c.astype(np.float32)
        f[:, :, 0] *= s
        f[:, :, 1] *= s
        assert np.array_equal(oc.scale_contour(c, s, s), f.astype(np.int32))
        # the rule the C-ABI host code implements: trunc(fl32(v) * fl32(s))
        want = (c.astype(np.float32) * np.float32(s)).astype(np.int32)
        assert np.array_equal(oc.scale_contour(c, s, s), want)


def test_geometry_rounding():
    g = oc.prepare_geometry(src_mag=40, target_mag=20, patch_size=256, step_size=None, downsamples=[1.0, 4.0])
    assert (g.level, g.read_w, g.patch_size_src, g.step_src, g.patch_size_level0) == (0, 512, 512, 512, 512)
    g = oc.prepare_geometry(src_mag=40, target_mag=10, patch_size=224, step_size=112, downsamples=[1.0, 4.0, 16.0])
    assert (g.level, g.read_w, g.patch_size_src, g.step_src) == (1, 224, 896, 448)
    with pytest.raises(ValueError):
        oc.prepare_geometry(src_mag=20, target_mag=40, patch_size=256, step_size=None, downsamples=[1.0])
