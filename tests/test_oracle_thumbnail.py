"""CPU: the thumbnail oracle (oracle/thumbnail.py) against the reference's own IWSI.get_thumbnail_at_power output (goldens made by
tests/golden/make_golden.py) and against cv2.resize(INTER_AREA) itself, for integer factors and for the general case
(level size not a multiple of the factor: OpenCV's fractional cell weights, restated operation by operation)."""
import numpy as np
import pytest

from oracle import thumbnail as ot
from tests.cases import THUMB_CASES, THUMB_GENERAL_CASES, case_spec


def _level0(case):
    from atlaspatch_b200.synthetic import render_region_host

    spec = case_spec(case)
    return render_region_host(spec, 0, 0, spec.width, spec.height)


@pytest.mark.parametrize("case", THUMB_CASES + THUMB_GENERAL_CASES, ids=lambda c: c["name"])
def test_oracle_reproduces_reference_thumbnails(case, golden_dir):
    from atlaspatch_b200.geometry import infer_mag

    gold = np.load(golden_dir / f"thumb_{case['name']}.npz")["thumb"]
    ds = ot.thumbnail_factor(infer_mag(case["mpp"]), 1.25)
    got = ot.thumbnail_reference_rule(_level0(case), ds)
    assert got.shape == gold.shape and np.array_equal(got, gold)


@pytest.mark.parametrize("W,H,ow,oh", [(1000, 777, 63, 49), (515, 300, 32, 19), (999, 1001, 125, 63), (641, 480, 40, 30), (640, 480, 40, 60)])
def test_general_area_restatement_equals_cv2(W, H, ow, oh):
    import cv2

    src = np.random.default_rng(W + H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = cv2.resize(src, (ow, oh), interpolation=cv2.INTER_AREA)
    sx, sy = W / ow, H / oh
    if sx == int(sx) and sy == int(sy):          # both integral (here 16 x 8): cv2's ResizeAreaFast, sum * (1.f / (fx fy))
        fx, fy = int(sx), int(sy)
        s = src.reshape(oh, fy, ow, fx, 3).astype(np.uint32).sum(axis=(1, 3))
        got = np.clip(np.rint(s.astype(np.float32) * np.float32(1.0 / (fx * fy))), 0, 255).astype(np.uint8)
    else:
        got = ot.area_resize_general(src, ow, oh)
    assert np.array_equal(got, want)


def test_thumbnail_size_uses_python_rounding():
    assert ot.thumbnail_size(5000, 3000, 48.0) == (104, 62)      # 104.17 -> 104, 62.5 -> 62 (half to even, iwsi.py:302-303)
    assert ot.thumbnail_size(8, 8, 16.0) == (1, 1)               # 0.5 -> 0 -> max(1, .)
