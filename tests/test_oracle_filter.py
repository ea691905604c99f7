"""CPU: the content-filter oracle (oracle/patch_filter.py) against cv2 itself (exhaustively) and against the rows the
reference keeps with fast_mode=False (tests/golden/filter_*.npz, made by make_golden.py --filter)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import patch_filter as pf
from tests.cases import COORD_CASES, FILTER_CASES, case_spec

GOLDEN = Path(__file__).parent / "golden"


def test_gray_and_hsv_match_cv2_on_every_colour():
    import cv2

    r, g, b = np.meshgrid(*(np.arange(256, dtype=np.uint8),) * 3, indexing="ij")
    img = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    assert np.array_equal(pf.rgb_to_gray(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
    s, v = pf.rgb_to_sv(img)
    assert np.array_equal(s, hsv[..., 1]) and np.array_equal(v, hsv[..., 2])


def test_halving_matches_cv2_resize():
    import cv2

    rng = np.random.default_rng(3)
    for hw in ((512, 512), (448, 448), (34, 34)):
        a = rng.integers(0, 256, (*hw, 3), dtype=np.uint8)
        assert np.array_equal(pf.halve_bilinear(a), cv2.resize(a, (hw[1] // 2, hw[0] // 2)))


def test_integer_ratio_shrink_matches_cv2_resize():
    import cv2

    rng = np.random.default_rng(4)
    for r in (2, 3, 4, 5, 6, 8):
        a = rng.integers(0, 256, (48 * r, 40 * r, 3), dtype=np.uint8)
        assert np.array_equal(pf.shrink_bilinear(a, r), cv2.resize(a, (40, 48))), r


@pytest.mark.parametrize("sizes", [(683, 256), (384, 256), (320, 256), (512, 256), (768, 256), (300, 224), (1000, 256), (257, 256),
                                   (640, 512), (341, 128)])
def test_general_linear_resize_matches_cv2(sizes):
    import cv2

    n, m = sizes
    img = np.random.default_rng(n).integers(0, 256, (n, n, 3), dtype=np.uint8)
    assert np.array_equal(pf.resize_linear_u8(img, m, m), cv2.resize(img, (m, m)))
    if n % m == 0:
        assert np.array_equal(pf.shrink_bilinear(img, n // m), cv2.resize(img, (m, m)))


def test_predicates_match_reference_functions():
    """utils/image.py restated: same decisions as cv2-based code on random and near-threshold patches."""
    import cv2

    rng = np.random.default_rng(5)
    for i in range(40):
        base = rng.integers(0, 256, 3)
        patch = np.clip(base + rng.integers(-30, 30, (32, 32, 3)), 0, 255).astype(np.uint8)
        bt, wt = int(rng.integers(1, 255)), int(rng.integers(1, 60))
        gray = cv2.cvtColor(patch, cv2.COLOR_RGB2GRAY)
        hsv = cv2.cvtColor(patch, cv2.COLOR_RGB2HSV)
        black = float((gray < bt).mean()) >= 0.7
        white = float(((hsv[..., 1] < wt) & (hsv[..., 2] >= 200)).mean()) >= 0.7
        assert pf.keep_patch(patch, bt, wt) == (not (black or white))


@pytest.mark.parametrize("fc", FILTER_CASES[:3], ids=lambda c: c["name"])
def test_oracle_keeps_what_the_reference_keeps(fc):
    from atlaspatch_b200.synthetic import render_region_host

    case = {c["name"]: c for c in COORD_CASES}[fc["coords"]]
    spec = case_spec(case)
    cand = np.load(GOLDEN / f"coords_{case['name']}.npz")["coords"]
    want = np.load(GOLDEN / f"filter_{fc['name']}.npz")["coords"]
    kept, counts = pf.filter_rows(lambda x, y, w, h: render_region_host(spec, x, y, w, h), cand, case["patch"], fc["black"], fc["white"])
    assert 0 < want.shape[0] < cand.shape[0]
    assert np.array_equal(kept, want)
