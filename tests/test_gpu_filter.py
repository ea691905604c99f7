"""GPU: ap_filter_patches (the --no-fast-mode content filter) against the reference's golden rows and the oracle's counts."""
from pathlib import Path

import numpy as np
import pytest

from oracle import patch_filter as pf
from tests.cases import COORD_CASES, FILTER_CASES, build_mask, case_spec

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _wsi(case):
    from atlaspatch_b200.slide import SyntheticWSI

    return SyntheticWSI(case_spec(case))


@pytest.mark.parametrize("fc", FILTER_CASES, ids=lambda c: c["name"])
def test_filter_matches_reference_golden(fc):
    import torch

    from atlaspatch_b200.extraction import filter_patches
    from atlaspatch_b200.synthetic import render_region_host

    case = {c["name"]: c for c in COORD_CASES}[fc["coords"]]
    wsi = _wsi(case)
    cand = np.load(GOLDEN / f"coords_{case['name']}.npz")["coords"]
    want = np.load(GOLDEN / f"filter_{fc['name']}.npz")["coords"]
    kept, kept_dev, counts = filter_patches(wsi.device_image, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(cand).cuda(),
                                            patch_size=case["patch"], black_threshold=fc["black"], white_threshold=fc["white"],
                                            return_counts=True)
    assert np.array_equal(kept, want)
    assert np.array_equal(kept_dev.cpu().numpy(), want)
    # per-candidate pixel counts, bit exact against the oracle on a spread of candidates
    idx = np.linspace(0, len(cand) - 1, 24).astype(int)
    _, ocounts = pf.filter_rows(lambda x, y, w, h: render_region_host(wsi.spec, x, y, w, h), cand[idx], case["patch"],
                                fc["black"], fc["white"])
    assert np.array_equal(counts[idx], ocounts)


def test_filter_edges_overhang_unaligned_pitch_and_empty():
    """Candidates hanging over the right/bottom border (zeros outside, like IWSI.extract), odd patch size, a slide whose
    pitch is not 16-byte aligned (scalar staging path), thresholds at their extremes, and n = 0."""
    import torch

    from atlaspatch_b200.extraction import filter_patches
    from atlaspatch_b200.synthetic import make_spec, render_region_host

    spec = make_spec(1500, 1100, seed=21)
    from atlaspatch_b200.slide import SyntheticWSI

    wsi = SyntheticWSI(spec)
    img = wsi.device_image
    rng = np.random.default_rng(9)
    for P, read in ((256, 256), (61, 61), (128, 256), (33, 66), (64, 192), (48, 192), (20, 160), (100, 267), (64, 171), (256, 683)):
        scale = read // P
        xs = np.concatenate([rng.integers(0, spec.width - 1, 40), [spec.width - 1, spec.width - read // 2, 0, 7]])
        ys = np.concatenate([rng.integers(0, spec.height + 40, 40), [spec.height - 1, 0, spec.height - read // 3, spec.height + 500]])
        rows = np.stack([xs, ys, np.full_like(xs, read), np.full_like(xs, read), np.zeros_like(xs)], 1).astype(np.int32)
        for bt, wt in ((142, 6), (50, 15), (1, 1), (255, 255)):
            want, wcounts = pf.filter_rows(lambda x, y, w, h: render_region_host(spec, x, y, w, h), rows, P, bt, wt)
            kept, _, counts = filter_patches(img, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(rows).cuda(), patch_size=P,
                                             black_threshold=bt, white_threshold=wt, return_counts=True)
            assert np.array_equal(counts, wcounts), (P, scale, bt, wt)
            assert np.array_equal(kept, want)
        # same slide with a pitch that breaks 16-byte alignment -> scalar staging loads
        pitch2 = wsi.pitch + 3
        img2 = torch.zeros((spec.height, pitch2), dtype=torch.uint8, device="cuda")
        img2[:, :wsi.pitch] = img
        want, wcounts = pf.filter_rows(lambda x, y, w, h: render_region_host(spec, x, y, w, h), rows, P, 142, 6)
        kept, _, counts = filter_patches(img2, wsi.w, wsi.h, pitch2, torch.from_numpy(rows).cuda(), patch_size=P,
                                         black_threshold=142, white_threshold=6, return_counts=True)
        assert np.array_equal(counts, wcounts) and np.array_equal(kept, want)
    kept, kept_dev = filter_patches(img, wsi.w, wsi.h, wsi.pitch, torch.empty((0, 5), dtype=torch.int32, device="cuda"), patch_size=256)
    assert kept.shape == (0, 5) and kept_dev.shape[0] == 0
    from atlaspatch_b200._lib import AtlasB200Error

    with pytest.raises(AtlasB200Error):  # a read smaller than the patch (up-sampling) does not occur in the reference and is refused
        bad = np.array([[0, 0, 128, 128, 0]], dtype=np.int32)
        filter_patches(img, wsi.w, wsi.h, wsi.pitch, torch.from_numpy(bad).cuda(), patch_size=256)


def test_service_no_fast_mode_matches_reference_golden():
    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide

    fc = FILTER_CASES[0]
    case = {c["name"]: c for c in COORD_CASES}[fc["coords"]]
    wsi = _wsi(case)
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=case["patch"], target_magnification=case["target_mag"],
                                                      step_size=case["step"], tissue_threshold=case["tissue_thresh"],
                                                      fast_mode=False, black_threshold=fc["black"], white_threshold=fc["white"]))
    res = svc.extract(wsi, build_mask(case, wsi.spec), slide=Slide(Path(wsi.path), mpp=wsi.spec.mpp))
    want = np.load(GOLDEN / f"filter_{fc['name']}.npz")["coords"]
    assert res.num_patches == want.shape[0] and np.array_equal(res.coords, want)
    assert np.array_equal(res.coords_device.cpu().numpy(), want)
