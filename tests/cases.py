"""Shared, seeded test cases (used by tests/golden/make_golden.py and by the tests).

Masks are synthetic: the slide's ground-truth tissue lattice, optionally
resampled to the <=1024 px thumbnail the segmentation service hands to extraction, or a noisy
blob mask that exercises many small contours, >10 holes and 1-pixel contours.
"""
from __future__ import annotations

import numpy as np

from atlaspatch_b200.synthetic import SyntheticSlideSpec, make_spec, render_region_host, truth_mask

COORD_CASES = [
    # BASELINE.json configs[0]: 8192^2, 256 px, stride 256
    dict(name="c0_8192_p256", width=8192, height=8192, seed=0, mpp=0.5, patch=256, step=256,
         target_mag=20, tissue_thresh=0.0, mask="truth"),
    dict(name="c0_8192_p256_s128", width=8192, height=8192, seed=1, mpp=0.5, patch=256, step=128,
         target_mag=20, tissue_thresh=0.0, mask="truth"),
    # 40x slide read at 20x: patch_src 512, step_src 512 (T9 geometry)
    dict(name="mag40_8192_p256", width=8192, height=8192, seed=2, mpp=0.25, patch=256, step=256,
         target_mag=20, tissue_thresh=0.0, mask="truth"),
    # 224 px patches (probe shift 56), CLI default tissue threshold of the dataclass (0.01)
    dict(name="p224_12000x9008", width=12000, height=9008, seed=3, mpp=0.5, patch=224, step=224,
         target_mag=20, tissue_thresh=0.01, mask="truth"),
    # noisy mask: hundreds of contours, >10 holes, single-pixel contours, non-integer scale
    dict(name="noisy_10000x7000", width=10000, height=7000, seed=4, mpp=0.5, patch=256, step=256,
         target_mag=20, tissue_thresh=0.0, mask="noisy", mask_hw=(437, 625)),
    dict(name="noisy_p512_s256", width=20000, height=20000, seed=5, mpp=0.5, patch=512, step=256,
         target_mag=20, tissue_thresh=0.0, mask="noisy", mask_hw=(1024, 1024)),
    # BASELINE.json configs[1]: 80000x60000; mask at the capped 1024x768 thumbnail (sx = 78.125)
    dict(name="c1_80000x60000_p256", width=80000, height=60000, seed=0, mpp=0.5, patch=256, step=256,
         target_mag=20, tissue_thresh=0.0, mask="truth_resized", mask_hw=(768, 1024)),
    # odd patch size -> half = 8, shift = 4; tiny patches, many candidates per contour
    dict(name="p17_2048", width=2048, height=2048, seed=6, mpp=0.5, patch=17, step=17,
         target_mag=20, tissue_thresh=0.0, mask="truth"),
    # patch_size 1 -> shift == 0 branch (single centre probe), utils/contours.py:33-35
    dict(name="p1_256", width=256, height=256, seed=7, mpp=0.5, patch=1, step=1,
         target_mag=20, tissue_thresh=0.0, mask="noisy", mask_hw=(64, 64)),
    # empty mask -> zero coords
    dict(name="empty_4096", width=4096, height=4096, seed=8, mpp=0.5, patch=256, step=256,
         target_mag=20, tissue_thresh=0.0, mask="empty", mask_hw=(256, 256)),
]

THUMB_CASES = [
    dict(name="8192", width=8192, height=8192, seed=0, mpp=0.5),          # f = 16
    dict(name="4096x2048_mag40", width=4096, height=2048, seed=1, mpp=0.25),  # f = 32
    dict(name="2000x3008_mag10", width=2000, height=3008, seed=2, mpp=1.0),   # f = 8
]

# a1 beyond power-of-two factors and dividing sizes: ds = 48 (60x slides: 1/48^2 is not exact in fp32, the ADVICE rounding case) and
# level sizes the factor does not divide (cv2's fractional INTER_AREA tables; output round(W/ds) x round(H/ds))
THUMB_GENERAL_CASES = [
    dict(name="4800x2400_mag60", width=4800, height=2400, seed=3, mpp=0.18),     # f = 48, dividing
    dict(name="4100x3001_mag20", width=4100, height=3001, seed=4, mpp=0.5),      # 256 x 188, scales 16.016 / 15.963
    dict(name="5000x3000_mag60", width=5000, height=3000, seed=5, mpp=0.18),     # 104 x 62 (round half even), scales 48.08 / 48.39
    dict(name="4097x4096_mag40", width=4097, height=4096, seed=6, mpp=0.25),     # x fractional, y integral -> general tables on both axes
]

# --no-fast-mode content filter (services/extraction.py:105-119): geometry from COORD_CASES + thresholds.  The synthetic
# tissue's gray level straddles 142 (about 70 % of a tissue patch below it) and the background's saturation straddles 6, so
# the non-default thresholds put hundreds of patches right at the 0.7 decision fraction.
FILTER_CASES = [
    dict(name="noisy_default", coords="noisy_10000x7000", black=50, white=15),      # config defaults (core/config.py:69-70)
    dict(name="truth_b142_w6", coords="c0_8192_p256_s128", black=142, white=6),
    dict(name="mag40_b142_w6", coords="mag40_8192_p256", black=142, white=6),       # 512 px read -> cv2.resize 2:1
    dict(name="p224_b143_w5", coords="p224_12000x9008", black=143, white=5),
]

FEATURE_CASE = dict(weight_seed=1234, slide=dict(width=4096, height=4096, seed=11, mpp=0.5), n=16)


# DINOv2 encoders (BASELINE.json configs[3..4]): patches cut from a synthetic slide at the patch size the config names
DINOV2_CASES = {
    "dinov2_large": dict(weight_seed=4321, slide=dict(width=4096, height=4096, seed=12, mpp=0.5), n=8, patch=224),
    "dinov2_giant": dict(weight_seed=777, slide=dict(width=4096, height=4096, seed=13, mpp=0.5), n=4, patch=512),
}


def dinov2_coords(name: str) -> np.ndarray:
    """Seeded (n, 5) int32 rows (x, y, patch, patch, 0): tissue, edge and background, one hanging over the slide border."""
    c = DINOV2_CASES[name]
    rng = np.random.default_rng(99)
    P, s = c["patch"], c["slide"]
    xy = np.stack([rng.integers(0, s["width"] - P, c["n"]), rng.integers(0, s["height"] - P, c["n"])], 1)
    xy[-1] = (s["width"] - P // 2, s["height"] - P // 3)          # overhang: zeros outside, like IWSI.extract
    return np.concatenate([xy, np.full((c["n"], 2), P), np.zeros((c["n"], 1))], 1).astype(np.int32)


def dinov2_patches(name: str) -> list[np.ndarray]:
    c = DINOV2_CASES[name]
    spec = make_spec(c["slide"]["width"], c["slide"]["height"], c["slide"]["seed"], mpp=c["slide"]["mpp"])
    return [render_region_host(spec, int(x), int(y), c["patch"], c["patch"]) for x, y in dinov2_coords(name)[:, :2]]


def _noisy_mask(h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.zeros((h, w), dtype=bool)
    for _ in range(rng.integers(3, 7)):
        cx, cy = rng.uniform(0.15, 0.85) * w, rng.uniform(0.15, 0.85) * h
        a, b = rng.uniform(0.05, 0.3) * w, rng.uniform(0.05, 0.3) * h
        m |= ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0
    m ^= rng.random((h, w)) < 0.02          # salt-and-pepper: 1-px contours and >10 holes
    for _ in range(14):                      # a few proper holes
        cx, cy, r = rng.uniform(0.2, 0.8) * w, rng.uniform(0.2, 0.8) * h, rng.uniform(0.01, 0.04) * min(w, h)
        m &= ~(((xx - cx) ** 2 + (yy - cy) ** 2) <= r * r)
    return m.astype(np.float32)


def build_mask(case: dict, spec: SyntheticSlideSpec) -> np.ndarray:
    kind = case["mask"]
    if kind == "truth":
        return truth_mask(spec)
    if kind == "truth_resized":
        from PIL import Image

        h, w = case["mask_hw"]
        t = (truth_mask(spec) * 255).astype(np.uint8)
        return np.asarray(Image.fromarray(t).resize((w, h), Image.Resampling.NEAREST), dtype=np.float32) / 255.0
    if kind == "noisy":
        h, w = case["mask_hw"]
        return _noisy_mask(h, w, case["seed"])
    if kind == "empty":
        h, w = case["mask_hw"]
        return np.zeros((h, w), dtype=np.float32)
    raise ValueError(kind)


def case_spec(case: dict) -> SyntheticSlideSpec:
    return make_spec(case["width"], case["height"], case["seed"], mpp=case["mpp"])


def feature_patches() -> list[np.ndarray]:
    """16 seeded 256x256 RGB patches cut from a synthetic slide (tissue, edge and background)."""
    s = FEATURE_CASE["slide"]
    spec = make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"])
    rng = np.random.default_rng(77)
    out = []
    for _ in range(FEATURE_CASE["n"]):
        x, y = int(rng.integers(0, spec.width - 256)), int(rng.integers(0, spec.height - 256))
        out.append(render_region_host(spec, x, y, 256, 256))
    return out


# --patch-size other than 256 with the torchvision ViTs (the preset resizes the PIL patch to 256 with Pillow's BILINEAR first)
VIT_RESIZE_CASES = dict(weight_seed=4242, slide=dict(width=4096, height=4096, seed=14, mpp=0.5), n=4, sizes=(224, 512))


def vit_resize_coords(P: int) -> np.ndarray:
    s = VIT_RESIZE_CASES["slide"]
    rng = np.random.default_rng(500 + P)
    n = VIT_RESIZE_CASES["n"]
    xy = np.stack([rng.integers(0, s["width"] - P, n), rng.integers(0, s["height"] - P, n)], 1)
    xy[-1] = (s["width"] - P // 2, s["height"] - P // 3)          # overhang
    return np.concatenate([xy, np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)


def vit_resize_patches(P: int) -> list[np.ndarray]:
    s = VIT_RESIZE_CASES["slide"]
    spec = make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"])
    return [render_region_host(spec, int(x), int(y), P, P) for x, y in vit_resize_coords(P)[:, :2]]


def sam2_input_image() -> np.ndarray:
    """The 1024 x 1024 uint8 image the segmentation service would hand to SAM2 for the 8192^2 synthetic slide (seed 0)."""
    from atlaspatch_b200.synthetic import sam2_benchmark_image

    return sam2_benchmark_image(8192, 8192, 0)
