"""Deterministic synthetic whole-slide images (integer arithmetic only).

The benchmark configs in BASELINE.json are quoted on *synthetic* RGB slides
(8192^2 ... 80000x60000, single pyramid level, mpp 0.5 => 20x).  A slide is a
pure function of ``(x, y, channel, spec)`` built from 32-bit integer hashes and
integer ellipse tests, so the very same pixels can be produced

* on the device, tile by tile, straight into HBM (``ap_synth_render`` in
  ``csrc/synth.cu`` -- the 14.4 GB level-0 image of an 80000x60000 slide never
  exists on the host), and
* on the host, lazily per region, by :func:`render_region_host` below, which is
  what the CPU oracle / CPU baseline read through the reference's ``IWSI.extract``
  contract (reference: atlas_patch/core/wsi/iwsi.py:59-85; out-of-bounds pixels
  are 0, like OpenSlide's ``read_region(...).convert("RGB")`` padding,
  atlas_patch/core/wsi/openslide_wsi.py:184-205).

``tests/test_synthetic.py`` asserts the two generators agree bit for bit.

Tissue geometry lives on a 16-pixel lattice (one thumbnail pixel at 1.25x for a
20x slide): ``tissue(x, y)`` only depends on ``(x >> 4, y >> 4)``, so the ground
truth mask at thumbnail resolution is exact.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

import numpy as np

CELL_SHIFT = 4  # tissue lattice = 16 px
_M1 = np.uint32(0x9E3779B1)
_M2 = np.uint32(0x85EBCA77)
_M3 = np.uint32(0xC2B2AE3D)


@dataclass(frozen=True)
class SyntheticSlideSpec:
    """Geometry of one synthetic slide.

    blobs: rows ``(cx, cy, a, b, c, s)`` -- centre and semi-axes in lattice cells,
           rotation as Q6 fixed point ``(c, s) = round(64*cos, 64*sin)``.
    holes: rows ``(cx, cy, r)`` in lattice cells (circles cut out of the tissue).
    """

    width: int
    height: int
    seed: int = 0
    mpp: float = 0.5
    blobs: tuple = field(default_factory=tuple)
    holes: tuple = field(default_factory=tuple)

    @property
    def cells_wh(self) -> tuple[int, int]:
        return (self.width + 15) >> CELL_SHIFT, (self.height + 15) >> CELL_SHIFT

    def blob_array(self) -> np.ndarray:
        return np.asarray(self.blobs, dtype=np.int32).reshape(-1, 6)

    def hole_array(self) -> np.ndarray:
        return np.asarray(self.holes, dtype=np.int32).reshape(-1, 3)


def make_spec(width: int, height: int, seed: int = 0, *, mpp: float = 0.5,
              n_blobs: int | None = None, n_holes: int | None = None) -> SyntheticSlideSpec:
    """Seeded slide: 3-6 elliptical tissue blobs and >=1 hole."""
    rng = random.Random(0x5EED0000 + seed)
    cw, ch = (width + 15) >> CELL_SHIFT, (height + 15) >> CELL_SHIFT
    nb = n_blobs if n_blobs is not None else rng.randint(3, 6)
    blobs = []
    for i in range(nb):
        if i == 0:  # one dominant piece of tissue, as on a resection slide
            cx, cy = int(cw * 0.42), int(ch * 0.5)
            a, b = max(2, int(cw * 0.30)), max(2, int(ch * 0.36))
        else:
            cx, cy = rng.randint(cw // 8, cw - cw // 8), rng.randint(ch // 8, ch - ch // 8)
            a, b = max(1, rng.randint(cw // 40 + 1, cw // 7 + 1)), max(1, rng.randint(ch // 40 + 1, ch // 7 + 1))
        ang = rng.randrange(0, 360)
        c = int(round(64 * np.cos(np.deg2rad(ang))))
        s = int(round(64 * np.sin(np.deg2rad(ang))))
        blobs.append((cx, cy, a, b, c, s))
    nh = n_holes if n_holes is not None else rng.randint(1, 3)
    holes = []
    for _ in range(nh):
        cx0, cy0, a0, b0 = blobs[0][:4]
        hx = cx0 + rng.randint(-a0 // 2, a0 // 2)
        hy = cy0 + rng.randint(-b0 // 2, b0 // 2)
        r = max(3, rng.randint(min(a0, b0) // 10 + 1, min(a0, b0) // 4 + 2))
        holes.append((hx, hy, r))
    return SyntheticSlideSpec(width, height, seed, mpp, tuple(blobs), tuple(holes))


# ---------------------------------------------------------------------------------------
# host generator (numpy, uint32 wrap-around arithmetic == the CUDA kernel's)
# ---------------------------------------------------------------------------------------
def _mix(u: np.ndarray) -> np.ndarray:
    u = u.astype(np.uint32, copy=True)
    u ^= u >> np.uint32(16)
    u *= np.uint32(0x7FEB352D)
    u ^= u >> np.uint32(15)
    u *= np.uint32(0x846CA68B)
    u ^= u >> np.uint32(16)
    return u


def tissue_cells(spec: SyntheticSlideSpec, cx: np.ndarray, cy: np.ndarray) -> np.ndarray:
    """Boolean tissue indicator for lattice-cell coordinates (broadcasting int arrays)."""
    X = np.asarray(cx, dtype=np.int64)
    Y = np.asarray(cy, dtype=np.int64)
    inside = np.zeros(np.broadcast(X, Y).shape, dtype=bool)
    for (bx, by, a, b, c, s) in spec.blob_array().astype(np.int64):
        dx, dy = X - bx, Y - by
        u = (dx * c + dy * s) >> 6
        v = (dy * c - dx * s) >> 6
        inside |= (u * u * (b * b) + v * v * (a * a)) <= (a * a) * (b * b)
    for (hx, hy, r) in spec.hole_array().astype(np.int64):
        dx, dy = X - hx, Y - hy
        inside &= ~((dx * dx + dy * dy) <= r * r)
    return inside


def truth_mask(spec: SyntheticSlideSpec) -> np.ndarray:
    """Ground-truth tissue mask on the 16-px lattice: float32 (H/16, W/16) in {0,1}."""
    cw, ch = spec.cells_wh
    m = tissue_cells(spec, np.arange(cw)[None, :], np.arange(ch)[:, None])
    return m.astype(np.float32)


def render_region_host(spec: SyntheticSlideSpec, x: int, y: int, w: int, h: int) -> np.ndarray:
    """RGB uint8 (h, w, 3) of the level-0 region at (x, y); out-of-bounds pixels are 0."""
    out = np.zeros((h, w, 3), dtype=np.uint8)
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + w, spec.width), min(y + h, spec.height)
    if x1 <= x0 or y1 <= y0:
        return out
    xs = np.arange(x0, x1, dtype=np.int64)[None, :]
    ys = np.arange(y0, y1, dtype=np.int64)[:, None]
    xu, yu = xs.astype(np.uint32), ys.astype(np.uint32)
    seed = np.uint32(spec.seed & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        hpx = _mix(xu * _M1 + yu * _M2 + seed * _M3)
        gcell = _mix((xu >> np.uint32(3)) * _M1 + (yu >> np.uint32(3)) * _M2 + (seed + np.uint32(1)) * _M3)
    t = tissue_cells(spec, xs >> CELL_SHIFT, ys >> CELL_SHIFT)
    r_t = 150 + (gcell & 63) + (hpx & 15)
    g_t = 60 + ((gcell >> 8) & 63) + ((hpx >> 4) & 15)
    b_t = 130 + ((gcell >> 16) & 63) + ((hpx >> 8) & 15)
    r_b = 232 + (hpx & 7)
    g_b = 232 + ((hpx >> 4) & 7)
    b_b = 232 + ((hpx >> 8) & 7)
    reg = out[y0 - y:y1 - y, x0 - x:x1 - x]
    reg[..., 0] = np.where(t, r_t, r_b).astype(np.uint8)
    reg[..., 1] = np.where(t, g_t, g_b).astype(np.uint8)
    reg[..., 2] = np.where(t, b_t, b_b).astype(np.uint8)
    return out


def sam2_benchmark_image(width: int = 8192, height: int = 8192, seed: int = 0) -> np.ndarray:
    """The 1024 x 1024 uint8 image the segmentation service hands to SAM2 for a synthetic slide: 1.25x thumbnail (exact area mean of
    16 x 16 blocks) -> PIL BILINEAR to 1024^2 (reference: services/segmentation.py:104-110).  Input data of BASELINE.json configs[2]."""
    from PIL import Image

    spec = make_spec(width, height, seed)
    lvl0 = render_region_host(spec, 0, 0, width, height)
    s = lvl0.reshape(height // 16, 16, width // 16, 16, 3).astype(np.uint32).sum(axis=(1, 3))
    thumb = np.clip(np.rint(s.astype(np.float32) * np.float32(1 / 256.0)), 0, 255).astype(np.uint8)
    return np.array(Image.fromarray(thumb).resize((1024, 1024), Image.Resampling.BILINEAR))
