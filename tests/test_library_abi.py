"""CPU tests: the C-ABI library builds/loads and exports every symbol include/atlaspatch_b200.h declares;
host-side logic (geometry, contour flattening) agrees with the oracle.  No compute calls (no GPU here)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from atlaspatch_b200 import build as b
    from atlaspatch_b200._lib import load_library

    b.build()
    return load_library()


def test_every_header_symbol_is_exported(lib):
    from atlaspatch_b200._lib import EXPORTED_SYMBOLS

    header = (ROOT / "include" / "atlaspatch_b200.h").read_text()
    declared = set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(EXPORTED_SYMBOLS)
    assert lib.ap_version() == 2
    # the ctypes struct declarations (and INTEGRATION.md's copy of them) match the structs compiled into the library
    import ctypes as C

    from atlaspatch_b200._lib import Sam2Desc, VitDesc

    assert lib.ap_sizeof(b"ap_vit_desc") == C.sizeof(VitDesc) == 92
    assert lib.ap_sizeof(b"ap_sam2_desc") == C.sizeof(Sam2Desc)
    assert lib.ap_sizeof(b"nope") == -1
    doc = (ROOT / "INTEGRATION.md").read_text()
    fields = re.search(r"class VitDesc\(C.Structure\):.*?_fields_ = \[(.*?)\]\s", doc, re.S).group(1)
    assert re.findall(r'\("(\w+)"', fields) == [f[0] for f in VitDesc._fields_]


def test_init_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from atlaspatch_b200._lib import AtlasB200Error, Context

    with pytest.raises(AtlasB200Error) as ei:
        Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_coords_capacity_is_host_only(lib):
    import ctypes as C

    xy = np.array([[10, 10], [10, 500], [700, 500], [700, 10], [5, 5]], dtype=np.int32)
    off = np.array([0, 4, 5], dtype=np.int32)
    cap = lib.ap_coords_capacity(xy.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), 2, 256)
    assert cap == 3 * 2 + 1      # ceil(691/256) x ceil(491/256) + the 1-vertex contour


def test_host_geometry_matches_oracle():
    from atlaspatch_b200 import geometry as g
    from oracle import coords as oc

    for src, tgt, p, s, ds in [(20, 20, 256, None, [1.0]), (40, 20, 256, 128, [1.0, 4.0, 16.0]), (40, 10, 224, 224, [1.0, 4.0]),
                               (40, 5, 255, 100, [1.0, 2.0, 4.0, 8.0]), (20, 20, 1, 1, [1.0])]:
        a = g.prepare_geometry(src_mag=src, target_mag=tgt, patch_size=p, step_size=s, downsamples=ds)
        b = oc.prepare_geometry(src_mag=src, target_mag=tgt, patch_size=p, step_size=s, downsamples=ds)
        assert (a.level, a.read_w, a.read_h, a.patch_size_src, a.step_src, a.patch_size_level0) == \
               (b.level, b.read_w, b.read_h, b.patch_size_src, b.step_src, b.patch_size_level0)
    with pytest.raises(ValueError):
        g.prepare_geometry(src_mag=20, target_mag=40, patch_size=256, step_size=None, downsamples=[1.0])
    assert [g.infer_mag(m) for m in (0.1, 0.17, 0.25, 0.5, 1.0, 2.0)] == [80, 60, 40, 20, 10, 5]


def test_flatten_contours_layout():
    from atlaspatch_b200.extraction import flatten_contours, mask_to_contours, scale_contours
    from oracle import coords as oc
    from tests.cases import build_mask

    mask = build_mask(dict(mask="noisy", mask_hw=(120, 160), seed=3), None)
    t, h = mask_to_contours(mask, tissue_area_thresh=0.0)
    t2, h2 = oc.mask_to_contours(mask, tissue_area_thresh=0.0)
    assert len(t) == len(t2) and all(np.array_equal(a, b) for a, b in zip(t, t2))
    assert [len(x) for x in h] == [len(x) for x in h2]
    ts = scale_contours(t, 13.7, 9.1)
    assert all(np.array_equal(a, oc.scale_contour(b, 13.7, 9.1)) for a, b in zip(ts, t))
    flat = flatten_contours(ts, [scale_contours(x, 13.7, 9.1) for x in h])
    assert flat.n_contours == len(t)
    assert flat.contour_offsets[-1] == flat.contour_xy.shape[0] == sum(c.shape[0] for c in t)
    assert flat.hole_first[-1] == sum(len(x) for x in h) == flat.hole_offsets.shape[0] - 1
    assert flat.contour_xy.dtype == np.int32 and flat.contour_xy.flags.c_contiguous
