"""CPU tests: the host-side segmentation steps (a2, a3, a5) against the reference's own functions when it is importable
(build container), and their invariants everywhere."""
import numpy as np
import pytest
from PIL import Image

from atlaspatch_b200 import segmentation as seg
from oracle.refimport import import_reference, reference_available


def test_mask_roundtrip_and_threshold():
    rng = np.random.default_rng(0)
    m = (rng.random((1024, 1024)) > 0.5).astype(np.float32)
    small = seg.resize_mask(m, (768, 1024))
    assert small.shape == (768, 1024) and set(np.unique(small)) <= {0.0, 1.0}
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    out, orig = seg.resize_for_sam(img)
    assert out.shape == (1024, 1024, 3) and orig == (600, 800)
    same, orig2 = seg.resize_for_sam(out)
    assert same is out and orig2 == (1024, 1024)
    t = seg.cap_thumbnail(Image.fromarray(rng.integers(0, 256, (3750, 5000, 3), dtype=np.uint8)))
    assert t.size == (1024, 768)


@pytest.mark.skipif(not reference_available(), reason="reference source only exists in the build container")
def test_host_steps_match_reference_functions():
    import_reference()
    from atlas_patch.services.segmentation import _SAM2Predictor

    ref = _SAM2Predictor.__new__(_SAM2Predictor)   # helpers only; no model is built
    ref.input_size = 1024
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (768, 1000, 3), dtype=np.uint8)
    a, sa = seg.resize_for_sam(img)
    b, sb = ref._resize_input_for_sam(img)
    assert sa == sb and np.array_equal(a, b)
    m = (rng.random((1024, 1024)) > 0.3).astype(np.float32)
    for shape in [(768, 1000), (512, 512), (313, 1024)]:
        assert np.array_equal(seg.resize_mask(m, shape), ref._resize_mask(m, shape))


def test_service_contract_with_a_stub_predictor():
    class W:  # minimal IWSI-like object
        def get_thumbnail_at_power(self, *, power, interpolation):
            return Image.fromarray(np.full((300, 500, 3), 200, np.uint8))

    svc = seg.B200SegmentationService(lambda im: np.where(np.arange(1024)[None, :] < 512, 1.0, -1.0) * np.ones((1024, 1)))
    mask = svc.segment_thumbnail(W())
    assert mask.data.shape == (300, 500) == mask.source_shape and mask.data.dtype == np.float32
    assert mask.data[:, :240].all() and not mask.data[:, 260:].any()
    with pytest.raises(NotImplementedError):
        seg.B200SegmentationService().segment_thumbnail(W())
