"""CPU: host-side logic of the service adapters (no GPU): the reference-style host-read fallback of the embedding service
(services/feature_embedding.py:81-96: read, cv2.resize to the patch size when the read differs) and the level guard of the device paths."""
from pathlib import Path

import numpy as np
import pytest

from atlaspatch_b200.services import (B200FeatureEmbeddingService, ExtractionConfig, ExtractionResult, Slide, _require_level0)


class _HostWSI:                     # no device_image: forces the host-read path
    mag, path = 40, "x.svs"

    def __init__(self):
        self.reads = []

    def extract(self, xy, lv, wh, mode="array"):
        self.reads.append((tuple(xy), lv, tuple(wh)))
        rng = np.random.default_rng(xy[0] * 7 + xy[1])
        return rng.integers(0, 256, (wh[1], wh[0], 3), dtype=np.uint8)


class _Recorder:
    name, embedding_dim, input_patch = "rec", 4, 256

    def __init__(self):
        self.patches = None

    def extract_batch(self, patches, *, batch_size=None):
        self.patches = [np.asarray(p) for p in patches]
        return np.zeros((len(patches), 4), np.float32)


def test_host_fallback_resizes_reads_to_the_patch_size_like_the_reference():
    import cv2

    coords = np.array([[0, 0, 512, 512, 0], [512, 256, 512, 512, 0]], dtype=np.int32)      # 40x slide read for 20x patches
    res = ExtractionResult(slide=Slide(Path("x.svs")), h5_path=None, num_patches=2, coords=coords, patch_size_level0=512)
    wsi, ext = _HostWSI(), _Recorder()
    svc = B200FeatureEmbeddingService(ext, ExtractionConfig(patch_size=256, target_magnification=20))
    out = svc.embed_features(res, wsi=wsi)
    assert out.features["rec"].shape == (2, 4) and wsi.reads == [((0, 0), 0, (512, 512)), ((512, 256), 0, (512, 512))]
    assert all(p.shape == (256, 256, 3) for p in ext.patches)
    want = cv2.resize(_HostWSI().extract((0, 0), 0, (512, 512)), (256, 256))
    assert np.array_equal(ext.patches[0], want)
    # without an extraction config the extractor's own input size is the target
    ext2 = _Recorder()
    B200FeatureEmbeddingService(ext2).embed_features(ExtractionResult(slide=Slide(Path("x.svs")), h5_path=None, num_patches=2,
                                                                      coords=coords, patch_size_level0=512), wsi=_HostWSI())
    assert all(p.shape == (256, 256, 3) for p in ext2.patches)


def test_device_paths_refuse_rows_on_other_pyramid_levels():
    _require_level0(np.zeros((0, 5), np.int32), "x")
    _require_level0(np.array([[0, 0, 256, 256, 0]], np.int32), "x")
    with pytest.raises(NotImplementedError, match="level > 0"):
        _require_level0(np.array([[0, 0, 256, 256, 0], [0, 0, 256, 256, 1]], np.int32), "x")


def test_large_reads_with_a_resizing_extractor_take_the_host_path():
    """A device-resident slide read at 2 x the patch size: the crop-preprocess extractors resample on the device, the families whose CUDA
    preprocess already resizes (DINOv2 / hub encoders) go through the reference's host sequence (read, cv2.resize, extract_batch)."""
    coords = np.array([[0, 0, 448, 448, 0], [448, 0, 448, 448, 0]], dtype=np.int32)

    class _DevWSI(_HostWSI):
        device_image, w, h, pitch = object(), 4096, 4096, 4096 * 3

    class _Resizing(_Recorder):
        input_patch, supports_large_reads = 224, False

        def embed_coords(self, *a, **k):
            raise AssertionError("device path must not be taken for reads larger than the patch")

    res = ExtractionResult(slide=Slide(Path("x.svs")), h5_path=None, num_patches=2, coords=coords, patch_size_level0=448)
    res.coords_device = object()
    wsi, ext = _DevWSI(), _Resizing()
    out = B200FeatureEmbeddingService(ext, ExtractionConfig(patch_size=224, target_magnification=20)).embed_features(res, wsi=wsi)
    assert out.features["rec"].shape == (2, 4) and len(wsi.reads) == 2 and all(p.shape == (224, 224, 3) for p in ext.patches)
