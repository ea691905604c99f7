"""GPU test of the service adapters: mask -> coords -> features for one small synthetic slide, against the oracle."""
from pathlib import Path

import numpy as np
import pytest

from oracle import coords as oc
from oracle import vit as ov
from oracle.weights import vit_state_dict

pytestmark = pytest.mark.gpu


def test_extract_then_embed_matches_oracle():
    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.services import B200FeatureEmbeddingService, B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec, render_region_host, truth_mask

    spec = make_spec(4096, 3072, seed=13)
    wsi = SyntheticWSI(spec)
    mask = truth_mask(spec)
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=256))
    res = svc.extract(wsi, mask, slide=Slide(Path(wsi.path), mpp=spec.mpp))
    want = oc.coords_from_mask(mask, level0_wh=(spec.width, spec.height), src_mag=20, target_mag=20, patch_size=256,
                               step_size=256, tissue_thresh=0.0)
    assert res.num_patches == want.shape[0] > 0 and np.array_equal(res.coords, want)
    assert res.patch_size_level0 == 256

    sd = vit_state_dict("vit_test_tiny", seed=5)
    ext = B200FeatureExtractor("vit_test_tiny", sd, max_batch=16)
    res = B200FeatureEmbeddingService(ext).embed_features(res, wsi=wsi)
    feats = res.features["vit_test_tiny"]
    assert feats.shape == (res.num_patches, 256)
    idx = np.linspace(0, res.num_patches - 1, 6).astype(int)
    ref = ov.extract_features([render_region_host(spec, int(x), int(y), 256, 256) for x, y in res.coords[idx, :2]], sd, "vit_test_tiny")
    rel = np.linalg.norm(feats[idx] - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-3, rel
    ext.cleanup()
