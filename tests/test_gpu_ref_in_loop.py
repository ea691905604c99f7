"""GPU: the UNMODIFIED reference's services with the B200 pieces swapped in at its own seams:

  seam #3  WSIFactory.register / map_extension -> the HBM-resident synthetic backend (atlaspatch_b200/ref_backend.py, a real IWSI subclass)
  seam #2  PatchFeatureExtractorRegistry.register -> B200FeatureExtractor behind FeatureExtractor.extract_batch
  seam #1  the reference's PatchExtractionService.extract vs B200PatchExtractionService.extract on the same slide / mask

Coordinates must be bit-exact, features within 1e-3 relative of the fp32 oracle, thumbnails equal to the inherited cv2 path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rp():
    from oracle import ref_pipeline

    if not ref_pipeline.reference_available():
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref: run oracle/make_ref.sh)")
    ref_pipeline.import_reference()
    return ref_pipeline


@pytest.fixture(scope="module")
def backend(rp):
    from atlas_patch.core.wsi.iwsi import IWSI
    from atlas_patch.core.wsi.wsi_factory import WSIFactory

    from atlaspatch_b200.ref_backend import register_synthetic_backend

    cls = register_synthetic_backend()
    assert issubclass(cls, IWSI) and WSIFactory.detect("x/y/slide.synth") == "synthetic"
    return WSIFactory


@pytest.mark.parametrize("wh", [(8192, 8192), (6000, 4100)])   # 4100 is not a multiple of 16: general INTER_AREA thumbnail
def test_reference_extraction_on_the_device_backend(rp, backend, tmp_path, wh):
    from atlas_patch.core.wsi.iwsi import IWSI

    from atlaspatch_b200.ref_backend import write_synth_descriptor
    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.synthetic import truth_mask

    path = write_synth_descriptor(tmp_path / "slideA.synth", wh[0], wh[1], seed=3, mpp=0.5)
    wsi = backend.load(str(path))                              # reference: WSIFactory.load -> SyntheticBackend(path=..., mpp=None)
    assert wsi.get_size() == wh and wsi.mag == 20 and wsi.metadata_attrs() == {"mpp": 0.5, "magnification": 20}
    thumb_dev = np.asarray(wsi.get_thumbnail_at_power(power=1.25))
    thumb_ref = np.asarray(IWSI.get_thumbnail_at_power(wsi, power=1.25))      # the reference's whole-level read + cv2.resize
    assert thumb_dev.shape == thumb_ref.shape and np.array_equal(thumb_dev, thumb_ref)
    mask = truth_mask(wsi.spec)
    extraction, _ = rp.reference_services(tmp_path / "ref", patch_size=256, target_mag=20, step_size=192, tissue_threshold=0.0)
    res_ref = extraction.extract(wsi, mask, slide=rp.reference_slide(path, mpp=0.5))
    ours = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=192))
    res = ours.extract(wsi, mask, slide=Slide(path, mpp=0.5))
    ref_rows = rp.read_h5(res_ref.h5_path)["coords"]
    assert res.num_patches == res_ref.num_patches > 100 and np.array_equal(res.coords, ref_rows)
    # --no-fast-mode: the reference reads every candidate through wsi.extract (HBM read-back) and filters on the host
    extraction_slow, _ = rp.reference_services(tmp_path / "ref_slow", patch_size=256, target_mag=20, step_size=256, fast_mode=False)
    extraction_slow.cfg.black_threshold, extraction_slow.cfg.white_threshold = 142, 6
    res_slow = extraction_slow.extract(wsi, mask, slide=rp.reference_slide(path, mpp=0.5))
    ours_slow = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=256, fast_mode=False,
                                                            black_threshold=142, white_threshold=6))
    res2 = ours_slow.extract(wsi, mask, slide=Slide(path, mpp=0.5))
    assert np.array_equal(res2.coords, rp.read_h5(res_slow.h5_path)["coords"]) and 0 < res2.num_patches
    wsi.cleanup()


def test_reference_embedding_service_with_the_b200_extractor(rp, backend, tmp_path):
    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.ref_backend import write_synth_descriptor
    from atlaspatch_b200.synthetic import render_region_host, truth_mask
    from oracle import vit as ov
    from oracle.weights import vit_state_dict

    path = write_synth_descriptor(tmp_path / "slideB.synth", 4096, 4096, seed=11, mpp=0.5)
    wsi = backend.load(str(path))
    sd = vit_state_dict("vit_b_16", seed=1234)
    builder = lambda: B200FeatureExtractor("vit_b_16", sd, max_batch=32, registry_name="b200_vit_b_16")   # noqa: E731
    extraction, embedding = rp.reference_services(tmp_path / "out", patch_size=256, target_mag=20, extractors={"b200_vit_b_16": builder},
                                                  feature_batch=32, device="cuda")
    res = extraction.extract(wsi, truth_mask(wsi.spec), slide=rp.reference_slide(path, mpp=0.5))
    assert res.num_patches > 40
    failures = embedding.embed_all([res], wsi_loader=type("L", (), {"open": staticmethod(lambda slide: backend.load(str(slide.path)))})())
    assert failures == []
    got = rp.read_h5(res.h5_path)
    feats = got["features"]["b200_vit_b_16"]
    assert feats.shape == (res.num_patches, 768) and feats.dtype == np.float32
    idx = np.linspace(0, res.num_patches - 1, 12).astype(int)
    patches = [render_region_host(wsi.spec, int(x), int(y), 256, 256) for x, y in got["coords"][idx, :2]]
    want = ov.extract_features(patches, sd, "vit_b_16")
    rel = np.linalg.norm(feats[idx] - want, axis=1) / np.linalg.norm(want, axis=1)
    assert rel.max() < 1e-3, rel
    assert res.metadata["feature_sets"] == ["b200_vit_b_16"]
    # a second embed_all finds the features complete and does nothing (services/feature_embedding.py:262-272)
    assert embedding.embed_all([res], wsi_loader=None) == []
