"""GPU: the slide scheduler (atlaspatch_b200/runner.py) with the real device services: .synth descriptors -> slides in HBM ->
thumbnail -> (SAM2 double: the slide's ground-truth lattice) -> coordinate kernels -> H5 -> encoder -> features in the H5,
then the skip-existing second run; plus an ordinary image file through ArrayWSI."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class TruthSegmentation:
    """segment_thumbnail double: the synthetic slide's exact tissue lattice at thumbnail resolution (SAM2 itself is covered by
    tests/test_gpu_sam2.py); it still pulls the 1.25x thumbnail through the device kernel like the real service."""

    def segment_thumbnail(self, wsi):
        from atlaspatch_b200.segmentation import Mask
        from atlaspatch_b200.synthetic import truth_mask

        thumb = wsi.get_thumbnail_at_power(power=1.25)
        m = truth_mask(wsi.spec)
        assert (thumb.height, thumb.width) == m.shape
        return Mask(data=m, source_shape=m.shape)


def test_runner_on_device_slides(tmp_path):
    from atlaspatch_b200 import storage
    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.ref_backend import write_synth_descriptor
    from atlaspatch_b200.runner import B200Runner, RunConfig, patch_h5_path
    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.slide import DeviceWSILoader
    from atlaspatch_b200.synthetic import make_spec, render_region_host
    from oracle import coords as oc
    from oracle import vit as ov
    from oracle.weights import vit_state_dict

    specs = {"s_a": (4096, 3072, 21), "s_b": (3072, 3072, 22), "s_c": (2048, 4096, 23)}
    slides = [Slide(write_synth_descriptor(tmp_path / f"{n}.synth", w, h, seed), mpp=0.5) for n, (w, h, seed) in specs.items()]
    (tmp_path / "broken.synth").write_text("{not json")
    slides.append(Slide(tmp_path / "broken.synth", mpp=0.5))
    ecfg = ExtractionConfig(patch_size=256, target_magnification=20, step_size=256)
    cfg = RunConfig(output_root=tmp_path / "out", extraction=ecfg, feature_extractors=["vit_test_tiny"], save_images=False)
    sd = vit_state_dict("vit_test_tiny", seed=5)
    runner = B200Runner(cfg, segmentation=TruthSegmentation(), extraction=B200PatchExtractionService(ecfg), wsi_loader=DeviceWSILoader(),
                        extractor_builders={"vit_test_tiny": lambda: B200FeatureExtractor("vit_test_tiny", sd, max_batch=16)})
    results, failures = runner.run(slides)
    assert [r.slide.stem for r in results] == ["s_a", "s_b", "s_c"] and [s.stem for s, _ in failures] == ["broken"]
    h5 = storage._h5py()
    for r, (n, (w, h, seed)) in zip(results, specs.items()):
        spec = make_spec(w, h, seed)
        from atlaspatch_b200.synthetic import truth_mask

        want = oc.coords_from_mask(truth_mask(spec), level0_wh=(w, h), src_mag=20, target_mag=20, patch_size=256, step_size=256,
                                   tissue_thresh=0.0)
        with h5.File(str(patch_h5_path(r.slide, cfg)), "r") as f:
            coords, feats = f["coords"][...], f["features/vit_test_tiny"][...]
            assert int(f.attrs["num_patches"]) == want.shape[0] and f.attrs["filename"] == f"{n}.synth"
        assert np.array_equal(coords, want) and feats.shape == (want.shape[0], 256)
        idx = np.linspace(0, want.shape[0] - 1, 4).astype(int)
        ref = ov.extract_features([render_region_host(spec, int(x), int(y), 256, 256) for x, y in want[idx, :2]], sd, "vit_test_tiny")
        rel = np.linalg.norm(feats[idx] - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert rel.max() < 1e-3, rel
    results2, failures2 = B200Runner(cfg, segmentation=None, extraction=None, wsi_loader=DeviceWSILoader(),
                                     extractor_builders={"vit_test_tiny": lambda: 1 / 0}).run(slides[:3])
    assert results2 == [] and failures2 == []          # complete outputs: nothing is opened, built or embedded again


def test_array_wsi_from_an_image_file(tmp_path):
    """An ordinary RGB image (the reference's ImageWSI case) uploaded to HBM: extract / thumbnail equal the host pixels / cv2."""
    import cv2
    from PIL import Image

    from atlaspatch_b200.slide import ArrayWSI

    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, (1000, 1300, 3), dtype=np.uint8)
    p = tmp_path / "photo.png"
    Image.fromarray(arr).save(p)
    wsi = ArrayWSI(p, mpp=1.0)                        # 10x -> thumbnail factor 8: 1300 / 8 = 162.5 -> 162, 1000 / 8 = 125
    assert wsi.get_size() == (1300, 1000) and wsi.mag == 10
    assert np.array_equal(wsi.extract((100, 200), 0, (64, 32)), arr[200:232, 100:164])
    edge = wsi.extract((1280, 990), 0, (64, 32))
    assert np.array_equal(edge[:10, :20], arr[990:, 1280:]) and not edge[10:].any() and not edge[:, 20:].any()
    thumb = np.asarray(wsi.get_thumbnail_at_power(power=1.25))
    assert np.array_equal(thumb, cv2.resize(arr, (162, 125), interpolation=cv2.INTER_AREA))
    with pytest.raises(ValueError):
        ArrayWSI(p, mpp=None)
