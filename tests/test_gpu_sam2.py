"""GPU parity tests of the SAM2 tissue-mask forward (a4).  Oracle: transformers' Sam2Model in fp32 on the CPU -- an independent
restatement; the reference's own `sam2` package and checkpoint are unavailable offline, so this row is not pinned by the
reference itself."""
import numpy as np
import pytest

from tests.cases import sam2_input_image

pytestmark = pytest.mark.gpu


def _iou(a, b):
    return (a & b).sum() / max((a | b).sum(), 1)


@pytest.fixture(scope="module")
def predictor():
    from atlaspatch_b200.sam2 import B200Sam2Predictor
    from oracle import sam2_hf

    p = B200Sam2Predictor(sam2_hf.sam2_state_dict(0))
    yield p
    p.close()


def test_hiera_t_matches_golden_logits(predictor, golden_dir):
    g = np.load(golden_dir / "sam2_hiera_t_lowres.npz")
    low_ref = g["low"].astype(np.float32)                     # stored as fp16: 1e-3 relative
    logits, low = predictor.predict_logits(sam2_input_image(), return_lowres=True)
    assert logits.shape == (1024, 1024) and low.shape == (256, 256) and np.isfinite(logits).all()
    rel = np.linalg.norm(low - low_ref) / np.linalg.norm(low_ref)
    assert rel < 2e-3, rel
    assert abs(int((logits > 0).sum()) - int(g["positives"])) < 2000
    assert _iou(low > 0, low_ref > 0) > 0.995


def test_hiera_t_matches_live_hf_model_stage_by_stage(predictor):
    import torch

    from oracle import sam2_hf

    model = sam2_hf.build_model(sam2_hf.sam2_state_dict(0))
    acts = {}
    bb = model.vision_encoder.backbone
    for i, blk in enumerate(bb.blocks):
        blk.register_forward_hook(lambda m, inp, out, i=i: acts.__setitem__(i, out.detach()))
    img = sam2_input_image()
    up_ref, low_ref = sam2_hf.predict_logits(model, img)
    up, low = predictor.predict_logits(img, return_lowres=True)
    for i in (0, 1, 3, 5, 10, 11):                            # windowed, q-pooled, global-attention and last blocks
        ref = acts[i][0].numpy()
        got = predictor.debug_buffer(f"blk{i}", (ref.shape[0] * ref.shape[1], ref.shape[2])).reshape(ref.shape)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-4, i
    assert np.linalg.norm(low - low_ref) / np.linalg.norm(low_ref) < 2e-3
    assert _iou(up > 0, up_ref > 0) > 0.999


def test_segmentation_service_end_to_end(predictor):
    """slide in HBM -> thumbnail kernel -> host PIL steps -> SAM2 kernels -> mask at thumbnail size -> coordinate kernels."""
    from atlaspatch_b200.extraction import extract_coords
    from atlaspatch_b200.segmentation import B200SegmentationService
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec
    from oracle import coords as oc

    spec = make_spec(8192, 8192, 0)
    wsi = SyntheticWSI(spec)
    mask = B200SegmentationService(predictor.predict_logits).segment_thumbnail(wsi)
    assert mask.data.shape == (512, 512) == mask.source_shape and set(np.unique(mask.data)) <= {0.0, 1.0}
    kw = dict(level0_wh=(spec.width, spec.height), src_mag=20, target_mag=20, patch_size=256, step_size=256, tissue_thresh=0.01)
    got = extract_coords(mask.data, **kw)
    assert np.array_equal(got, oc.coords_from_mask(mask.data, **kw))


def test_batch_prediction_equals_single_predictions(predictor):
    """predict_batch / segment_batch (services/segmentation.py:142-180,216-229): the pipelined batch entry returns, image by image,
    exactly what the single-image entry returns, and the service threads the thumbnails."""
    from atlaspatch_b200.segmentation import B200SegmentationService
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    img = sam2_input_image()
    imgs = np.stack([img, img[::-1].copy(), img[:, ::-1].copy(), img])
    single = [predictor.predict_logits(i) for i in imgs]
    batch = predictor.predict_logits_batch(imgs)
    assert batch.shape == (4, 1024, 1024)
    for a, b in zip(single, batch):
        assert np.array_equal(a, b)
    assert predictor.predict_logits_batch(np.zeros((0, 1024, 1024, 3), np.uint8)).shape == (0, 1024, 1024)
    svc = B200SegmentationService(predictor.predict_logits)
    wsis = [SyntheticWSI(make_spec(4096, 4096, s)) for s in (1, 2, 3)]
    masks = svc.segment_batch(wsis)
    for w, m in zip(wsis, masks):
        one = svc.segment_thumbnail(w)
        assert m.source_shape == one.source_shape == (256, 256) and np.array_equal(m.data, one.data)


def test_hiera_large_matches_live_hf_model():
    """BASELINE.json configs[2]: SAM2 Hiera-L (embed 144, blocks 2/6/36/4, windows 8/4/16/8, global blocks 23/33/43) on the
    1024 x 1024 thumbnail, mask IoU against the fp32 restatement."""
    from atlaspatch_b200.sam2 import HIERA_L, B200Sam2Predictor
    from oracle import sam2_hf

    sd = sam2_hf.sam2_state_dict(1, "large")
    model = sam2_hf.build_model(sd, "large")
    img = sam2_input_image()
    up_ref, low_ref = sam2_hf.predict_logits(model, img)
    pred = B200Sam2Predictor(sd, config=HIERA_L)
    up, low = pred.predict_logits(img, return_lowres=True)
    pred.close()
    rel = np.linalg.norm(low - low_ref) / np.linalg.norm(low_ref)
    print("hiera-L low-res rel-l2", rel, "IoU", _iou(up > 0, up_ref > 0))
    assert rel < 5e-3
    assert _iou(up > 0, up_ref > 0) > 0.998
