"""Run every BASELINE.json config once on one GPU and print one summary line each (what ran, parity check, device time).

    python tools/run_configs.py [c0 c1 c2 c3 c4]

c3 / c4 are the per-GPU share of their 8-GPU configs (one 40000^2 slide of the 8; 1/8 of the 100 k patches): slides / row ranges
are independent across ranks (sharding.py), so the 8-GPU run is this workload on every rank plus, for c4, one all-gather.
Parity checks use the oracles (dev tool, not product code)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.services import B200FeatureEmbeddingService, B200PatchExtractionService, ExtractionConfig, Slide  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, render_region_host, truth_mask  # noqa: E402
from tests.cases import COORD_CASES, build_mask  # noqa: E402

which = set(sys.argv[1:]) or {"c0", "c1", "c2", "c3", "c4"}
GOLD = ROOT / "tests" / "golden"


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def capped_mask(spec):
    """The slide's true tissue mask at the <= 1024 px thumbnail size the segmentation service hands to extraction."""
    from PIL import Image

    m = (truth_mask(spec) * 255).astype(np.uint8)
    f = 1024 / max(m.shape)
    if f < 1:
        m = np.asarray(Image.fromarray(m).resize((round(m.shape[1] * f), round(m.shape[0] * f)), Image.Resampling.NEAREST))
    return m.astype(np.float32) / 255.0


def embed_slide(name, weights, W, H, patch, step, seed, sample_rows=None, check=None):
    wsi = SyntheticWSI(make_spec(W, H, seed))
    wsi.device_image
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=patch, target_magnification=20, step_size=step))
    mask = capped_mask(wsi.spec)
    svc.extract(wsi, mask, slide=Slide(Path(wsi.path), mpp=0.5))   # warm (cv2 import, CUDA module load)
    res, t_coords = timed(lambda: svc.extract(wsi, mask, slide=Slide(Path(wsi.path), mpp=0.5)))
    ext = B200FeatureExtractor(name, weights, input_patch=patch, max_batch=127)
    rows = res.coords_device if sample_rows is None else res.coords_device[:sample_rows].contiguous()
    ext.embed_coords(wsi.device_image, W, H, wsi.pitch, rows[:254].contiguous())   # warm
    feats, t = timed(lambda: ext.embed_coords(wsi.device_image, W, H, wsi.pitch, rows))
    out = {"slide": f"{W}x{H}", "encoder": name, "patch": patch, "step": step, "coords": res.num_patches, "coords_ms": round(t_coords, 1),
           "embedded": int(rows.shape[0]), "embed_s": round(t / 1000, 2), "patches_per_s": round(rows.shape[0] / t * 1000)}
    if check is not None:
        idx = np.linspace(0, rows.shape[0] - 1, 4).astype(int)
        patches = [render_region_host(wsi.spec, int(x), int(y), patch, patch) for x, y in res.coords[idx, :2]]
        out["max_rel_err_vs_oracle"] = float(rel_rows(feats[idx].cpu().numpy(), check(patches)).max())
    ext.cleanup()
    return out


if "c0" in which:
    from oracle import coords as oc

    case = COORD_CASES[0]
    spec = make_spec(8192, 8192, 0)
    wsi = SyntheticWSI(spec)
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=256))
    res, t = timed(lambda: svc.extract(wsi, build_mask(case, spec), slide=Slide(Path(wsi.path), mpp=0.5)))
    gold = np.load(GOLD / "coords_c0_8192_p256.npz")["coords"]
    print(json.dumps({"config": "c0", "coords": res.num_patches, "equal_to_reference_list": bool(np.array_equal(res.coords, gold)),
                      "ms_incl_host_contours": round(t, 2)}), flush=True)

if "c1" in which:
    from atlaspatch_b200.weights import vit_state_dict
    from oracle import vit as ov

    sd = vit_state_dict("vit_b_16", seed=1234)
    r = embed_slide("vit_b_16", sd, 80000, 60000, 256, 256, 0, check=lambda p: ov.extract_features(p, sd, "vit_b_16"))
    print(json.dumps({"config": "c1", **r}), flush=True)

if "c2" in which:
    from atlaspatch_b200.sam2 import HIERA_L, B200Sam2Predictor
    from oracle import sam2_hf
    from tests.cases import sam2_input_image

    sd = sam2_hf.sam2_state_dict(0, "large")
    pred = B200Sam2Predictor(sd, config=HIERA_L)
    img = sam2_input_image()
    pred.predict_logits(img)
    t0 = time.perf_counter()
    up = pred.predict_logits(img)
    ms = (time.perf_counter() - t0) * 1000
    up_ref, _ = sam2_hf.predict_logits(sam2_hf.build_model(sd, "large"), img)
    a, b = up > 0, up_ref > 0
    print(json.dumps({"config": "c2", "model": "sam2 hiera-large", "ms_incl_h2d_d2h": round(ms, 1),
                      "mask_iou_vs_transformers_fp32": float((a & b).sum() / max((a | b).sum(), 1))}), flush=True)
    pred.close()

if "c3" in which:
    from oracle import dinov2_hf

    sd = dinov2_hf.dinov2_state_dict("dinov2_large", seed=4321)
    r = embed_slide("dinov2_large", sd, 40000, 40000, 224, 224, 3, check=lambda p: dinov2_hf.extract_features(p, sd, "dinov2_large"))
    print(json.dumps({"config": "c3 (one of the 8 slides = one rank's share)", **r}), flush=True)

if "c4" in which:
    from oracle import dinov2_hf

    sd = dinov2_hf.dinov2_state_dict("dinov2_giant", seed=777)
    # 512 px patches, stride 256: ~100 k patches need a ~83000 x 83000 tissue area; one rank of 8 embeds 12.5 k of them
    r = embed_slide("dinov2_giant", sd, 90000, 80000, 512, 256, 5, sample_rows=12500)
    print(json.dumps({"config": "c4 (1/8 of the rows = one rank's share)", **r}), flush=True)
