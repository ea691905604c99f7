"""Feature error of the DINOv2 giant/large CUDA path against the transformers golden features as a function of the number of
leading layers whose weights are kept as fp16 hi/lo pairs (ap_vit_desc.precise_layers)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec  # noqa: E402
from oracle import dinov2_hf  # noqa: E402
from tests.cases import DINOV2_CASES, dinov2_coords  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dinov2_giant"
levels = [int(v) for v in sys.argv[2:]] or [0, 1, 8, 40]
case = DINOV2_CASES[name]
g = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / f"{name}.npz")
s = case["slide"]
wsi = SyntheticWSI(make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"]))
sd = dinov2_hf.dinov2_state_dict(name, seed=case["weight_seed"])
rows = torch.from_numpy(dinov2_coords(name)).cuda()
for pl in levels:
    ext = B200FeatureExtractor(name, sd, input_patch=case["patch"], max_batch=16, precise_layers=pl)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    rel = np.linalg.norm(got - g["feats"], axis=1) / np.linalg.norm(g["feats"], axis=1)
    print(f"{name} precise_layers={pl}: rel = {np.array2string(rel, precision=5)} max {rel.max():.2e}", flush=True)
    ext.cleanup()
    del ext
    torch.cuda.empty_cache()
