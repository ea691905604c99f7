"""Time ap_filter_patches on the BASELINE configs[1] slide (80000x60000, 256 px patches) and report HBM throughput."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402
from atlaspatch_b200.extraction import extract_coords, filter_patches  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, truth_mask  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (80000, 60000)
spec = make_spec(W, H, 0)
wsi = SyntheticWSI(spec)
img = wsi.device_image
mask = np.ones((H // 256, W // 256), dtype=np.float32)  # every grid cell is a candidate: the filter does all the work
_, rows = extract_coords(mask, level0_wh=(W, H), src_mag=20, target_mag=20, patch_size=256, return_device=True)
n = rows.shape[0]
ctx = Context.get(0)
for _ in range(2):
    kept, _ = filter_patches(img, W, H, wsi.pitch, rows, patch_size=256)
ctx.profile(True, ["coords"])
reps = 5
for _ in range(reps):
    filter_patches(img, W, H, wsi.pitch, rows, patch_size=256)
torch.cuda.synchronize()
ms, cnt = ctx.profile_read()["coords"]
ctx.profile(False)
t = ms / cnt
print(json.dumps({"candidates": n, "kept": int(kept.shape[0]), "ms": t, "GB/s": n * 256 * 256 * 3 / t / 1e6}))
