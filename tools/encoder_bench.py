"""Device-timed throughput of any registered encoder on an HBM-resident synthetic slide (embed_coords fast path).

    python tools/encoder_bench.py dinov2_large|hibou_l|midnight|phikon_v2|... 224 [n_patches] [max_batch] [fast|strict]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec  # noqa: E402

GFLOP = {"vit_b_16": 35.13, "vit_l_16": 123.11, "dinov2_large": 162.02, "dinov2_giant": 598.78}
name, P = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1016
mb = int(sys.argv[4]) if len(sys.argv) > 4 else 127
precision = sys.argv[5] if len(sys.argv) > 5 else "fast"     # "fast" | "strict" (B200FeatureExtractor precision preset)
from atlaspatch_b200.encoder import FAMILY_RECIPES  # noqa: E402

# hub families: same FLOPs as the architecture they share (+ 4 register tokens: 261 / 257 of the per-token work)
GFLOP.update({"midnight": 598.78, "openmidnight": 598.78 * 261 / 257, "phikon_v1": 35.13, "phikon_v2": 123.11, "hibou_l": 162.02 * 261 / 257,
              "hibou_b": 46.7 * 261 / 257, "plip": 8.8, "quilt_b_32": 8.8, "quilt_b_16": 35.2})
if name in FAMILY_RECIPES:
    from oracle.hub_families import state_dict as make_sd
elif name.startswith("dinov2"):
    from oracle.dinov2_hf import dinov2_state_dict as make_sd
else:
    from oracle.weights import vit_state_dict as make_sd
sd = make_sd(name, seed=1)
ext = B200FeatureExtractor(name, sd, input_patch=P, max_batch=mb, precision=precision)
del sd
wsi = SyntheticWSI(make_spec(20000, 20000, 3))
rng = np.random.default_rng(0)
rows = np.concatenate([rng.integers(0, 20000 - P, (n, 2)), np.full((n, 2), P), np.zeros((n, 1))], 1).astype(np.int32)
rows = torch.from_numpy(rows).cuda()
img = wsi.device_image
out = torch.empty((n, ext.embedding_dim), dtype=torch.float32, device="cuda")
ctx = Context.get(0)
for _ in range(2):
    ext.embed_coords(img, wsi.w, wsi.h, wsi.pitch, rows, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    ext.embed_coords(img, wsi.w, wsi.h, wsi.pitch, rows, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ctx.profile(True)
ext.embed_coords(img, wsi.w, wsi.h, wsi.pitch, rows, out=out)
torch.cuda.synchronize()
prof = {k: round(v[0], 2) for k, v in ctx.profile_read().items() if v[1]}
ctx.profile(False)
pps = n / ms * 1000
print(json.dumps({"encoder": name, "patch": P, "patches": n, "chunk": mb, "ms": ms, "patches_per_s": pps,
                  "model_tflops": pps * GFLOP.get(name, 0) / 1000, "per_class_ms": prof}))
