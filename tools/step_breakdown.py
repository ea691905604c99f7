"""Per-GEMM-shape timing inside a real encoder step (CUDA events per launch): which GEMM loses what against the sustained peak."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec  # noqa: E402
from oracle.weights import vit_state_dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vit_b_16"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 127
n = chunk * 16
for kv in sys.argv[3:]:                      # library options key=int (ap_set_option), e.g. fold_ln=0
    Context.get(0).set_option(kv.split("=")[0], int(kv.split("=")[1]))
ext = B200FeatureExtractor(name, vit_state_dict(name, seed=1), max_batch=chunk)
wsi = SyntheticWSI(make_spec(30000, 30000, 3))
rng = np.random.default_rng(0)
rows = torch.from_numpy(np.concatenate([rng.integers(0, 30000 - 256, (n, 2)), np.full((n, 2), 256), np.zeros((n, 1))], 1).astype(np.int32)).cuda()
out = torch.empty((n, ext.embedding_dim), dtype=torch.float32, device="cuda")
ctx = Context.get(0)
for _ in range(3):
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows, out=out)
torch.cuda.synchronize()
ctx.profile(True, ["gemm"])
for _ in range(3):
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows, out=out)
shapes = ctx.profile_read_gemm_shapes()
ctx.profile(False)
M = chunk * 197
EPI = {0: "bias_f16", 1: "gelu_f16", 2: "resid_f32", 3: "bias_f32", 4: "swiglu"}
tot = 0.0
for (N, K, epi), (ms, cnt) in sorted(shapes.items(), key=lambda kv: -kv[1][0]):
    us = ms / cnt * 1000
    print(f"N={N:5d} K={K:5d} {EPI[epi]:10s} launches {cnt:5d}  avg {us:8.1f} us  share-of-gemm-time {ms:8.2f} ms"
          f"  (full-M launches ~{2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s if M={M})")
    tot += ms
print("total gemm ms per 3 steps:", round(tot, 2))
