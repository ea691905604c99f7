"""Which patches carry the largest feature error of the shipped ViT-B/16 setting?  Prints the worst rows of a survey with what
distinguishes them (position, share of background / black pixels, feature norm) and the error after each encoder layer count."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, render_region_host  # noqa: E402
from oracle import vit as ov  # noqa: E402
from oracle.weights import vit_state_dict  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spec = make_spec(6000, 5000, seed=41)
wsi = SyntheticWSI(spec)
rng = np.random.default_rng(1)
xy = [(int(rng.integers(-200, spec.width - 56)), int(rng.integers(-200, spec.height - 56))) for _ in range(n)]
xy[:4] = [(-128, -128), (spec.width - 64, 100), (100, spec.height - 40), (spec.width - 30, spec.height - 30)]
coords = torch.tensor([[x, y, 256, 256, 0] for x, y in xy], dtype=torch.int32, device="cuda")
sd = vit_state_dict("vit_b_16", seed=1234)
patches = [render_region_host(spec, x, y, 256, 256) for x, y in xy]
want = ov.extract_features(patches, sd, "vit_b_16")
ext = B200FeatureExtractor("vit_b_16", sd, max_batch=127)
got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, coords).cpu().numpy()
rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
order = np.argsort(-rel)
def stats(p):
    c = p[16:240, 16:240].astype(np.float32)
    g = c.mean(axis=2)
    return dict(black=float((g < 5).mean()), white=float((g > 235).mean()), std=float(g.std()), mean=float(g.mean()))
print("worst rows:")
for i in order[:12]:
    print(f"  row {i:4d} xy {xy[i]} rel {rel[i]:.3e} |f| {np.linalg.norm(want[i]):.2f} {stats(patches[i])}")
print("typical rows:")
for i in order[n // 2: n // 2 + 6]:
    print(f"  row {i:4d} xy {xy[i]} rel {rel[i]:.3e} |f| {np.linalg.norm(want[i]):.2f} {stats(patches[i])}")
b = np.array([stats(p)["white"] for p in patches]); s_ = np.array([stats(p)["std"] for p in patches])
print("corr(rel, white share) =", np.corrcoef(rel, b)[0, 1], " corr(rel, pixel std) =", np.corrcoef(rel, s_)[0, 1],
      " corr(rel, 1/|f|) =", np.corrcoef(rel, 1 / np.linalg.norm(want, axis=1))[0, 1])
