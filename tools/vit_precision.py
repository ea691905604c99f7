"""Feature error of the ViT-B/16 CUDA path vs the fp32 oracle over many patch kinds, per precision setting
(precise_layers x which GEMMs of those layers carry hi/lo split weights)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, render_region_host  # noqa: E402
from oracle import vit as ov  # noqa: E402
from oracle.weights import vit_state_dict  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
# optional: "shipped" = only the library default (precise_layers 1, all four GEMMs), then attention variants to compare
# (ap_set_option "attn_variant": 0 one-pass softmax, 512 two-pass, 32 the round-1 pipeline)
shipped_only = len(sys.argv) > 2 and sys.argv[2] == "shipped"
attn_variants = [int(v) for v in sys.argv[3:]] or [0]
spec = make_spec(6000, 5000, seed=41)
wsi = SyntheticWSI(spec)
rng = np.random.default_rng(1)
xy = [(int(rng.integers(-200, spec.width - 56)), int(rng.integers(-200, spec.height - 56))) for _ in range(n)]
xy[:4] = [(-128, -128), (spec.width - 64, 100), (100, spec.height - 40), (spec.width - 30, spec.height - 30)]   # mostly black
coords = torch.tensor([[x, y, 256, 256, 0] for x, y in xy], dtype=torch.int32, device="cuda")
ctx = Context.get(0)
import os  # noqa: E402
for kv in filter(None, os.environ.get("AP_OPTS", "").split(",")):     # e.g. AP_OPTS=fold_ln=0,precise_aw_layers=1
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    print("option", kv)
for seed in (1234, 7):
    sd = vit_state_dict("vit_b_16", seed=seed)
    want = ov.extract_features([render_region_host(spec, x, y, 256, 256) for x, y in xy], sd, "vit_b_16")
    settings = ((1, 15), (2, 15)) if shipped_only else ((0, 15), (1, 15), (1, 5), (1, 10), (1, 1), (1, 4), (1, 3), (1, 12), (2, 15))
    for pl, mask in settings:
        for av in attn_variants:
            ctx.set_option("precise_mask", mask)
            ctx.set_option("attn_variant", av)
            ext = B200FeatureExtractor("vit_b_16", sd, max_batch=127, precise_layers=pl)
            got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, coords).cpu().numpy()
            rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
            print(f"seed {seed} precise_layers {pl} mask {mask:2d} attn_variant {av} ({n} patches): mean {rel.mean():.2e} p90 {np.quantile(rel, 0.9):.2e} "
                  f"p99 {np.quantile(rel, 0.99):.2e} max {rel.max():.2e} (black-ish rows {np.array2string(rel[:4], precision=5)})", flush=True)
            ext.cleanup()
    ctx.set_option("precise_mask", 15)
    ctx.set_option("attn_variant", 0)
