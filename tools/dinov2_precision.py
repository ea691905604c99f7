"""Feature error of the DINOv2 giant/large CUDA path against the transformers golden features as a function of the number of
leading "precise" layers (ap_vit_desc.precise_layers) and of what is split there (ap_set_option "precise_kind": 0 = weights as
fp16 hi/lo pairs, 1 = the A operands of in_proj / out_proj / mlp.0 as hi/lo pairs + mlp.3's weights), with the throughput.

    python tools/dinov2_precision.py dinov2_giant 0:8 0:20 0:40 1:8 1:16 1:20 1:24        # kind:layers"""
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
levels = [tuple(int(x) for x in (v.split(":") + ["0"])[:3]) for v in sys.argv[2:]] or [(0, 8, 0), (0, 20, 0), (1, 8, 0), (1, 20, 0)]
from atlaspatch_b200._lib import Context  # noqa: E402

ctx = Context.get(0)
case = DINOV2_CASES[name]
g = np.load(Path(__file__).resolve().parents[1] / "tests" / "golden" / f"{name}.npz")
s = case["slide"]
wsi = SyntheticWSI(make_spec(s["width"], s["height"], s["seed"], mpp=s["mpp"]))
sd = dinov2_hf.dinov2_state_dict(name, seed=case["weight_seed"])
rows = torch.from_numpy(dinov2_coords(name)).cuda()
big = torch.from_numpy(np.concatenate([dinov2_coords(name)[:-1]] * 64)[:127 * 2]).cuda()     # throughput: two full chunks
for kind, pl, aw in levels:      # kind : precise layers : leading layers with A AND W split (kind 0 only)
    ctx.set_option("precise_kind", kind)
    ctx.set_option("precise_aw_layers", aw)
    ext = B200FeatureExtractor(name, sd, input_patch=case["patch"], max_batch=127, precise_layers=pl)
    got = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows).cpu().numpy()
    rel = np.linalg.norm(got - g["feats"], axis=1) / np.linalg.norm(g["feats"], axis=1)
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, big)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, big)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} precise_kind={kind} precise_layers={pl} aw_layers={aw}: rel = {np.array2string(rel, precision=5)} max {rel.max():.2e} mean {rel.mean():.2e}"
          f"  {big.shape[0] / e0.elapsed_time(e1) * 1000:.0f} patches/s", flush=True)
    ext.cleanup()
    del ext
    torch.cuda.empty_cache()
ctx.set_option("precise_kind", 0)
ctx.set_option("precise_aw_layers", -1)
