"""One encoder forward chunk (127 patches) twice -- a small target for `ncu -k regex:gemm_tcgen05 -s 53 -c 4` (layer-1 GEMMs of the
second chunk).  Options key=int are passed to ap_set_option (e.g. fold_ln=0); model=<name> picks the encoder (vit_b_16 default,
dinov2_large / dinov2_giant with their patch sizes 224 / 512)."""
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

model, chunk = "vit_b_16", 127
for kv in sys.argv[1:]:
    if kv.startswith("model="):
        model = kv.split("=")[1]
    elif kv.startswith("chunk="):
        chunk = int(kv.split("=")[1])
    else:
        Context.get(0).set_option(kv.split("=")[0], int(kv.split("=")[1]))
P = 256
if model.startswith("dinov2"):
    from oracle import dinov2_hf  # noqa: E402
    P = 512 if model == "dinov2_giant" else 224
    ext = B200FeatureExtractor(model, dinov2_hf.dinov2_state_dict(model, seed=1), input_patch=P, max_batch=chunk)
else:
    ext = B200FeatureExtractor(model, vit_state_dict(model, seed=1), max_batch=chunk)
wsi = SyntheticWSI(make_spec(8192, 8192, 3))
rng = np.random.default_rng(0)
rows = torch.from_numpy(np.concatenate([rng.integers(0, 8192 - P, (chunk, 2)), np.full((chunk, 2), P), np.zeros((chunk, 1))], 1).astype(np.int32)).cuda()
for _ in range(2):
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows)
torch.cuda.synchronize()
