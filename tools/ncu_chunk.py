"""One encoder forward chunk (127 patches) twice -- a small target for `ncu -k regex:gemm_tcgen05 -s 53 -c 4` (layer-1 GEMMs of the
second chunk).  Options key=int are passed to ap_set_option (e.g. fold_ln=0)."""
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

for kv in sys.argv[1:]:
    Context.get(0).set_option(kv.split("=")[0], int(kv.split("=")[1]))
ext = B200FeatureExtractor("vit_b_16", vit_state_dict("vit_b_16", seed=1), max_batch=127)
wsi = SyntheticWSI(make_spec(8192, 8192, 3))
rng = np.random.default_rng(0)
rows = torch.from_numpy(np.concatenate([rng.integers(0, 8192 - 256, (127, 2)), np.full((127, 2), 256), np.zeros((127, 1))], 1).astype(np.int32)).cuda()
for _ in range(2):
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows)
torch.cuda.synchronize()
