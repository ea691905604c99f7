"""Time the SAM2 forward (H2D + forward + D2H through ap_sam2_predict_host) for Hiera-T / Hiera-L with seeded weights, and its kernel classes."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402
from atlaspatch_b200.sam2 import HIERA_L, HIERA_T, B200Sam2Predictor  # noqa: E402
from oracle import sam2_hf  # noqa: E402
from tests.cases import sam2_input_image  # noqa: E402

img = sam2_input_image()
ctx = Context.get(0)
only = sys.argv[1] if len(sys.argv) > 1 else None   # "tiny" | "large"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for variant, cfg in (("tiny", HIERA_T), ("large", HIERA_L)):
    if only and variant != only:
        continue
    pred = B200Sam2Predictor(sam2_hf.sam2_state_dict(0, variant), config=cfg)
    for _ in range(1 if reps < 2 else 2):
        pred.predict_logits(img)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        pred.predict_logits(img)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1000
    gflop = 210.9 if variant == "tiny" else 1625.4
    print(f"hiera-{variant}: {ms:.1f} ms per 1024x1024 thumbnail = {gflop / ms:.1f} TFLOP/s", flush=True)
    pred.close()
