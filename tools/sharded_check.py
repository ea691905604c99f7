"""torchrun --nproc-per-node N tools/sharded_check.py: intra-slide sharding on N GPUs (BASELINE.json configs[4] layout).

Every rank renders the same synthetic slide, embeds its contiguous range of the coordinate rows with a tiny encoder and all-gathers
the (N, D) matrix over NCCL; rank 0 also embeds all rows alone and the two matrices must be identical."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200.encoder import B200FeatureExtractor  # noqa: E402
from atlaspatch_b200.services import B200FeatureEmbeddingService, B200PatchExtractionService, ExtractionConfig, Slide  # noqa: E402
from atlaspatch_b200.slide import SyntheticWSI  # noqa: E402
from atlaspatch_b200.synthetic import make_spec, truth_mask  # noqa: E402
from oracle.weights import vit_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = make_spec(8192, 8192, 0)
wsi = SyntheticWSI(spec)
res = B200PatchExtractionService(ExtractionConfig(patch_size=256, target_magnification=20, step_size=128)).extract(
    wsi, truth_mask(spec), slide=Slide(Path(wsi.path), mpp=spec.mpp))
ext = B200FeatureExtractor("vit_test_tiny", vit_state_dict("vit_test_tiny", seed=3), max_batch=64, device=local)
svc = B200FeatureEmbeddingService(ext)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
sharded = svc.embed_features(res, wsi=wsi, sharded=True).features[ext.name].copy()
t1.record(); torch.cuda.synchronize()
ok = True
if rank == 0:
    alone = svc.embed_features(res, wsi=wsi).features[ext.name]
    ok = sharded.shape == alone.shape == (res.num_patches, 256) and np.array_equal(sharded, alone)
    print(f"world {world}: {res.num_patches} rows, sharded == single-rank: {ok}, {t0.elapsed_time(t1):.1f} ms", flush=True)
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
