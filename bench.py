#!/usr/bin/env python
"""bench.py -- patches/s embedded (256 px, ViT-B/16) on synthetic 80000x60000 slides, one slide per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this framework (B200, CUDA path)
    python bench.py --impl reference --steps 3 --warmup 1          # the reference's CPU pipeline (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                     # N slides on N GPUs (weak scaling)

A step = one pass of the hot path over one batch of `--batch` patch coordinates of the slide resident in HBM:
fused crop/centre/im2col -> ViT-B/16 forward -> (batch, 768) fp32 features.  Successive steps walk through the
slide's coordinate list, so every step reads different pixels (batch * 150 KB >> L2).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "patches/sec embedded (256px, ViT-B/16)"
UNIT = "patches/s"
WEIGHT_SEED = 1234
# algorithmic FLOPs per 256-px patch, ViT-B/16 @224, 197 tokens: 2 x MAC
GEMM_MAC_PER_TOKEN_LAYER = 768 * 2304 + 768 * 768 + 2 * 768 * 3072
GEMM_FLOP_PER_PATCH = 2 * (197 * GEMM_MAC_PER_TOKEN_LAYER * 12 + 196 * 768 * 768)
ATTN_FLOP_PER_PATCH = 2 * 12 * 12 * 2 * 197 * 197 * 64
MODEL_FLOP_PER_PATCH = GEMM_FLOP_PER_PATCH + ATTN_FLOP_PER_PATCH  # 35.13 GFLOP
PATCH_BYTES = 256 * 256 * 3
# patches per forward chunk: 508 x 197 tokens = 391 CTA-pair row tiles -> 1173 / 3519 / 4692 tiles = 15.9 / 47.6 / 63.4 waves of 74 pairs.
# Chunk sweep of round 2 (gpurun_out/r02/bench36_*, same box): 127: 22.95 k patches/s (e2e 22.9 k), 254: 24.1 k (23.7 k), 508: 24.9 k (24.0 k),
# 762: 24.9 k (23.2 k), 1016: 25.2 k (22.9 k): fewer launch gaps and residual-stream tails per patch; beyond 508 the host-patch (e2e) path
# loses its H2D / compute overlap granularity.
CHUNK = 508


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=2032, help="patches per step (b200 arm); 4 forward chunks of 508")
    ap.add_argument("--chunk", type=int, default=CHUNK, help="patches per forward chunk (workspace size)")
    ap.add_argument("--width", type=int, default=80000)
    ap.add_argument("--height", type=int, default=60000)
    ap.add_argument("--e2e-steps", type=int, default=10, help="steps of the host-buffer (e2e) measurement; each step embeds DIFFERENT patches")
    ap.add_argument("--aux", default="c2,c3,c4", help="comma list of the other BASELINE.json configs measured beside the headline "
                    "(c2 SAM2 Hiera-L, c3 one 40000^2 slide per rank with dinov2_large, c4 the 100k-patch slide with dinov2_giant split "
                    "by row range + all-gather); 'none' to skip")
    ap.add_argument("--c4-cap", type=int, default=12500, help="rows of the C4 slide each rank embeds at most (12 500 = one rank's share "
                    "of 100 k at 8 GPUs; with fewer ranks the run is a bounded sample of the same slide)")
    ap.add_argument("--cpu-sample", type=int, default=96, help="patches timed for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-stride", type=int, default=4,
                    help="time every Nth GEMM launch of the timed region with CUDA events (0 = none: roofline from the extra step only)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=int (ap_set_option), for A/B measurements")
    ap.add_argument("--ref-sample", type=int, default=32, help="patches per step of the reference arm")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def slide_mask_coords_cpu(width, height, seed):
    """Mask the segmentation service would hand to extraction: truth lattice resampled to the <=1024 thumbnail."""
    from PIL import Image

    from atlaspatch_b200.synthetic import make_spec, truth_mask

    spec = make_spec(width, height, seed)
    t = (truth_mask(spec) * 255).astype(np.uint8)
    th, tw = t.shape
    s = min(1.0, 1024.0 / max(th, tw))
    mh, mw = max(1, int(round(th * s))), max(1, int(round(tw * s)))
    if (mh, mw) != (th, tw):
        t = np.asarray(Image.fromarray(t).resize((mw, mh), Image.Resampling.NEAREST))
    return spec, t.astype(np.float32) / 255.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        # "under load" = upper half of the samples (idle samples before/after the region would bias the median)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU pipeline.  Preferred: the UNMODIFIED reference installed under baseline/_ref by
# oracle/make_ref.sh (kind "reference": its PatchFeatureEmbeddingService / PatchFeatureExtractor / H5PatchWriter run as they are);
# fallback when that install is absent: the port in oracle/reference_loop.py (kind "port").
# ---------------------------------------------------------------------------------------------------------
def use_all_host_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms must use the host cores the process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["MKL_NUM_THREADS"] = str(n)
    import torch

    torch.set_num_threads(n)
    return n


class ReferenceArm:
    """One slide's embedding through the reference's own services, `n` coordinate rows at a time."""

    def __init__(self, width, height, seed):
        import tempfile as _tf

        from atlaspatch_b200.weights import vit_state_dict

        self.cores = use_all_host_cores()
        self.spec, mask = slide_mask_coords_cpu(width, height, seed)
        gold = ROOT / "tests" / "golden" / "coords_c1_80000x60000_p256.npz"
        if (width, height, seed) == (80000, 60000, 0) and gold.exists():
            self.coords = np.load(gold)["coords"]     # produced by the reference itself on this exact slide / mask
        else:
            from oracle import coords as oc

            self.coords = oc.coords_from_mask(mask, level0_wh=(width, height), src_mag=20, target_mag=20, patch_size=256,
                                              step_size=256, tissue_thresh=0.0)
        sd = vit_state_dict("vit_b_16", seed=WEIGHT_SEED)
        self.kind = "port"
        self.tmp = _tf.TemporaryDirectory(prefix="ap_ref_")
        try:
            from oracle import ref_pipeline as rp

            if not rp.reference_available():
                raise RuntimeError("reference not installed (baseline/_ref)")
            rp.import_reference()
            self.rp = rp
            self.wsi = rp.host_synthetic_wsi_class()(self.spec)
            _, self.embedding = rp.reference_services(self.tmp.name, patch_size=256, target_mag=20,
                                                      extractors={"vit_b_16": rp.reference_vit_builder(sd, num_workers=4)},
                                                      feature_batch=32, num_workers=4)
            self.extractor = self.embedding.registry.create("vit_b_16")
            self.kind = "reference"
            self.what = ("the unmodified reference (baseline/_ref): PatchFeatureEmbeddingService._embed_with_extractor -> per-row H5 "
                         "coordinate read, IWSI.extract, H5PatchWriter.append_features, PatchFeatureExtractor.extract_batch (new "
                         "DataLoader(num_workers=4) per 32 patches), torchvision vit_b_16 fp32")
        except Exception as e:  # noqa: BLE001
            from atlaspatch_b200.synthetic import render_region_host
            from oracle import reference_loop as rl

            self.model, self.preprocess = rl.build_vit_b_16(sd)
            self.read = lambda x, y, w, h: render_region_host(self.spec, x, y, w, h)  # noqa: E731
            self.what = f"oracle/reference_loop.py (port; the reference install was unavailable: {e})"
        self._n = 0

    def run(self, start, n) -> float:
        rows = self.coords[start:start + n]
        t0 = time.perf_counter()
        if self.kind == "reference":
            self._n += 1
            slide = self.rp.reference_slide(Path(self.tmp.name) / f"step{self._n}.synth", mpp=self.spec.mpp)
            res = self.rp.write_reference_coords(self.embedding, self.wsi, slide, rows, patch_size_level0=256)
            t0 = time.perf_counter()                      # the coordinate file exists before embedding starts, as in the reference
            self.embedding._embed_with_extractor(result=res, wsi=self.wsi, extractor=self.extractor)
            dt = time.perf_counter() - t0
            feats = self.rp.read_h5(res.h5_path)["features"]["vit_b_16"]
            os.remove(res.h5_path)
        else:
            from oracle import reference_loop as rl

            feats = rl.embed_slide_rows(self.model, self.preprocess, self.read, rows, patch_size=256, feature_batch=32, num_workers=4)
            dt = time.perf_counter() - t0
        assert feats.shape == (rows.shape[0], 768)
        return dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    arm = ReferenceArm(args.width, args.height, 0)
    n = args.ref_sample
    last = max(1, len(arm.coords) - n)
    for i in range(args.warmup):
        arm.run((i * n) % last, n)
    t = 0.0
    for i in range(args.steps):
        t += arm.run(((args.warmup + i) * n) % last, n)
    value = args.steps * n / t
    sample = (f"{args.steps} steps x {n} consecutive coordinate rows of the {args.width}x{args.height} slide ({len(arm.coords)} patches "
              f"total) through {arm.what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single synthetic {args.width}x{args.height} RGB slide, 256px patches stride 256, ViT-B/16 "
                               "random-init (seeded), the reference's CPU pipeline on the host cores",
                   "patches_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def host_patches_from_slide(wsi, rows_np, P):
    """The (P, P, 3) uint8 host patches a caller of extract_batch would hold for `rows_np`: gathered on the device in blocks, one
    D2H copy per block (zeros outside the slide, like IWSI.extract)."""
    import torch

    img, W, H = wsi.device_image, wsi.w, wsi.h
    n = int(rows_np.shape[0])
    out = np.empty((n, P, P, 3), dtype=np.uint8)
    blk = 512
    for s0 in range(0, n, blk):
        m = min(blk, n - s0)
        buf = torch.zeros((m, P, P, 3), dtype=torch.uint8, device="cuda")
        for i, (x, y) in enumerate(rows_np[s0:s0 + m, :2].tolist()):
            x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + P, W), min(y + P, H)
            if x1 > x0 and y1 > y0:
                buf[i, y0 - y:y1 - y, x0 - x:x1 - x] = img[y0:y1, x0 * 3:x1 * 3].reshape(y1 - y0, x1 - x0, 3)
        out[s0:s0 + m] = buf.cpu().numpy()
    return [out[i] for i in range(n)]


def _cuda_time(fn):
    import torch

    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


def _max_over_ranks(ms, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(v, world):
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def _capped_mask(spec):
    from PIL import Image

    from atlaspatch_b200.synthetic import truth_mask

    m = (truth_mask(spec) * 255).astype(np.uint8)
    f = 1024 / max(m.shape)
    if f < 1:
        m = np.asarray(Image.fromarray(m).resize((round(m.shape[1] * f), round(m.shape[0] * f)), Image.Resampling.NEAREST))
    return m.astype(np.float32) / 255.0


def _golden_check(ext, name):
    """max relative row error of the extractor on the committed golden rows of `name` (tests/golden/<name>.npz, produced by
    transformers' Dinov2Model; slide / coordinates of tests/cases.py: DINOV2_CASES)."""
    import torch

    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    case = GOLDEN_CASES[name]
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    sl = case["slide"]
    w = SyntheticWSI(make_spec(sl["width"], sl["height"], sl["seed"], mpp=sl["mpp"]))
    rng = np.random.default_rng(99)
    P = case["patch"]
    xy = np.stack([rng.integers(0, sl["width"] - P, case["n"]), rng.integers(0, sl["height"] - P, case["n"])], 1)
    xy[-1] = (sl["width"] - P // 2, sl["height"] - P // 3)
    rows = np.concatenate([xy, np.full((case["n"], 2), P), np.zeros((case["n"], 1))], 1).astype(np.int32)
    got = ext.embed_coords(w.device_image, w.w, w.h, w.pitch, torch.from_numpy(rows).cuda()).cpu().numpy()
    w.cleanup()
    return float(_rel_rows(got, g["feats"]).max())


# the golden cases of tests/cases.py: DINOV2_CASES (kept in step by tests/test_bench_config.py)
GOLDEN_CASES = {
    "dinov2_large": dict(weight_seed=4321, slide=dict(width=4096, height=4096, seed=12, mpp=0.5), n=8, patch=224),
    "dinov2_giant": dict(weight_seed=777, slide=dict(width=4096, height=4096, seed=13, mpp=0.5), n=4, patch=512),
}
GFLOP_PER_PATCH = {"dinov2_large": 162.02, "dinov2_giant": 598.78}


def aux_c2_sam2(ctx, rank, world):
    """BASELINE.json configs[2]: SAM2 Hiera-L forward on the 1024 x 1024 thumbnail; mask IoU against the golden mask of the fp32
    restatement (tests/golden/sam2_hiera_l_mask.npz, transformers' Sam2Model -- the reference's sam2 package is not installable
    offline)."""
    import torch

    from atlaspatch_b200.sam2 import HIERA_L, B200Sam2Predictor
    from atlaspatch_b200.synthetic import sam2_benchmark_image
    from atlaspatch_b200.weights import sam2_state_dict

    pred = B200Sam2Predictor(sam2_state_dict(1, "large"), config=HIERA_L, device=torch.cuda.current_device())
    img = sam2_benchmark_image()
    pred.predict_logits(img)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        up = pred.predict_logits(img)
    ms = (time.perf_counter() - t0) * 1000 / reps
    pred.close()
    out = {"model": "sam2 hiera-large, 1024x1024, box prompt = whole image", "ms_per_thumbnail_incl_h2d_d2h": ms,
           "gflop_per_forward": 1625.4, "tflops": 1625.4 / ms}
    gpath = ROOT / "tests" / "golden" / "sam2_hiera_l_mask.npz"
    if gpath.exists():
        g = np.load(gpath)
        ref = np.unpackbits(g["mask_bits"])[:1024 * 1024].reshape(1024, 1024).astype(bool)
        a = up > 0
        out["mask_iou_vs_golden"] = float((a & ref).sum() / max((a | ref).sum(), 1))
        out["positives"] = [int(a.sum()), int(g["positives"])]
    out["ms_max_over_ranks"] = _max_over_ranks(ms, world)
    return out


def _encoder_for(name, patch, seed, chunk=254):
    import torch

    from atlaspatch_b200.encoder import B200FeatureExtractor
    from atlaspatch_b200.weights import dinov2_state_dict

    sd = dinov2_state_dict(name, seed=seed)     # seeded input data: every rank draws the same values (no collective needed)
    ext = B200FeatureExtractor(name, sd, input_patch=patch, max_batch=chunk, device=torch.cuda.current_device())
    del sd
    return ext


def aux_c3_slides(ctx, rank, world, local_rank, peaks):
    """BASELINE.json configs[3]: a batch of `world` synthetic 40000 x 40000 slides, "ViT-L/14" = dinov2_large, 224 px patches, one
    slide per rank (sharding.assign_slides; with 8 ranks this is the named 8-slide batch).  Per rank: slide in HBM -> coords of the
    whole slide -> features of every coordinate; no collective on the data path."""
    import torch

    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.sharding import assign_slides
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    name = "dinov2_large"
    slides = [dict(width=40000, height=40000, seed=100 + i) for i in range(world)]
    mine = assign_slides([sl["width"] * sl["height"] for sl in slides], world)[rank]
    ext = _encoder_for(name, 224, GOLDEN_CASES[name]["weight_seed"])
    golden_err = _golden_check(ext, name)
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=224, target_magnification=20, step_size=224))
    n_patches, ms_total = 0, 0.0
    for i in mine:
        sl = slides[i]
        wsi = SyntheticWSI(make_spec(sl["width"], sl["height"], sl["seed"]))
        wsi.device_image
        res = svc.extract(wsi, _capped_mask(wsi.spec), slide=Slide(Path(wsi.path), mpp=0.5))
        rows = res.coords_device.contiguous()
        ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows[:254].contiguous(), read_size=224)   # warm
        _, ms = _cuda_time(lambda: ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows, read_size=224))
        n_patches += int(rows.shape[0])
        ms_total += ms
        wsi.cleanup()
        del wsi, rows, res
        torch.cuda.empty_cache()
    ext.cleanup()
    total = _sum_over_ranks(n_patches, world)
    ms_max = _max_over_ranks(ms_total, world)
    pps = total / (ms_max / 1000.0)
    return {"workload": f"{world} synthetic 40000x40000 slides (seeds 100..), dinov2_large (ViT-L/14), 224px patches stride 224, one slide "
                        "per rank, whole slide embedded", "slides": world, "patches_total": int(total), "patches_rank0": n_patches,
            "embed_ms_max_over_ranks": ms_max, "patches_per_s": pps, "model_tflops_per_gpu": pps / world * GFLOP_PER_PATCH[name] / 1000,
            "frac_of_sustained_peak": pps / world * GFLOP_PER_PATCH[name] / 1000 / peaks["tflops_sustained"],
            "max_rel_err_vs_golden": golden_err, "tolerance": 1e-3}


def c4_slide_spec():
    """82000 x 80000 slide covered by tissue except for three holes: 512 px patches at stride 256 give ~100 k coordinates
    (BASELINE.json configs[4]: "100k-patch slide")."""
    from atlaspatch_b200.synthetic import SyntheticSlideSpec

    W, H = 82000, 80000
    cw, ch = (W + 15) >> 4, (H + 15) >> 4
    blobs = ((cw // 2, ch // 2, int(0.75 * cw), int(0.75 * ch), 64, 0),)        # an ellipse that contains the whole rectangle
    holes = ((cw // 3, ch // 3, ch // 12), (2 * cw // 3, ch // 2, ch // 16), (cw // 2, 4 * ch // 5, ch // 20))
    return SyntheticSlideSpec(W, H, 5, 0.5, blobs, holes)


def aux_c4_intra_slide(ctx, rank, world, local_rank, peaks, cap):
    """BASELINE.json configs[4]: ONE synthetic slide, 512 px patches at stride 256 (overlap) -> ~100 k patches, dinov2_giant.  Every
    rank holds the slide and the coordinate list, embeds its contiguous row range (sharding.row_range) and the (N, 1536) fp32
    matrix is all-gathered over NCCL (sharding.gather_rows) -- the path's only data-path collective, timed here.  With fewer than 8
    ranks each rank embeds at most `cap` rows of its range (a bounded sample of the same slide; at 8 ranks the whole slide)."""
    import torch
    import torch.distributed as dist

    from atlaspatch_b200.services import B200PatchExtractionService, ExtractionConfig, Slide
    from atlaspatch_b200.sharding import gather_rows, row_range
    from atlaspatch_b200.slide import SyntheticWSI
    from atlaspatch_b200.synthetic import make_spec

    name = "dinov2_giant"
    ext = _encoder_for(name, 512, GOLDEN_CASES[name]["weight_seed"])
    golden_err = _golden_check(ext, name)
    wsi = SyntheticWSI(c4_slide_spec())
    wsi.device_image
    svc = B200PatchExtractionService(ExtractionConfig(patch_size=512, target_magnification=20, step_size=256))
    res = svc.extract(wsi, _capped_mask(wsi.spec), slide=Slide(Path(wsi.path), mpp=0.5))
    n_all = res.num_patches
    b, e = row_range(n_all, rank, world)
    e = min(e, b + cap)
    rows = res.coords_device[b:e].contiguous()
    ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows[:127].contiguous(), read_size=512)   # warm
    local, ms = _cuda_time(lambda: ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, rows, read_size=512))
    n_local = int(rows.shape[0])
    out = {"workload": f"one synthetic {wsi.w}x{wsi.h} slide (tissue nearly everywhere), 512px patches stride 256, dinov2_giant (ViT-g/14, SwiGLU, 40 layers), rows split "
                       "by contiguous range over the ranks", "coords_total": int(n_all), "rows_per_rank_cap": int(cap)}
    gather_ms, identical = None, None
    if world > 1:
        sizes = [min(row_range(n_all, r, world)[1], row_range(n_all, r, world)[0] + cap) - row_range(n_all, r, world)[0] for r in range(world)]
        full = all(sz == row_range(n_all, r, world)[1] - row_range(n_all, r, world)[0] for r, sz in enumerate(sizes))
        if full:
            dist.barrier()
            gathered, gather_ms = _cuda_time(lambda: gather_rows(local, n_all))
        else:   # bounded sample: gather the equal-sized blocks that were embedded
            n_eq = min(sizes)
            dist.barrier()
            gathered, gather_ms = _cuda_time(lambda: gather_rows(local[:n_eq].contiguous(), n_eq * world))
        gather_ms = _max_over_ranks(gather_ms, world)
        # bit-identity of the sharded result with a single-rank run: rank 0 recomputes the first rows of rank 1's block itself
        if rank == 0:
            b1 = row_range(n_all, 1, world)[0]
            k = min(16, sizes[1])
            mine = ext.embed_coords(wsi.device_image, wsi.w, wsi.h, wsi.pitch, res.coords_device[b1:b1 + k].contiguous(), read_size=512)
            off = (b1 if full else min(sizes))
            identical = bool(torch.equal(mine, gathered[off:off + k]))
        out["gather"] = {"collective": "all_gather (NCCL) of the fp32 feature blocks", "ms_max_over_ranks": gather_ms,
                         "bytes_total": int((n_all if full else min(sizes) * world) * 1536 * 4), "whole_slide": bool(full),
                         "rows_identical_to_single_rank_run": identical}
    ext.cleanup()
    wsi.cleanup()
    total = _sum_over_ranks(n_local, world)
    ms_max = _max_over_ranks(ms, world)
    pps = total / (ms_max / 1000.0)
    out.update({"patches_embedded": int(total), "embed_ms_max_over_ranks": ms_max, "patches_per_s": pps,
                "patches_per_s_incl_gather": total / ((ms_max + (gather_ms or 0.0)) / 1000.0),
                "model_tflops_per_gpu": pps / world * GFLOP_PER_PATCH[name] / 1000,
                "frac_of_sustained_peak": pps / world * GFLOP_PER_PATCH[name] / 1000 / peaks["tflops_sustained"],
                "precise_layers": "library default for > 32 layers", "max_rel_err_vs_golden": golden_err, "tolerance": 1e-3})
    return out


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def main_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU pipeline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()

    from atlaspatch_b200._lib import Context
    from atlaspatch_b200.encoder import B200FeatureExtractor, vit_state_dict_names
    from atlaspatch_b200.extraction import extract_coords
    from atlaspatch_b200.slide import SyntheticWSI

    ctx = Context.get(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    peaks = measured_peaks()

    # ---- slide resident in HBM, thumbnail, coords (one slide per rank, seed = rank) ----------------------
    spec, mask = slide_mask_coords_cpu(args.width, args.height, rank)
    wsi = SyntheticWSI(spec, ctx=ctx)
    image = wsi.device_image
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wsi.thumbnail_at_power_device(1.25)  # warm
    ev0.record()
    thumb = wsi.thumbnail_at_power_device(1.25)
    ev1.record()
    torch.cuda.synchronize()
    thumb_ms = ev0.elapsed_time(ev1)
    thumb_bytes = args.width * args.height * 3 + thumb.numel()
    t0 = time.perf_counter()
    coords_np, coords_dev = extract_coords(mask, level0_wh=(args.width, args.height), src_mag=20, target_mag=20, patch_size=256,
                                           step_size=256, tissue_thresh=0.0, ctx=ctx, return_device=True)
    coords_first_ms = 1000 * (time.perf_counter() - t0)  # includes the cv2 import and the CUDA module load
    t0 = time.perf_counter()
    extract_coords(mask, level0_wh=(args.width, args.height), src_mag=20, target_mag=20, patch_size=256, step_size=256,
                   tissue_thresh=0.0, ctx=ctx, return_device=True)
    coords_ms = 1000 * (time.perf_counter() - t0)
    n_coords = int(coords_np.shape[0])
    assert n_coords > 0
    # --no-fast-mode content filter over every candidate (ap_filter_patches; both kernels timed with CUDA events on their stream)
    from atlaspatch_b200.extraction import filter_patches

    filter_patches(image, args.width, args.height, wsi.pitch, coords_dev, patch_size=256)  # warm
    ctx.profile(True, ["coords"])
    kept_np, _ = filter_patches(image, args.width, args.height, wsi.pitch, coords_dev, patch_size=256)
    torch.cuda.synchronize()
    filter_ms = ctx.profile_read()["coords"][0]
    ctx.profile(False)
    filter_bytes = n_coords * 256 * 256 * 3

    # ---- encoder weights: rank 0 builds the seeded state_dict, NCCL broadcast to the other ranks ----------
    names = vit_state_dict_names(12)
    if rank == 0:
        from atlaspatch_b200.weights import vit_state_dict  # seeded random-init weights (input data), shared with the CPU arm

        sd = vit_state_dict("vit_b_16", seed=WEIGHT_SEED)
    else:
        sd = None
    if world > 1:
        shapes = [None]
        if rank == 0:
            shapes = [{k: tuple(sd[k].shape) for k in names}]
        dist.broadcast_object_list(shapes, src=0)
        out = {}
        for k in names:
            t = sd[k].cuda() if rank == 0 else torch.empty(shapes[0][k], dtype=torch.float32, device="cuda")
            dist.broadcast(t, src=0)
            out[k] = t.cpu()
        sd = out
    ext = B200FeatureExtractor("vit_b_16", sd, max_batch=args.chunk, device=local_rank)
    del sd

    B = args.batch
    reps = (B + n_coords - 1) // n_coords + 1
    coords_ring = coords_dev.repeat(reps + 1, 1).contiguous() if n_coords < 2 * B else coords_dev
    n_ring = int(coords_ring.shape[0])
    feats = torch.empty((B, 768), dtype=torch.float32, device="cuda")

    def step(i):
        s = (i * B) % (n_ring - B + 1)
        ext.embed_coords(image, wsi.w, wsi.h, wsi.pitch, coords_ring[s:s + B], out=feats)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    # live CUDA events for the dominant kernel inside the timed region: every --profile-stride-th GEMM launch (the stride, 4, is
    # coprime with the 49 GEMM launches of a chunk, so over a step every shape is sampled evenly); timing EVERY launch costs the
    # step ~2 % because an event record between two kernels breaks their programmatic-dependent-launch overlap
    stride = max(args.profile_stride, 0)
    if stride:
        ctx.set_option("profile_stride", stride)
        ctx.profile(True, classes=("gemm",))
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.set_option("profile_stride", 1)
    ctx.profile(True)                        # one extra, untimed step with every kernel class timed: per-class breakdown
    step(args.warmup + args.steps)
    prof_all = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * args.steps * B / (ms_max / 1000.0)

    # ---- e2e: the reference-facing plug-in call with HOST patches (H2D + D2H inside the timed region) -----
    # every step embeds DIFFERENT patches: a pool of e2e_steps x B patches read from the slide beforehand (not timed), as a
    # caller of extract_batch would hold them
    n_e2e = B
    e2e_steps = max(1, args.e2e_steps)
    pool_n = min(n_coords, e2e_steps * B) if n_coords >= B else B
    pool_rows = coords_np[:pool_n] if n_coords >= B else np.concatenate([coords_np] * (B // n_coords + 1))[:B]
    pool = host_patches_from_slide(wsi, pool_rows, 256)
    ext.extract_batch(pool[:256], batch_size=32)  # warm
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        s0 = (i * B) % (pool_n - B + 1)
        f_host = ext.extract_batch(pool[s0:s0 + B], batch_size=32)
    e2e_s = time.perf_counter() - t0
    assert f_host.shape == (n_e2e, 768)
    del pool
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * n_e2e / float(t_e2e.item())

    # ---- the other BASELINE.json configs, measured beside the headline (every rank takes part) ------------
    ext.cleanup()
    del ext, feats, coords_ring
    aux_cfg = {} if args.aux.strip().lower() in ("", "none") else {k.strip(): True for k in args.aux.split(",")}
    aux_out = {}
    def _aux(key, fn, *a):
        # The aux configs must not take the headline down with them.  On one GPU an exception is recorded in the line; with several
        # ranks it propagates (the other ranks are inside the same collectives: swallowing it on one rank would hang the rest).
        if key not in aux_cfg:
            return
        if world > 1:
            aux_out[key] = fn(*a)
            return
        try:
            aux_out[key] = fn(*a)
        except Exception as exc:  # noqa: BLE001
            aux_out[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    _aux("c2", aux_c2_sam2, ctx, rank, world)
    wsi.cleanup()
    del wsi, image
    torch.cuda.empty_cache()
    _aux("c3", aux_c3_slides, ctx, rank, world, local_rank, peaks)
    _aux("c4", aux_c4_intra_slide, ctx, rank, world, local_rank, peaks, args.c4_cap)

    if rank != 0:
        if world > 1:
            dist.barrier()           # rank 0 is still timing the CPU baseline and printing the line
            dist.destroy_process_group()
        return 0

    gemm_ms, gemm_n = prof["gemm"]
    if stride == 0:                          # nothing timed inside the region: fall back to the extra profiled step
        gemm_ms, gemm_n = prof_all["gemm"]
        patches_timed = B
    else:
        gemm_ms, gemm_n = gemm_ms * stride, gemm_n * stride   # sampled mean x all launches
        patches_timed = args.steps * B
    achieved_tflops = (patches_timed * GEMM_FLOP_PER_PATCH) / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else None
    roofline = {
        "bound": "tensor", "kernel": f"gemm_tcgen05_kernel (all 49 launches per {args.chunk}-patch chunk: conv_proj, in_proj, out_proj, mlp.0, mlp.3)",
        "achieved": achieved_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
        "frac": (achieved_tflops / peaks["tflops_sustained"]) if achieved_tflops else None,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 49 GEMM launches of one forward chunk (ncu,
        # profiles/r02_ncu_gemm_dram_traffic.csv at the default 508-patch chunk: 405.2 MB read + 418.7 MB written; at 127 patches
        # 103.7 + 64.7 MB, profiles/r01_ncu_gemm_dram_traffic.csv); algorithmic bytes A + W + out (+ resid + the fp16 copy of the
        # residual stream the folded LayerNorm needs) average 233 MB per launch per 127 patches, the difference is served by L2
        "traffic": (824.0e6 if args.chunk == 508 else 168.4e6 * args.chunk / 127.0),
        "traffic_unit": "B/launch (ncu, profiles/r02_ncu_gemm_dram_traffic.csv; other chunk sizes: scaled from the 127-patch capture)",
        "peak_source": peaks["source"] + ", sustained bf16 cuBLAS figure (kernel timed inside a long step)",
        "algorithmic_flop_per_launch": GEMM_FLOP_PER_PATCH * patches_timed / max(gemm_n, 1),
        "avg_launch_ms": gemm_ms / max(gemm_n, 1), "launches_timed": gemm_n // max(stride, 1) if stride else gemm_n,
        "launch_sampling": (f"every {stride}th GEMM launch of the timed region" if stride else "the extra profiled step only"),
        "kernel_share_of_step": gemm_ms / (ms * patches_timed / (args.steps * B)) if ms > 0 else None,
        "per_class_ms_one_step": {k: round(v[0], 3) for k, v in prof_all.items() if v[1]},
        "whole_model": {"tflops": value / world * MODEL_FLOP_PER_PATCH / 1e12,
                        "frac_of_sustained_peak": value / world * MODEL_FLOP_PER_PATCH / 1e12 / peaks["tflops_sustained"]},
        "thumbnail_hbm": {"bound": "hbm", "achieved": thumb_bytes / (thumb_ms / 1000.0) / 1e9, "peak": peaks["hbm_gbs"],
                          "unit": "GB/s", "frac": thumb_bytes / (thumb_ms / 1000.0) / 1e9 / peaks["hbm_gbs"], "ms": thumb_ms},
        # the filter is bound by the SM's ALU pipe (ncu: 72 % ALU, 74 % issue slots; profiles/r01_ncu_full_filter_count.csv), the
        # HBM fraction is reported because its algorithmic work is bytes (patch^2 * 3 per candidate, read exactly once)
        "content_filter_hbm": {"bound": "hbm", "achieved": filter_bytes / (filter_ms / 1000.0) / 1e9, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": filter_bytes / (filter_ms / 1000.0) / 1e9 / peaks["hbm_gbs"], "ms": filter_ms,
                               "candidates": n_coords, "kept": int(kept_np.shape[0])},
    }

    cpu_baseline = None
    if not args.no_cpu_baseline:     # rank 0, at every N (the other ranks are idle at the final barrier meanwhile)
        arm = ReferenceArm(args.width, args.height, 0)
        arm.run(0, 32)  # warm
        dt = arm.run(32, args.cpu_sample)
        cpu_baseline = {"value": args.cpu_sample / dt, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                        "sample": f"{args.cpu_sample} consecutive coordinate rows (after a 32-patch warm-up) of the same slide through "
                                  f"{arm.what}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate/residual/LayerNorm/softmax", "data": "synthetic",
        "config": {"workload": f"single synthetic {args.width}x{args.height} RGB slide per GPU resident in HBM, 256px patches "
                               f"stride 256 ({n_coords} coords on rank 0), ViT-B/16 random-init (seeded)",
                   "patches_per_step": B, "forward_chunk": args.chunk, "l2": "inputs larger than L2 (each step reads a different "
                   f"{B * 150528 / 1e6:.0f} MB of the 14.4 GB slide)", "parallelism": f"slides sharded 1 per GPU x{world}, no steady-state collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_e2e * PATCH_BYTES, "d2h_bytes_per_step": n_e2e * 768 * 4,
                "api": "B200FeatureExtractor.extract_batch(list of host uint8 patches) -> host float32 features",
                "patches_per_step": n_e2e, "steps": e2e_steps, "distinct_patches": int(pool_n)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "aux": {"coords": n_coords, "coords_ms_incl_host_contours": coords_ms, "coords_first_call_ms": coords_first_ms,
                "thumbnail_ms": thumb_ms, "content_filter_ms": filter_ms, **aux_out},
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
