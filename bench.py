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
# patches per forward chunk: 127 x 197 tokens = 98 CTA-pair row tiles -> 294 / 882 / 1176 tiles = 3.97 / 11.9 / 15.9 waves of 74 pairs
CHUNK = 127


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=2032, help="patches per step (b200 arm); 16 forward chunks of 127")
    ap.add_argument("--chunk", type=int, default=CHUNK, help="patches per forward chunk (workspace size)")
    ap.add_argument("--width", type=int, default=80000)
    ap.add_argument("--height", type=int, default=60000)
    ap.add_argument("--e2e-steps", type=int, default=3)
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
# CPU arm: the reference's own CPU pipeline (oracle port; /root/reference is absent on the GPU box)
# ---------------------------------------------------------------------------------------------------------
def cpu_pipeline_setup(width, height, seed):
    import torch

    from atlaspatch_b200.synthetic import render_region_host
    from oracle import reference_loop as rl
    from oracle.weights import vit_state_dict

    spec, mask = slide_mask_coords_cpu(width, height, seed)
    gold = ROOT / "tests" / "golden" / "coords_c1_80000x60000_p256.npz"
    if (width, height, seed) == (80000, 60000, 0) and gold.exists():
        coords = np.load(gold)["coords"]     # produced by the reference itself on this exact slide / mask
    else:
        from oracle import coords as oc

        coords = oc.coords_from_mask(mask, level0_wh=(width, height), src_mag=20, target_mag=20, patch_size=256,
                                     step_size=256, tissue_thresh=0.0)
    model, preprocess = rl.build_vit_b_16(vit_state_dict("vit_b_16", seed=WEIGHT_SEED))
    read = lambda x, y, w, h: render_region_host(spec, x, y, w, h)  # noqa: E731
    return spec, coords, model, preprocess, read, torch.get_num_threads()


def run_cpu_sample(coords, model, preprocess, read, start, n):
    from oracle import reference_loop as rl

    rows = coords[start:start + n]
    t0 = time.perf_counter()
    feats = rl.embed_slide_rows(model, preprocess, read, rows, patch_size=256, feature_batch=32, num_workers=4)
    dt = time.perf_counter() - t0
    assert feats.shape == (rows.shape[0], 768)
    return dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec, coords, model, preprocess, read, threads = cpu_pipeline_setup(args.width, args.height, 0)
    n = args.ref_sample
    for i in range(args.warmup):
        run_cpu_sample(coords, model, preprocess, read, (i * n) % max(1, len(coords) - n), n)
    t = 0.0
    for i in range(args.steps):
        t += run_cpu_sample(coords, model, preprocess, read, ((args.warmup + i) * n) % max(1, len(coords) - n), n)
    value = args.steps * n / t
    sample = (f"{args.steps} steps x {n} consecutive coordinate rows of the {args.width}x{args.height} slide "
              f"({len(coords)} patches total), batch 32, DataLoader num_workers=4, fp32")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single synthetic {args.width}x{args.height} RGB slide, 256px patches stride 256, ViT-B/16 "
                               "random-init (seeded), reference CPU pipeline port (oracle/reference_loop.py)",
                   "patches_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


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
    n_e2e = B
    host_rows = coords_np[:n_e2e] if n_coords >= n_e2e else np.concatenate([coords_np] * (n_e2e // n_coords + 1))[:n_e2e]
    host_patches = []
    img3 = image  # (H, pitch) uint8
    for (x, y, _rw, _rh, _lv) in host_rows.tolist():     # patches a caller would have read from the slide (not timed)
        host_patches.append(wsi.extract((x, y), 0, (256, 256)))
    ext.extract_batch(host_patches[:256], batch_size=32)  # warm
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        f_host = ext.extract_batch(host_patches, batch_size=32)
    e2e_s = time.perf_counter() - t0
    assert f_host.shape == (n_e2e, 768)
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * args.e2e_steps * n_e2e / float(t_e2e.item())

    if rank != 0:
        if world > 1:
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
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel (all 49 launches per 127-patch chunk: conv_proj, in_proj, out_proj, mlp.0, mlp.3)",
        "achieved": achieved_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
        "frac": (achieved_tflops / peaks["tflops_sustained"]) if achieved_tflops else None,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 49 GEMM launches of one 127-patch chunk (ncu,
        # profiles/r01_ncu_gemm_dram_traffic.csv: 103.7 MB read + 64.7 MB written); algorithmic bytes A + W + out (+ resid + the fp16
        # copy of the residual stream the folded LayerNorm needs) average 233 MB per launch, the difference is served by L2
        "traffic": 168.4e6, "traffic_unit": "B/launch (ncu, profiles/r01_ncu_gemm_dram_traffic.csv)",
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
    if world == 1 and not args.no_cpu_baseline:
        _spec, c_coords, model, preprocess, read, threads = cpu_pipeline_setup(args.width, args.height, 0)
        run_cpu_sample(c_coords, model, preprocess, read, 0, 32)  # warm
        dt = run_cpu_sample(c_coords, model, preprocess, read, 32, args.cpu_sample)
        cpu_baseline = {"value": args.cpu_sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_sample} consecutive coordinate rows (after a 32-patch warm-up) of the same slide through "
                                  "oracle/reference_loop.py: per-row read, new DataLoader(num_workers=4) per 32 patches, "
                                  "torchvision vit_b_16 fp32 on the host cores"}

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
                "patches_per_step": n_e2e, "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "aux": {"coords": n_coords, "coords_ms_incl_host_contours": coords_ms, "coords_first_call_ms": coords_first_ms,
                "thumbnail_ms": thumb_ms, "content_filter_ms": filter_ms},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
