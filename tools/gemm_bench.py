"""GEMM micro-benchmark on a GPU box: the four per-layer shapes of ViT-B/16 at one forward chunk, per kernel flavour and
diagnostic mode.  python tools/gemm_bench.py [chunk]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402

ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 127
M = chunk * 197
SHAPES = [("qkv", 2304, 768, 0), ("proj", 768, 768, 2), ("fc1", 3072, 768, 1), ("fc2", 768, 3072, 2)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def bench(N, K, epi, reps=20):
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float16 if epi < 2 else torch.float32)
    resid = out
    for _ in range(3):
        ctx.check(ctx.lib.ap_gemm_f16(ctx.handle, P(A), P(W), P(bias), P(resid), P(out), M, N, K, epi, None))
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.check(ctx.lib.ap_gemm_f16(ctx.handle, P(A), P(W), P(bias), P(resid), P(out), M, N, K, epi, None))
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    return ms, 2.0 * M * N * K / ms / 1e9


print(f"M = {M} ({chunk} patches x 197 tokens); L2 flushed between launches")
for cg in (2, 1):
    for dbg, label in ((0, "full"), (1, "no-epilogue"), (2, "no-mma"), (3, "no-mma no-epi (TMA only)"), (4, "no-tma"), (5, "no-tma no-epi (MMA only)")):
        ctx.set_option("gemm_cta_group", cg)
        ctx.set_option("gemm_debug", dbg)
        row = []
        for name, N, K, epi in SHAPES:
            ms, tf = bench(N, K, epi)
            row.append(f"{name} {ms*1000:7.1f}us {tf:7.1f}TF")
        print(f"cg{cg} {label:28s} | " + " | ".join(row))
ctx.set_option("gemm_debug", 0)
# cuBLAS reference (fp16 in, fp32 accumulate) for the same shapes
for name, N, K, epi in SHAPES:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * 0.05).half()
    for _ in range(3):
        A @ W.T
    torch.cuda.synchronize()
    tot = 0
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A @ W.T; e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms = tot / 20
    print(f"cuBLAS {name}: {ms*1000:.1f} us {2.0*M*N*K/ms/1e9:.1f} TF (no epilogue)")
