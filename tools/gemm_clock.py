"""Sustained GEMM throughput with the SM clock sampled under load (is the kernel clock- or pipe-limited?)."""
import ctypes as C, subprocess, sys, time, tempfile, os
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
M = 127 * 197

def sample_clock(fn, seconds=2.0):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=f)
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    p.terminate(); p.wait(); f.flush(); f.seek(0)
    rows = [l.split(",") for l in f.read().splitlines() if "," in l]
    os.unlink(f.name)
    clk = sorted(float(r[0]) for r in rows)[len(rows)//2:]
    pw = sorted(float(r[1]) for r in rows)[len(rows)//2:]
    return ms / n, float(np.median(clk)), float(np.median(pw))

for name, N, K, epi in [("qkv", 2304, 768, 0), ("fc2", 768, 3072, 2)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); W = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda"); out = torch.zeros(M, N, device="cuda", dtype=torch.float16 if epi < 2 else torch.float32)
    for dbg, label in ((0, "full"), (1, "no-epilogue"), (5, "MMA only")):
        ctx.set_option("gemm_debug", dbg)
        ms, clk, pw = sample_clock(lambda: ctx.lib.ap_gemm_f16(ctx.handle, P(A), P(W), P(bias), P(out), P(out), M, N, K, epi, None))
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"{name} {label:12s}: {ms*1000:7.1f} us  {tf:7.1f} TFLOP/s  sm {clk:.0f} MHz  {pw:.0f} W  -> {tf*1e12/(148*clk*1e6):.0f} FLOP/clk/SM ({tf*1e12/(148*clk*1e6)/8192:.2f} of 8192)")
    ctx.set_option("gemm_debug", 0)
    ms, clk, pw = sample_clock(lambda: A @ W.T)
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"{name} cuBLAS      : {ms*1000:7.1f} us  {tf:7.1f} TFLOP/s  sm {clk:.0f} MHz  {pw:.0f} W  -> {tf*1e12/(148*clk*1e6):.0f} FLOP/clk/SM ({tf*1e12/(148*clk*1e6)/8192:.2f} of 8192)")
