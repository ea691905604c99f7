"""Diagnostics for the tcgen05 GEMM on a GPU box: structured inputs that localise descriptor / layout errors.
Usage (under gpurun): python tools/gemm_debug.py > gpurun_out/gemm_debug.log 2>&1"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402

ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())


def run(A, W, bias, epi=3, resid=None):
    M, K = A.shape
    N = W.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if epi >= 2 else torch.float16)
    rc = ctx.lib.ap_gemm_f16(ctx.handle, P(A), P(W), P(bias), P(resid) if resid is not None else None, P(out), M, N, K, epi, None)
    if rc:
        print("rc", rc, ctx.lib.ap_last_error(ctx.handle).decode())
    torch.cuda.synchronize()
    return out.float()


def report(tag, got, ref):
    err = (got - ref).abs()
    bad = ~(err <= 1e-2 * (1 + ref.abs()))
    print(f"[{tag}] shape {tuple(got.shape)} max_err {err[torch.isfinite(err)].max().item() if torch.isfinite(err).any() else float('nan'):.4g} "
          f"nan {int(torch.isnan(got).sum())} bad {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print("   bad rows (first 24):", rows[:24].tolist(), "... count", rows.numel())
        print("   bad cols (first 24):", cols[:24].tolist(), "... count", cols.numel())
        print("   got[0:4,0:8]", got[0:4, 0:8].tolist())
        print("   ref[0:4,0:8]", ref[0:4, 0:8].tolist())


torch.manual_seed(0)
for (M, N, K) in [(128, 256, 64), (128, 256, 256), (128, 128, 64), (256, 512, 128), (300, 256, 768)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ W.float().T + bias
    got = run(A, W, bias)
    report(f"rand {M}x{N}x{K}", got, ref)
    if not torch.allclose(got, ref, rtol=1e-2, atol=1e-2):
        for s in range(0, K, 16):   # which 16-wide K slices are wrong?
            As = torch.zeros_like(A)
            As[:, s:s + 16] = A[:, s:s + 16]
            g2 = run(As, W, bias)
            r2 = As.float() @ W.float().T + bias
            report(f"   kslice {s}", g2, r2)
        break
print("done")
