"""Precision of the attention kernels against a float64 reference (rel-l2 over all outputs and worst row), both flavours."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
torch.manual_seed(0)
for (B, S, heads, amp, coherent) in [(8, 197, 12, 1.0, 0), (8, 257, 16, 1.0, 0), (8, 257, 16, 1.0, 1), (8, 197, 12, 1.0, 1), (8, 257, 16, 3.0, 0), (8, 256, 16, 1.0, 0)]:
    D = heads * 64
    qkv = torch.randn(B * S, 3 * D, device="cuda") * amp
    if coherent:   # near-identical tokens: one common row + small per-token part (an 83 % black patch)
        qkv = qkv[:1].repeat(B * S, 1) + 0.02 * torch.randn(B * S, 3 * D, device="cuda")
    qkv = qkv.half()
    q, k, v = qkv.double().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    for mode in (2, 1):
        ctx.set_option("attn_mode", mode)
        out = torch.empty((B * S, D), device="cuda", dtype=torch.float16)
        ctx.check(ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None))
        torch.cuda.synchronize()
        err = out.double() - ref
        rel = (err.norm() / ref.norm()).item()
        row = (err.norm(dim=1) / ref.norm(dim=1)).max().item()
        rnd = ((ref.half().double() - ref).norm() / ref.norm()).item()
        print(f"B{B} S{S} h{heads} amp {amp} coherent {coherent} mode {mode}: rel-l2 {rel:.3e} worst row {row:.3e} (fp16 rounding of the exact result alone: {rnd:.3e})")
ctx.set_option("attn_mode", 2)
