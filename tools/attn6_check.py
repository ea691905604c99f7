"""Timing and precision of the free-standing-O attention pipeline (attention_tc6_kernel) by variant and by the number of exponential
pairs per 16 evaluated on the FMA pipe ("attn_emu"), against the previous pipeline (variant 32) and a float64 reference."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
torch.manual_seed(0)
cases = [(32, 0), (0, 0), (64, 0), (0, 4), (2048, 0), (2048, 4)]
for (B, S, heads) in [(127, 197, 12), (508, 197, 12), (127, 256, 16), (508, 50, 12)]:
    D = heads * 64
    qkv = torch.randn(B * S, 3 * D, device="cuda").half()
    q, k, v = qkv[: 8 * S].double().view(8, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(8 * S, D)
    out = torch.empty((B * S, D), device="cuda", dtype=torch.float16)
    for variant, emu in cases:
        ctx.set_option("attn_variant", variant)
        ctx.set_option("attn_emu", emu)
        for _ in range(3):
            ctx.check(ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
        e1.record()
        torch.cuda.synchronize()
        err = out[: 8 * S].double() - ref
        print(f"B{B} S{S} h{heads} variant {variant} emu {emu}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us   rel-l2 {(err.norm() / ref.norm()).item():.3e} "
              f"worst row {(err.norm(dim=1) / ref.norm(dim=1)).max().item():.3e}", flush=True)
ctx.set_option("attn_variant", 0)
ctx.set_option("attn_emu", 0)
