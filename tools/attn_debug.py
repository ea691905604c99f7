"""Diagnostics for the tcgen05 attention kernel: error vs torch for a few shapes / layout variants."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context  # noqa: E402

ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())


def ref_attn(qkv, B, S, heads):
    D = heads * 64
    q, k, v = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    o = torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * S, D)


torch.manual_seed(0)
for mode in (2, 1):
    ctx.set_option("attn_mode", mode)
    for variant in ((0, 1) if mode == 2 else (0,)):
        ctx.set_option("attn_variant", variant)
        for (B, S, heads) in [(1, 64, 1), (1, 197, 1), (2, 197, 12), (3, 16, 2), (1, 1, 1), (2, 130, 3), (1, 256, 2)]:
            qkv = (torch.randn(B * S, 3 * heads * 64, device="cuda") * 1.5).half()
            out = torch.full((B * S, heads * 64), float("nan"), device="cuda", dtype=torch.float16)
            rc = ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
            if rc:
                print("rc", rc, ctx.lib.ap_last_error(ctx.handle).decode())
            torch.cuda.synchronize()
            ref = ref_attn(qkv, B, S, heads)
            err = (out.float() - ref).abs()
            print(f"mode {mode} variant {variant} B{B} S{S} h{heads}: max err {err.max().item():.4g} nan {int(torch.isnan(out).sum())} "
                  f"ref scale {ref.abs().max().item():.3g}")
ctx.set_option("attn_mode", 2)
ctx.set_option("attn_variant", 0)
# timing at the encoder shape
B, S, heads = 127, 197, 12
qkv = (torch.randn(B * S, 3 * heads * 64, device="cuda") * 1.0).half()
out = torch.empty((B * S, heads * 64), device="cuda", dtype=torch.float16)
for mode, variant in ((2, 0), (2, 2), (2, 4), (2, 6), (2, 8), (2, 10), (1, 0)):
    ctx.set_option("attn_mode", mode)
    ctx.set_option("attn_variant", variant)
    for _ in range(3):
        ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
    e1.record()
    torch.cuda.synchronize()
    print(f"mode {mode} variant {variant} (2=no max pass, 4=no exp, 8=no pass 2): {e0.elapsed_time(e1) / 20 * 1000:.1f} us per call (B=127, S=197, 12 heads)")
ctx.set_option("attn_variant", 0)
ctx.set_option("attn_mode", 2)
