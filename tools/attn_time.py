import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
for (B, S, heads, mode, variant) in [(127, 197, 12, 2, 0), (127, 197, 12, 2, 16), (127, 257, 16, 2, 0), (127, 257, 16, 2, 16), (127, 256, 16, 2, 0),
                                     (127, 257, 16, 1, 0), (127, 257, 24, 2, 0), (508, 50, 12, 2, 0), (127, 197, 12, 1, 0)]:
    ctx.set_option("attn_mode", mode)
    ctx.set_option("attn_variant", variant)
    qkv = (torch.randn(B * S, 3 * heads * 64, device="cuda")).half()
    out = torch.empty((B * S, heads * 64), device="cuda", dtype=torch.float16)
    for _ in range(3):
        ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None)
    e1.record(); torch.cuda.synchronize()
    print(f"B{B} S{S} h{heads} mode {mode} variant {variant} ({'tcgen05' if mode == 2 and S <= 257 else 'mma.sync'}): {e0.elapsed_time(e1) / 10 * 1000:.1f} us")
ctx.set_option("attn_variant", 0)
