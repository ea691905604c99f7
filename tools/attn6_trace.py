"""Prints the per-tile hand-over timeline block 0 of attention_tc6_kernel records (attn_variant bit 128): clock cycles relative to the
first event, for the MMA thread, softmax warps 4 / 8 and epilogue warp 12."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
B, S, heads = 127, 197, 12
emu = int(sys.argv[1]) if len(sys.argv) > 1 else 0
qkv = torch.randn(B * S, 3 * heads * 64, device="cuda").half()
out = torch.empty((B * S, heads * 64), device="cuda", dtype=torch.float16)
ctx.set_option("attn_emu", emu)
for v in (0, 0, 128):
    ctx.set_option("attn_variant", v)
    ctx.check(ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None))
torch.cuda.synchronize()
ctx.set_option("attn_variant", 0)
tr = np.zeros((4, 256), dtype=np.int64)
ctx.lib.ap_debug_attn6_trace.argtypes = [C.c_void_p, C.c_void_p]
ctx.check(ctx.lib.ap_debug_attn6_trace(ctx.handle, tr.ctypes.data_as(C.c_void_p)))
t0 = tr[tr > 0].min()
r = lambda x: int(x - t0) if x > 0 else -1
print("emu", emu)
print("tile | MMA: loop-top  p_full  o_empty  issued | softmax(wg): wait-start s_full pass1-done P-published | epilogue: wait-start o_full regs-loaded stored")
for i in range(22):
    wg, k = i & 1, i >> 1
    m = [r(tr[2, 4 * i + j]) for j in range(4)]
    s = [r(tr[wg, 4 * k + j]) for j in range(4)]
    e = [r(tr[3, 4 * i + j]) for j in range(4)]
    print(f"{i:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | wg{wg} {s[0]:7d} {s[1]:7d} {s[2]:7d} {s[3]:7d} | {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d}")
