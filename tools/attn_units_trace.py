"""Per-unit hand-over timeline block 0 of attention_units_kernel records (attn_variant bit 128): clock cycles relative to the first
event, for the MMA warp, softmax warps 4 / 8 (warpgroups 0 / 1) and epilogue warp 12.   python tools/attn_units_trace.py [S heads]"""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
P = lambda t: C.c_void_p(t.data_ptr())
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = 127
qkv = torch.randn(B * S, 3 * heads * 64, device="cuda").half()
out = torch.empty((B * S, heads * 64), device="cuda", dtype=torch.float16)
for v in (2048, 2048, 2048 | 128):
    ctx.set_option("attn_variant", v)
    ctx.check(ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None))
torch.cuda.synchronize()
ctx.set_option("attn_variant", 0)
tr = np.zeros((4, 256), dtype=np.int64)
ctx.lib.ap_debug_attn_units_trace.argtypes = [C.c_void_p, C.c_void_p]
ctx.check(ctx.lib.ap_debug_attn_units_trace(ctx.handle, tr.ctypes.data_as(C.c_void_p)))
t0 = tr[tr > 0].min()
r = lambda x: int(x - t0) if x > 0 else -1
KB = 2 if S > 128 else 1
print(f"S {S} heads {heads}")
print("unit (tile,blk) wg | MMA: loop-top  p_full  o_empty-ok  issued(+S c+3) | softmax: wait-start  s_full  P-published")
for c in range(40):
    if KB == 2:
        tile, blk = 2 * (c >> 2) + (c & 1), (c >> 1) & 1
    else:
        tile, blk = c, 0
    wg = tile & 1
    m = [r(tr[2, 4 * c + j]) for j in range(4)]
    s = [r(tr[wg, 4 * c + j]) for j in range(3)]
    print(f"{c:3d} ({tile:2d},{'ab'[blk]}) wg{wg} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {s[0]:7d} {s[1]:7d} {s[2]:7d}")
print("tile | epilogue: wait-start o_full regs-loaded")
for i in range(12):
    e = [r(tr[3, 4 * i + j]) for j in range(3)]
    print(f"{i:3d} | {e[0]:7d} {e[1]:7d} {e[2]:7d}")
