"""One attention launch at the encoder shape (for ncu captures): python tools/attn_one.py [B S heads]"""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from atlaspatch_b200._lib import Context
ctx = Context.get(0)
B, S, heads = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (127, 197, 12)
P = lambda t: C.c_void_p(t.data_ptr())
qkv = (torch.randn(B * S, 3 * heads * 64, device="cuda")).half()
out = torch.empty((B * S, heads * 64), device="cuda", dtype=torch.float16)
for _ in range(3):
    ctx.check(ctx.lib.ap_attention_f16(ctx.handle, P(qkv), P(out), B, S, heads, None))
torch.cuda.synchronize()
print("ok")
