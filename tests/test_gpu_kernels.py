"""GPU parity tests of the building-block kernels, called through the C ABI (ctypes)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from atlaspatch_b200._lib import Context

    return Context.get(0)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


GEMM_CASES = [
    # M, N, K, epilogue
    (128, 256, 64, 0), (128, 256, 128, 0), (256, 256, 768, 0), (200, 512, 256, 0),
    (197 * 4, 2304, 768, 0), (197 * 4, 3072, 768, 1), (197 * 4, 768, 3072, 2), (196 * 4, 768, 768, 3),
    (1000, 128, 192, 0), (77, 384, 64, 2), (25216, 768, 768, 2), (5000, 2304, 768, 0),
]


@pytest.mark.parametrize("cg", [2, 1])
@pytest.mark.parametrize("M,N,K,epi", GEMM_CASES)
def test_gemm_tcgen05(ctx, M, N, K, epi, cg):
    ctx.set_option("gemm_cta_group", cg)   # 2: CTA-pair 256x256 tiles (cta_group::2), 1: 128xBN tiles
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = A.float() @ W.float().T + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + resid
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16 if epi in (0, 1) else torch.float32)
    ctx.check(ctx.lib.ap_gemm_f16(ctx.handle, _p(A), _p(W), _p(bias), _p(resid) if epi == 2 else None, _p(out), M, N, K, epi, _stream()))
    torch.cuda.synchronize()
    got = out.float()
    assert torch.isfinite(got).all()
    tol = 2e-3 if epi in (0, 1) else 2e-5   # fp16 output rounding vs fp32 output
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ctx.set_option("gemm_cta_group", 2)
    assert err <= tol * max(scale, 1.0), f"max abs err {err} (scale {scale})"


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (1000, 768, 768), (257 * 3, 4608, 1536)])
def test_gemm_split_operands(ctx, M, N, K):
    """hi/lo split operands: each split removes that operand's fp16 rounding, so the error against the fp32 product of the ORIGINAL
    fp32 operands must fall: none > W-split ~ A-split > both (~fp32)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.double() @ W.double().T + bias.double())

    def split(x):
        hi = x.half()
        return hi, (x - hi.float()).half()

    A_hi, A_lo = split(A)
    W_hi, W_lo = split(W)
    ops = {0: (A_hi, W_hi), 1: (A_hi, torch.cat([W_hi, W_lo], 1)), 2: (torch.cat([A_hi, A_lo], 1), W_hi),
           3: (torch.cat([A_hi, A_lo], 1), torch.cat([W_hi, W_lo], 1))}
    err = {}
    for mode, (a, w) in ops.items():
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
        ctx.check(ctx.lib.ap_gemm_f16_split(ctx.handle, _p(a.contiguous()), _p(w.contiguous()), _p(bias), None, _p(out), M, N, K, 3, mode, _stream()))
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        err[mode] = ((out.double() - ref).norm() / ref.norm()).item()
    print("split-operand GEMM rel-l2 errors:", err)
    assert err[0] < 5e-4 and err[1] < 0.8 * err[0] and err[2] < 0.8 * err[0] and err[3] < 0.05 * err[0], err
    assert 0.7 < err[1] / err[2] < 1.4, err          # the two operands contribute alike


def test_gemm_rejects_bad_shapes(ctx):
    from atlaspatch_b200._lib import AtlasB200Error

    A = torch.zeros(128, 100, device="cuda", dtype=torch.float16)
    W = torch.zeros(256, 100, device="cuda", dtype=torch.float16)
    b = torch.zeros(256, device="cuda")
    o = torch.zeros(128, 256, device="cuda", dtype=torch.float16)
    with pytest.raises(AtlasB200Error):
        ctx.check(ctx.lib.ap_gemm_f16(ctx.handle, _p(A), _p(W), _p(b), None, _p(o), 128, 256, 100, 0, _stream()))


@pytest.mark.parametrize("rows,D", [(1, 256), (197 * 3, 768), (1000, 1024), (37, 1536), (5, 128)])
def test_layernorm(ctx, rows, D):
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 1.5
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps=1e-6)
    out = torch.empty(rows, D, device="cuda", dtype=torch.float16)
    ctx.check(ctx.lib.ap_layernorm_f16(ctx.handle, _p(x), D, _p(gamma), _p(beta), 1e-6, _p(out), rows, D, _stream()))
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


DEFAULT_ATTN_EMU = 0   # library default of ap_set_option("attn_emu")


# (64, 197, 12), (30, 257, 16): several jobs per CTA (the two-buffer tile pipeline wraps its stages / phases many times);
# (200, 50, 12), (300, 130, 2): one / two query tiles with fewer than 128 keys; 257 = 256 MMA keys + the class token as the extra key
@pytest.mark.parametrize("amp", [1.5, 6.0])
@pytest.mark.parametrize("B,S,heads", [(1, 197, 12), (3, 197, 4), (2, 257, 16), (2, 16, 2), (1, 1, 1), (2, 64, 3), (1, 272, 2),
                                       (64, 197, 12), (30, 257, 16), (200, 50, 12), (300, 130, 2), (5, 256, 3), (3, 129, 1)])
def test_attention(ctx, B, S, heads, amp):
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(B * 100 + S + heads)
    qkv = (torch.randn(B * S, 3 * D, device="cuda", generator=g) * amp).half()   # amp 6: logits of +-100, one-hot-like rows
    out = torch.full((B * S, D), float("nan"), device="cuda", dtype=torch.float16)
    ctx.check(ctx.lib.ap_attention_f16(ctx.handle, _p(qkv), _p(out), B, S, heads, _stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v          # (B, heads, S, 64)
    ref = ref.permute(0, 2, 1, 3).reshape(B * S, D)
    got = out.float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())


# variant 16: the generic run-time (two-pass) kernel instead of the register-resident instantiations; 32: the round-1/2 pipeline with
# O inside the score buffer instead of the free-standing-O pipeline (<= 208 keys); 64: no flipped second query tile; 2048 / 1024: force the key-block-unit pipeline (attention_units.cu) /
# the whole-tile pipeline (attention_tc6_kernel) whatever the sequence length; 512: two-pass softmax (exact row maximum first) instead of
# the one-pass softmax with a lazily moved reference maximum;
# emu: exponential pairs per 16 evaluated by the FMA-pipe polynomial instead of MUFU
@pytest.mark.parametrize("variant,emu", [(16, 0), (32, 0), (48, 0), (64, 0), (0, 0), (0, 4), (0, 6), (0, 8), (64, 8), (512, 0), (512, 4), (2048, 0), (2048, 4), (2112, 0), (1024, 0), (1024, 4)])
@pytest.mark.parametrize("B,S,heads,amp", [(64, 197, 12, 2.0), (30, 257, 16, 2.0), (3, 197, 4, 2.0), (200, 50, 12, 6.0), (300, 130, 2, 6.0),
                                           (150, 197, 12, 6.0), (7, 208, 3, 1.0), (9, 193, 5, 3.0), (40, 256, 4, 6.0), (33, 144, 3, 6.0)])
def test_attention_kernel_variants(ctx, B, S, heads, amp, variant, emu):
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(B + S + heads)
    qkv = (torch.randn(B * S, 3 * D, device="cuda", generator=g) * amp).half()
    out = torch.full((B * S, D), float("nan"), device="cuda", dtype=torch.float16)
    ctx.set_option("attn_variant", variant)
    ctx.set_option("attn_emu", emu)
    try:
        ctx.check(ctx.lib.ap_attention_f16(ctx.handle, _p(qkv), _p(out), B, S, heads, _stream()))
        torch.cuda.synchronize()
    finally:
        ctx.set_option("attn_variant", 0)
        ctx.set_option("attn_emu", DEFAULT_ATTN_EMU)
    q, k, v = qkv.float().view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * S, D)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())
