"""GPU parity, per kernel group (K1..K9), through the C ABI.  The checker is the CPU oracle
(oracle/msa_ref.py) or a float64 restatement of the single op on the same (already rounded) inputs.

Tolerances (norm-relative, max|a-b| / max|b|):
  fp32 path : 2e-5 per op (FFMA, fp32 accumulate; the whole-model gate is 1e-4)
  bf16 path : inputs are rounded to bf16 first and the checker consumes the SAME rounded inputs, so
              what remains is fp32-accumulation order plus one bf16 rounding of the output: 1e-2.
"""
import math

import numpy as np
import pytest
import torch

from oracle import msa_ref as O

pytestmark = pytest.mark.gpu

H, D = 12, 768


@pytest.fixture(scope="module")
def L():
    from rnamsm_b200 import _lib
    _lib.device_check(torch.device("cuda:0"))
    return _lib


def rel(a, b):
    return O.rel_err(a.detach().double().cpu(), b.detach().double().cpu())


def tol(code):
    return {0: 2e-5, 1: 1e-2, 2: 2e-3}[code]


def gen(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("R,C,pad_cols,pad_rows", [(5, 9, 0, 0), (8, 70, 3, 2), (3, 130, 0, 1), (33, 64, 64 - 1, 0)])
def test_embed_layernorm(L, R, C, pad_cols, pad_rows):
    sd = O.make_weights(3, num_layers=1)
    tokens = O.make_tokens(R, C, 7, pad_cols=pad_cols, pad_rows=pad_rows)
    # sprinkle pads in the middle of rows too: positions must skip them (modules.py:286-291)
    if pad_cols or pad_rows:
        tokens[0, 1, 2] = O.PAD_IDX
    ref, pm = O.embed(sd, tokens)
    dev = "cuda"
    x = torch.empty(R * C, D, device=dev)
    pad = torch.empty(R * C, dtype=torch.uint8, device=dev)
    cu = {k: sd[k].to(dev) for k in ("embed_tokens.weight", "embed_positions.weight", "msa_position_embedding",
                                      "emb_layer_norm_before.weight", "emb_layer_norm_before.bias")}
    rp = cu["msa_position_embedding"].reshape(-1).contiguous()
    tok_dev = tokens[0].to(dev)
    L.check(L.lib.rnamsm_embed_layernorm(L.ptr(tok_dev), R, C, L.ptr(cu["embed_tokens.weight"]), O.VOCAB,
                                         L.ptr(cu["embed_positions.weight"]), cu["embed_positions.weight"].shape[0],
                                         L.ptr(rp), L.ptr(cu["emb_layer_norm_before.weight"]),
                                         L.ptr(cu["emb_layer_norm_before.bias"]), D, O.PAD_IDX, O.LN_EPS, L.ptr(x),
                                         L.ptr(pad), L.stream_ptr()))
    assert rel(x.view(R, C, D), ref[0]) < 2e-6
    want_pad = tokens[0].eq(O.PAD_IDX).reshape(-1).to(torch.uint8)
    assert torch.equal(pad.cpu(), want_pad)
    if pm is not None:
        assert float(x.view(R, C, D)[tokens[0].eq(O.PAD_IDX).to(dev)].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("rows", [1, 7, 1000])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_layernorm(L, rows, code):
    x = gen((rows, D), 1, 3.0) + 0.5
    w, b = 1 + 0.1 * gen((D,), 2), 0.1 * gen((D,), 3)
    ref = O.layer_norm(x.double(), w.double(), b.double())
    y = torch.empty(rows, D, dtype=L.torch_dtype(code), device="cuda")
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()      # named: a temporary would be freed before the launch
    L.check(L.lib.rnamsm_layernorm(L.ptr(xd), L.ptr(wd), L.ptr(bd), L.ptr(y), code, rows, D, O.LN_EPS, 0, 0,
                                   L.stream_ptr()))
    assert rel(y, ref) < (2e-6 if code == 0 else 5e-3 if code == 1 else 6e-4)
    if rows == 1000:       # token transpose on the way out: row r*C + c -> row c*R + r
        R_, C_ = 40, 25
        yt = torch.empty(rows, D, dtype=L.torch_dtype(code), device="cuda")
        L.check(L.lib.rnamsm_layernorm(L.ptr(xd), L.ptr(wd), L.ptr(bd), L.ptr(yt), code, rows, D, O.LN_EPS, R_, C_,
                                       L.stream_ptr()))
        assert torch.equal(yt.view(C_, R_, D), y.view(R_, C_, D).transpose(0, 1))


# ------------------------------------------------------------------------------------------ K3/K6/K8
def run_linear(L, x, W, bias, code, epi, q_scale=1.0, q_cols=0, row_mask=None, out=None):
    M, K = x.shape
    N = W.shape[0]
    dt = L.torch_dtype(code)
    xd, Wd, bd = x.to(dt).cuda(), W.to(dt).cuda(), bias.cuda()
    if out is None:
        out = torch.empty(M, N, dtype=dt, device="cuda")
    L.check(L.lib.rnamsm_linear(L.ptr(xd), L.ptr(Wd), L.ptr(bd), M, N, K, code, epi, q_scale, q_cols,
                                L.ptr(row_mask), L.ptr(out), L.stream_ptr()))
    return out, xd.double().cpu(), Wd.double().cpu()


@pytest.mark.parametrize("M,N,K", [(1, 768, 768), (100, 2304, 768), (128, 768, 768), (300, 768, 3072),
                                   (1000, 3072, 768), (257, 768, 768)])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_linear_bias_and_qscale(L, M, N, K, code):
    x, W, bias = gen((M, K), 1), gen((N, K), 2, 0.05), gen((N,), 3, 0.1)
    mask = (torch.rand(M, generator=torch.Generator().manual_seed(4)) < 0.2).to(torch.uint8)
    q_cols = 768 if N == 2304 else 0
    out, xr, Wr = run_linear(L, x, W, bias, code, 0, 0.37, q_cols, mask.cuda() if q_cols else None)
    ref = xr @ Wr.T + bias.double()
    if q_cols:
        ref[:, :q_cols] *= 0.37
        ref[mask.bool(), :q_cols] = 0
    assert rel(out, ref) < tol(code)


@pytest.mark.parametrize("M,N,K", [(130, 3072, 768), (64, 768, 768)])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_linear_gelu(L, M, N, K, code):
    x, W, bias = gen((M, K), 5), gen((N, K), 6, 0.08), gen((N,), 7, 0.1)
    out, xr, Wr = run_linear(L, x, W, bias, code, 1)
    ref = O.gelu_erf(xr @ Wr.T + bias.double())
    assert rel(out, ref) < tol(code)


@pytest.mark.parametrize("M,N,K", [(200, 768, 3072), (129, 768, 768)])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_linear_residual(L, M, N, K, code):
    x, W, bias = gen((M, K), 8), gen((N, K), 9, 0.05), gen((N,), 10, 0.1)
    resid = gen((M, N), 11)
    out = resid.clone().cuda()
    out, xr, Wr = run_linear(L, x, W, bias, code, 2, out=out)
    ref = resid.double() + xr @ Wr.T + bias.double()
    assert out.dtype == torch.float32
    assert rel(out, ref) < (2e-5 if code == 0 else 1e-5)   # fp32 output in both modes


@pytest.mark.parametrize("M,N,K,epi", [(1, 768, 768, 0), (300, 2304, 768, 0), (1000, 3072, 768, 1), (257, 768, 3072, 2),
                                       (70000, 768, 768, 2)])
def test_linear_tf32_split(L, M, N, K, epi):
    """The 'tf32x3' mode's tensor-core linear (three kind::tf32 MMAs on hi / lo operand halves) against float64 on the
    SAME fp32 inputs, every epilogue, tails, q scale / row mask.  The operand split is exact to 2^-23; the remaining
    error is the tensor core's truncating fp32 accumulation (a few 1e-6 norm-relative, ~10x the FFMA kernel's, 1000x
    below the fp16 path's) -- hence a fast high-precision mode, not the <= 1e-4 whole-model parity path."""
    x, W, bias = gen((M, K), 21, 2.0), gen((N, K), 22, 0.05), gen((N,), 23, 0.1)
    mask = (torch.rand(M, generator=torch.Generator().manual_seed(4)) < 0.2).to(torch.uint8)
    q_cols = 768 if N == 2304 else 0
    xd, Wd, bd = x.cuda(), W.cuda(), bias.cuda()
    resid = gen((M, N), 24)
    out = resid.clone().cuda() if epi == 2 else torch.empty(M, N, device="cuda")
    nb = L.lib.rnamsm_linear_tf32_scratch_bytes(M, N, K)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    md = mask.cuda()
    L.check(L.lib.rnamsm_linear_tf32(L.ptr(xd), L.ptr(Wd), L.ptr(bd), M, N, K, epi, 0.37, q_cols, L.ptr(md) if q_cols else None,
                                     L.ptr(out), L.ptr(scratch), nb, L.stream_ptr()))
    ref = x.double() @ W.double().T + bias.double()
    if epi == 0 and q_cols:
        ref[:, :q_cols] *= 0.37
        ref[mask.bool(), :q_cols] = 0
    elif epi == 0:
        pass
    elif epi == 1:
        ref = O.gelu_erf(ref)
    else:
        ref = resid.double() + ref
    e = rel(out, ref)
    # the FFMA kernel on the same inputs, for scale
    out2 = resid.clone().cuda() if epi == 2 else torch.empty(M, N, device="cuda")
    L.check(L.lib.rnamsm_linear(L.ptr(xd), L.ptr(Wd), L.ptr(bd), M, N, K, 0, epi, 0.37, q_cols, L.ptr(md) if q_cols else None,
                                L.ptr(out2), L.stream_ptr()))
    print(f"[tf32x3 {M}x{N}x{K} epi {epi}] err {e:.2e} (FFMA kernel {rel(out2, ref):.2e})")
    assert e < (1.5e-5 if K <= 768 else 5e-5)       # grows with the length of the accumulation chain (K / 8 steps x 3)


@pytest.mark.parametrize("M,K,tr", [(129, 768, None), (256, 768, None), (1000, 3072, (40, 25)), (3000, 768, (30, 100)),
                                    (40000, 768, None)])
@pytest.mark.parametrize("code,ycode", [(2, 2), (1, 1), (2, 1)], ids=["f16", "bf16", "f16_to_bf16"])
def test_linear_residual_layernorm(L, M, K, tr, code, ycode):
    """The fused NormalizedResidualBlock epilogue: the residual stream must equal the plain residual GEMM's bit for bit,
    and the LayerNorm rows must equal the stand-alone LayerNorm kernel's on that stream bit for bit (same arithmetic),
    including the transposed output order, tails (M % 256 != 0) and more m-blocks than CTA pairs (40000 rows)."""
    N = D
    x, W, bias = gen((M, K), 8), gen((N, K), 9, 0.05), gen((N,), 10, 0.1)
    resid = gen((M, N), 11)
    lw, lb = (1 + 0.1 * gen((N,), 12)).cuda(), (0.1 * gen((N,), 13)).cuda()
    dt = L.torch_dtype(code)
    xd, Wd, bd = x.to(dt).cuda(), W.to(dt).cuda(), bias.cuda()
    plain = resid.clone().cuda()
    L.check(L.lib.rnamsm_linear(L.ptr(xd), L.ptr(Wd), L.ptr(bd), M, N, K, code, 2, 1.0, 0, None, L.ptr(plain), L.stream_ptr()))
    y_plain = torch.empty(M, N, dtype=L.torch_dtype(ycode), device="cuda")
    trR, trC = tr if tr else (0, 0)
    L.check(L.lib.rnamsm_layernorm(L.ptr(plain), L.ptr(lw), L.ptr(lb), L.ptr(y_plain), ycode, M, N, O.LN_EPS, trR, trC,
                                   L.stream_ptr()))
    fused = resid.clone().cuda()
    y = torch.full((M, N), float("nan"), dtype=L.torch_dtype(ycode), device="cuda")
    counters = torch.zeros(2 * ((M + 255) // 256), dtype=torch.int32, device="cuda")
    for _ in range(2):                         # twice on the same counters: the kernel must leave them zero
        fused.copy_(resid)
        L.check(L.lib.rnamsm_linear_residual_layernorm(L.ptr(xd), L.ptr(Wd), L.ptr(bd), M, N, K, code, L.ptr(fused), L.ptr(lw),
                                                       L.ptr(lb), O.LN_EPS, L.ptr(y), ycode, trR, trC, L.ptr(counters),
                                                       L.stream_ptr()))
    torch.cuda.synchronize()
    assert int(counters.abs().sum()) == 0
    ref = resid.double() + xd.double().cpu() @ Wd.double().cpu().T + bias.double()
    assert rel(fused, ref) < 1e-5
    assert torch.equal(fused, plain)
    assert torch.equal(y, y_plain)
    ln_ref = O.layer_norm(ref, lw.double().cpu(), lb.double().cpu())
    if tr:
        ln_ref = ln_ref.view(trR, trC, N).transpose(0, 1).reshape(M, N)
    assert rel(y, ln_ref) < (5e-3 if ycode == 1 else 6e-4)


# ------------------------------------------------------------------------------------------ K4/K5/K6
def make_qkv(R, C, seed, code, L, scale=1.0):
    dt = L.torch_dtype(code)
    qkv = (gen((R, C, 3 * D), seed) * scale).to(dt).cuda()
    return qkv, qkv.double().cpu()


@pytest.mark.parametrize("R,C", [(4, 9), (7, 36), (33, 130), (64, 257), (20, 300)])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_row_attention_chain(L, R, C, code):
    """K4 (split-K tied logits) -> K5 (softmax + key mask) -> K6 (AV)."""
    qkv, q64 = make_qkv(R, C, 21, code, L, 0.4)
    q = q64[..., :D].view(R, C, H, 64)
    k = q64[..., D:2 * D].view(R, C, H, 64)
    v = q64[..., 2 * D:].view(R, C, H, 64)
    logits_ref = torch.einsum("rihd,rjhd->hij", q, k)
    for n_splits in sorted({1, L.lib.rnamsm_row_attn_splits(R, C, H, code), min(R, 3)}):
        if (n_splits - 1) * math.ceil(R / n_splits) >= R:
            continue                                   # would leave an empty split: rejected by the ABI
        splits = n_splits
        partial = torch.empty(splits, H, C, C, device="cuda")
        L.check(L.lib.rnamsm_row_attn_logits(L.ptr(qkv), R, C, H, code, L.ptr(partial), splits, L.stream_ptr()))
        assert rel(partial.sum(0), logits_ref) < 2e-5, f"splits={splits}"
    key_pad = torch.zeros(C, dtype=torch.uint8)
    key_pad[-2:] = 1
    logit_scale = 0.37
    masked = (logits_ref * logit_scale).float().masked_fill(key_pad.bool()[None, None, :], -10000)
    probs_ref = masked.double().softmax(-1)
    pmap = torch.empty(H, C, C, device="cuda")
    ldp = (C + 7) // 8 * 8 if code != 0 else C
    plp = torch.full((H, C, ldp), 7.0, dtype=L.torch_dtype(code), device="cuda") if code != 0 else None
    key_pad_dev = key_pad.cuda()
    L.check(L.lib.rnamsm_row_softmax(L.ptr(partial), splits, H, C, L.ptr(key_pad_dev), logit_scale, L.ptr(pmap),
                                     L.ptr(plp), ldp, code, L.stream_ptr()))
    assert rel(pmap, probs_ref) < 2e-5
    if plp is not None:
        assert rel(plp[..., :C], probs_ref) < 5e-3
        assert float(plp[..., C:].abs().sum()) == 0.0            # zero-filled padding columns
    p_in = plp if plp is not None else pmap
    ctx = torch.empty(R * C, D, dtype=L.torch_dtype(code), device="cuda")
    L.check(L.lib.rnamsm_row_attn_av(L.ptr(p_in), ldp, L.ptr(qkv), R, C, H, code, L.ptr(ctx), L.stream_ptr()))
    ctx_ref = torch.einsum("hij,rjhd->rihd", p_in[..., :C].double().cpu(), v).reshape(R * C, D)
    assert rel(ctx, ctx_ref) < tol(code)


@pytest.mark.parametrize("R,C,with_pad", [(2, 5, False), (7, 36, True), (33, 16, False), (64, 48, True), (130, 64, True),
                                          (512, 36, False), (100, 65, True), (61, 100, False), (300, 128, True),
                                          (1, 20, False), (515, 127, True)])
@pytest.mark.parametrize("code", [1, 2], ids=["bf16", "f16"])
def test_row_attention_short(L, R, C, with_pad, code):
    """K4 + K5 + K6 in one cooperative launch (C <= 128) against float64, and against the three-kernel chain."""
    qkv, q64 = make_qkv(R, C, 23, code, L, 0.4)
    q = q64[..., :D].view(R, C, H, 64)
    k = q64[..., D:2 * D].view(R, C, H, 64)
    v = q64[..., 2 * D:].view(R, C, H, 64)
    logits_ref = torch.einsum("rihd,rjhd->hij", q, k)
    key_pad = torch.zeros(C, dtype=torch.uint8)
    if with_pad:
        key_pad[-2:] = 1
    logit_scale = 1.0 / math.sqrt(R)
    masked = (logits_ref * logit_scale).float().masked_fill(key_pad.bool()[None, None, :], -10000)
    probs_ref = masked.double().softmax(-1)
    chunks = L.lib.rnamsm_row_attn_short_chunks(R, C, H)
    assert 1 <= chunks <= R and chunks * H <= 148
    ldp = (C + 7) // 8 * 8
    key_pad_dev = key_pad.cuda() if with_pad else None
    dt = L.torch_dtype(code)
    for n_chunks in sorted({chunks, 1, min(R, 5)}):
        if (n_chunks - 1) * math.ceil(R / n_chunks) >= R:
            continue
        partial = torch.full((n_chunks, H, C, C), float("nan"), device="cuda")
        pmap = torch.full((H, C, C), float("nan"), device="cuda")
        plp = torch.full((H, C, ldp), 7.0, dtype=dt, device="cuda")
        ctx = torch.full((R * C, D), float("nan"), dtype=dt, device="cuda")
        for _ in range(2):                              # twice: the barrier counter must come back to zero
            L.check(L.lib.rnamsm_row_attn_short(L.ptr(qkv), R, C, H, code, L.ptr(key_pad_dev), logit_scale, L.ptr(partial),
                                                n_chunks, L.ptr(pmap), L.ptr(plp), ldp, L.ptr(ctx), L.stream_ptr()))
        torch.cuda.synchronize()
        # (one chunk = one tensor-core accumulator over all R x 64 products: its truncating fp32 adds reach 3e-5 at R = 512)
        assert rel(partial.sum(0), logits_ref) < (2e-5 if R // n_chunks <= 128 else 6e-5), f"chunks={n_chunks}"
        assert rel(pmap, probs_ref) < (2e-5 if R // n_chunks <= 128 else 1e-4)
        assert rel(plp[..., :C], probs_ref) < 5e-3
        assert float(plp[..., C:].float().abs().sum()) == 0.0
        ctx_ref = torch.einsum("hij,rjhd->rihd", plp[..., :C].double().cpu(), v).reshape(R * C, D)
        assert rel(ctx, ctx_ref) < tol(code)
    # the three-kernel chain on the same partial sums agrees to fp32 rounding (its softmax is compiled separately)
    pmap2 = torch.empty(H, C, C, device="cuda")
    plp2 = torch.empty(H, C, ldp, dtype=dt, device="cuda")
    L.check(L.lib.rnamsm_row_softmax(L.ptr(partial), n_chunks, H, C, L.ptr(key_pad_dev), logit_scale, L.ptr(pmap2),
                                     L.ptr(plp2), ldp, code, L.stream_ptr()))
    assert rel(pmap, pmap2) < 1e-6 and rel(plp, plp2) < 5e-3
    ctx2 = torch.empty(R * C, D, dtype=dt, device="cuda")
    L.check(L.lib.rnamsm_row_attn_av(L.ptr(plp2), ldp, L.ptr(qkv), R, C, H, code, L.ptr(ctx2), L.stream_ptr()))
    assert rel(ctx, ctx2.double().cpu()) < 1e-3


def test_row_attention_short_on_two_streams(L):
    """Launches of the one-launch kernel that overlap in time (two streams) take different barrier counters from the
    per-device ring: both streams must reproduce the result of a lone launch, bit for bit."""
    R, C, code = 96, 36, 2
    dt = L.torch_dtype(code)
    ldp = (C + 7) // 8 * 8
    chunks = L.lib.rnamsm_row_attn_short_chunks(R, C, H)
    sets = []
    for seed in (41, 42):
        qkv, _ = make_qkv(R, C, seed, code, L, 0.4)
        sets.append(dict(qkv=qkv, partial=torch.empty(chunks, H, C, C, device="cuda"), pmap=torch.empty(H, C, C, device="cuda"),
                         plp=torch.empty(H, C, ldp, dtype=dt, device="cuda"), ctx=torch.empty(R * C, D, dtype=dt, device="cuda")))

    def launch(b, stream):
        L.check(L.lib.rnamsm_row_attn_short(L.ptr(b["qkv"]), R, C, H, code, None, 1.0 / math.sqrt(R), L.ptr(b["partial"]), chunks,
                                            L.ptr(b["pmap"]), L.ptr(b["plp"]), ldp, L.ptr(b["ctx"]), stream))

    want = []
    for b in sets:
        launch(b, L.stream_ptr())
        torch.cuda.synchronize()
        want.append((b["pmap"].clone(), b["ctx"].clone()))
        b["pmap"].zero_(); b["ctx"].zero_()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(40):
        for b, st in zip(sets, streams):
            launch(b, st.cuda_stream)
    torch.cuda.synchronize()
    for b, (pm, cx) in zip(sets, want):
        assert torch.equal(b["pmap"], pm) and torch.equal(b["ctx"], cx)


# ------------------------------------------------------------------------------------------ K7
@pytest.mark.parametrize("R,C,with_pad", [(2, 5, False), (9, 7, True), (64, 3, False), (65, 4, True), (130, 6, True),
                                          (300, 2, False), (257, 3, True), (300, 7, True), (520, 13, False),
                                          (1030, 2, True)])
@pytest.mark.parametrize("code", [0, 1, 2], ids=["f32", "bf16", "f16"])
def test_column_attention(L, R, C, with_pad, code):
    qkv, q64 = make_qkv(R, C, 31, code, L, 0.5)
    q = q64[..., :D].view(R, C, H, 64)
    k = q64[..., D:2 * D].view(R, C, H, 64)
    v = q64[..., 2 * D:].view(R, C, H, 64)
    att = torch.einsum("ichd,jchd->hcij", q, k)
    pad = None
    if with_pad:
        pad = torch.zeros(R, C, dtype=torch.bool)
        pad[-1, :] = True
        pad[R // 2, 0] = True
        pad[:, C - 1] = True                      # a fully padded column -> uniform attention
        att = att.masked_fill(pad.T[None, :, None, :], -10000)
    ctx_ref = torch.einsum("hcij,jchd->ichd", att.softmax(-1), v).reshape(R * C, D)
    ctx = torch.empty(R * C, D, dtype=L.torch_dtype(code), device="cuda")
    pad_u8 = pad.to(torch.uint8).cuda() if pad is not None else None
    L.check(L.lib.rnamsm_col_attn(L.ptr(qkv), R, C, H, code, 0, L.ptr(pad_u8), L.ptr(ctx), L.stream_ptr()))
    assert rel(ctx, ctx_ref) < tol(code)
    if code != 0:          # same problem with q|k|v handed over column-major [C, R, 3D]; ctx stays token-major
        qkv_t = qkv.transpose(0, 1).contiguous()
        ctx2 = torch.empty_like(ctx)
        L.check(L.lib.rnamsm_col_attn(L.ptr(qkv_t), R, C, H, code, 1, L.ptr(pad_u8), L.ptr(ctx2), L.stream_ptr()))
        assert torch.equal(ctx2, ctx)


# ------------------------------------------------------------------------------------------ K9b
def test_vocab_proj(L):
    h, E, b = gen((77, D), 1), gen((12, D), 2), gen((12,), 3)
    out = torch.empty(77, 12, device="cuda")
    hd, Ed, bd = h.cuda(), E.cuda(), b.cuda()
    L.check(L.lib.rnamsm_vocab_proj(L.ptr(hd), L.ptr(Ed), L.ptr(bd), 77, 12, D, L.ptr(out), L.stream_ptr()))
    assert rel(out, h.double() @ E.double().T + b.double()) < 2e-6


def test_error_paths(L):
    x = torch.zeros(4, 100, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 128"):
        L.check(L.lib.rnamsm_layernorm(L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), 0, 4, 100, 1e-5, 0, 0, L.stream_ptr()), "ln")
    with pytest.raises(RuntimeError, match="R=1"):
        L.check(L.lib.rnamsm_col_attn(L.ptr(x), 1, 4, 12, 0, 0, None, L.ptr(x), L.stream_ptr()), "col")


# ------------------------------------------------------------------------------------------ contact head (8f row 2)
@pytest.mark.parametrize("C,K", [(9, 24), (36, 120), (70, 120), (131, 120)])
def test_contact_head_kernel(L, C, K):
    """rnamsm_contact_head vs the oracle's symmetrize + APC + logistic regression (modules.py:347-366)."""
    g = torch.Generator().manual_seed(5)
    maps = torch.rand(1, K // 12, 12, C, C, generator=g).softmax(-1)
    sd = {"contact_head.regression.weight": gen((1, K), 6), "contact_head.regression.bias": gen((1,), 7)}
    ref = O.contact_head({k: v.double() for k, v in sd.items()}, None, maps.double())[0]
    Ls = C - 1
    md = maps[0].reshape(K, C, C).contiguous().cuda()
    w, b = sd["contact_head.regression.weight"].reshape(-1).cuda(), sd["contact_head.regression.bias"].cuda()
    out = torch.empty(Ls, Ls, device="cuda")
    ws = torch.empty(K * Ls + K, device="cuda")
    L.check(L.lib.rnamsm_contact_head(L.ptr(md), K, C, 1, Ls, L.ptr(w), L.ptr(b), L.ptr(out), L.ptr(ws), L.stream_ptr()))
    assert rel(out, ref) < 1e-5
    assert torch.equal(out, out.T) or rel(out, out.T) < 1e-6      # contacts are symmetric


# ------------------------------------------------------------------------------------------ SS input packing (8f row 3)
def test_ss_input_packing_matches_reference(golden_dir):
    """pack_ss_input vs the [1,128,L,L] tensor the reference's own SS pre-processing builds (oracle/gen_golden_ss.py)."""
    import os
    import rnamsm_b200 as pkg
    g = np.load(os.path.join(golden_dir, "ss_pack.npz"))
    seq, atp, want = str(g["seq"]), g["atp"], g["x"]
    Ls = len(seq)
    ra = torch.zeros(1, 10, 12, Ls + 1, Ls + 1)
    ra[0, :, :, 1:, 1:] = torch.from_numpy(atp).view(10, 12, Ls, Ls)       # maps with the BOS row / column back in
    x = pkg.pack_ss_input(ra.cuda(), seq)
    assert tuple(x.shape) == want.shape == (1, 128, Ls, Ls) and x.dtype == torch.float32
    assert torch.equal(x.cpu(), torch.from_numpy(want))                    # bit-exact: copies and 0/1


# ------------------------------------------------------------------------------------------ RSA input packing (8f row 4)
def test_rsa_input_packing_matches_reference(golden_dir):
    """pack_rsa_input vs the [1,773,L] tensor obtained by executing the reference's own packing statements
    (oracle/gen_golden_rsa.py), bit-exact; and the embedding-only variant vs the oracle."""
    import os
    import rnamsm_b200 as pkg
    g = np.load(os.path.join(golden_dir, "rsa_pack.npz"))
    seq, emb, want = str(g["seq"]), g["emb"], g["x"]
    Ls = len(seq)
    rep = torch.zeros(1, 3, Ls + 1, 768)
    rep[0, 0, 1:] = torch.from_numpy(emb)                                   # MSA row 0 with the BOS column back in
    rep[0, 1:] = 7.0                                                        # other rows must not be read
    x = pkg.pack_rsa_input(rep.cuda(), seq, g["mu_emb"], g["std_emb"], g["mu_oh"], g["std_oh"])
    assert tuple(x.shape) == want.shape == (1, 773, Ls) and x.dtype == torch.float32
    assert torch.equal(x.cpu(), torch.from_numpy(want))
    x0 = pkg.pack_rsa_input(rep.cuda(), seq, g["mu_emb"], g["std_emb"])
    assert torch.equal(x0.cpu(), torch.from_numpy(O.rsa_input(emb, seq, g["mu_emb"], g["std_emb"])))
    with pytest.raises(ValueError):
        pkg.pack_rsa_input(rep.cuda(), seq[:-1], g["mu_emb"], g["std_emb"])

