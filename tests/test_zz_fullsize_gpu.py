"""Parity at BASELINE.json's FULL sizes (cfg2 512x256, cfg4 4096x128, cfg5 1024x1024), where neither the CPU oracle nor
the reference can hold the problem: every kernel is run once at the full shape through the C ABI and
checked on SUB-BLOCKS against a float64 restatement of the same op on the same (already rounded) inputs -- column
attention is independent per (column, head), tied logits per head, AV per (row, head), linear / LayerNorm per token --
plus the size-independent properties of the whole forward (softmax rows sum to 1, nothing non-finite, the fp16
tensor-core path agrees with the independent fp32 FFMA path).  The sub-block restatements themselves are checked on the
CPU against the full einsum expressions that tests/test_gpu_ops.py uses (`test_subblock_references_cpu`, not a GPU test).

The file sorts last on purpose: these cases allocate tens of GB and take a few seconds each.
"""
import pytest
import torch

from oracle import msa_ref as O

H, D, HD = 12, 768, 64
F16 = 2
DENSE_RC = (1024, 1024)        # token grid of the dense-kernel case: cfg5


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------ sub-block restatements (device-agnostic)
def ref_col_attn_block(qkv_cm, R, C, c, h):
    """Column attention of column c, head h from column-major q|k|v [C*R, 3D] (token c*R + r): [R, 64] float64.
    q is already scaled (modules.py:905 is applied by the QKV epilogue)."""
    blk = qkv_cm.view(C, R, 3 * D)[c].double()
    q, k, v = (blk[:, s * D + h * HD:s * D + (h + 1) * HD] for s in range(3))
    return (q @ k.T).softmax(-1) @ v


def ref_tied_logits_head(qkv, R, C, h):
    """sum_r sum_d q[r,i,h,d] k[r,j,h,d] for one head from token-major q|k|v [R*C, 3D]: [C, C] float64."""
    t = qkv.view(R, C, 3 * D)
    q = t[:, :, h * HD:(h + 1) * HD].double().permute(1, 0, 2).reshape(C, R * HD)
    k = t[:, :, D + h * HD:D + (h + 1) * HD].double().permute(1, 0, 2).reshape(C, R * HD)
    return q @ k.T


def ref_av_block(probs_h, qkv, R, C, r, h):
    """ctx[r, :, h, :] = P[h] @ v[r, :, h, :]: [C, 64] float64 (probs_h: [C, >=C])."""
    v = qkv.view(R, C, 3 * D)[r, :, 2 * D + h * HD:2 * D + (h + 1) * HD].double()
    return probs_h[:, :C].double() @ v


def test_subblock_references_cpu():
    """The sub-block restatements above == the full einsum expressions of tests/test_gpu_ops.py (float64, CPU)."""
    R, C = 6, 5
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(R * C, 3 * D, generator=g, dtype=torch.float64) * 0.3
    t = qkv.view(R, C, 3 * D)
    q, k, v = (t[..., s * D:(s + 1) * D].reshape(R, C, H, HD) for s in range(3))
    logits = torch.einsum("rihd,rjhd->hij", q, k)
    probs = logits.softmax(-1)
    ctx_row = torch.einsum("hij,rjhd->rihd", probs, v)
    ctx_col = torch.einsum("hcij,jchd->ichd", torch.einsum("ichd,jchd->hcij", q, k).softmax(-1), v)
    qkv_cm = t.transpose(0, 1).reshape(C * R, 3 * D).contiguous()
    for h in (0, 7, 11):
        assert torch.allclose(ref_tied_logits_head(qkv, R, C, h), logits[h], rtol=1e-12, atol=1e-12)
        for r in (0, R - 1):
            assert torch.allclose(ref_av_block(probs[h], qkv, R, C, r, h), ctx_row[r, :, h], rtol=1e-12, atol=1e-12)
        for c in (0, C - 1):
            assert torch.allclose(ref_col_attn_block(qkv_cm, R, C, c, h), ctx_col[:, c, h], rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------------------- GPU, full size
@pytest.fixture(scope="module")
def L():
    from rnamsm_b200 import _lib
    _lib.device_check(torch.device("cuda:0"))
    return _lib


def need_gb(n):
    """These cases are sized for a B200 (180 GB); on a smaller or busy device they are skipped, not failed."""
    torch.cuda.empty_cache()
    free = torch.cuda.mem_get_info()[0] / 2 ** 30
    if free < n:
        pytest.skip(f"needs ~{n} GiB of free device memory, {free:.0f} GiB available")


def randn_f16(shape, seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).half()


@pytest.mark.gpu
@pytest.mark.parametrize("R,C", [(512, 256), (4096, 128), (1024, 1024)], ids=["cfg2", "cfg4", "cfg5"])
def test_column_attention_full_size(L, R, C):
    """K7 at the full shape (column-major q|k|v, the production layout): three (column, head) problems checked."""
    need_gb(24)
    qkv_cm = randn_f16((C * R, 3 * D), 31, 0.5)
    ctx = torch.empty(R * C, D, dtype=torch.float16, device="cuda")
    L.check(L.lib.rnamsm_col_attn(L.ptr(qkv_cm), R, C, H, F16, 1, None, L.ptr(ctx), L.stream_ptr()))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ctx).all())
    got = ctx.view(R, C, D)
    for c, h in ((0, 0), (C // 2 - 1, 5), (C - 1, 11)):
        ref = ref_col_attn_block(qkv_cm, R, C, c, h)
        assert rel(got[:, c, h * HD:(h + 1) * HD], ref) < 3e-3, (c, h)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C", [(512, 256), (4096, 128), (1024, 1024)], ids=["cfg2", "cfg4", "cfg5"])
def test_tied_row_attention_full_size(L, R, C):
    """K4 -> K5 -> K6 at the full shape: logits of three heads, softmax rows, AV of three (row, head) blocks."""
    need_gb(24)
    qkv = randn_f16((R * C, 3 * D), 21, 0.4)
    splits = L.lib.rnamsm_row_attn_splits(R, C, H, F16)
    partial = torch.empty(splits, H, C, C, device="cuda")
    L.check(L.lib.rnamsm_row_attn_logits(L.ptr(qkv), R, C, H, F16, L.ptr(partial), splits, L.stream_ptr()))
    logits = partial.sum(0)
    heads = (0, 7, 11)
    refs = {h: ref_tied_logits_head(qkv, R, C, h) for h in heads}
    for h in heads:
        # fp32 accumulation on the tensor cores over R*64 (up to 262 144) products in up to ~2 700 MMA steps per split:
        # allow for a truncating accumulator (<= steps * 2^-24 relative); a layout or indexing error would be O(1)
        assert rel(logits[h], refs[h]) < 1e-3, h
    scale = 1.0 / (R ** 0.5)
    pmap = torch.empty(H, C, C, device="cuda")
    ldp = (C + 7) // 8 * 8
    plp = torch.empty(H, C, ldp, dtype=torch.float16, device="cuda")
    L.check(L.lib.rnamsm_row_softmax(L.ptr(partial), splits, H, C, None, scale, L.ptr(pmap), L.ptr(plp), ldp, F16,
                                     L.stream_ptr()))
    assert float((pmap.sum(-1) - 1).abs().max()) < 1e-5 and float(pmap.min()) >= 0.0
    assert rel(pmap, (partial.double().sum(0) * scale).softmax(-1)) < 5e-5      # K5 on the logits K4 produced
    for h in heads:
        assert rel(pmap[h], (refs[h] * scale).softmax(-1)) < 1e-2, h             # ... and end to end from q, k
    assert rel(plp[..., :C], pmap) < 5e-3
    ctx = torch.empty(R * C, D, dtype=torch.float16, device="cuda")
    L.check(L.lib.rnamsm_row_attn_av(L.ptr(plp), ldp, L.ptr(qkv), R, C, H, F16, L.ptr(ctx), L.stream_ptr()))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ctx).all())
    got = ctx.view(R, C, D)
    for r, h in ((0, 0), (R // 2 + 1, 7), (R - 1, 11)):
        assert rel(got[r, :, h * HD:(h + 1) * HD], ref_av_block(plp[h], qkv, R, C, r, h)) < 2e-3, (r, h)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C", [(4096, 128), (1024, 100)], ids=["cfg4", "deep_ragged"])
def test_tied_row_attention_one_launch_full_size(L, R, C):
    """The one-launch kernel for <= 128 columns (rnamsm_row_attn_short) at BASELINE config 4's full shape and at a
    width that is neither a multiple of 16 nor of 64: float64 logits of three heads (from q, k), its own softmax on
    them, and the context of three (row, head) blocks."""
    need_gb(24)
    qkv = randn_f16((R * C, 3 * D), 21, 0.4)
    chunks = L.lib.rnamsm_row_attn_short_chunks(R, C, H)
    assert chunks >= 1
    partial = torch.empty(chunks, H, C, C, device="cuda")
    pmap = torch.empty(H, C, C, device="cuda")
    ldp = (C + 7) // 8 * 8
    plp = torch.empty(H, C, ldp, dtype=torch.float16, device="cuda")
    ctx = torch.empty(R * C, D, dtype=torch.float16, device="cuda")
    scale = 1.0 / (R ** 0.5)
    L.check(L.lib.rnamsm_row_attn_short(L.ptr(qkv), R, C, H, F16, None, scale, L.ptr(partial), chunks, L.ptr(pmap), L.ptr(plp),
                                        ldp, L.ptr(ctx), L.stream_ptr()))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ctx).all())
    heads = (0, 7, 11)
    refs = {h: ref_tied_logits_head(qkv, R, C, h) for h in heads}
    logits = partial.sum(0)
    for h in heads:
        assert rel(logits[h], refs[h]) < 1e-3, h
        assert rel(pmap[h], (refs[h] * scale).softmax(-1)) < 1e-2, h
    assert float((pmap.sum(-1) - 1).abs().max()) < 1e-5 and float(pmap.min()) >= 0.0
    assert rel(pmap, (partial.double().sum(0) * scale).softmax(-1)) < 5e-5
    assert rel(plp[..., :C], pmap) < 5e-3 and float(plp[..., C:].float().abs().sum()) == 0.0
    got = ctx.view(R, C, D)
    for r, h in ((0, 0), (R // 2 + 1, 7), (R - 1, 11)):
        assert rel(got[r, :, h * HD:(h + 1) * HD], ref_av_block(plp[h], qkv, R, C, r, h)) < 2e-3, (r, h)


@pytest.mark.gpu
def test_dense_linear_and_layernorm_full_size(L):
    """K2 / K3 / K8 over the 1 048 576 tokens of cfg5 (outputs of 6.4 GB: offsets beyond 2^32 bytes): LayerNorm with the
    token transpose, fc1 + GELU, fc2 + residual, checked on sampled tokens from the first, middle and last tiles."""
    need_gb(24)
    R, C = DENSE_RC
    T, Fdim = R * C, 3072
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((T, D), generator=g, device="cuda") * 3.0 + 0.5
    w_ln = 1 + 0.1 * torch.randn(D, generator=g, device="cuda")
    b_ln = 0.1 * torch.randn(D, generator=g, device="cuda")
    rows = torch.tensor([0, 1, 127, 128, 255, 256, T // 2 - 1, T // 2, T - 257, T - 256, T - 1], device="cuda")

    y = torch.empty(T, D, dtype=torch.float16, device="cuda")
    L.check(L.lib.rnamsm_layernorm(L.ptr(x), L.ptr(w_ln), L.ptr(b_ln), L.ptr(y), F16, T, D, O.LN_EPS, 0, 0, L.stream_ptr()))
    ref = torch.nn.functional.layer_norm(x[rows].double(), (D,), w_ln.double(), b_ln.double(), O.LN_EPS)
    assert rel(y[rows], ref) < 6e-4
    yt = torch.empty(T, D, dtype=torch.float16, device="cuda")      # token (r, c) -> row c * R + r
    L.check(L.lib.rnamsm_layernorm(L.ptr(x), L.ptr(w_ln), L.ptr(b_ln), L.ptr(yt), F16, T, D, O.LN_EPS, R, C, L.stream_ptr()))
    for r, c in ((0, 0), (0, 1), (1, 0), (R // 2 - 1, C // 2 + 1), (R - 1, 0), (R - 1, C - 1)):
        assert torch.equal(yt[c * R + r], y[r * C + c]), (r, c)
    del yt

    W1 = (torch.randn((Fdim, D), generator=g, device="cuda") * 0.05).half()
    b1 = torch.randn(Fdim, generator=g, device="cuda") * 0.1
    hbuf = torch.empty(T, Fdim, dtype=torch.float16, device="cuda")
    L.check(L.lib.rnamsm_linear(L.ptr(y), L.ptr(W1), L.ptr(b1), T, Fdim, D, F16, 1, 1.0, 0, None, L.ptr(hbuf), L.stream_ptr()))
    ref = O.gelu_erf(y[rows].double() @ W1.double().T + b1.double())
    assert rel(hbuf[rows], ref) < 2e-3

    W2 = (torch.randn((D, Fdim), generator=g, device="cuda") * 0.05).half()
    b2 = torch.randn(D, generator=g, device="cuda") * 0.1
    before = x[rows].clone()
    L.check(L.lib.rnamsm_linear(L.ptr(hbuf), L.ptr(W2), L.ptr(b2), T, D, Fdim, F16, 2, 1.0, 0, None, L.ptr(x), L.stream_ptr()))
    torch.cuda.synchronize()
    ref = before.double() + hbuf[rows].double() @ W2.double().T + b2.double()
    assert rel(x[rows], ref) < 3e-5                                  # fp32 residual stream
    assert bool(torch.isfinite(x).all())


def _model(precision, layers, embed_positions_msa=True):
    import rnamsm_b200 as pkg
    vocab = pkg.Vocab(pkg.Alphabet())
    model = pkg.MSATransformer(vocab, num_layers=layers, embed_positions_msa=embed_positions_msa, precision=precision)
    model.load_state_dict(O.make_weights(42, num_layers=layers, embed_positions_msa=embed_positions_msa), strict=True)
    return model.eval().cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,layers,epm", [(512, 256, 10, True), (4096, 128, 2, False), (1024, 1024, 2, True)],
                         ids=["cfg2-10layers", "cfg4-2layers", "cfg5-2layers"])
def test_whole_forward_full_size_fp16_vs_fp32_path(R, C, layers, epm):
    """The production fp16 tensor-core forward against the independent fp32 FFMA forward (itself pinned to the reference at
    the sizes the oracle can hold) at the full BASELINE shapes, with the 16-bit gate of the north star (2e-2), plus the
    size-independent properties of the exported maps."""
    need_gb(48)
    tokens = O.make_tokens(R, C, 11).cuda()
    out16 = _model("fp16", layers, epm)(tokens, repr_layers=[layers], need_head_weights=True, want_logits=False)
    maps16, emb16 = out16["row_attentions"], out16["representations"][layers][0, 0].clone()
    assert tuple(maps16.shape) == (1, layers, H, C, C)
    assert bool(torch.isfinite(maps16).all()) and bool(torch.isfinite(emb16).all())
    assert float(maps16.min()) >= 0.0 and float((maps16.sum(-1) - 1).abs().max()) < 1e-4
    out32 = _model("fp32", layers, epm)(tokens, repr_layers=[layers], need_head_weights=True, want_logits=False)
    assert rel(emb16, out32["representations"][layers][0, 0]) < 2e-2
    assert rel(maps16, out32["row_attentions"]) < 2e-2
