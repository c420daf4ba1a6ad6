"""GPU parity of the whole hot path against the committed golden vectors (outputs of the
reference's own fp32 forward, oracle/gen_golden.py) and against the CPU oracle on seeded inputs.

Gates (BASELINE.json north_star): fp32 path max|d|/max|ref| <= 1e-4 on emb and atp;
16-bit tensor-core path <= 2e-2 plus identical row-argmax on >= 99 % of map rows.

The production 16-bit path is precision="fp16" (fp16 operands, fp32 accumulate / residual /
softmax): same tcgen05 kind::f16 rate as bf16, three more mantissa bits.  Measured on BASELINE
config 1 (2DRB_1, whose redundant rows make the tied logits add up coherently): all-bf16 operands
5.1e-2 on the maps (fails the gate, as SURVEY.md 6 predicted for naive bf16), bf16 with an fp16
tied-row block 2.7e-2, fp16 5.4e-3.  So the 2e-2 gate is enforced on "fp16"; the "bf16" mode (bf16
everywhere except the fp16 tied-row block) is still shipped for range-critical checkpoints and is
gated at its own measured envelope (emb 2e-2, maps 6e-2, row-argmax >= 99 %).
"""
import os

import numpy as np
import pytest
import torch

from oracle import msa_ref as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2            # the north_star 16-bit gate; enforced on the production fp16 path
BF16_MAP_TOL = 6e-2        # envelope of the optional bf16 mode on the attention maps
PRECISIONS = ["fp32", "fp16", "bf16"]


def tols(precision, sharp_deep=False):
    """(tolerance on representations / logits, tolerance on attention maps)."""
    if precision == "fp32":
        return FP32_TOL, FP32_TOL
    if precision == "fp16":
        t = BF16_TOL_DEEP_SHARP if sharp_deep else BF16_TOL
        return t, t
    return (BF16_TOL_DEEP_SHARP if sharp_deep else BF16_TOL), max(BF16_MAP_TOL, BF16_TOL_DEEP_SHARP if sharp_deep else 0)

# Whole-model bf16 gate on SHARPENED 10-layer weight sets.  `sharpen` (our own stress knob, not a
# reference configuration) multiplies q/k so the logits span tens of nats; a random-init 10-layer
# network then amplifies any perturbation ~10x per unit of sharpen (measured on the reference
# itself, fp32 vs fp64: sharpen 2 -> 6e-6, 4 -> 3e-4, 8 -> 7e-2 on the maps; oracle/gen_golden.py).
# The north_star gate (<= 2e-2, random-init weights) is enforced on every sharpen == 1 case and,
# for sharpened weights, per layer (layer.npz at sharpen 8, mid-size 3-layer at sharpen 3) where
# the kernels are measured without the chaotic amplification; deep sharpened stacks get the
# amplification-scaled bound below plus the row-argmax agreement gate.
BF16_TOL_DEEP_SHARP = 8e-2


@pytest.fixture(scope="module")
def pkg():
    import rnamsm_b200
    return rnamsm_b200


def build(pkg, g_or_seed, layers=10, sharpen=1.0, precision="fp32", embed_positions_msa=True):
    vocab = pkg.Vocab(pkg.Alphabet())
    model = pkg.MSATransformer(vocab, num_layers=layers, embed_positions_msa=embed_positions_msa, precision=precision)
    sd = O.make_weights(int(g_or_seed), num_layers=layers, sharpen=float(sharpen),
                        embed_positions_msa=embed_positions_msa)
    model.load_state_dict(sd, strict=True)
    return model.eval().cuda(), sd


def argmax_agreement(a, b):
    a = torch.as_tensor(a).reshape(-1, a.shape[-1])
    b = torch.as_tensor(b).reshape(-1, b.shape[-1])
    return float((a.argmax(-1) == b.argmax(-1)).float().mean())


@pytest.mark.parametrize("name", ["tiny", "tiny_pad", "ragged_sharp", "single_row", "batch2_pad", "mid_sharp"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_model_vs_reference_golden(pkg, golden_dir, name, precision):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    layers = int(g["layers"])
    model, _ = build(pkg, g["wseed"], layers, g["sharpen"], precision)
    tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
    out = model(tokens, repr_layers=[0, 1, layers], need_head_weights=True)
    rows = torch.from_numpy(g["rep_rows"]).cuda()
    tol, map_tol = tols(precision, sharp_deep=float(g["sharpen"]) >= 3 and layers >= 10)
    errs = {
        "rep0": O.rel_err(out["representations"][0][:, rows].cpu(), g["rep0"]),
        "rep1": O.rel_err(out["representations"][1][:, rows].cpu(), g["rep1"]),
        "rep_last": O.rel_err(out["representations"][layers][:, rows].cpu(), g["rep_last"]),
        "logits": O.rel_err(out["logits"][:, rows].cpu(), g["logits"]),
    }
    ra = out["row_attentions"].cpu()
    if "row_attentions" in g:
        ref_ra = torch.from_numpy(g["row_attentions"])
        errs["row_attn"] = O.rel_err(ra, ref_ra)
        agree = argmax_agreement(ra, ref_ra)
    else:
        flat = ra.reshape(ra.shape[0], -1, ra.shape[-2], ra.shape[-1])
        sel = flat[:, g["row_attentions_sel_idx"]]
        errs["row_attn"] = O.rel_err(sel, g["row_attentions_sel"])
        errs["row_attn_mean"] = O.rel_err(flat.mean(1), g["row_attentions_mean"])
        agree = argmax_agreement(sel, torch.from_numpy(g["row_attentions_sel"]))
    print(f"[{name}/{precision}] {errs} argmax_agree={agree:.4f}")
    assert errs["rep0"] < 1e-5                       # K1 is fp32 in every mode
    assert max(v for k, v in errs.items() if not k.startswith("row_attn")) < tol, errs
    assert max(v for k, v in errs.items() if k.startswith("row_attn")) < map_tol, errs
    if precision != "fp32" and float(g["sharpen"]) > 1:
        assert agree >= 0.99
    if "atp" in g:
        emb, atp = pkg.extract_features(out, model.vocab, layers)
        assert emb.dtype == np.float32 and atp.dtype == np.float32
        assert emb.shape == g["emb"].shape and atp.shape == g["atp"].shape
        assert O.rel_err(emb, g["emb"]) < tol and O.rel_err(atp, g["atp"]) < map_tol


@pytest.mark.parametrize("precision", PRECISIONS)
def test_config1_2drb(pkg, golden_dir, precision):
    """BASELINE config 1: the shipped 2DRB_1 MSA (first 512 rows), emb (35,768) + atp (120,35,35)."""
    g = np.load(os.path.join(golden_dir, "2DRB_1.npz"))
    model, _ = build(pkg, g["wseed"], 10, g["sharpen"], precision)
    tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
    out = model(tokens, repr_layers=[10], need_head_weights=True, want_logits=False)
    emb, atp = pkg.extract_features(out, model.vocab, 10)
    assert emb.shape == (35, 768) and atp.shape == (120, 35, 35)
    e_emb, e_atp = O.rel_err(emb, g["emb"]), O.rel_err(atp, g["atp"])
    agree = argmax_agreement(atp, g["atp"])
    print(f"[2DRB_1/{precision}] emb {e_emb:.3e} atp {e_atp:.3e} argmax {agree:.4f}")
    tol, map_tol = tols(precision)
    assert e_emb < tol and e_atp < map_tol
    assert agree >= 0.99
    assert atp.min() >= 0 and atp.sum(-1).max() <= 1 + 1e-4


@pytest.mark.parametrize("precision", PRECISIONS)
def test_layer_vs_reference_layer(pkg, golden_dir, precision):
    """AxialTransformerLayer.forward against the top-level reference layer's vectors."""
    g = np.load(os.path.join(golden_dir, "layer.npz"))
    sd = O.make_weights(int(g["wseed"]), num_layers=1, sharpen=float(g["sharpen"]))
    layer = pkg.AxialTransformerLayer(768, 3072, 12).set_precision(precision)
    layer.load_state_dict({k[len("layers.0."):]: v for k, v in sd.items() if k.startswith("layers.0.")}, strict=True)
    layer = layer.eval().cuda()
    x = torch.from_numpy(g["x"]).cuda()
    pad = torch.from_numpy(g["pad"]).cuda()
    tol, map_tol = tols(precision)
    for tag, pm in (("nopad", None), ("pad", pad)):
        y, col, row = layer(x, self_attn_padding_mask=pm, need_head_weights=True)
        assert col is None and tuple(row.shape) == (12, 1, x.shape[1], x.shape[1])
        e = (O.rel_err(y.cpu(), g[f"y_{tag}"]), O.rel_err(row.cpu(), g[f"row_{tag}"]))
        print(f"[layer/{tag}/{precision}] x {e[0]:.3e} row {e[1]:.3e}")
        assert e[0] < tol and e[1] < map_tol
        y2 = layer(x, self_attn_padding_mask=pm)
        assert torch.equal(y2, y)          # deterministic; same result without head weights


@pytest.mark.parametrize("precision", PRECISIONS)
def test_submodules_match_fused_layer(pkg, golden_dir, precision):
    """Row / Column / FFN residual blocks called one by one == the fused layer driver."""
    g = np.load(os.path.join(golden_dir, "layer.npz"))
    sd = O.make_weights(int(g["wseed"]), num_layers=1, sharpen=float(g["sharpen"]))
    layer = pkg.AxialTransformerLayer(768, 3072, 12).set_precision(precision)
    layer.load_state_dict({k[len("layers.0."):]: v for k, v in sd.items() if k.startswith("layers.0.")}, strict=True)
    layer = layer.eval().cuda()
    x = torch.from_numpy(g["x"]).cuda()
    pad = torch.from_numpy(g["pad"]).cuda()
    y_fused, _, row_fused = layer(x, self_attn_padding_mask=pad, need_head_weights=True)
    h, row = layer.row_self_attention(x, self_attn_padding_mask=pad)
    h, col = layer.column_self_attention(h, self_attn_padding_mask=pad)
    h = layer.feed_forward_layer(h)
    tol = 1e-5 if precision == "fp32" else 1e-2     # 16-bit: the un-fused path rounds the block outputs to 16 bits
    assert O.rel_err(row.cpu(), row_fused.cpu()) < tol
    assert O.rel_err(h.cpu(), y_fused.cpu()) < tol
    with pytest.raises(NotImplementedError):
        layer.row_self_attention.layer(x, self_attn_mask=torch.ones(1, device="cuda"))


def pair_scores(maps):
    """Scores of all residue pairs i < j from the head-averaged, symmetrised, APC-corrected maps [K, L, L] (the features
    the contact head regresses on, modules.py:347-366 / utils/tensor.py:98-113)."""
    m = torch.as_tensor(np.asarray(maps), dtype=torch.float64)
    f = O.apc(O.symmetrize(m)).mean(0)
    L_ = f.shape[-1]
    iu = torch.triu_indices(L_, L_, offset=1)
    return iu, f[iu[0], iu[1]]


def top_pairs(maps, k=None):
    """Top-L residue pairs (i < j): SURVEY.md 8d's contact-pair gate."""
    iu, score = pair_scores(maps)
    order = torch.argsort(score, descending=True)[:(k or int(iu.max()) + 1)]
    return {(int(iu[0][o]), int(iu[1][o])) for o in order}


def ambiguous_pairs(ref_maps, rel_margin=5e-3):
    """How many pairs sit within rel_margin x (score range) of the reference's own top-L cut: those may legitimately
    fall on either side of it under any finite-precision arithmetic (near-uniform maps have many such ties)."""
    iu, score = pair_scores(ref_maps)
    L_ = int(iu.max()) + 1
    srt = torch.sort(score, descending=True).values
    cut = 0.5 * (srt[L_ - 1] + srt[L_])
    return int(((score - cut).abs() <= rel_margin * (srt[0] - srt[-1])).sum())


@pytest.mark.parametrize("precision,min_common", [("fp32", 1.0), ("fp16", 1.0), ("bf16", 0.97)])
def test_top_L_contact_pairs_config1(pkg, golden_dir, precision, min_common):
    """north_star: 'identical argmax contact pairs' in the 16-bit path.  BASELINE config 1 (2DRB_1): the top-L pairs of
    the exported maps must be THE SAME SET as the reference's for fp32 and for the production fp16 path; the bf16 mode
    may differ in at most one pair of 35 (the reference itself cast to bf16 loses one of 119).  The same
    pairs must come out of the device contact head (rnamsm_contact_head) fed with our maps vs the oracle's head fed
    with the reference's maps."""
    g = np.load(os.path.join(golden_dir, "2DRB_1.npz"))
    model, sd = build(pkg, g["wseed"], 10, g["sharpen"], precision)
    tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
    out = model(tokens, repr_layers=[10], need_head_weights=True, return_contacts=True, want_logits=False)
    _, atp = pkg.extract_features(out, model.vocab, 10)
    ours, ref = top_pairs(atp), top_pairs(g["atp"])
    common = len(ours & ref) / len(ref)
    print(f"[2DRB_1/{precision}] top-L pairs in common: {len(ours & ref)}/{len(ref)}")
    assert common >= min_common
    # the learned head on the device: logistic regression over the 120 APC'd maps, top-L of ITS scores
    w = sd["contact_head.regression.weight"].double().view(-1)
    b = sd["contact_head.regression.bias"].double()
    feat = O.apc(O.symmetrize(torch.as_tensor(g["atp"], dtype=torch.float64)))
    ref_c = torch.sigmoid((feat * w[:, None, None]).sum(0) + b)
    got_c = out["contacts"][0].double().cpu()
    assert got_c.shape == ref_c.shape

    def pairs(c):
        L_ = c.shape[-1]
        iu = torch.triu_indices(L_, L_, offset=1)
        o = torch.argsort(c[iu[0], iu[1]], descending=True)[:L_]
        return {(int(iu[0][k]), int(iu[1][k])) for k in o}
    assert len(pairs(got_c) & pairs(ref_c)) / ref_c.shape[-1] >= min_common


def test_top_L_contact_pairs_sharpened(pkg):
    """Same gate on the chunk-threshold-crossing 96 x 200 shape of test_mid_size_vs_oracle (sharpen 3), fp16 path against
    the fp32 oracle: the same top-L pair set except for pairs tied with the cut (random-init maps are close to uniform,
    so a few of the 19 701 pairs always sit within rounding of the L-th score)."""
    model, sd = build(pkg, 42, 3, 3.0, "fp16")
    tokens = O.make_tokens(96, 200, 8)
    ref = O.forward(sd, tokens, repr_layers=[3], need_head_weights=True, num_layers=3, want_logits=False)
    out = model(tokens.cuda(), repr_layers=[3], need_head_weights=True, want_logits=False)
    ours = top_pairs(out["row_attentions"][0, :, :, 1:, 1:].reshape(-1, 199, 199).cpu())
    want = top_pairs(ref["row_attentions"][0, :, :, 1:, 1:].reshape(-1, 199, 199))
    amb = ambiguous_pairs(ref["row_attentions"][0, :, :, 1:, 1:].reshape(-1, 199, 199))
    print(f"[mid/fp16] top-L pairs in common: {len(ours & want)}/{len(want)} ({amb} pairs within 0.5 % of the cut)")
    # identical up to the pairs the reference itself places within 0.5 % (of the score range) of its own top-L cut
    assert len(ours & want) >= len(want) - (amb + 1) // 2 and len(want) - len(ours & want) <= 3


def test_fp16_range_stress_and_watch(pkg):
    """fp16 range safety: every fp32 -> fp16 conversion in the kernels is cvt.rn.satfinite, the residual
    stream / LayerNorm statistics / softmax stay fp32.  Stress: scale the FFN's first layer so the post-GELU hidden
    reaches (a) 2.3e4 -- inside the range: the range watch must report no saturation and parity must hold -- and
    (b) 8e4 -- beyond 65504: the watch must report it, every output must stay finite (no inf / nan: the clamp
    saturates instead of overflowing) and the damage stays local: the maps' row-argmax still agrees with the UNCLAMPED
    fp32 oracle on >= 98 % of rows (the fp32 oracle with the same clamp applied agrees on 99.5 %)."""
    from rnamsm_b200 import _lib as L
    tokens = O.make_tokens(24, 40, 3)

    def run(scale):
        sd = O.make_weights(5, num_layers=2, sharpen=2.0)
        for l in range(2):
            sd[f"layers.{l}.feed_forward_layer.layer.fc1.weight"] *= scale
            sd[f"layers.{l}.feed_forward_layer.layer.fc2.weight"] /= scale      # keep the residual stream O(1)
        vocab = pkg.Vocab(pkg.Alphabet())
        model = pkg.MSATransformer(vocab, num_layers=2, precision="fp16")
        model.load_state_dict(sd, strict=True)
        model = model.eval().cuda()
        ref = O.forward(sd, tokens, repr_layers=[2], need_head_weights=True, num_layers=2, want_logits=False)
        with L.RangeWatch() as w:
            out = model(tokens.cuda(), repr_layers=[2], need_head_weights=True, want_logits=False)
        return out, ref, w

    out, ref, w = run(8000.0)
    print(f"[range] scale 8000: max |16-bit activation| {w.max_abs:.3g}, saturated {w.saturated}")
    assert w.saturated == 0 and 1.5e4 < w.max_abs < 65504
    assert O.rel_err(out["representations"][2].cpu(), ref["representations"][2]) < BF16_TOL
    assert O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"]) < BF16_TOL
    out, ref, w = run(28000.0)
    print(f"[range] scale 28000: max finite |activation| {w.max_abs:.3g}, saturated {w.saturated}")
    assert w.saturated > 0                                     # the watch sees the clamp ...
    assert torch.isfinite(out["representations"][2]).all() and torch.isfinite(out["row_attentions"]).all()
    assert argmax_agreement(out["row_attentions"].cpu(), ref["row_attentions"]) >= 0.98   # ... and it stays local


def test_check_fp16_range_falls_back_to_bf16(pkg):
    """MSATransformer.check_fp16_range: in-range weights keep fp16; weights that drive the FFN hidden past 65504 are
    detected on the first forward and the model switches to bf16, whose result is then within the bf16 envelope."""
    tokens = O.make_tokens(24, 40, 3)
    vocab = pkg.Vocab(pkg.Alphabet())
    for scale, want in ((1.0, "fp16"), (28000.0, "bf16")):
        sd = O.make_weights(5, num_layers=2, sharpen=2.0)
        for l in range(2):
            sd[f"layers.{l}.feed_forward_layer.layer.fc1.weight"] *= scale
            sd[f"layers.{l}.feed_forward_layer.layer.fc2.weight"] /= scale
        model = pkg.MSATransformer(vocab, num_layers=2, precision="fp16")
        model.load_state_dict(sd, strict=True)
        model = model.eval().cuda()
        rep = model.check_fp16_range(tokens.cuda())
        print(f"[range guard] scale {scale:g}: {rep}")
        assert rep["precision"] == want == model.precision and (rep["saturated"] > 0) == (want == "bf16")
        assert all(m.precision == want for m in model.modules() if hasattr(m, "precision"))
        out = model(tokens.cuda(), repr_layers=[2], need_head_weights=True, want_logits=False)
        ref = O.forward(sd, tokens, repr_layers=[2], need_head_weights=True, num_layers=2, want_logits=False)
        assert torch.isfinite(out["representations"][2]).all()
        assert O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"]) < BF16_TOL_DEEP_SHARP
    assert pkg.MSATransformer(vocab, num_layers=1, precision="fp32").eval().cuda().check_fp16_range(tokens.cuda())["saturated"] == 0


def test_bf16_mode_beats_the_reference_cast_to_bf16(pkg, golden_dir):
    """Why the 'bf16' mode is gated at 6e-2 and not at the 2e-2 the fp16 path meets: on BASELINE config 1 bf16 WEIGHTS
    alone (every activation exact) already put 4.4e-2 on the maps, and the reference itself cast to bf16 9.3e-2
    (tests/golden/precision_sites_2DRB_1.json, produced by tests/tools/precision_sites.py) -- no bf16-operand path can
    meet 2e-2 on this input.  Our bf16 mode must stay well below the reference's own bf16 error."""
    import json
    fx = json.load(open(os.path.join(golden_dir, "precision_sites_2DRB_1.json")))["results"]
    ref_bf16 = fx["oracle cast to bf16 (model.bfloat16())"]["atp"]
    weights_only = fx["only weights bf16"]["atp"]
    assert weights_only > BF16_TOL and ref_bf16 > BF16_TOL          # the premise, from the committed measurement
    g = np.load(os.path.join(golden_dir, "2DRB_1.npz"))
    model, _ = build(pkg, g["wseed"], 10, g["sharpen"], "bf16")
    tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
    out = model(tokens, repr_layers=[10], need_head_weights=True, want_logits=False)
    _, atp = pkg.extract_features(out, model.vocab, 10)
    e = O.rel_err(atp, g["atp"])
    print(f"[2DRB_1/bf16] atp {e:.3e} vs reference-cast-to-bf16 {ref_bf16:.3e}, bf16 weights alone {weights_only:.3e}")
    assert e < 0.5 * ref_bf16 and e < BF16_MAP_TOL


TF32X3_TOL = 1e-3      # 'tf32x3': 20x tighter than the 16-bit gate; measured values are printed


@pytest.mark.parametrize("name", ["tiny_pad", "ragged_sharp", "single_row", "mid_sharp", "2DRB_1"])
def test_tf32x3_mode_vs_reference_golden(pkg, golden_dir, name):
    """precision='tf32x3' (fp32 storage / attention / LayerNorm, nn.Linear layers as three-term tf32 products on the
    tensor cores) against the reference's own vectors: between the fp32 path (<= 1e-4) and the fp16 path (<= 2e-2)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    layers = int(g["layers"])
    model, _ = build(pkg, g["wseed"], layers, g["sharpen"], "tf32x3")
    tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
    out = model(tokens, repr_layers=[layers], need_head_weights=True)
    emb, atp = pkg.extract_features(out, model.vocab, layers)
    e_emb = O.rel_err(emb, g["emb"])
    e_atp = O.rel_err(atp, g["atp"]) if "atp" in g else float("nan")
    print(f"[{name}/tf32x3] emb {e_emb:.3e} atp {e_atp:.3e}")
    assert e_emb < TF32X3_TOL and not e_atp >= TF32X3_TOL
    assert torch.isfinite(out["logits"]).all()


def test_deep_msa_without_row_positions(pkg):
    """R > 1024 is rejected with the row-position embedding (model.py:354-359) and accepted without
    (BASELINE config 4 runs with embed_positions_msa=False); checked against the oracle at 2 layers."""
    model, sd = build(pkg, 3, 2, 2.0, "fp32", embed_positions_msa=False)
    tokens = O.make_tokens(1030, 6, 9)
    ref = O.forward(sd, tokens, repr_layers=[2], need_head_weights=True, num_layers=2, want_logits=False)
    out = model(tokens.cuda(), repr_layers=[2], need_head_weights=True, want_logits=False)
    assert O.rel_err(out["representations"][2].cpu(), ref["representations"][2]) < FP32_TOL
    assert O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"]) < FP32_TOL
    model2, _ = build(pkg, 3, 1, 1.0, "fp32")
    with pytest.raises(RuntimeError, match="1024"):
        model2(tokens.cuda())


@pytest.mark.parametrize("precision", PRECISIONS)
def test_mid_size_vs_oracle(pkg, precision):
    """A chunk-threshold-crossing shape (R*C > 16384) against the un-chunked oracle, 3 layers."""
    model, sd = build(pkg, 42, 3, 3.0, precision)
    tokens = O.make_tokens(96, 200, 8)
    ref = O.forward(sd, tokens, repr_layers=[3], need_head_weights=True, num_layers=3, want_logits=False)
    out = model(tokens.cuda(), repr_layers=[3], need_head_weights=True, want_logits=False)
    e_rep = O.rel_err(out["representations"][3].cpu(), ref["representations"][3])
    e_att = O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"])
    agree = argmax_agreement(out["row_attentions"].cpu(), ref["row_attentions"])
    print(f"[mid/{precision}] rep {e_rep:.3e} att {e_att:.3e} argmax {agree:.4f}")
    tol, map_tol = tols(precision)
    assert e_rep < tol and e_att < map_tol
    if precision != "fp32":
        assert agree >= 0.99


def test_contacts_and_api_surface(pkg):
    model, sd = build(pkg, 1, 2, 3.0, "fp32")
    tokens = O.make_tokens(5, 12, 3)
    ref = O.forward(sd, tokens, need_head_weights=True, return_contacts=True, num_layers=2)
    out = model(tokens.cuda(), return_contacts=True)
    assert set(out) == {"logits", "representations", "row_attentions", "contacts"}
    assert O.rel_err(out["contacts"].cpu(), ref["contacts"]) < 1e-4
    assert tuple(model.predict_contacts(tokens.cuda()).shape) == (1, 11, 11)
    assert tuple(model.get_sequence_attention(tokens).shape) == (1, 2, 12, 12, 12)
    model.max_tokens_per_msa_(1 << 20)
    with pytest.raises(AssertionError):
        model(tokens[0].cuda())
    with pytest.raises(RuntimeError, match="CUDA"):
        model(tokens)                    # CPU tensor: no fallback
    model.train()
    with pytest.raises(RuntimeError, match="eval"):
        model(tokens.cuda())


@pytest.mark.parametrize("pads", [0, 2])
def test_streamed_extraction_equals_forward(pkg, pads):
    """extract_features_streamed (layer-by-layer, D2H overlapped) == forward + extract_features, bit for bit."""
    model, _ = build(pkg, 4, 3, 2.0, "fp16")
    tokens = O.make_tokens(33, 41, 5, pad_cols=pads)
    out = model(tokens.cuda(), repr_layers=[3], need_head_weights=True, want_logits=False)
    emb, atp = pkg.extract_features(out, model.vocab, 3)
    atp_h = torch.empty((3 * 12, 40, 40), dtype=torch.float32).pin_memory()
    emb_h = torch.empty((40, 768), dtype=torch.float32).pin_memory()
    pkg.extract_features_streamed(model, tokens.pin_memory(), atp_h, emb_h)
    assert np.array_equal(atp_h.numpy(), atp) and np.array_equal(emb_h.numpy(), emb)


# ------------------------------------------------------------------------------ many short MSAs per pass (8f row 4)
BATCH_SHAPES = [(33, 41, 2), (7, 36, 0), (140, 20, 0), (260, 70, 0), (2, 9, 1)]     # (R, C, padded columns)


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_forward_batch_equals_single_forwards(pkg, precision):
    """forward_batch (token-local kernels once over all MSAs, attention per MSA with its own 1/sqrt(R)) is
    bit-identical to one forward per MSA -- and therefore NOT the reference's padded-batch result."""
    model, _ = build(pkg, 9, 3, 2.0, precision)
    toks = [O.make_tokens(R, C, 20 + i, pad_cols=p).cuda() for i, (R, C, p) in enumerate(BATCH_SHAPES)]
    outs = model.forward_batch(toks, need_head_weights=True)
    assert len(outs) == len(toks)
    for t, o in zip(toks, outs):
        one = model(t, repr_layers=[3], need_head_weights=True, want_logits=False)
        d_rep = (o["representations"][3] - one["representations"][3]).abs().max().item()
        d_att = (o["row_attentions"] - one["row_attentions"]).abs().max().item()
        assert o["representations"][3].shape == one["representations"][3].shape
        assert d_rep == 0.0 and d_att == 0.0, (tuple(t.shape), d_rep, d_att)
    no_maps = model.forward_batch(toks[:2])
    assert "row_attentions" not in no_maps[0]
    assert torch.equal(no_maps[1]["representations"][3], outs[1]["representations"][3])


def test_forward_batch_vs_oracle_and_routing(pkg):
    """The batch entry against the CPU oracle (per-MSA tied scaling), the one-by-one route taken by the fp32 path
    and by single-row inputs, batched extraction, and the error behaviour."""
    model, sd = build(pkg, 9, 2, 2.0, "fp16")
    shapes = [(12, 30, 0), (40, 17, 3), (1, 22, 0)]
    toks = [O.make_tokens(R, C, 40 + i, pad_cols=p) for i, (R, C, p) in enumerate(shapes)]
    outs = model.forward_batch([t.cuda() for t in toks[:2]], need_head_weights=True)
    for t, o in zip(toks, outs):
        ref = O.forward(sd, t, repr_layers=[2], need_head_weights=True, num_layers=2, want_logits=False)
        assert O.rel_err(o["representations"][2].cpu(), ref["representations"][2]) < BF16_TOL
        assert O.rel_err(o["row_attentions"].cpu(), ref["row_attentions"]) < BF16_TOL
    feats = pkg.extract_features_batch(model, toks, token_budget=600)          # forces several groups + the R = 1 route
    for t, (emb, atp) in zip(toks, feats):
        one = model(t.cuda(), repr_layers=[2], need_head_weights=True, want_logits=False)
        emb1, atp1 = pkg.extract_features(one, model.vocab, 2)
        assert np.array_equal(emb, emb1) and np.array_equal(atp, atp1)
    atp_h = [torch.empty((2 * 12, C - 1, C - 1), dtype=torch.float32).pin_memory() for _, C, _ in shapes]
    emb_h = [torch.empty((C - 1, 768), dtype=torch.float32).pin_memory() for _, C, _ in shapes]
    pkg.extract_features_batch_streamed(model, [t.pin_memory() for t in toks], atp_h, emb_h, token_budget=1100)
    for (emb, atp), a, e in zip(feats, atp_h, emb_h):
        assert np.array_equal(a.numpy(), atp) and np.array_equal(e.numpy(), emb)
    m32, _ = build(pkg, 9, 2, 2.0, "fp32")
    o32 = m32.forward_batch([toks[0].cuda()], need_head_weights=True)[0]
    one = m32(toks[0].cuda(), repr_layers=[2], need_head_weights=True, want_logits=False)
    assert torch.equal(o32["row_attentions"], one["row_attentions"])
    with pytest.raises(RuntimeError, match="CUDA"):
        model.forward_batch([toks[0]])
    assert model.forward_batch([]) == []


def test_inference_cli_writes_reference_file_formats(pkg, tmp_path):
    """python -m rnamsm_b200.inference (the hydra-free RNA_MSM_Inference.py): <id>_atp.npy (120, L, L) f32 and
    <id>_emb.npy (L, 768) f32, equal to forward + extract_features on the same tokens."""
    from rnamsm_b200 import inference
    rng = np.random.default_rng(0)
    L_, N_ = 23, 40
    msa_dir = tmp_path / "results"
    msa_dir.mkdir()
    seqs = ["".join(rng.choice(list("AGCU-"), L_)) for _ in range(N_)]
    with open(msa_dir / "toy.a2m_msa2", "w") as f:
        for i, s_ in enumerate(seqs):
            f.write(f">s{i}\n{s_}\n")
    (tmp_path / "rna_id.txt").write_text("toy\n")
    inference.main(["--root_path", str(tmp_path), "--MSA_path", "results", "--MSA_list", "rna_id.txt",
                    "--max_seqs_per_msa", "16", "--sample_method", "diversity-max"])
    atp = np.load(msa_dir / "toy_atp.npy")
    emb = np.load(msa_dir / "toy_emb.npy")
    assert atp.shape == (120, L_, L_) and atp.dtype == np.float32 and emb.shape == (L_, 768) and emb.dtype == np.float32
    assert atp.min() >= 0 and atp.sum(-1).max() <= 1 + 1e-4
    # same numbers as the module API on the tokens the ingest selected
    model, vocab = inference.build_model(None)
    tokens, rows = pkg.ingest_msa(str(msa_dir / "toy.a2m_msa2"), vocab, 16, sample_method="diversity-max")
    assert tokens.shape == (16, L_ + 1) and rows[0] == 0
    out = model(tokens.unsqueeze(0), repr_layers=[10], need_head_weights=True, want_logits=False)
    emb2, atp2 = pkg.extract_features(out, vocab, 10)
    assert np.array_equal(emb2, emb) and np.array_equal(atp2, atp)
