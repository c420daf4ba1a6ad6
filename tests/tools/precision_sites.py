"""Diagnostic (CPU, no GPU needed): WHICH 16-bit roundings put the error on the exported maps of BASELINE config 1
(2DRB_1, 512 redundant rows)?  Emulates the 16-bit pipeline of csrc/ (fp32 accumulate, fp32 residual stream /
LayerNorm statistics / softmax) on the CPU with the oracle's functions and rounds one group of tensors at a time
to bf16 or fp16.  Also runs the oracle with EVERYTHING cast to bf16 (what `model.bfloat16()` would do to the
reference itself).  Lives under tests/ because it uses the oracle; numbers are quoted in DESIGN.md section 2.

    python tests/tools/precision_sites.py [--rows 512]
"""
import argparse
import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import msa_ref as O  # noqa: E402

SITES = ["w_row", "xn_row", "qk_row", "v_row", "probs_row", "ctx_row",
         "w_col", "xn_col", "qkv_col", "p_col", "ctx_col", "w_ffn", "xn_ffn", "hidden"]


def make_round(cfg):
    """cfg: {site: torch dtype}; returns rd(site, tensor) that rounds through that dtype (or passes through)."""
    def rd(site, t):
        dt = cfg.get(site)
        return t if dt is None else t.to(dt).float()
    return rd


def forward_emulated(sd, tokens, rd):
    """Oracle forward (un-chunked, fp32) with roundings at the tensors the CUDA path stores in 16 bits."""
    x, pm = O.embed(sd, tokens)
    x = x.permute(1, 2, 0, 3).contiguous()
    R, C, B, D = x.shape
    H, d = O.NUM_HEADS, D // O.NUM_HEADS
    maps = []
    for l in range(O.NUM_LAYERS):
        p = f"layers.{l}."
        # ---- row block
        pfx = p + "row_self_attention.layer."
        xn = rd("xn_row", O.layer_norm(x, sd[p + "row_self_attention.layer_norm.weight"], sd[p + "row_self_attention.layer_norm.bias"]))
        W = lambda n: rd("w_row", sd[pfx + n + ".weight"])
        q = rd("qk_row", (F.linear(xn, W("q_proj"), sd[pfx + "q_proj.bias"]) * (d ** -0.5))).view(R, C, B, H, d)
        k = rd("qk_row", F.linear(xn, W("k_proj"), sd[pfx + "k_proj.bias"])).view(R, C, B, H, d)
        v = rd("v_row", F.linear(xn, W("v_proj"), sd[pfx + "v_proj.bias"])).view(R, C, B, H, d)
        attn = torch.einsum("rinhd,rjnhd->hnij", q, k) / math.sqrt(R)
        probs = attn.softmax(-1)
        maps.append(probs)
        ctx = rd("ctx_row", torch.einsum("hnij,rjnhd->rinhd", rd("probs_row", probs), v).contiguous().view(R, C, B, D))
        x = x + F.linear(ctx, W("out_proj"), sd[pfx + "out_proj.bias"])
        # ---- column block
        pfx = p + "column_self_attention.layer."
        xn = rd("xn_col", O.layer_norm(x, sd[p + "column_self_attention.layer_norm.weight"], sd[p + "column_self_attention.layer_norm.bias"]))
        W = lambda n: rd("w_col", sd[pfx + n + ".weight"])
        q = rd("qkv_col", F.linear(xn, W("q_proj"), sd[pfx + "q_proj.bias"]) * (d ** -0.5)).view(R, C, B, H, d)
        k = rd("qkv_col", F.linear(xn, W("k_proj"), sd[pfx + "k_proj.bias"])).view(R, C, B, H, d)
        v = rd("qkv_col", F.linear(xn, W("v_proj"), sd[pfx + "v_proj.bias"])).view(R, C, B, H, d)
        pc = torch.einsum("icnhd,jcnhd->hcnij", q, k).softmax(-1)
        ctx = rd("ctx_col", torch.einsum("hcnij,jcnhd->icnhd", rd("p_col", pc), v).contiguous().view(R, C, B, D))
        del pc
        x = x + F.linear(ctx, W("out_proj"), sd[pfx + "out_proj.bias"])
        # ---- FFN
        pfx = p + "feed_forward_layer.layer."
        xn = rd("xn_ffn", O.layer_norm(x, sd[p + "feed_forward_layer.layer_norm.weight"], sd[p + "feed_forward_layer.layer_norm.bias"]))
        h = rd("hidden", F.gelu(F.linear(xn, rd("w_ffn", sd[pfx + "fc1.weight"]), sd[pfx + "fc1.bias"])))
        x = x + F.linear(h, rd("w_ffn", sd[pfx + "fc2.weight"]), sd[pfx + "fc2.bias"])
    x = O.layer_norm(x, sd["emb_layer_norm_after.weight"], sd["emb_layer_norm_after.bias"])
    att = torch.stack([m.permute(1, 0, 2, 3) for m in maps], 1)       # [B, N, H, C, C]
    emb = x.permute(2, 0, 1, 3)[0, 0, 1:].numpy()
    atp = att[0, :, :, 1:, 1:].reshape(-1, C - 1, C - 1).numpy()
    return emb, atp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=512)
    a = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "2DRB_1.npz"))
    tokens = torch.from_numpy(g["tokens"].astype(np.int64))[:, :a.rows]
    sd = O.make_weights(int(g["wseed"]), sharpen=float(g["sharpen"]))
    torch.set_grad_enabled(False)
    t0 = time.time()
    ref_emb, ref_atp = forward_emulated(sd, tokens, make_round({}))
    if a.rows == 512:
        print(f"emulation with no rounding vs committed golden: emb {O.rel_err(ref_emb, g['emb']):.2e} atp {O.rel_err(ref_atp, g['atp']):.2e}"
              f" ({time.time() - t0:.0f} s per forward)", flush=True)

    def report(name, cfg):
        emb, atp = forward_emulated(sd, tokens, make_round(cfg))
        print(f"{name:34s} emb {O.rel_err(emb, ref_emb):.2e}  atp {O.rel_err(atp, ref_atp):.2e}", flush=True)

    bf, hf = torch.bfloat16, torch.float16
    report("all sites bf16 (bf16_pure)", {s: bf for s in SITES})
    report("all sites fp16 (production)", {s: hf for s in SITES})
    report("bf16, row block fp16 ('bf16' mode)", {s: (hf if s.endswith("_row") else bf) for s in SITES})
    for s in SITES:
        report(f"only {s} bf16", {s: bf})
    for grp, ss in (("weights", ["w_row", "w_col", "w_ffn"]), ("LayerNorm outputs", ["xn_row", "xn_col", "xn_ffn"]),
                    ("row block", [s for s in SITES if s.endswith("_row")]),
                    ("column block", ["w_col", "xn_col", "qkv_col", "p_col", "ctx_col"]),
                    ("ffn", ["w_ffn", "xn_ffn", "hidden"])):
        report(f"only {grp} bf16", {s: bf for s in ss})
    # the reference itself cast to bf16 (weights AND every activation / accumulator output in bf16)
    sd16 = O.to_dtype(sd, torch.bfloat16)
    out = O.forward(sd16, tokens, repr_layers=[10], need_head_weights=True, want_logits=False)
    emb, atp = O.extract_features(out)
    print(f"{'oracle cast to bf16 (model.bfloat16())':34s} emb {O.rel_err(np.asarray(emb, dtype=np.float32), ref_emb):.2e}  "
          f"atp {O.rel_err(np.asarray(atp, dtype=np.float32), ref_atp):.2e}", flush=True)


if __name__ == "__main__":
    main()
