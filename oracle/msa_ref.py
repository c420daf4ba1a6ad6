"""CPU oracle for the RNA-MSM MSA-transformer forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``rnamsm_b200``) never routes through
this file and fails loudly when its CUDA library is missing.

What this is: a plain-PyTorch (CPU, fp32 or fp64) restatement of the
reference algorithm for the hot path ``tokens [B,R,C] -> emb (L,768) +
atp (120,L,L)``.  Every function cites the reference file:line it follows
(paths relative to the reference checkout, yikunpku/RNA-MSM).

Parity pinning: the reference ships no numeric golden vectors for this path
(its ``results/*.npy`` need the unpublished checkpoint + ``hhfilter``), so the
oracle is pinned against *outputs of the reference itself run in the build
container*: ``oracle/gen_golden.py`` imports the reference's own
``msm.MSATransformer`` / ``modules.AxialTransformerLayer`` from the read-only
checkout, loads the seeded weights produced by :func:`make_weights`, and writes
``tests/golden/*.npz``.  ``tests/test_oracle.py`` re-checks this restatement
against those committed vectors on every run (and against the live reference
import whenever the checkout is present).  The shipped ``*_atp.npy`` /
``*_emb.npy`` fixtures pin the *format* (shape/dtype/index order) only.

The restatement is deliberately un-chunked and layer-streaming: it never keeps
the column-attention probabilities alive (the reference does, model.py:388-390)
so that 512x256-sized MSAs fit in host RAM.  Chunked vs un-chunked reference
results differ by ~1e-6; with padding the reference's chunked
path has a per-chunk masking quirk (modules.py:727-738) -- this oracle, like
the CUDA path, follows the UN-chunked semantics.
"""
from __future__ import annotations

import math
import re
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# Alphabet (msm/data.py:166-172, msm/constants.py:10-12)
# ----------------------------------------------------------------------------
PREPEND_TOKS = ("<cls>", "<pad>", "<eos>", "<unk>")
STANDARD_TOKS = ("A", "G", "C", "U", "X", "N", "-")
APPEND_TOKS = ("<mask>",)
ALL_TOKS = PREPEND_TOKS + STANDARD_TOKS + APPEND_TOKS
TOK_TO_IDX = {t: i for i, t in enumerate(ALL_TOKS)}
CLS_IDX, PAD_IDX, EOS_IDX, UNK_IDX, MASK_IDX = 0, 1, 2, 3, 11
VOCAB = len(ALL_TOKS)  # 12
PREPEND_BOS, APPEND_EOS = True, False

# Model hyper-parameters used by RNA_MSM_Inference.py:36-43
EMBED_DIM = 768
NUM_HEADS = 12
HEAD_DIM = EMBED_DIM // NUM_HEADS
FFN_DIM = 4 * EMBED_DIM
NUM_LAYERS = 10
MAX_SEQLEN = 1024          # LearnedPositionalEmbedding(max_seqlen, ...) -> 1026 rows
MAX_MSA_ROWS = 1024        # msa_position_embedding rows, model.py:293-296
LN_EPS = 1e-5              # nn.LayerNorm default (model.py:331-332, modules.py:382)


# ----------------------------------------------------------------------------
# Parameters: names/shapes follow the top-level model.MSATransformer state dict
# (model.py:288-334; 275 tensors, 95 911 301 parameters)
# ----------------------------------------------------------------------------
def state_dict_spec(num_layers: int = NUM_LAYERS, embed_dim: int = EMBED_DIM,
                    ffn_dim: Optional[int] = None, embed_positions_msa: bool = True,
                    max_seqlen: int = MAX_SEQLEN) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (name, shape, kind) list; kind in {linear_w, linear_b, emb, ln_w, ln_b, rowpos, tied, zeros}."""
    D = embed_dim
    Fd = ffn_dim or 4 * D
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    if embed_positions_msa:
        spec.append(("msa_position_embedding", (1, MAX_MSA_ROWS, 1, 1), "rowpos"))
    spec.append(("embed_tokens.weight", (VOCAB, D), "emb"))
    for l in range(num_layers):
        for blk in ("row_self_attention", "column_self_attention"):
            for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
                spec.append((f"layers.{l}.{blk}.layer.{proj}.weight", (D, D), "linear_w"))
                spec.append((f"layers.{l}.{blk}.layer.{proj}.bias", (D,), "linear_b"))
            spec.append((f"layers.{l}.{blk}.layer_norm.weight", (D,), "ln_w"))
            spec.append((f"layers.{l}.{blk}.layer_norm.bias", (D,), "ln_b"))
        spec.append((f"layers.{l}.feed_forward_layer.layer.fc1.weight", (Fd, D), "linear_w"))
        spec.append((f"layers.{l}.feed_forward_layer.layer.fc1.bias", (Fd,), "linear_b"))
        spec.append((f"layers.{l}.feed_forward_layer.layer.fc2.weight", (D, Fd), "linear_w"))
        spec.append((f"layers.{l}.feed_forward_layer.layer.fc2.bias", (D,), "linear_b"))
        spec.append((f"layers.{l}.feed_forward_layer.layer_norm.weight", (D,), "ln_w"))
        spec.append((f"layers.{l}.feed_forward_layer.layer_norm.bias", (D,), "ln_b"))
    spec.append(("contact_head.regression.weight", (1, num_layers * NUM_HEADS), "linear_w"))
    spec.append(("contact_head.regression.bias", (1,), "linear_b"))
    spec.append(("embed_positions.weight", (max_seqlen + PAD_IDX + 1, D), "emb"))
    spec.append(("emb_layer_norm_before.weight", (D,), "ln_w"))
    spec.append(("emb_layer_norm_before.bias", (D,), "ln_b"))
    spec.append(("emb_layer_norm_after.weight", (D,), "ln_w"))
    spec.append(("emb_layer_norm_after.bias", (D,), "ln_b"))
    spec.append(("lm_head.weight", (VOCAB, D), "tied"))
    spec.append(("lm_head.bias", (VOCAB,), "zeros"))
    spec.append(("lm_head.dense.weight", (D, D), "linear_w"))
    spec.append(("lm_head.dense.bias", (D,), "linear_b"))
    spec.append(("lm_head.layer_norm.weight", (D,), "ln_w"))
    spec.append(("lm_head.layer_norm.bias", (D,), "ln_b"))
    return spec


def make_weights(seed: int = 42, *, num_layers: int = NUM_LAYERS, perturb: bool = True,
                 sharpen: float = 1.0, embed_positions_msa: bool = True,
                 dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (the trained checkpoint is not available offline).

    Follows the reference init recipe, model.py:89-101 (Linear/Embedding N(0,0.02),
    padding row zeroed, LayerNorm 1/0) and model.py:293-296 (row-position scalars
    0.01*randn).  With ``perturb=True`` biases and LayerNorm affine parameters are
    additionally randomised so that every parameter influences the output (zero
    biases / unit gains would hide indexing bugs).  ``sharpen`` multiplies the
    q/k projection weights of every attention block so that attention logits span
    several nats, like the trained model's maps (results/2DRB_1_atp.npy spans
    2e-20..1), which makes parity sensitive to logit/softmax errors.

    Values are drawn per tensor from a CPU ``torch.Generator`` seeded with
    ``seed`` in spec order, so they are reproducible across machines with the same
    torch build (build container and GPU box share the image).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in state_dict_spec(num_layers, embed_positions_msa=embed_positions_msa):
        if kind == "linear_w":
            w = torch.randn(shape, generator=g) * 0.02
            if sharpen != 1.0 and (".q_proj." in name or ".k_proj." in name):
                w = w * sharpen
        elif kind == "linear_b":
            w = torch.randn(shape, generator=g) * 0.02 if perturb else torch.zeros(shape)
        elif kind == "emb":
            w = torch.randn(shape, generator=g) * 0.02
            w[PAD_IDX].zero_()
        elif kind == "ln_w":
            w = 1.0 + 0.1 * torch.randn(shape, generator=g) if perturb else torch.ones(shape)
        elif kind == "ln_b":
            w = 0.1 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        elif kind == "rowpos":
            w = 0.01 * torch.randn(shape, generator=g)
        elif kind == "tied":
            w = sd["embed_tokens.weight"]
        elif kind == "zeros":
            w = 0.02 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[name] = w.to(dtype) if kind != "tied" else w
    if dtype != torch.float32:
        sd["lm_head.weight"] = sd["embed_tokens.weight"]
    return sd


def make_tokens(R: int, C: int, seed: int = 0, *, batch: int = 1, pad_cols: int = 0,
                pad_rows: int = 0, gap_p: float = 0.0) -> torch.Tensor:
    """Synthetic MSA token grid: uniform over A,G,C,U,X,N,- (4..10),
    column 0 = <cls>.  ``pad_cols`` trailing columns of every row and ``pad_rows``
    trailing rows are set to <pad> (ragged-batch style) to exercise the padding
    semantics (model.py:346-367, modules.py:767-784, 911-915)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    tok = torch.randint(4, 11, (batch, R, C), generator=g, dtype=torch.int64)
    if gap_p > 0:
        gaps = torch.rand((batch, R, C), generator=g) < gap_p
        tok[gaps] = TOK_TO_IDX["-"]
    tok[:, :, 0] = CLS_IDX
    if pad_cols:
        tok[:, :, C - pad_cols:] = PAD_IDX
    if pad_rows:
        tok[:, R - pad_rows:, 1:] = PAD_IDX
    return tok


# ----------------------------------------------------------------------------
# MSA ingest for config 1 (utils/align.py:292-317, utils/tokenization.py:107-129)
# ----------------------------------------------------------------------------
def read_a2m(path: str, max_seqs: int = 512) -> Tuple[List[str], torch.Tensor]:
    """Parse an .a2m_msa2 (FASTA) file the way ``MSA.from_fasta`` does
    (utils/align.py:304-316: drop lowercase/'.'/'*', T->U, IUPAC ambiguity -> X) and
    tokenise like ``Vocab.encode`` (utils/tokenization.py:107-129; unknown -> <unk>,
    <cls> prepended).  Documented deviation: the reference sub-samples with the
    external ``hhfilter`` binary (utils/align.py:165-173), unavailable offline; we
    keep the FIRST ``max_seqs`` rows (what the reference does after hhfilter when
    it still returns too many, utils/align.py:172-173)."""
    names: List[str] = []
    seqs: List[str] = []
    cur: List[str] = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:])
                cur = []
            elif line:
                cur.append(line.strip())
    if names:
        seqs.append("".join(cur))
    clean = []
    for s in seqs:
        s = re.sub(r"([a-z]|\.|\*)", "", s)
        s = re.sub(r"[T]", "U", s)
        s = re.sub(r"[RYKMSWBDHVN]", "X", s)
        clean.append(s)
    clean = clean[:max_seqs]
    L = len(clean[0])
    assert all(len(s) == L for s in clean), "MSA rows must have equal length"
    tok = torch.full((len(clean), L + 1), UNK_IDX, dtype=torch.int64)
    tok[:, 0] = CLS_IDX
    for r, s in enumerate(clean):
        for c, ch in enumerate(s):
            tok[r, c + 1] = TOK_TO_IDX.get(ch, UNK_IDX)
    return names[:max_seqs], tok


# ----------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------
def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact erf GELU: modules.py:11-20 (LM head) and nn.GELU() (modules.py:416)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm(768), eps 1e-5, biased variance (modules.py:382, model.py:331-332)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def positions_from_tokens(tokens2d: torch.Tensor) -> torch.Tensor:
    """LearnedPositionalEmbedding index rule, modules.py:286-291:
    pos = cumsum(tok != pad) * (tok != pad) + pad_idx, per row along columns."""
    mask = tokens2d.ne(PAD_IDX).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + PAD_IDX


def embed(sd: Dict[str, torch.Tensor], tokens: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """model.py:346-367: token + learned-position + per-row scalar embedding, LayerNorm,
    zero the padded positions.  Returns x [B,R,C,D] and the padding mask (or None)."""
    assert tokens.ndim == 3                                     # model.py:344
    B, R, C = tokens.shape
    padding_mask = tokens.eq(PAD_IDX)
    if not padding_mask.any():
        padding_mask = None                                     # model.py:347-348
    x = sd["embed_tokens.weight"][tokens.long()]                # model.py:349
    pos = positions_from_tokens(tokens.view(B * R, C))
    x = x + sd["embed_positions.weight"][pos].view(B, R, C, -1)  # model.py:351-353
    p_row = sd.get("msa_position_embedding")
    if p_row is not None:
        if R > MAX_MSA_ROWS:                                    # model.py:354-359
            raise RuntimeError(
                "Using model with MSA position embedding trained on maximum MSA "
                f"depth of 1024, but received {R} alignments.")
        x = x + p_row[:, :R]                                    # (1,R,1,1) scalar per row
    x = layer_norm(x, sd["emb_layer_norm_before.weight"], sd["emb_layer_norm_before.bias"])
    if padding_mask is not None:
        x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))     # model.py:366-367
    return x, padding_mask


def row_attention(sd, pfx: str, xn: torch.Tensor, padding_mask: Optional[torch.Tensor]):
    """Tied row attention, un-chunked path: modules.py:752-821.
    xn [R,C,B,D] (already layer-normed) -> (out [R,C,B,D], probs [H,B,C,C])."""
    R, C, B, D = xn.shape
    H, d = NUM_HEADS, D // NUM_HEADS
    scaling = (d ** -0.5) / math.sqrt(R)                        # modules.py:713-715
    q = F.linear(xn, sd[pfx + "q_proj.weight"], sd[pfx + "q_proj.bias"]).view(R, C, B, H, d)
    k = F.linear(xn, sd[pfx + "k_proj.weight"], sd[pfx + "k_proj.bias"]).view(R, C, B, H, d)
    q = q * scaling                                             # modules.py:766
    if padding_mask is not None:                                # modules.py:767-772
        q = q * (1 - padding_mask.permute(1, 2, 0).unsqueeze(3).unsqueeze(4).to(q))
    attn = torch.einsum("rinhd,rjnhd->hnij", q, k)              # modules.py:774
    if padding_mask is not None:                                # modules.py:780-784
        attn = attn.masked_fill(padding_mask[:, 0].unsqueeze(0).unsqueeze(2), -10000)
    probs = attn.softmax(-1)                                    # modules.py:818
    v = F.linear(xn, sd[pfx + "v_proj.weight"], sd[pfx + "v_proj.bias"]).view(R, C, B, H, d)
    ctx = torch.einsum("hnij,rjnhd->rinhd", probs, v).contiguous().view(R, C, B, D)
    out = F.linear(ctx, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])  # modules.py:797-800
    return out, probs


def column_attention(sd, pfx: str, xn: torch.Tensor, padding_mask: Optional[torch.Tensor],
                     col_chunk: int = 0, return_probs: bool = False):
    """Column attention (attends over MSA rows per column): modules.py:875-924.
    Columns are independent, so the oracle may process them in chunks purely to bound
    memory (no effect on values); probabilities are not retained unless asked."""
    R, C, B, D = xn.shape
    H, d = NUM_HEADS, D // NUM_HEADS
    if R == 1:                                                  # modules.py:882-894
        out = F.linear(F.linear(xn, sd[pfx + "v_proj.weight"], sd[pfx + "v_proj.bias"]),
                       sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])
        probs = torch.ones(H, C, B, 1, 1, dtype=xn.dtype) if return_probs else None
        return out, probs
    if col_chunk <= 0:
        col_chunk = max(1, min(C, (1 << 27) // max(1, H * B * R * R)))  # ~0.5 GiB fp32 of probs
    outs, plist = [], []
    for c0 in range(0, C, col_chunk):
        xs = xn[:, c0:c0 + col_chunk]
        Cc = xs.shape[1]
        q = F.linear(xs, sd[pfx + "q_proj.weight"], sd[pfx + "q_proj.bias"]).view(R, Cc, B, H, d)
        k = F.linear(xs, sd[pfx + "k_proj.weight"], sd[pfx + "k_proj.bias"]).view(R, Cc, B, H, d)
        v = F.linear(xs, sd[pfx + "v_proj.weight"], sd[pfx + "v_proj.bias"]).view(R, Cc, B, H, d)
        q = q * (d ** -0.5)                                     # modules.py:905
        attn = torch.einsum("icnhd,jcnhd->hcnij", q, k)         # modules.py:907
        if padding_mask is not None:                            # modules.py:911-915
            pm = padding_mask[:, :, c0:c0 + col_chunk]
            attn = attn.masked_fill(pm.permute(2, 0, 1).unsqueeze(0).unsqueeze(3), -10000)
        probs = attn.softmax(-1)
        ctx = torch.einsum("hcnij,jcnhd->icnhd", probs, v).contiguous().view(R, Cc, B, D)
        outs.append(F.linear(ctx, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"]))
        if return_probs:
            plist.append(probs)
    out = torch.cat(outs, 1) if len(outs) > 1 else outs[0]
    return out, (torch.cat(plist, 1) if return_probs else None)


def feed_forward(sd, pfx: str, xn: torch.Tensor) -> torch.Tensor:
    """FeedForwardNetwork.forward, modules.py:423-427 (exact-erf GELU, dropout = identity)."""
    h = F.linear(xn, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"])
    h = F.gelu(h)                                               # nn.GELU() == erf form
    return F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])


def axial_layer(sd, l: int, x: torch.Tensor, padding_mask: Optional[torch.Tensor],
                return_col_probs: bool = False):
    """AxialTransformerLayer.forward (modules.py:242-267) with the three
    NormalizedResidualBlocks (modules.py:385-401) written out: x + f(LN(x))."""
    p = f"layers.{l}."
    xn = layer_norm(x, sd[p + "row_self_attention.layer_norm.weight"], sd[p + "row_self_attention.layer_norm.bias"])
    out, row_probs = row_attention(sd, p + "row_self_attention.layer.", xn, padding_mask)
    x = x + out
    xn = layer_norm(x, sd[p + "column_self_attention.layer_norm.weight"], sd[p + "column_self_attention.layer_norm.bias"])
    out, col_probs = column_attention(sd, p + "column_self_attention.layer.", xn, padding_mask,
                                      return_probs=return_col_probs)
    x = x + out
    xn = layer_norm(x, sd[p + "feed_forward_layer.layer_norm.weight"], sd[p + "feed_forward_layer.layer_norm.bias"])
    x = x + feed_forward(sd, p + "feed_forward_layer.layer.", xn)
    return x, col_probs, row_probs


def lm_head(sd, x: torch.Tensor) -> torch.Tensor:
    """RobertaLMHead.forward, modules.py:313-319."""
    h = F.linear(x, sd["lm_head.dense.weight"], sd["lm_head.dense.bias"])
    h = gelu_erf(h)
    h = layer_norm(h, sd["lm_head.layer_norm.weight"], sd["lm_head.layer_norm.bias"])
    return F.linear(h, sd["lm_head.weight"]) + sd["lm_head.bias"]


def symmetrize(x):
    """utils/tensor.py:98-100."""
    return x + x.transpose(-1, -2)


def apc(x):
    """Average product correction, utils/tensor.py:103-113."""
    a1 = x.sum(-1, keepdim=True)
    a2 = x.sum(-2, keepdim=True)
    a12 = x.sum((-1, -2), keepdim=True)
    return x - a1 * a2 / a12


def contact_head(sd, tokens: torch.Tensor, row_attentions: torch.Tensor) -> torch.Tensor:
    """ContactPredictionHead.forward, modules.py:347-366 (prepend_bos=True, append_eos=False)."""
    att = row_attentions[..., 1:, 1:]
    B, Ln, H, S, _ = att.shape
    att = att.reshape(B, Ln * H, S, S)
    att = apc(symmetrize(att)).permute(0, 2, 3, 1)
    w = sd["contact_head.regression.weight"].to(att)
    b = sd["contact_head.regression.bias"].to(att)
    return torch.sigmoid(F.linear(att, w, b).squeeze(3))


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, repr_layers: Iterable[int] = (),
            need_head_weights: bool = False, return_contacts: bool = False,
            num_layers: Optional[int] = None, want_logits: bool = True) -> Dict[str, object]:
    """MSATransformer.forward, model.py:338-416, layer-streaming (column probabilities
    are dropped after each layer; the top-level reference never returns them,
    model.py:406-410).  ``sd`` tensors decide the dtype (fp32 or fp64 oracle)."""
    if return_contacts:
        need_head_weights = True
    if num_layers is None:
        num_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    x, padding_mask = embed(sd, tokens)
    repr_layers = set(repr_layers)
    reps: Dict[int, torch.Tensor] = {}
    if 0 in repr_layers:
        reps[0] = x
    row_attn: List[torch.Tensor] = []
    x = x.permute(1, 2, 0, 3)                                   # B,R,C,D -> R,C,B,D  (model.py:379)
    for l in range(num_layers):
        x, _, rp = axial_layer(sd, l, x, padding_mask)
        if need_head_weights:
            row_attn.append(rp.permute(1, 0, 2, 3))             # H,B,C,C -> B,H,C,C (model.py:392)
        if (l + 1) in repr_layers:
            reps[l + 1] = x.permute(2, 0, 1, 3)
    x = layer_norm(x, sd["emb_layer_norm_after.weight"], sd["emb_layer_norm_after.bias"])
    x = x.permute(2, 0, 1, 3)                                   # model.py:396-397
    if num_layers in repr_layers:
        reps[num_layers] = x                                    # post-LN overwrite, model.py:400-401
    result: Dict[str, object] = {"representations": reps}
    if want_logits:
        result["logits"] = lm_head(sd, x)
    if need_head_weights:
        result["row_attentions"] = torch.stack(row_attn, 1)     # B,N,H,C,C (model.py:409)
        if return_contacts:
            result["contacts"] = contact_head(sd, tokens, result["row_attentions"])
    return result


def extract_features(result: Dict[str, object], num_layers: int = NUM_LAYERS) -> Tuple[np.ndarray, np.ndarray]:
    """RNA_MSM_Inference.py:150-166 post-processing: strip the BOS row/column, flatten
    (layer, head) -> 120 maps ``(120, L, L)`` f32, and row 0 of the final representation
    ``(L, 768)`` f32."""
    att = result["row_attentions"]
    start = int(PREPEND_BOS)
    end = att.size(-1) - int(APPEND_EOS)
    att = att[..., start:end, start:end]
    L = att.size(-1)
    atp = att.reshape(-1, L, L).float().cpu().numpy()
    emb = result["representations"][num_layers]
    end = emb.size(-2) - int(APPEND_EOS)
    emb = emb[:, 0, start:end, :].squeeze(0).float().cpu().numpy()
    return emb, atp


def rsa_input(emb: np.ndarray, seq: str, mu_emb, std_emb, mu_oh=None, std_oh=None) -> np.ndarray:
    """Input tensor of the downstream RSA predictor, _downstream_tasks/RSA/predict.py:131-141 (one-hot:
    ``one_hot_encode`` :67-74, sklearn ``OneHotEncoder(handle_unknown='ignore')`` over ``ACGU`` = an all-zero
    row for any other letter): ``[1, 4 + D + 1, L]`` f32.  The embedding z-score is float32 arithmetic (numpy
    float32 arrays), the one-hot z-score float64, the concatenation float64 and the final cast back to f32 --
    kept in that order because parity with the reference's tensor is bit-exact.  ``mu_oh=None``: the
    embedding-only predictor (``models/RNA-MSM_Emb``), no one-hot channels."""
    emb = (np.asarray(emb, dtype=np.float32) - np.asarray(mu_emb, dtype=np.float32)) / np.asarray(std_emb, dtype=np.float32)
    parts = [emb, np.ones((emb.shape[0], 1))]
    if mu_oh is not None:
        oh = np.array([[1.0 if ch == a else 0.0 for a in "ACGU"] for ch in seq], dtype=np.float64).reshape(-1, 4)
        parts.insert(0, (oh - np.asarray(mu_oh, dtype=np.float64)) / np.asarray(std_oh, dtype=np.float64))
    x = np.expand_dims(np.concatenate(parts, axis=1), 0)
    return np.ascontiguousarray(x.transpose(0, 2, 1)).astype(np.float32)


def to_dtype(sd: Dict[str, torch.Tensor], dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    out = {k: v.to(dtype) for k, v in sd.items()}
    out["lm_head.weight"] = out["embed_tokens.weight"]
    return out


def flops(R: int, C: int, num_layers: int = NUM_LAYERS, D: int = EMBED_DIM) -> float:
    """Algorithmic FLOPs of one forward (multiply-add = 2)."""
    return num_layers * (32.0 * R * C * D * D + 4.0 * R * C * D * (R + C)) + 2.0 * R * C * (D * D + VOCAB * D)


def rel_err(a, b) -> float:
    """Norm-relative error used by every parity gate: max|a-b| / max|b|."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
