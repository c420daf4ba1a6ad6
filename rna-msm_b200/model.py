"""``MSATransformer`` -- drop-in for the reference's top-level ``model.MSATransformer``
(model.py:258-433; ``msm/model.py:206-423`` is the same network): same constructor arguments,
same 275 state-dict keys (so the published checkpoint loads with ``strict=True``), same
``forward(tokens, repr_layers, need_head_weights, return_contacts)`` result dict.

The forward body is one call into ``rnamsm_msa_forward`` per MSA: K1 (embedding + LayerNorm) ->
10 x [LN -> QKV GEMM -> tied logits -> softmax/map export -> AV -> out-proj+residual ->
LN -> QKV GEMM -> flash column attention -> out-proj+residual -> LN -> fc1+GELU -> fc2+residual]
-> final LN (-> LM head).  The Lightning training shell of the reference (model.py:29-255) is out
of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .alphabet import Vocab
from .modules import (AxialTransformerLayer, ColumnSelfAttention, ContactPredictionHead, LearnedPositionalEmbedding,
                      RobertaLMHead, RowSelfAttention, _PrecisionMixin, _infer_only)


class MSATransformer(nn.Module, _PrecisionMixin):
    def __init__(
        self,
        vocab: Vocab,
        optimizer_config=None,          # accepted and ignored: training shell is out of scope
        contact_train_data=None,        # accepted and ignored
        embed_dim: int = 768,
        num_attention_heads: int = 12,
        num_layers: int = 12,
        embed_positions_msa: bool = True,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.1,
        max_tokens_per_msa: int = 2 ** 14,
        max_seqlen: int = 1024,
        precision: Optional[str] = None,
    ):
        super().__init__()
        self.vocab = vocab
        self.embed_dim = embed_dim
        self.num_attention_heads = num_attention_heads
        self.num_layers = num_layers
        self.embed_positions_msa = embed_positions_msa
        self.dropout = dropout
        self.attention_dropout = attention_dropout
        self.activation_dropout = activation_dropout
        self.max_tokens_per_msa = max_tokens_per_msa

        self.embed_tokens = nn.Embedding(len(vocab), embed_dim, padding_idx=vocab.pad_idx)
        if embed_positions_msa:
            # one SCALAR per MSA row, as in the shipped model/checkpoint (model.py:293-296)
            self.msa_position_embedding = nn.Parameter(0.01 * torch.randn(1, 1024, 1, 1), requires_grad=True)
        else:
            self.register_parameter("msa_position_embedding", None)
        self.dropout_module = nn.Dropout(dropout)
        self.layers = nn.ModuleList([
            AxialTransformerLayer(embedding_dim=embed_dim, ffn_embedding_dim=4 * embed_dim,
                                  num_attention_heads=num_attention_heads, dropout=dropout,
                                  attention_dropout=attention_dropout, activation_dropout=activation_dropout,
                                  max_tokens_per_msa=max_tokens_per_msa)
            for _ in range(num_layers)
        ])
        self.contact_head = ContactPredictionHead(num_layers * num_attention_heads, vocab.prepend_bos,
                                                  vocab.append_eos, eos_idx=vocab.eos_idx)
        self.contact_head.requires_grad_(False)
        self.embed_positions = LearnedPositionalEmbedding(max_seqlen, embed_dim, vocab.pad_idx)
        self.emb_layer_norm_before = nn.LayerNorm(embed_dim)
        self.emb_layer_norm_after = nn.LayerNorm(embed_dim)
        self.lm_head = RobertaLMHead(embed_dim=embed_dim, output_dim=len(self.vocab), weight=self.embed_tokens.weight)
        self.init_weights()
        if precision is not None:
            self.set_precision(precision)
        self._wstruct = None

    # -- reference API -------------------------------------------------------------------------
    def init_weights(self):
        """model.py:89-101."""
        for module in self.modules():
            if isinstance(module, nn.Linear):
                nn.init.normal_(module.weight, std=0.02)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)
            elif isinstance(module, nn.Embedding):
                nn.init.normal_(module.weight, std=0.02)
                if module.padding_idx is not None:
                    module.weight.data[module.padding_idx].zero_()
            elif isinstance(module, nn.LayerNorm) and module.elementwise_affine:
                nn.init.ones_(module.weight)
                nn.init.zeros_(module.bias)

    @property
    def device(self) -> torch.device:
        return self.embed_tokens.weight.device

    def max_tokens_per_msa_(self, value: int) -> None:
        """Kept callable for API parity (model.py:418-428).  The reference uses it to chunk
        attention so that activations fit in memory; these kernels never materialise those
        activations, so the value is recorded and has no effect on results or memory."""
        self.max_tokens_per_msa = value
        for module in self.modules():
            if isinstance(module, (RowSelfAttention, ColumnSelfAttention)):
                module.max_tokens_per_msa = value

    @torch.no_grad()
    def check_fp16_range(self, tokens, fallback: Optional[str] = "bf16") -> Dict[str, object]:
        """Guard for weights never validated in fp16 (the trained RNA-MSM checkpoint is not available offline,
        RNA_MSM_Inference.py:133-135): runs ONE forward of ``tokens`` under the library's range watch, which scans every
        16-bit activation the forward writes (LayerNorm outputs, q|k|v, attention contexts, the post-GELU hidden) for the
        type's largest finite value -- where the fp16 kernels clamp (``cvt.rn.satfinite``).  Returns
        ``{"saturated": n, "max_abs": m, "precision": p}`` (n, m of the fp16 pass); if anything saturated and ``fallback``
        is given, the model is switched to that precision (default ``'bf16'``: fp32's exponent range; its tied row block
        is still fp16, so the watch runs once more and ``'bf16_pure'`` is selected if that block is what overflowed).
        A debugging pass: it adds one scan per activation tensor, so call it once per checkpoint / input family."""
        if self.precision not in ("fp16",):
            return {"saturated": 0, "max_abs": float("nan"), "precision": self.precision}

        def watched():
            with L.RangeWatch() as w:
                self(tokens, repr_layers=[self.num_layers], need_head_weights=True, want_logits=False)
            return w

        w = watched()
        if w.saturated and fallback is not None:
            self.set_precision(fallback)
            if fallback == "bf16" and watched().saturated:
                self.set_precision("bf16_pure")
        return {"saturated": w.saturated, "max_abs": w.max_abs, "precision": self.precision}

    def get_sequence_attention(self, tokens):
        return self(tokens.to(device=self.device), need_head_weights=True)["row_attentions"]

    def predict_contacts(self, tokens):
        return self(tokens, return_contacts=True)["contacts"]

    def check_msa_shape(self, R: int, Cc: int, n_nonpad_max: int) -> None:
        """The reference's input limits, raised identically by every entry point (forward, forward_batch, the streamed
        and the sharded paths): depth <= 1024 with the learned row positions (model.py:354-359) and non-pad length
        within the learned column positions (F.embedding raises IndexError, modules.py:292)."""
        if self.msa_position_embedding is not None and R > 1024:
            raise RuntimeError(
                "Using model with MSA position embedding trained on maximum MSA "
                f"depth of 1024, but received {R} alignments.")
        if n_nonpad_max + self.vocab.pad_idx >= self.embed_positions.weight.shape[0]:
            raise IndexError(
                f"sequence length {Cc} exceeds the {self.embed_positions.max_positions} learned positions")

    def check_tokens(self, tokens: torch.Tensor):
        """-> (has_pad, padding_mask) after check_msa_shape; one host sync, as model.py:347."""
        R, Cc = tokens.shape[-2:]
        padding_mask = tokens.eq(self.vocab.pad_idx)
        has_pad = bool(padding_mask.any())
        n_nonpad_max = int((~padding_mask).sum(-1).max()) if has_pad else Cc
        self.check_msa_shape(R, Cc, n_nonpad_max)
        return has_pad, padding_mask

    # -- C-ABI plumbing ------------------------------------------------------------------------
    def c_weights(self, code: int) -> L.ModelWeights:
        key = (code, self._row_code) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._wstruct is not None and self._wstruct[0] == key:
            return self._wstruct[1]
        keep = []

        def f32(t):
            if t is None:
                return None
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        layer_array = (L.LayerWeights * self.num_layers)(*[layer.c_weights(code) for layer in self.layers])
        keep.append(layer_array)
        m = L.ModelWeights()
        m.num_layers, m.embed_dim, m.num_heads = self.num_layers, self.embed_dim, self.num_attention_heads
        m.ffn_dim, m.vocab = 4 * self.embed_dim, len(self.vocab)
        m.n_pos, m.pad_idx = self.embed_positions.weight.shape[0], self.vocab.pad_idx
        m.ln_eps = float(self.emb_layer_norm_before.eps)
        m.tok_emb = f32(self.embed_tokens.weight)
        m.pos_emb = f32(self.embed_positions.weight)
        m.row_pos = f32(self.msa_position_embedding.reshape(-1)) if self.msa_position_embedding is not None else None
        m.ln_before_w, m.ln_before_b = f32(self.emb_layer_norm_before.weight), f32(self.emb_layer_norm_before.bias)
        m.ln_after_w, m.ln_after_b = f32(self.emb_layer_norm_after.weight), f32(self.emb_layer_norm_after.bias)
        m.lm_dense_w, m.lm_dense_b = f32(self.lm_head.dense.weight), f32(self.lm_head.dense.bias)
        m.lm_ln_w, m.lm_ln_b = f32(self.lm_head.layer_norm.weight), f32(self.lm_head.layer_norm.bias)
        m.lm_bias = f32(self.lm_head.bias)
        m.layers = C.cast(layer_array, C.POINTER(L.LayerWeights))
        if code != L.F32:                                              # LM-head dense GEMM on the tensor cores
            w16 = self.lm_head.dense.weight.detach().to(L.torch_dtype(code)).contiguous()
            keep.append(w16)
            m.lm_dense_w16 = w16.data_ptr()
        self._wstruct = (key, m, keep)
        return m

    # -- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, tokens, repr_layers: Iterable[int] = (), need_head_weights: bool = False,
                return_contacts: bool = False, want_logits: bool = True) -> Dict[str, object]:
        """model.py:338-416.  ``want_logits=False`` (extension) skips the LM head, which the
        inference script never reads."""
        if return_contacts:
            need_head_weights = True
        _infer_only(self)
        assert tokens.ndim == 3                                            # model.py:344
        L.require_cuda(tokens, "tokens")
        if tokens.device != self.device:
            raise RuntimeError(f"tokens on {tokens.device} but model on {self.device}")
        L.device_check(tokens.device)
        B, R, Cc = tokens.shape
        D, H, N = self.embed_dim, self.num_attention_heads, self.num_layers
        tokens = tokens.long().contiguous()
        has_pad, _ = self.check_tokens(tokens)                             # model.py:347, 354-359; modules.py:292
        repr_layers = set(repr_layers)
        code = self._code
        dev = tokens.device
        x = torch.empty((B, R, Cc, D), dtype=torch.float32, device=dev)
        row_att = torch.empty((B, N, H, Cc, Cc), dtype=torch.float32, device=dev) if need_head_weights else None
        logits = torch.empty((B, R, Cc, len(self.vocab)), dtype=torch.float32, device=dev) if want_logits else None
        reps: Dict[int, torch.Tensor] = {
            l: torch.empty((B, R, Cc, D), dtype=torch.float32, device=dev) for l in sorted(repr_layers) if 0 <= l < N}
        with torch.cuda.device(dev):
            m = self.c_weights(code)
            fcode = self._fwd_code
            nbytes = L.lib.rnamsm_workspace_bytes(R, Cc, D, H, 4 * D, fcode) + ((R * Cc + 255) // 256) * 256
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            st = L.stream_ptr()
            for b in range(B):
                rep_ptrs = (C.c_void_p * (N + 1))(*[reps[l][b].data_ptr() if l in reps else None for l in range(N + 1)])
                L.check(L.lib.rnamsm_msa_forward(
                    C.byref(m), L.ptr(tokens[b]), R, Cc, int(has_pad), fcode, L.ptr(x[b]),
                    L.ptr(row_att[b]) if row_att is not None else None, rep_ptrs,
                    L.ptr(logits[b]) if logits is not None else None, L.ptr(ws), nbytes, st), "msa_forward")
        if N in repr_layers:
            reps[N] = x                                                    # post-LN, model.py:400-401
        result: Dict[str, object] = {"logits": logits, "representations": reps}
        if need_head_weights:
            result["row_attentions"] = row_att                             # B x N x H x C x C
            if return_contacts:
                result["contacts"] = self.contact_head(tokens, row_att)
        return result

    @torch.no_grad()
    def forward_batch(self, tokens_list, need_head_weights: bool = False):
        """Several MSAs of different shapes in ONE pass (``rnamsm_msa_forward_batch``).

        ``tokens_list``: ``[1, R_i, C_i]`` (or ``[R_i, C_i]``) int64 grids.  Returns one dict per MSA with
        ``representations[num_layers]`` ``[1, R_i, C_i, D]`` and (optionally) ``row_attentions``
        ``[1, N, H, C_i, C_i]`` -- bit-identical to ``forward`` called on each MSA alone (each MSA keeps its own
        ``1/sqrt(R_i)`` in the tied attention, modules.py:713-715), unlike the reference's padded ``[B, R, C]``
        batch.  The token-local kernels run once over all MSAs' tokens, which is what makes short alignments
        (256 x 51) fill the GPU.  16-bit precisions; the fp32 parity path and single-row inputs go one by one."""
        _infer_only(self)
        grids = []
        for t in tokens_list:
            t = t[0] if t.ndim == 3 and t.shape[0] == 1 else t
            assert t.ndim == 2, "forward_batch takes one MSA per entry ([1,R,C] or [R,C])"
            L.require_cuda(t, "tokens")
            if t.device != self.device:
                raise RuntimeError(f"tokens on {t.device} but model on {self.device}")
            grids.append(t.long().contiguous())
        if not grids:
            return []
        N, D, H = self.num_layers, self.embed_dim, self.num_attention_heads
        code = self._code
        if code == L.dtype_code("fp32") or any(g.shape[0] < 2 for g in grids):
            outs = []
            for g in grids:
                r = self.forward(g.unsqueeze(0), repr_layers=[N], need_head_weights=need_head_weights, want_logits=False)
                outs.append({"representations": {N: r["representations"][N]}, **(
                    {"row_attentions": r["row_attentions"]} if need_head_weights else {})})
            return outs
        dev = self.device
        L.device_check(dev)
        n = len(grids)
        Rs, Cs = [g.shape[0] for g in grids], [g.shape[1] for g in grids]
        flat = torch.cat([g.reshape(-1) for g in grids])
        is_pad = flat.eq(self.vocab.pad_idx)
        sizes = [R * Cc for R, Cc in zip(Rs, Cs)]
        pad_counts = torch.stack([c.sum() for c in is_pad.split(sizes)]).tolist()      # one host sync for the batch
        for g, cnt, R, Cc in zip(grids, pad_counts, Rs, Cs):
            self.check_msa_shape(R, Cc, int((~g.eq(self.vocab.pad_idx)).sum(-1).max()) if cnt else Cc)
        has_pad = bytes(1 if c else 0 for c in pad_counts)
        T = sum(sizes)
        x = torch.empty((T, D), dtype=torch.float32, device=dev)
        maps = [torch.empty((1, N, H, Cc, Cc), dtype=torch.float32, device=dev) for Cc in Cs] if need_head_weights else None
        with torch.cuda.device(dev):
            m = self.c_weights(code)
            Ra, Ca = (C.c_int * n)(*Rs), (C.c_int * n)(*Cs)
            nbytes = L.lib.rnamsm_batch_workspace_bytes(n, Ra, Ca, D, H, 4 * D, code)
            if nbytes == 0:
                raise RuntimeError("rnamsm_batch_workspace_bytes: " + L.lib.rnamsm_last_error().decode(errors="replace"))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            map_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in maps]) if maps is not None else None
            L.check(L.lib.rnamsm_msa_forward_batch(C.byref(m), n, L.ptr(flat), Ra, Ca, has_pad, code, L.ptr(x), map_ptrs,
                                                   L.ptr(ws), nbytes, L.stream_ptr()), "msa_forward_batch")
        outs = []
        for i, xi in enumerate(x.split(sizes)):
            o = {"representations": {N: xi.view(1, Rs[i], Cs[i], D)}}
            if maps is not None:
                o["row_attentions"] = maps[i]
            outs.append(o)
        return outs
