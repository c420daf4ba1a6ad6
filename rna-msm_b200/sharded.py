"""One deep MSA across the GPUs of a box (BASELINE configs 4 and 5).

What shards how (follows the reference's math, not BASELINE.json's wording):

* tied row attention (modules.py:752-821) sums its logits over MSA rows -> shard ROWS: rank g owns
  rows [g R/n, (g+1) R/n), computes the partial logits of its rows (scaled with the GLOBAL row
  count, modules.py:713-715), one sum over ranks of the [H, C, C] fp32 logits per layer, softmax
  replicated, AV + out-projection local.
* column attention (modules.py:875-924) attends OVER rows (einsum "icnhd,jcnhd->hcnij", :907), so it
  is not local to a row shard: it shards by COLUMNS.  Per layer the LayerNorm output of the column
  block travels row-shard -> column-shard (all-to-all, 16-bit), the block's contribution
  out_proj(attn(.)) travels back (all-to-all, 16-bit) and is added to the fp32 residual stream,
  which never leaves its row shard.  K/V all-gather was rejected: (n-1) x more bytes.
* embedding, LayerNorms, FFN, final LayerNorm: token-local, stay in the row shard.

``ShardedMSAForward`` is the host-side schedule: shard plan, collectives, re-layouts.  What one
rank computes between collectives is delegated to an ops object: ``CudaShardOps`` (the C-ABI
kernels; the product) -- tests inject a CPU implementation to exercise this schedule under
``gloo`` with world_size 2.  Collectives here are plain ``torch.distributed`` (NCCL over NVLink on
the box); the fused peer-memory variants plug in at the two marked sites.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist


class ShardPlan:
    """Row / column ranges of every rank for an R x C token grid over n ranks."""

    def __init__(self, R: int, C: int, world: int, rank: int):
        if world < 1 or not (0 <= rank < world):
            raise ValueError(f"bad world/rank {world}/{rank}")
        if R % world or C % world:
            raise ValueError(
                f"sharded forward needs depth R={R} and columns C={C} divisible by the number of ranks {world}.  "
                "Padding COLUMNS with <pad> is result-neutral (pad keys are masked, modules.py:780-784); padding ROWS "
                "is NOT: align_scaling uses the padded depth (1/sqrt(R), modules.py:713-715), so the maps and "
                "embeddings of a row-padded MSA differ from the unpadded one -- drop rows to a multiple instead")
        self.R, self.C, self.world, self.rank = R, C, world, rank
        self.Rn, self.Cn = R // world, C // world
        self.r0, self.c0 = rank * self.Rn, rank * self.Cn

    def rows(self, rank: Optional[int] = None) -> slice:
        g = self.rank if rank is None else rank
        return slice(g * self.Rn, (g + 1) * self.Rn)

    def cols(self, rank: Optional[int] = None) -> slice:
        g = self.rank if rank is None else rank
        return slice(g * self.Cn, (g + 1) * self.Cn)

    def bytes_per_layer(self, D: int = 768, H: int = 12, act_bytes: int = 2, splits: int = 1) -> Dict[str, int]:
        """NVLink bytes one rank SENDS per layer (what the scaling model in DESIGN.md uses)."""
        n = self.world
        a2a = self.Rn * self.C * D * act_bytes * (n - 1) // n
        ar = 2 * (n - 1) * splits * H * self.C * self.C * 4 // n           # ring all-reduce volume per rank
        return {"all_to_all_fwd": a2a, "all_to_all_back": a2a, "logit_all_reduce": ar}


def pad_to_shards(tokens: torch.Tensor, world: int, col_multiple: int, pad_idx: int) -> torch.Tensor:
    """``[1, R, C]`` tokens padded with ``<pad>`` rows / columns so that ``world`` divides the depth and
    ``world * col_multiple`` the width (uneven shards).  Result-neutral for the real tokens PROVIDED the caller
    keeps the TRUE depth in ``align_scaling`` (``1/sqrt(R)``, modules.py:713-715) -- which both schedules below do: a pad
    row has ``q = 0`` (modules.py:767-772: no contribution to the tied logits) and is a masked key of the column attention
    (modules.py:911-915, ``exp(-10000 - max)`` is exactly 0 in fp32); a pad column is a masked key of the row attention
    (modules.py:780-784) and an independent, discarded item of the column attention."""
    _, R, C = tokens.shape
    Rp = -(-R // world) * world
    cm = world * max(1, col_multiple)
    Cp = -(-C // cm) * cm
    if Rp == R and Cp == C:
        return tokens
    out = torch.full((1, Rp, Cp), pad_idx, dtype=tokens.dtype, device=tokens.device)
    out[:, :R, :C] = tokens
    return out


class ShardedMSAForward:
    """MSATransformer.forward (model.py:338-416) for ONE MSA sharded over a process group.

    Every rank passes the same full ``tokens [1, R, C]``.  Returns the reference's result dict with
    ``row_attentions [1, N, H, C, C]`` complete on every rank and ``representations[N]`` holding this
    rank's ROW SHARD ``[1, R/n, C, D]`` (rank 0 owns MSA row 0, the source of ``*_emb.npy``);
    ``gather_rows=True`` all-gathers the full ``[1, R, C, D]``."""

    def __init__(self, ops, num_layers: int, group=None):
        self.ops = ops
        self.num_layers = num_layers
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # -- collectives (the two sites a fused peer-memory kernel replaces) --------------------------
    def _sum_over_ranks(self, partial: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
        return partial

    def _exchange(self, send: torch.Tensor) -> torch.Tensor:
        """send [n, ...] (chunk d goes to rank d) -> recv [n, ...] (chunk s came from rank s)."""
        if self.world == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, need_head_weights: bool = True, gather_rows: bool = False,
                pad_idx: int = 1) -> Dict[str, object]:
        assert tokens.ndim == 3 and tokens.shape[0] == 1, "one MSA per call: tokens [1, R, C]"
        _, Rt, Ct = tokens.shape                                                    # the TRUE shape: scaling, outputs
        tokens = pad_to_shards(tokens, self.world, 1, pad_idx)
        _, R, C = tokens.shape
        plan = ShardPlan(R, C, self.world, self.rank)
        n, Rn, Cn = plan.world, plan.Rn, plan.Cn
        ops, N = self.ops, self.num_layers
        tok = tokens[0]
        pad_full = tok.eq(pad_idx)
        has_pad = bool(pad_full.any())
        pad_rows = pad_full[plan.rows()].contiguous() if has_pad else None          # [Rn, C] this rank's rows
        pad_cols = pad_full[:, plan.cols()].contiguous() if has_pad else None       # [R, Cn] this rank's columns
        key_pad = pad_full[0].contiguous() if has_pad else None                     # MSA row 0, modules.py:780-784

        x = ops.embed(tok[plan.rows()].contiguous(), plan.r0, Rt)                   # [Rn*C, D] fp32
        D = x.shape[-1]
        maps = ops.new_maps(N, C) if need_head_weights else None
        for l in range(N):
            # ---- tied row attention on the row shard ----------------------------------------------
            partial = ops.row_logits(l, x, Rn, C, pad_rows, Rt)                     # [S, H, C, C] fp32, local rows
            partial = self._sum_over_ranks(partial)                                 # <-- fused P2P site 1
            ops.row_finish(l, x, partial, Rn, C, key_pad, Rt, None if maps is None else maps[l])
            # ---- column attention on the column shard ---------------------------------------------
            xn = ops.col_prepare(l, x, Rn, C)                                       # [Rn*C, D] 16-bit (fp32 path: fp32)
            send = xn.view(Rn, n, Cn, D).permute(1, 0, 2, 3).contiguous()           # [n(dest), Rn, Cn, D]
            recv = self._exchange(send)                                             # [n(src), Rn, Cn, D] == [R, Cn, D]
            delta = ops.col_block(l, recv.view(R * Cn, D), R, Cn, pad_cols)         # [R*Cn, D] == [n(dest), Rn, Cn, D]
            back = self._exchange(delta.view(n, Rn, Cn, D))                         # [n(src cols), Rn, Cn, D]  <-- site 2
            ops.add_delta(x, back, Rn, n, Cn)                                       # x[r, (s, c), :] += back[s, r, c, :]
            # ---- feed-forward on the row shard ----------------------------------------------------
            ops.ffn(l, x, Rn * C)
        ops.final_ln(x, Rn * C)
        rep = x.view(1, Rn, C, D)
        valid = max(0, min(Rn, Rt - plan.r0))                                       # real rows of this shard
        if gather_rows and n > 1:
            full = [torch.empty_like(rep) for _ in range(n)]
            dist.all_gather(full, rep.contiguous(), group=self.group)
            rep = torch.cat(full, 1)[:, :Rt, :Ct]
        elif (R, C) != (Rt, Ct):
            rep = rep[:, :valid, :Ct].contiguous()
        out: Dict[str, object] = {"logits": None, "representations": {N: rep}, "row_shard": (plan.r0, plan.r0 + valid)}
        if maps is not None:
            out["row_attentions"] = maps.view(1, N, maps.shape[1], C, C)
            if C != Ct:
                out["row_attentions"] = out["row_attentions"][..., :Ct, :Ct].contiguous()
        return out


class CudaShardOps:
    """What one rank computes between collectives, through the C ABI (no fallback)."""

    def __init__(self, model):
        from . import _lib as L
        from .modules import _linear
        self.L, self._linear, self.m = L, _linear, model
        self.code = model._code                  # column block / FFN operand type
        self.row_code = model._row_code          # tied row block operand type
        self.D, self.H = model.embed_dim, model.num_attention_heads
        self.dev = model.device
        self._qkv = None
        self._splits = 1

    # -- helpers -------------------------------------------------------------------------------------
    def _ln(self, x, ln: torch.nn.LayerNorm, code: int, T: int) -> torch.Tensor:
        L = self.L
        y = torch.empty((T, self.D), dtype=L.torch_dtype(code), device=x.device)
        L.check(L.lib.rnamsm_layernorm(L.ptr(x), L.ptr(ln.weight), L.ptr(ln.bias), L.ptr(y), code, T, self.D,
                                       float(ln.eps), 0, 0, L.stream_ptr()), "layernorm")
        return y

    @staticmethod
    def _u8(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None if mask is None else mask.to(torch.uint8).contiguous()

    # -- phases ---------------------------------------------------------------------------------------
    def new_maps(self, N: int, C: int) -> torch.Tensor:
        return torch.empty((N, self.H, C, C), dtype=torch.float32, device=self.dev)

    def embed(self, tokens_rows: torch.Tensor, r0: int, R_global: int) -> torch.Tensor:
        L, m = self.L, self.m
        Rn, C = tokens_rows.shape
        m.check_msa_shape(R_global, C, int(tokens_rows.ne(m.vocab.pad_idx).sum(-1).max()))   # this rank's rows
        x = torch.empty((Rn * C, self.D), dtype=torch.float32, device=self.dev)
        row_pos = None
        if m.msa_position_embedding is not None:
            table = m.msa_position_embedding.detach().reshape(-1).float()
            row_pos = torch.zeros(Rn, dtype=torch.float32, device=self.dev)           # rows past the table: <pad> rows only
            have = max(0, min(Rn, table.numel() - r0))
            row_pos[:have] = table[r0:r0 + have]
        self._keep = row_pos
        toks = tokens_rows.long().contiguous()
        L.check(L.lib.rnamsm_embed_layernorm(
            L.ptr(toks), Rn, C, L.ptr(m.embed_tokens.weight), m.embed_tokens.weight.shape[0],
            L.ptr(m.embed_positions.weight), m.embed_positions.weight.shape[0], L.ptr(row_pos),
            L.ptr(m.emb_layer_norm_before.weight), L.ptr(m.emb_layer_norm_before.bias), self.D, m.vocab.pad_idx,
            float(m.emb_layer_norm_before.eps), L.ptr(x), None, L.stream_ptr()), "embed_layernorm")
        return x

    def row_logits(self, l, x, Rn, C, pad_rows, R_global) -> torch.Tensor:
        L, H, D, code = self.L, self.H, self.D, self.row_code
        blk = self.m.layers[l].row_self_attention
        w_qkv, b_qkv, _, _ = blk.layer._pack(code)
        xn = self._ln(x, blk.layer_norm, code, Rn * C)
        q_scale = 0.125 if code != L.F32 else 0.125 / math.sqrt(R_global)           # align_scaling uses the GLOBAL depth
        self._pad_rows_u8 = self._u8(pad_rows)
        self._qkv = self._linear(xn, w_qkv, b_qkv, code, L.EPI_BIAS, q_scale, D, self._pad_rows_u8)
        self._splits = L.lib.rnamsm_row_attn_splits(Rn, C, H, code)
        partial = torch.empty((self._splits, H, C, C), dtype=torch.float32, device=x.device)
        L.check(L.lib.rnamsm_row_attn_logits(L.ptr(self._qkv), Rn, C, H, code, L.ptr(partial), self._splits,
                                             L.stream_ptr()), "row_attn_logits")
        return partial

    def row_finish(self, l, x, partial, Rn, C, key_pad, R_global, map_out) -> None:
        L, H, D, code = self.L, self.H, self.D, self.row_code
        blk = self.m.layers[l].row_self_attention
        _, _, w_out, b_out = blk.layer._pack(code)
        pmap = map_out if map_out is not None else torch.empty((H, C, C), dtype=torch.float32, device=x.device)
        logit_scale = 1.0 / math.sqrt(R_global) if code != L.F32 else 1.0
        if code != L.F32:
            ldp = (C + 7) // 8 * 8
            plp = torch.empty((H, C, ldp), dtype=L.torch_dtype(code), device=x.device)
        else:
            ldp, plp = C, None
        kp = self._u8(key_pad)
        L.check(L.lib.rnamsm_row_softmax(L.ptr(partial), partial.shape[0], H, C, L.ptr(kp), float(logit_scale),
                                         L.ptr(pmap), L.ptr(plp), ldp, code, L.stream_ptr()), "row_softmax")
        ctx = torch.empty((Rn * C, D), dtype=L.torch_dtype(code), device=x.device)
        L.check(L.lib.rnamsm_row_attn_av(L.ptr(plp if plp is not None else pmap), ldp, L.ptr(self._qkv), Rn, C, H, code,
                                         L.ptr(ctx), L.stream_ptr()), "row_attn_av")
        self._linear(ctx, w_out, b_out, code, L.EPI_BIAS_RESIDUAL, out=x)
        self._qkv = None

    def col_prepare(self, l, x, Rn, C) -> torch.Tensor:
        return self._ln(x, self.m.layers[l].column_self_attention.layer_norm, self.code, Rn * C)

    def col_block(self, l, xn_cols, R, Cn, pad_cols) -> torch.Tensor:
        L, H, D, code = self.L, self.H, self.D, self.code
        blk = self.m.layers[l].column_self_attention
        w_qkv, b_qkv, w_out, b_out = blk.layer._pack(code)
        xn_cols = xn_cols.contiguous()
        if R == 1:                                                                   # modules.py:882-894
            v = self._linear(xn_cols, w_qkv[2 * D:], b_qkv[2 * D:], code)
            return self._linear(v, w_out, b_out, code)
        qkv = self._linear(xn_cols, w_qkv, b_qkv, code, L.EPI_BIAS, 0.125, D, None)
        ctx = torch.empty((R * Cn, D), dtype=L.torch_dtype(code), device=xn_cols.device)
        pu8 = self._u8(pad_cols)
        L.check(L.lib.rnamsm_col_attn(L.ptr(qkv), R, Cn, H, code, 0, L.ptr(pu8), L.ptr(ctx), L.stream_ptr()), "col_attn")
        return self._linear(ctx, w_out, b_out, code)                                 # bias only: the delta travels back

    def add_delta(self, x, back, Rn, n, Cn) -> None:
        x.view(Rn, n, Cn, self.D).add_(back.permute(1, 0, 2, 3))                     # fp32 += 16-bit, strided read

    def ffn(self, l, x, T) -> None:
        L, code = self.L, self.code
        blk = self.m.layers[l].feed_forward_layer
        w1, b1, w2, b2 = blk.layer._pack(code)
        xn = self._ln(x, blk.layer_norm, code, T)
        h = self._linear(xn, w1, b1, code, L.EPI_BIAS_GELU)
        self._linear(h, w2, b2, code, L.EPI_BIAS_RESIDUAL, out=x)

    def final_ln(self, x, T) -> None:
        L, ln = self.L, self.m.emb_layer_norm_after
        L.check(L.lib.rnamsm_layernorm(L.ptr(x), L.ptr(ln.weight), L.ptr(ln.bias), L.ptr(x), L.F32, T, self.D,
                                       float(ln.eps), 0, 0, L.stream_ptr()), "layernorm")


_FUSED_CACHE = {}


def sharded_forward(model, tokens: torch.Tensor, group=None, need_head_weights: bool = True,
                    gather_rows: bool = False, fused: bool = False, host_out=None,
                    gather_maps: bool = True) -> Dict[str, object]:
    """Convenience: run ``model`` (an eval-mode ``MSATransformer`` replicated on every rank's GPU) on one
    MSA sharded over ``group``.  ``fused=True`` selects the peer-memory schedule (16-bit path); there
    ``host_out`` (a :class:`ShardedHostOutput`) makes every rank copy the map rows it owns straight to the shared
    host buffer and ``gather_maps=False`` skips assembling the full maps on every device."""
    L = __import__("rnamsm_b200")._lib
    L.require_cuda(tokens, "tokens")
    L.device_check(tokens.device)
    with torch.cuda.device(tokens.device):
        if fused:
            key = (id(model), id(group))
            if key not in _FUSED_CACHE:
                _FUSED_CACHE[key] = FusedShardedForward(model, group)
            return _FUSED_CACHE[key].forward(tokens, need_head_weights=need_head_weights, pad_idx=model.vocab.pad_idx,
                                             host_out=host_out, gather_maps=gather_maps)
        return ShardedMSAForward(CudaShardOps(model), model.num_layers, group).forward(
            tokens, need_head_weights=need_head_weights, gather_rows=gather_rows, pad_idx=model.vocab.pad_idx)


# =====================================================================================================
# Fused peer-memory schedule: the three exchanges above without NCCL on the data path
# =====================================================================================================
class PeerBuffer:
    """A device buffer every rank of the group can address: allocated by librnamsm_b200 (cudaMalloc),
    exported as a CUDA IPC handle, handles all-gathered over torch.distributed, peers' buffers mapped."""

    def __init__(self, nbytes: int, group=None):
        import ctypes as C
        from . import _lib as L
        self.L, self.C = L, C
        self.group = group
        self.nbytes = max(int(nbytes), 256)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        own = C.c_void_p()
        L.check(L.lib.rnamsm_peer_alloc(self.nbytes, C.byref(own)), "peer_alloc")
        self.own = own.value
        self._imported = []
        ptrs = [None] * self.world
        ptrs[self.rank] = self.own
        if self.world > 1:
            handle = (C.c_ubyte * 64)()
            L.check(L.lib.rnamsm_ipc_export(self.own, handle), "ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            for g, h in enumerate(handles):
                if g == self.rank:
                    continue
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                p = C.c_void_p()
                L.check(L.lib.rnamsm_ipc_import(buf, C.byref(p)), "ipc_import")
                ptrs[g] = p.value
                self._imported.append(p.value)
        self.ptrs = ptrs
        self.ptr_array = (C.c_void_p * self.world)(*ptrs)

    def offset_array(self, byte_offset: int):
        """void*[n] of every rank's buffer + byte_offset."""
        return (self.C.c_void_p * self.world)(*[p + byte_offset for p in self.ptrs])

    def tensor(self, dtype: torch.dtype, shape, byte_offset: int = 0) -> torch.Tensor:
        """This rank's own memory as a torch tensor (no copy).  The view must not outlive the buffer: the fused
        forward hands out CLONES of anything it returns."""
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        assert byte_offset + nbytes <= self.nbytes

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.own + byte_offset, False),
                                      "version": 2}
        t = torch.as_tensor(h, device="cuda")
        t._rnamsm_keepalive = (h, self)
        return t.view(dtype).view(*shape)

    def close_imports(self):
        for p in self._imported:
            self.L.lib.rnamsm_ipc_close(p)
        self._imported = []

    def free_own(self):
        if self.own:
            self.L.lib.rnamsm_peer_free(self.own)
            self.own = None

    @staticmethod
    def close_all(bufs, group=None):
        """Unmap every peer's memory on every rank BEFORE any owner frees it (a barrier in between)."""
        torch.cuda.synchronize()
        for b in bufs:
            b.close_imports()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.barrier(group=group)
        for b in bufs:
            b.free_own()


class SharedHostBuffer:
    """fp32 host memory visible to every rank of the box: a POSIX shared-memory file mapped by all ranks and
    page-locked in each rank's CUDA context, so each GPU can DMA into it over its own PCIe link."""

    def __init__(self, numel: int, group=None, tag: str = "maps"):
        import mmap
        import os
        import uuid
        from . import _lib as L
        self.L = L
        self.numel = int(numel)
        self.nbytes = self.numel * 4
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        names = [f"/dev/shm/rnamsm_{tag}_{os.getpid()}_{uuid.uuid4().hex}" if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(names, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.path = names[0]
        size = max(self.nbytes, mmap.PAGESIZE)
        if rank == 0:
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize < size + (64 << 20):
                raise RuntimeError(f"/dev/shm has {st.f_bavail * st.f_frsize} bytes free, need {size}")
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_EXCL, 0o600)
            os.ftruncate(fd, size)
        if world > 1:
            dist.barrier(group=group)
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self._mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        os.close(fd)
        self.tensor = torch.frombuffer(self._mm, dtype=torch.float32, count=self.numel)
        if rank == 0:
            self.tensor.zero_()                   # touch every page before it is page-locked
        if world > 1:
            dist.barrier(group=group)
        if rank == 0:
            os.unlink(self.path)                  # the mappings keep the memory alive; nothing is left behind
        self._registered = False
        if torch.cuda.is_available():
            L.check(L.lib.rnamsm_host_register(self.tensor.data_ptr(), size), "host_register")
            self._registered = True

    def close(self):
        if self._registered:
            self.L.lib.rnamsm_host_unregister(self.tensor.data_ptr())
            self._registered = False


class ShardedHostOutput:
    """Host-side results of the sharded forward, in the reference's file layouts (RNA_MSM_Inference.py:150-166):
    ``atp`` ``[(N*H), L, L]`` fp32 in memory shared by the ranks (every rank writes the query rows it owns; complete on
    return from ``forward``) and ``emb`` ``[L, D]`` fp32 pinned on the rank that owns MSA row 0 (rank 0)."""

    def __init__(self, num_layers: int, heads: int, C: int, D: int, start: int = 1, end_strip: int = 0, group=None):
        self.N, self.H, self.C, self.D, self.start = num_layers, heads, C, D, start
        self.Ls = C - start - end_strip
        self.shared = SharedHostBuffer(num_layers * heads * self.Ls * self.Ls, group, "atp")
        self.atp = self.shared.tensor.view(num_layers * heads, self.Ls, self.Ls)
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.emb = torch.empty((self.Ls, D), dtype=torch.float32).pin_memory() if rank == 0 else None

    def close(self):
        self.shared.close()


class FusedShardedForward:
    """Same partition as ShardedMSAForward, 16-bit path only, with the exchanges fused into kernels that
    address peer memory over NVLink (csrc/peer.cu, umma_gemm.cu):

        tied logits    rnamsm_row_softmax_p2p : pull owned query rows of every rank's partial logits, softmax,
                                                push 16-bit rows to all ranks; the fp32 map rows stay with their
                                                owner, which copies them to the host over its own PCIe link
        row -> column  rnamsm_layernorm_push  : LayerNorm rows stored straight into the column owner's
                                                [C/n, R, D] buffer
        column -> row  rnamsm_linear_residual_scatter : out-projection GEMM whose epilogue TMA-stores the 16-bit
                                                result into the row owner's receive buffer (or TMA-reduce-adds
                                                fp32 into its residual stream); rnamsm_add_layernorm adds it
                                                on the FFN's LayerNorm pass

    Between phases: ``rnamsm_peer_barrier``, a flag barrier in peer memory launched on the stream (4 per layer; no
    host-launched collective on the data path).  Results are fresh tensors (clones of the persistent peer
    buffers): ``representations[N]`` = this rank's row shard, ``row_attentions`` = the full maps on every rank when
    ``gather_maps`` (one NCCL all-gather of the owners' rows at the end), ``row_attentions_rows`` = the owned query
    rows ``[N, H, C/n, C]`` otherwise."""

    def __init__(self, model, group=None):
        from . import _lib as L
        from .modules import _linear
        self.L, self._linear, self.m, self.group = L, _linear, model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > L_MAX_PEERS:
            raise ValueError(f"at most {L_MAX_PEERS} ranks")
        self.code, self.row_code = model._code, model._row_code
        if self.code == L.F32:
            raise ValueError("the fused peer-memory schedule exists for the 16-bit path; use ShardedMSAForward for fp32")
        self.ops = CudaShardOps(model)
        self._bufs = {}
        self._flags = None
        self._epoch = [0, 0]                       # barrier epochs of the compute stream / the copy stream
        self._side = None
        import os
        # column->row payload: 16-bit into a receive buffer (default) or fp32 TMA reduce-add straight into x
        self.scatter_fp32 = os.environ.get("RNAMSM_SCATTER_FP32", "0") == "1"
        # RNAMSM_NCCL_BARRIER=1: the round-1 ordering (a 4-byte NCCL all-reduce per phase), kept for A/B timing
        self.nccl_barrier = os.environ.get("RNAMSM_NCCL_BARRIER", "0") == "1"
        self._nccl_flag = None

    def _barrier(self, which: int = 0, stream=None):
        """Stream-ordered barrier across the ranks (flag set ``which``: 0 = compute stream, 1 = copy stream)."""
        if self.world == 1:
            return
        L = self.L
        if self.nccl_barrier and which == 0:
            dist.all_reduce(self._nccl_flag, group=self.group)
            return
        self._epoch[which] += 1
        L.check(L.lib.rnamsm_peer_barrier(self._flags.offset_array(128 * which), self.world, self.rank,
                                          self._epoch[which], stream if stream is not None else L.stream_ptr()),
                "peer_barrier")

    def _buffers(self, R, C, N):
        key = (R, C, N)
        if key in self._bufs:
            return self._bufs[key]
        if self._bufs:                             # a new shape: every result handed out so far was a clone
            PeerBuffer.close_all([b for old in self._bufs.values() for b in old["all"]], self.group)
        self._bufs = {}
        L, n = self.L, self.world
        D, H = self.m.embed_dim, self.m.num_attention_heads
        Rn, Cn = R // n, C // n
        splits = L.lib.rnamsm_row_attn_splits(Rn, C, H, self.row_code)
        # Every split is one more [H, C/n, C] fp32 slab that each rank pulls from each peer over NVLink in the P2P
        # softmax: at 1024 x 1024 on 8 GPUs the wave-efficient 3 splits made that 132 MB per rank and layer (190 us at the
        # measured 700 GB/s) to save ~10 % of a 180 us GEMM.  Keep the pull under ~32 MB; short alignments keep their splits.
        per_split = (n - 1) * H * Cn * C * 4
        splits = max(1, min(splits, (32 << 20) // max(1, per_split))) if n > 1 else splits
        ldp = (C + 7) // 8 * 8
        b = {
            "x": PeerBuffer(Rn * C * D * 4, self.group),
            "xn_cols": PeerBuffer(Cn * R * D * 2, self.group),
            "partial": PeerBuffer(splits * H * C * C * 4, self.group),
            "probs": PeerBuffer(H * C * ldp * 2, self.group),
            "delta": PeerBuffer(Rn * C * D * 2, self.group),
        }
        b["all"] = list(b.values())
        b["splits"], b["ldp"] = splits, ldp
        self._bufs[key] = b
        return b

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, need_head_weights: bool = True, pad_idx: int = 1, host_out=None,
                gather_maps: bool = True) -> Dict[str, object]:
        L, m, ops = self.L, self.m, self.ops
        assert tokens.ndim == 3 and tokens.shape[0] == 1, "one MSA per call: tokens [1, R, C]"
        _, Rt, Ct = tokens.shape                                                    # the TRUE shape: scaling, outputs
        tokens = pad_to_shards(tokens, self.world, 16, pad_idx)                     # (16: TMA box rows of the scatter GEMM)
        _, R, C = tokens.shape
        plan = ShardPlan(R, C, self.world, self.rank)
        n, Rn, Cn, g = plan.world, plan.Rn, plan.Cn, plan.rank
        N, D, H = m.num_layers, m.embed_dim, m.num_attention_heads
        if Cn % 16:
            raise ValueError(f"fused schedule needs C / ranks = {Cn} to be a multiple of 16 (TMA box rows)")
        code, row_code = self.code, self.row_code
        dt, row_dt = L.torch_dtype(code), L.torch_dtype(row_code)
        if self._flags is None:
            self._flags = PeerBuffer(256, self.group)       # two sets of n uint32 flags (zeroed by peer_alloc)
            self._nccl_flag = torch.zeros(1, dtype=torch.float32, device=tokens.device)
            self._side = torch.cuda.Stream()
            if n > 1:
                dist.barrier(group=self.group)              # every rank's flag array exists and is zero
        B = self._buffers(R, C, N)
        splits, ldp = B["splits"], B["ldp"]
        x = B["x"].tensor(torch.float32, (Rn * C, D))
        xn_cols = B["xn_cols"].tensor(dt, (Cn * R, D))
        partial = B["partial"].tensor(torch.float32, (splits, H, C, C))
        probs = B["probs"].tensor(row_dt, (H, C, ldp))
        delta = B["delta"].tensor(dt, (Rn * C, D))
        want_maps = need_head_weights or host_out is not None
        # this rank's maps: full [N, H, C, C] addressing, only the owned query rows [g*Cn, (g+1)*Cn) are ever written
        maps = torch.empty((N, H, C, C), dtype=torch.float32, device=tokens.device) if want_maps else None
        st = L.stream_ptr()
        main, side = torch.cuda.current_stream(), self._side
        i0, i1 = g * Cn, (g + 1) * Cn

        tok = tokens[0]
        pad_full = tok.eq(pad_idx)
        has_pad = bool(pad_full.any())
        pad_rows = CudaShardOps._u8(pad_full[plan.rows()]) if has_pad else None
        pad_cols = CudaShardOps._u8(pad_full[:, plan.cols()]) if has_pad else None
        key_pad = CudaShardOps._u8(pad_full[0]) if has_pad else None

        x.copy_(ops.embed(tok[plan.rows()].contiguous(), plan.r0, Rt))
        self._barrier()                                   # every rank's buffers are initialised / the previous call is over
        logit_scale = 1.0 / math.sqrt(Rt)                 # align_scaling with the true depth (pad rows add nothing)
        for l in range(N):
            layer = m.layers[l]
            # ---- tied row attention ------------------------------------------------------------------
            blk = layer.row_self_attention
            w_qkv, b_qkv, w_out, b_out = blk.layer._pack(row_code)
            xn = ops._ln(x, blk.layer_norm, row_code, Rn * C)
            qkv = self._linear(xn, w_qkv, b_qkv, row_code, L.EPI_BIAS, 0.125, D, pad_rows)
            L.check(L.lib.rnamsm_row_attn_logits(L.ptr(qkv), Rn, C, H, row_code, L.ptr(partial), splits, st), "row_attn_logits")
            self._barrier()                               # all partial logits written
            L.check(L.lib.rnamsm_row_softmax_p2p(B["partial"].ptr_array, n, g, splits, H, C, L.ptr(key_pad),
                                                 float(logit_scale), L.ptr(maps[l]) if maps is not None else None,
                                                 B["probs"].ptr_array, ldp, row_code, st),
                    "row_softmax_p2p")
            if host_out is not None:                      # the owned rows of this layer's maps -> shared host memory
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    L.check(L.lib.rnamsm_copy_map_rows_d2h(L.ptr(maps[l]), H, C, i0, i1, host_out.start, host_out.Ls,
                                                           host_out.atp[l * H].data_ptr(), side.cuda_stream),
                            "copy_map_rows_d2h")
            self._barrier()                               # every rank's probabilities complete
            ctx = torch.empty((Rn * C, D), dtype=row_dt, device=x.device)
            L.check(L.lib.rnamsm_row_attn_av(L.ptr(probs), ldp, L.ptr(qkv), Rn, C, H, row_code, L.ptr(ctx), st), "row_attn_av")
            self._linear(ctx, w_out, b_out, row_code, L.EPI_BIAS_RESIDUAL, out=x)
            # ---- column attention ----------------------------------------------------------------------
            blk = layer.column_self_attention
            w_qkv, b_qkv, w_out, b_out = blk.layer._pack(code)
            ln = blk.layer_norm
            L.check(L.lib.rnamsm_layernorm_push(L.ptr(x), L.ptr(ln.weight), L.ptr(ln.bias), B["xn_cols"].ptr_array, n, Rn, C,
                                                R, plan.r0, D, float(ln.eps), code, st), "layernorm_push")
            self._barrier()                               # this rank's column shard has arrived
            qkv_c = self._linear(xn_cols, w_qkv, b_qkv, code, L.EPI_BIAS, 0.125, D, None)       # [Cn, R, 3D]
            ctx_c = torch.empty((R * Cn, D), dtype=dt, device=x.device)                          # token-major [R, Cn, D]
            L.check(L.lib.rnamsm_col_attn(L.ptr(qkv_c), R, Cn, H, code, 1, L.ptr(pad_cols), L.ptr(ctx_c), st), "col_attn")
            if self.scatter_fp32:
                L.check(L.lib.rnamsm_linear_residual_scatter(L.ptr(ctx_c), L.ptr(w_out), L.ptr(b_out), R, Cn, D, D, code,
                                                             B["x"].ptr_array, n, Rn, C, plan.c0, 0, st),
                        "linear_residual_scatter")
                self._barrier()                           # every contribution has been reduced into x
                ops.ffn(l, x, Rn * C)
                continue
            L.check(L.lib.rnamsm_linear_residual_scatter(L.ptr(ctx_c), L.ptr(w_out), L.ptr(b_out), R, Cn, D, D, code,
                                                         B["delta"].ptr_array, n, Rn, C, plan.c0, 1, st),
                    "linear_residual_scatter")
            self._barrier()                               # every column owner's 16-bit contribution has arrived
            # ---- feed-forward (its LayerNorm pass also adds the column block's contribution to x) --------
            blk = layer.feed_forward_layer
            w1, b1, w2, b2 = blk.layer._pack(code)
            ln = blk.layer_norm
            xn_f = torch.empty((Rn * C, D), dtype=dt, device=x.device)
            L.check(L.lib.rnamsm_add_layernorm(L.ptr(x), L.ptr(delta), code, L.ptr(ln.weight), L.ptr(ln.bias), L.ptr(xn_f),
                                               code, Rn * C, D, float(ln.eps), st), "add_layernorm")
            hdn = self._linear(xn_f, w1, b1, code, L.EPI_BIAS_GELU)
            self._linear(hdn, w2, b2, code, L.EPI_BIAS_RESIDUAL, out=x)
        ops.final_ln(x, Rn * C)
        valid = max(0, min(Rn, Rt - plan.r0))             # real rows of this shard
        rep = x.view(1, Rn, C, D)[:, :valid, :Ct].clone()  # a fresh tensor: the peer buffer is rewritten by the next call
        out: Dict[str, object] = {"logits": None, "representations": {N: rep}, "row_shard": (plan.r0, plan.r0 + valid)}
        if host_out is not None:
            if g == 0:                                    # rank 0 owns MSA row 0 = the source of *_emb.npy
                if host_out.start + host_out.Ls > Ct:
                    raise ValueError(f"host_out was built for C={host_out.C}, tokens have C={Ct}")
                host_out.emb.copy_(rep[0, 0, host_out.start:host_out.start + host_out.Ls], non_blocking=True)
            with torch.cuda.stream(side):                 # returns once EVERY rank's rows have landed in host memory
                self._barrier(1, side.cuda_stream)
            maps.record_stream(side)
            side.synchronize()
            main.synchronize()
        if need_head_weights:
            rows = maps[:, :, i0:i1, :]
            if gather_maps:
                if n > 1:
                    parts = [torch.empty((N, H, Cn, C), dtype=torch.float32, device=x.device) for _ in range(n)]
                    dist.all_gather(parts, rows.contiguous(), group=self.group)
                    out["row_attentions"] = torch.cat(parts, 2).view(1, N, H, C, C)
                else:
                    out["row_attentions"] = maps.view(1, N, H, C, C)
                if C != Ct:
                    out["row_attentions"] = out["row_attentions"][..., :Ct, :Ct].contiguous()
            else:
                hi = max(i0, min(i1, Ct))                 # real query rows this rank owns
                out["row_attentions_rows"] = rows[:, :, :hi - i0, :Ct] if C != Ct else rows
                out["row_attentions_range"] = (i0, hi)
        return out


L_MAX_PEERS = 8
