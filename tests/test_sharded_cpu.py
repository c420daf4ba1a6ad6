"""CPU, world_size 2, gloo: the host-side schedule of the sharded forward (rna-msm_b200/sharded.py:
shard plan, logit sum over ranks, row<->column all-to-all re-layouts, delta add) driven by an ops
object built from the CPU oracle, against the un-sharded oracle forward.  The CUDA ops are covered by
tests/test_gpu_sharded.py; what is tested here is exactly the code that differs between 1 and N ranks.
"""
import math
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import msa_ref as O  # noqa: E402

LAYERS = 2


class OracleShardOps:
    """Same phase interface as rnamsm_b200.sharded.CudaShardOps, computed with the oracle's torch code."""

    def __init__(self, sd, dtype=torch.float64):
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        self.dt = dtype
        self.H, self.d = O.NUM_HEADS, O.HEAD_DIM

    def new_maps(self, N, C):
        return torch.empty(N, self.H, C, C, dtype=self.dt)

    def embed(self, tokens_rows, r0, R_global):
        sd = dict(self.sd)
        Rn = tokens_rows.shape[0]
        if "msa_position_embedding" in sd:
            sd["msa_position_embedding"] = sd["msa_position_embedding"][:, r0:r0 + Rn]
        x, _ = O.embed(sd, tokens_rows.unsqueeze(0))
        return x[0].reshape(-1, x.shape[-1]).contiguous()

    def _p(self, l, blk):
        return f"layers.{l}.{blk}."

    def row_logits(self, l, x, Rn, C, pad_rows, R_global):
        sd, p = self.sd, self._p(l, "row_self_attention")
        D = x.shape[-1]
        xn = O.layer_norm(x, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"]).view(Rn, C, D)
        scaling = (self.d ** -0.5) / math.sqrt(R_global)                   # GLOBAL depth, modules.py:713-715
        q = F.linear(xn, sd[p + "layer.q_proj.weight"], sd[p + "layer.q_proj.bias"]).view(Rn, C, self.H, self.d) * scaling
        k = F.linear(xn, sd[p + "layer.k_proj.weight"], sd[p + "layer.k_proj.bias"]).view(Rn, C, self.H, self.d)
        if pad_rows is not None:
            q = q * (1 - pad_rows.to(q).view(Rn, C, 1, 1))
        self._v = F.linear(xn, sd[p + "layer.v_proj.weight"], sd[p + "layer.v_proj.bias"]).view(Rn, C, self.H, self.d)
        return torch.einsum("rihd,rjhd->hij", q, k).unsqueeze(0).contiguous()

    def row_finish(self, l, x, partial, Rn, C, key_pad, R_global, map_out):
        sd, p = self.sd, self._p(l, "row_self_attention")
        att = partial.sum(0)
        if key_pad is not None:
            att = att.masked_fill(key_pad.view(1, 1, C), -10000)
        probs = att.softmax(-1)
        if map_out is not None:
            map_out.copy_(probs)
        ctx = torch.einsum("hij,rjhd->rihd", probs, self._v).reshape(Rn * C, -1)
        x.add_(F.linear(ctx, sd[p + "layer.out_proj.weight"], sd[p + "layer.out_proj.bias"]))

    def col_prepare(self, l, x, Rn, C):
        sd, p = self.sd, self._p(l, "column_self_attention")
        return O.layer_norm(x, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])

    def col_block(self, l, xn_cols, R, Cn, pad_cols):
        p = self._p(l, "column_self_attention")
        pm = None if pad_cols is None else pad_cols.unsqueeze(0)
        out, _ = O.column_attention(self.sd, p + "layer.", xn_cols.view(R, Cn, 1, -1), pm)
        return out.reshape(R * Cn, -1).contiguous()

    def add_delta(self, x, back, Rn, n, Cn):
        x.view(Rn, n, Cn, -1).add_(back.permute(1, 0, 2, 3))

    def ffn(self, l, x, T):
        sd, p = self.sd, self._p(l, "feed_forward_layer")
        xn = O.layer_norm(x, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])
        x.add_(O.feed_forward(sd, p + "layer.", xn))

    def final_ln(self, x, T):
        x.copy_(O.layer_norm(x, self.sd["emb_layer_norm_after.weight"], self.sd["emb_layer_norm_after.bias"]))


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rnamsm_b200.sharded import ShardedMSAForward, ShardPlan
        R, C, pad_cols, pad_rows = case
        sd = O.make_weights(5, num_layers=LAYERS, sharpen=2.0)
        tokens = O.make_tokens(R, C, 3, pad_cols=pad_cols, pad_rows=pad_rows)
        fwd = ShardedMSAForward(OracleShardOps(sd), LAYERS)
        out = fwd.forward(tokens, need_head_weights=True, gather_rows=True)
        ref = O.forward(O.to_dtype(sd, torch.float64), tokens, repr_layers=[LAYERS], need_head_weights=True,
                        num_layers=LAYERS, want_logits=False)
        got_rep, ref_rep = out["representations"][LAYERS], ref["representations"][LAYERS]
        if R % world or C % world:
            # internal <pad> rows change what the column attention computes AT <pad> positions (all keys masked: uniform
            # over 6 rows instead of 5) -- values nobody reads; every real token must agree
            keep = tokens.ne(O.PAD_IDX).unsqueeze(-1).to(ref_rep.dtype)
            got_rep, ref_rep = got_rep * keep, ref_rep * keep
        e_rep = O.rel_err(got_rep, ref_rep)
        e_att = O.rel_err(out["row_attentions"], ref["row_attentions"])
        Rp, Cp = -(-R // world) * world, -(-C // world) * world          # uneven shapes are padded inside forward()
        plan = ShardPlan(Rp, Cp, world, rank)
        valid = max(0, min(plan.Rn, R - plan.r0))
        ok_plan = out["row_shard"] == (plan.r0, plan.r0 + valid) and plan.rows(rank).start == rank * (Rp // world)
        ok_plan = ok_plan and tuple(out["representations"][LAYERS].shape[1:3]) == (R, C) \
            and tuple(out["row_attentions"].shape[-2:]) == (C, C)
        q.put((rank, e_rep, e_att, ok_plan))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(8, 10, 0, 0), (6, 12, 3, 2), (7, 9, 0, 0), (5, 11, 2, 1)],
                         ids=["nopad", "pad", "uneven", "uneven_pad"])
def test_sharded_schedule_world2_gloo(case):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29610 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, e_rep, e_att, ok_plan in results:
        assert ok_plan
        assert e_rep < 1e-10 and e_att < 1e-10, (rank, e_rep, e_att)     # fp64 on both sides: exact up to summation order


def test_shard_plan_and_traffic_model():
    from rnamsm_b200.sharded import ShardPlan
    p = ShardPlan(1024, 1024, 8, 3)
    assert (p.Rn, p.Cn, p.r0, p.c0) == (128, 128, 384, 384)
    b = p.bytes_per_layer()
    assert b["all_to_all_fwd"] == 128 * 1024 * 768 * 2 * 7 // 8                 # 176 MB per rank per direction
    assert b["logit_all_reduce"] == 2 * 7 * 12 * 1024 * 1024 * 4 // 8
    with pytest.raises(ValueError):
        ShardPlan(1001, 1024, 8, 0)
    with pytest.raises(ValueError):
        ShardPlan(8, 8, 2, 2)


def test_pad_to_shards_shapes_and_content():
    """Uneven shapes are padded with <pad> up to world | R and world * col_multiple | C; even shapes pass through."""
    from rnamsm_b200.sharded import pad_to_shards
    tok = O.make_tokens(7, 9, 1)
    assert pad_to_shards(tok, 1, 1, O.PAD_IDX) is tok
    out = pad_to_shards(tok, 2, 1, O.PAD_IDX)
    assert tuple(out.shape) == (1, 8, 10) and torch.equal(out[:, :7, :9], tok)
    assert bool((out[:, 7:, :] == O.PAD_IDX).all()) and bool((out[:, :, 9:] == O.PAD_IDX).all())
    assert tuple(pad_to_shards(tok, 8, 16, O.PAD_IDX).shape) == (1, 8, 128)
    even = O.make_tokens(8, 32, 1)
    assert pad_to_shards(even, 2, 16, O.PAD_IDX) is even
