#!/usr/bin/env python
"""bench.py -- RNA-MSM MSA-transformer forward throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg1|cfg4|cfg5] [--precision fp16|bf16|fp32]

A "step" is one full forward of the hot path over one synthetic MSA: int64 token grid ->
emb (L,768) + 120 tied-row attention maps, with random-init weights of the real architecture
(10 layers, D=768, 12 heads; the trained checkpoint is not available offline).  At N=1 the
workload is BASELINE config[1]: depth 512 x L 256 (token grid R x C = 512 x 256, 131 072 tokens).
N > 1 (torchrun, one rank per GPU): independent MSAs are data-parallel with NO collective
 -- every rank runs its own MSA of the same shape, weak scaling;
value = all ranks' tokens / max-over-ranks device time.

Printed JSON (rank 0, one line):
  value        tokens/s with the tokens already resident in HBM (device-timed, CUDA events)
  e2e          same metric through the public API `MSATransformer.forward` + `extract_features`
               with pinned HOST tokens in and emb/atp copied back to the host every step
  roofline     dominant kernel class (the tcgen05 dense GEMM): algorithmic FLOPs / live
               CUDA-event time per launch vs the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline the oracle (port of the reference's fp32 CPU forward) timed on this box's cores on
               a bounded sample: ONE of the 10 AxialTransformerLayers at the full shape, x10
  --impl reference : the CPU arm alone, same metric/config, one sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MSA tokens/sec (depth x L) full forward incl. emb+atp"
WORKLOADS = {
    # name: (R, C, embed_positions_msa, description)
    "cfg1": (512, 36, True, "cfg1 2DRB_1-shaped MSA depth 512 x L 35 (+BOS)"),
    "cfg2": (512, 256, True, "cfg2 synthetic MSA depth 512 x L 256 (token grid 512x256), batch 1"),
    "cfg4": (4096, 128, False, "cfg4 synthetic deep MSA depth 4096 x L 128, embed_positions_msa=False"),
    "cfg5": (1024, 1024, True, "cfg5 synthetic long MSA depth 1024 x L 1024 token grid"),
    # cfg3: 64 independent MSAs, depth 256, L ~ U{50..500} (seeded), farmed over the ranks longest-first
    "cfg3": (256, None, True, "cfg3 batch of 64 synthetic MSAs, depth 256, L 50-500 (RNAcmap3-like), one B=1 forward each, "
                              "longest-processing-time-first assignment over the ranks, no collective"),
}


def measured_traffic(workload, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class, from the committed
    ncu --set full capture of the same workload (profiles/r01e_traffic.json); None for other workloads."""
    p = os.path.join(ROOT, "profiles", "r01e_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("workload") == workload and d.get("precision") == precision:
            return d["dense_gemm"]["dram_bytes_per_launch"], d["dense_gemm"]["algorithmic_bytes_per_launch"]
    except Exception:
        pass
    return None, None


def total_flops(R, C):
    """Algorithmic FLOPs of one forward: N [32 T D^2 + 4 T D (R + C)] + 2 T (D^2 + V D)."""
    T = R * C
    return NL * (32.0 * T * D * D + 4.0 * T * D * (R + C)) + 2.0 * T * (D * D + V * D)


def synthetic_tokens(R, C, seed):
    """Synthetic MSA: uniform over A,G,C,U,X,N,- (indices 4..10), column 0 = <cls>, no <pad>."""
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(4, 11, (1, R, C), generator=g, dtype=torch.int64)
    t[:, :, 0] = 0
    return t


def cfg3_lengths(n=64, seed=3):
    import random
    rng = random.Random(seed)
    return [rng.randint(50, 500) + 1 for _ in range(n)]          # token columns C = L + 1 (BOS)


def lpt_assign(costs, world):
    """Longest-processing-time-first: job indices per rank."""
    loads, jobs = [0.0] * world, [[] for _ in range(world)]
    for i in sorted(range(len(costs)), key=lambda i: -costs[i]):
        g = min(range(world), key=lambda r: loads[r])
        loads[g] += costs[i]
        jobs[g].append(i)
    return jobs
D, H, F, NL, V = 768, 12, 3072, 10, 12


def flops_breakdown(R, C):
    """Algorithmic FLOPs per forward by kernel class (multiply-add = 2)."""
    T = R * C
    return {
        "linear_qkv": NL * 2 * (2.0 * T * D * 3 * D),
        "linear_out_resid": NL * 2 * (2.0 * T * D * D),
        "linear_fc1_gelu": NL * (2.0 * T * D * F),
        "linear_fc2_resid": NL * (2.0 * T * D * F),
        "row_logits": NL * (2.0 * R * C * C * D),
        "row_av": NL * (2.0 * R * C * C * D),
        "col_attn": NL * (4.0 * R * R * C * D),
    }


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]          # upper half ~ samples taken under load
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    "hbm": d["hbm_gbs"], "source": "measured"}
        except Exception:
            pass
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(R, C, embed_positions_msa, threads=None):
    """One AxialTransformerLayer of the CPU oracle (port of modules.py:242-267) at the full R x C
    shape + the embedding, fp32, all host threads.  Returns (seconds for the sample, threads)."""
    import torch
    from oracle import msa_ref as O
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core (count reported in the JSON)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    sd = O.make_weights(42, num_layers=1, embed_positions_msa=embed_positions_msa)
    tokens = O.make_tokens(R, C, 0)
    with torch.no_grad():
        t0 = time.perf_counter()
        x, pm = O.embed(sd, tokens)
        x = x.permute(1, 2, 0, 3)
        t1 = time.perf_counter()
        x, _, rp = O.axial_layer(sd, 0, x, pm)
        t2 = time.perf_counter()
    return (t1 - t0), (t2 - t1), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    R, C, epm, desc = WORKLOADS[args.workload]
    if C is None:                                        # cfg3: bounded sample = the median-length MSA of the batch
        lens = cfg3_lengths()
        C = sorted(lens)[len(lens) // 2]
    tokens = R * C
    times = []
    threads = os.cpu_count()
    for i in range(args.warmup + args.steps):
        t_emb, t_layer, threads = cpu_reference_sample(R, C, epm)
        if i >= args.warmup:
            times.append(t_emb + NL * t_layer)          # one layer timed, x10 layers extrapolated
    per_fwd = sum(times) / len(times)
    value = tokens / per_fwd
    sample = (f"oracle port of the reference fp32 CPU forward: embedding + ONE of {NL} AxialTransformerLayers at the "
              f"full {R}x{C} shape per step, extrapolated x{NL}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_fwd * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "R": R, "C": C, "tokens_per_step": tokens, "layers": NL},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import rnamsm_b200 as pkg
    from rnamsm_b200 import _lib
    # NB: the measured arm never touches oracle/: weights are the model's own random init (the reference
    # recipe, model.py:89-101), tokens are generated here.  Only cpu_reference_sample() imports the oracle.

    R, C, epm, desc = WORKLOADS[args.workload]
    farm = C is None                                   # cfg3: a list of MSAs per rank instead of one shape
    if farm:
        lens = cfg3_lengths()
        mine = lpt_assign([total_flops(R, c) for c in lens], world)[rank]
        my_C = [lens[i] for i in mine]
        C = max(lens)
        tokens_per_step = R * sum(lens)                # whole job (all ranks)
    else:
        tokens_per_step = R * C
    vocab = pkg.Vocab(pkg.Alphabet())
    torch.manual_seed(42)                              # seed_everything(42), RNA_MSM_Inference.py:17
    model = pkg.MSATransformer(vocab, num_layers=NL, embed_positions_msa=epm, precision=args.precision)
    model = model.eval().cuda()
    tok_host = synthetic_tokens(R, C, seed=100 + rank).pin_memory()
    tok_dev = tok_host.cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard = bool(args.shard) and world > 1
    if shard:
        from rnamsm_b200.sharded import sharded_forward
        tok_host = synthetic_tokens(R, C, seed=100).pin_memory()   # the SAME MSA on every rank
        tok_dev = tok_host.cuda()

    if farm:
        farm_host = [synthetic_tokens(R, c, seed=200 + i).pin_memory() for i, c in zip(mine, my_C)]
        farm_dev = [t.cuda() for t in farm_host]
        # --batch-tokens > 0: short MSAs grouped into forward_batch passes; results per MSA
        # are bit-identical to the one-forward-per-MSA farm (tests/test_gpu_model.py)
        groups = pkg.plan_batches([(R, c) for c in my_C], args.batch_tokens) if args.batch_tokens > 0 else None

    def step_device():
        if shard:
            return sharded_forward(model, tok_dev, fused=args.fused)
        if farm:
            out = None
            if groups is not None:
                for g in groups:
                    out = model.forward_batch([farm_dev[i] for i in g], need_head_weights=True)
                return out
            for t in farm_dev:
                out = model(t, repr_layers=[NL], need_head_weights=True, want_logits=False)
            return out
        return model(tok_dev, repr_layers=[NL], need_head_weights=True, want_logits=False)

    emb_host = torch.empty((C - 1, D), dtype=torch.float32).pin_memory()
    atp_host = torch.empty((NL * H, C - 1, C - 1), dtype=torch.float32).pin_memory()

    if farm:                                           # per-MSA views of the pinned result buffers
        farm_atp = [atp_host.view(-1)[:NL * H * (c - 1) ** 2].view(NL * H, c - 1, c - 1) for c in my_C]
        farm_emb = [emb_host.view(-1)[:(c - 1) * D].view(c - 1, D) for c in my_C]

    def step_e2e():
        if farm and groups is not None:                # pinned host tokens in, every MSA's emb + atp back in pinned memory
            pkg.extract_features_batch_streamed(model, farm_host, farm_atp, farm_emb, args.batch_tokens)
            return
        if farm:
            for th in farm_host:                       # contiguous pinned views sized for this MSA
                c = th.shape[-1]
                pkg.extract_features_streamed(model, th, atp_host.view(-1)[:NL * H * (c - 1) ** 2].view(NL * H, c - 1, c - 1),
                                              emb_host.view(-1)[:(c - 1) * D].view(c - 1, D))
            return
        if shard:
            out = sharded_forward(model, tok_host.cuda(non_blocking=True), fused=args.fused)
            if rank == 0:                                           # rank 0 owns MSA row 0 and writes the files
                att = out["row_attentions"][..., 1:, 1:].reshape(-1, C - 1, C - 1)
                atp_host.copy_(att, non_blocking=True)
                emb_host.copy_(out["representations"][NL][0, 0, 1:, :], non_blocking=True)
            torch.cuda.synchronize()
            return
        # public API: forward + the reference's slicing, D2H of every layer's maps overlapped with the next layer
        pkg.extract_features_streamed(model, tok_host, atp_host, emb_host)

    # ---- warm-up, then the device-timed region with per-kernel-class events recording -----------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.lib.rnamsm_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = _lib.lib.rnamsm_launch_count() - launches0
    # per-kernel-class CUDA events over a second, identical region of K steps: two event records around each of the
    # ~132 launches per forward cost ~0.5 ms, which would be charged to `value` if they sat in the region above
    _lib.profile_enable(True)
    barrier()
    for _ in range(args.steps):
        step_device()
    barrier()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)

    # ---- end-to-end region (host tokens in, emb + atp back on the host, every step) --------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish() if sampler else None

    if world > 1:
        tt = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(tt[0]), float(tt[1])

    if rank == 0:
        ms_per_step = ms_total / args.steps
        jobs = 1 if (shard or farm) else world   # sharded / cfg3: tokens_per_step already is the whole job
        value = jobs * tokens_per_step / (ms_per_step * 1e-3)
        e2e_value = jobs * tokens_per_step / (e2e_s / args.steps)
        peaks = measured_peaks()
        if farm:                                        # this rank's MSAs
            fl = {}
            for c in my_C:
                for k, v in flops_breakdown(R, c).items():
                    fl[k] = fl.get(k, 0.0) + v
        else:
            fl = flops_breakdown(R, C)
        if shard:                                       # per-rank share of the one sharded MSA
            fl = {k: v / world for k, v in fl.items()}
        gemm_classes = ["linear_qkv", "linear_out_resid", "linear_fc1_gelu", "linear_fc2_resid"]
        gemm_ms = sum(prof[k][0] for k in gemm_classes)
        gemm_launches = sum(prof[k][1] for k in gemm_classes)
        gemm_flops_per_step = sum(fl[k] for k in gemm_classes)
        kernel_ms_total = sum(v[0] for v in prof.values())
        shares = {k: round(v[0] / kernel_ms_total, 4) for k, v in prof.items() if v[1]}
        tflops = {k: round(fl[k] * args.steps / (prof[k][0] * 1e-3) / 1e12, 1) for k in fl if prof.get(k, (0, 0))[0] > 0}
        achieved = gemm_flops_per_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_sustained"] if args.precision != "fp32" else None
        roofline = {
            "kernel": "umma_gemm_kernel<DENSE> (2-CTA tcgen05 dense linear: QKV / out-proj / fc1+GELU / fc2)"
                      if args.precision != "fp32" else "sgemm_kernel<LinearProb> (fp32 FFMA)",
            "bound": "tensor", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
            "frac": round(achieved / peak, 4) if peak else None,
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (cuBLAS bf16; kind::f16 runs fp16 and bf16 at the same "
                           "rate; kernel timed inside a long step)",
            "flops_per_launch": gemm_flops_per_step / max(1, gemm_launches / args.steps),
            "avg_launch_ms": gemm_ms / max(1, gemm_launches), "launches": gemm_launches,
            "share_of_kernel_time": round(gemm_ms / kernel_ms_total, 4) if kernel_ms_total else None,
            "traffic": measured_traffic(args.workload, args.precision)[0] if not (shard or farm) else None,
            "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01e_traffic.json)",
            "algorithmic_bytes_per_launch": measured_traffic(args.workload, args.precision)[1] if not (shard or farm) else None,
            "class_time_share": shares, "class_tflops": tflops,
            "whole_forward_tflops_per_gpu": round((sum(total_flops(R, c) for c in my_C) if farm else total_flops(R, C) / (world if shard else 1))
                                                  * args.steps / (ms_total * 1e-3) / 1e12, 2),
        }
        cpu = None
        if world == 1:                                  # contract: the CPU baseline is timed at N=1 only
            if farm:                                    # bounded sample: the median-length MSA of the batch
                c_med = sorted(lens)[len(lens) // 2]
                t_emb, t_layer, threads = cpu_reference_sample(R, c_med, epm)
                tokens_per_step_cpu = R * c_med
            else:
                t_emb, t_layer, threads = cpu_reference_sample(R, C, epm)
                tokens_per_step_cpu = tokens_per_step
            per_fwd = t_emb + NL * t_layer
            cpu = {"value": tokens_per_step_cpu / per_fwd, "unit": "tokens/s", "cores": threads, "kind": "port",
                   "sample": f"embedding + ONE of {NL} AxialTransformerLayers of the oracle (fp32 torch CPU port of "
                             f"modules.py:242-267) at the full {R}x{c_med if farm else C} shape = {t_emb + t_layer:.1f} s, x{NL} layers"}
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if shard else "weak",
            "vs_baseline": None, "dtype": {"fp16": "fp16", "bf16": "bf16", "bf16_pure": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": desc, "R": R, "C": C, "tokens_per_step_per_gpu": tokens_per_step, "layers": NL,
                       "batch_tokens": (args.batch_tokens if farm else None),
                       "embed_dim": D, "heads": H, "weights": "random-init (reference recipe model.py:89-101, seed 42)",
                       "precision": f"{args.precision}: 16-bit operands on tcgen05 (kind::f16), fp32 accumulate / residual "
                                    "stream / LayerNorm / softmax / exported maps" if args.precision != "fp32"
                                    else "fp32 FFMA parity path",
                       "parallelism": (f"one MSA sharded over {world} GPUs: rows (tied row attention, fp32 logit all-reduce) "
                                       f"<-> columns (column attention, 16-bit all-to-all), "
                                       + ("fused peer-memory kernels over NVLink" if args.fused else "NCCL over NVLink")) if shard
                                      else f"dp{world} independent MSAs, no collective",
                       "l2": "no explicit flush: per-step working set (>= 1.4 GB activations + 183 MB weights at "
                             "cfg2) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": (sum(t.numel() for t in farm_host) if farm else tok_host.numel()) * 8,
                    "d2h_bytes_per_step": (sum(NL * H * (c - 1) ** 2 + (c - 1) * D for c in my_C) if farm
                                           else atp_host.numel() + emb_host.numel()) * 4,
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "bf16_pure", "fp32"])
    ap.add_argument("--batch-tokens", type=int, default=0,
                    help="cfg3 only: group the MSAs into forward_batch passes of at most this many tokens "
                         "(0 = one forward per MSA, the reference's B=1 loop)")
    ap.add_argument("--shard", action="store_true",
                    help="N > 1: ONE deep MSA sharded over the ranks (rows for tied row attention, columns for column "
                         "attention; NCCL all-reduce + all-to-all) -> strong scaling.  Default: independent MSAs, weak.")
    ap.add_argument("--fused", action="store_true",
                    help="with --shard: the peer-memory schedule (softmax pulling/pushing logits over NVLink, LayerNorm "
                         "pushing rows to the column owners, out-projection reducing into the row owners' residual)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
