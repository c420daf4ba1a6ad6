#!/usr/bin/env python
"""Time the dense tcgen05 GEMM (rnamsm_linear) and the tied row-attention GEMMs through the C ABI
at forward-pass shapes.  CUDA events, L2-sized inputs.  Usage: python tools/gemm_bench.py [R C]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnamsm_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    M, D, F, H = R * C, 768, 3072, 12
    dev = "cuda"
    st = L.stream_ptr()
    for code, name in ((L.BF16, "bf16"), (L.F16, "fp16")):
        dt = L.torch_dtype(code)
        x = torch.randn(M, D, device=dev).to(dt)
        h = torch.randn(M, F, device=dev).to(dt)
        res = torch.zeros(M, D, device=dev)
        for (N, K, epi, label) in ((3 * D, D, 0, "qkv"), (D, D, 2, "out+resid"), (F, D, 1, "fc1+gelu"), (D, F, 2, "fc2+resid")):
            W = (torch.randn(N, K, device=dev) * 0.02).to(dt)
            bias = torch.zeros(N, device=dev)
            inp = h if K == F else x
            out = res if epi == 2 else torch.empty(M, N, device=dev, dtype=dt)
            fn = lambda: L.check(L.lib.rnamsm_linear(L.ptr(inp), L.ptr(W), L.ptr(bias), M, N, K, code, epi, 1.0, 0, None,
                                                     L.ptr(out), st))
            ms = timeit(fn)
            print(f"{name} {label:10s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        qkv = torch.randn(M, 3 * D, device=dev).to(dt)
        splits = L.lib.rnamsm_row_attn_splits(R, C, H, code)
        partial = torch.empty(splits, H, C, C, device=dev)
        ms = timeit(lambda: L.check(L.lib.rnamsm_row_attn_logits(L.ptr(qkv), R, C, H, code, L.ptr(partial), splits, st)))
        print(f"{name} row_logits R={R} C={C} splits={splits}: {ms:.3f} ms  {2.0 * R * C * C * D / ms / 1e9:.0f} TFLOP/s", flush=True)
        ldp = (C + 7) // 8 * 8
        probs = torch.rand(H, C, ldp, device=dev).to(dt)
        ctx = torch.empty(M, D, device=dev, dtype=dt)
        ms = timeit(lambda: L.check(L.lib.rnamsm_row_attn_av(L.ptr(probs), ldp, L.ptr(qkv), R, C, H, code, L.ptr(ctx), st)))
        print(f"{name} row_av     R={R} C={C}: {ms:.3f} ms  {2.0 * R * C * C * D / ms / 1e9:.0f} TFLOP/s", flush=True)
        for cm in (0, 1):
            ms = timeit(lambda: L.check(L.lib.rnamsm_col_attn(L.ptr(qkv), R, C, H, code, cm, None, L.ptr(ctx), st)))
            print(f"{name} col_attn   R={R} C={C} col_major={cm}: {ms:.3f} ms  {4.0 * R * R * C * D / ms / 1e9:.0f} TFLOP/s",
                  flush=True)
        if code == L.BF16:
            print("gemm pairs:", L.lib.rnamsm_gemm_pairs(), flush=True)
        break_after = os.environ.get("GEMM_BENCH_ONE")
        if break_after:
            break


if __name__ == "__main__":
    main()
