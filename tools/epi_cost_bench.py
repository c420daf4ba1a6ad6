#!/usr/bin/env python
"""What the GELU epilogue costs: the fc1 GEMM (M x 3072 x 768) with the bias-only and the bias + GELU epilogue, and the QKV
shape for reference.  Usage: python tools/epi_cost_bench.py [M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnamsm_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    st = L.stream_ptr()
    K = 768
    x = torch.randn(M, K, device="cuda").half()
    for N, epi, label in ((3072, 0, "fc1 shape, bias only"), (3072, 1, "fc1 shape, bias + GELU"), (2304, 0, "qkv shape, bias only"),
                          (2304, 1, "qkv shape, bias + GELU")):
        W = (torch.randn(N, K, device="cuda") * 0.02).half()
        bias = torch.zeros(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        fn = lambda: L.check(L.lib.rnamsm_linear(L.ptr(x), L.ptr(W), L.ptr(bias), M, N, K, L.F16, epi, 1.0, 0, None, L.ptr(out), st))
        us = timeit(fn)
        print(f"M={M} N={N} K={K} {label}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
