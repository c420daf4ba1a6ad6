#!/usr/bin/env python
"""Stand-alone timing of the residual GEMM + LayerNorm pair, un-fused (rnamsm_linear + rnamsm_layernorm) against the
fused rnamsm_linear_residual_layernorm, at forward-pass shapes.  Usage: python tools/lnfuse_bench.py [M] [fused-only]"""
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
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    fused_only = len(sys.argv) > 2
    D, F = 768, 3072
    st = L.stream_ptr()
    code = L.F16
    for K in (D, F):
        x = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(D, K, device="cuda") * 0.02).half()
        bias = torch.zeros(D, device="cuda")
        lw, lb = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
        resid = torch.zeros(M, D, device="cuda")
        y = torch.empty(M, D, device="cuda", dtype=torch.float16)
        cnt = torch.zeros(2 * ((M + 255) // 256), dtype=torch.int32, device="cuda")
        fl = 2.0 * M * D * K

        def plain():
            L.check(L.lib.rnamsm_linear(L.ptr(x), L.ptr(W), L.ptr(bias), M, D, K, code, 2, 1.0, 0, None, L.ptr(resid), st))

        def ln():
            L.check(L.lib.rnamsm_layernorm(L.ptr(resid), L.ptr(lw), L.ptr(lb), L.ptr(y), code, M, D, 1e-5, 0, 0, st))

        def fused():
            L.check(L.lib.rnamsm_linear_residual_layernorm(L.ptr(x), L.ptr(W), L.ptr(bias), M, D, K, code, L.ptr(resid), L.ptr(lw),
                                                           L.ptr(lb), 1e-5, L.ptr(y), code, 0, 0, L.ptr(cnt), st))
        if fused_only:
            for _ in range(3):
                fused()
            torch.cuda.synchronize()
            continue
        tp, tl, tf = timeit(plain), timeit(ln), timeit(fused)
        print(f"M={M} K={K}: residual GEMM {tp*1e3:.0f} us ({fl/tp/1e9:.0f} TF/s) + LayerNorm {tl*1e3:.0f} us = {(tp+tl)*1e3:.0f} us;"
              f"  fused {tf*1e3:.0f} us ({fl/tf/1e9:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
