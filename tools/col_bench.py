#!/usr/bin/env python
"""Stand-alone timing of the column attention (rnamsm_col_attn, column-major q|k|v) at forward-pass shapes.
Usage: python tools/col_bench.py [R C]..."""
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
    args = [int(v) for v in sys.argv[1:]] or [512, 256, 4096, 128, 1024, 1024, 256, 300]
    shapes = list(zip(args[0::2], args[1::2]))
    D, H = 768, 12
    st = L.stream_ptr()
    for R, C in shapes:
        qkv = (torch.randn(R * C, 3 * D, device="cuda") * 0.5).half()
        ctx = torch.empty(R * C, D, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: L.check(L.lib.rnamsm_col_attn(L.ptr(qkv), R, C, H, L.F16, 1, None, L.ptr(ctx), st)))
        print(f"col_attn R={R} C={C}: {ms:.3f} ms  {4.0 * R * R * C * D / ms / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
