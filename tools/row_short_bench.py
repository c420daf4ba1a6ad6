#!/usr/bin/env python
"""Tied row attention of short alignments: the one-launch kernel (rnamsm_row_attn_short) against the three-kernel chain
(logits, softmax, AV) at forward-pass shapes.  Usage: python tools/row_short_bench.py [R C]..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnamsm_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=50):
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
    args = [int(v) for v in sys.argv[1:]] or [512, 36, 256, 64, 256, 100, 1024, 128, 4096, 128]
    D, H = 768, 12
    st = L.stream_ptr()
    for R, C in zip(args[0::2], args[1::2]):
        qkv = (torch.randn(R * C, 3 * D, device="cuda") * 0.4).half()
        ctx = torch.empty(R * C, D, device="cuda", dtype=torch.float16)
        ldp = (C + 7) // 8 * 8
        pmap = torch.empty(H, C, C, device="cuda")
        plp = torch.empty(H, C, ldp, device="cuda", dtype=torch.float16)
        chunks = L.lib.rnamsm_row_attn_short_chunks(R, C, H)
        splits = L.lib.rnamsm_row_attn_splits(R, C, H, L.F16)
        partial = torch.empty(max(chunks, splits), H, C, C, device="cuda")
        sc = 1.0 / R ** 0.5

        def short():
            L.check(L.lib.rnamsm_row_attn_short(L.ptr(qkv), R, C, H, L.F16, None, sc, L.ptr(partial), chunks, L.ptr(pmap),
                                                L.ptr(plp), ldp, L.ptr(ctx), st))

        def chain():
            L.check(L.lib.rnamsm_row_attn_logits(L.ptr(qkv), R, C, H, L.F16, L.ptr(partial), splits, st))
            L.check(L.lib.rnamsm_row_softmax(L.ptr(partial), splits, H, C, None, sc, L.ptr(pmap), L.ptr(plp), ldp, L.F16, st))
            L.check(L.lib.rnamsm_row_attn_av(L.ptr(plp), ldp, L.ptr(qkv), R, C, H, L.F16, L.ptr(ctx), st))

        t_s, t_c = timeit(short), timeit(chain)
        gb = (R * C * 4 * D * 2) / 1e9
        print(f"R={R} C={C}: one launch {t_s:.1f} us ({gb / t_s * 1e6:.0f} GB/s of q|k|v in + ctx out), chain {t_c:.1f} us "
              f"({chunks} chunks / {splits} splits)", flush=True)


if __name__ == "__main__":
    main()
