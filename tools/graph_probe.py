#!/usr/bin/env python
"""Does a CUDA graph of the whole forward help short alignments?  Captures the rnamsm_msa_forward call (one C call =
~110 kernel launches) in a torch.cuda.CUDAGraph and compares replay with the eager call.
Usage: python tools/graph_probe.py [R C]..."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rnamsm_b200 as pkg  # noqa: E402
from rnamsm_b200 import _lib as L  # noqa: E402


def timeit(fn, iters):
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
    args = [int(v) for v in sys.argv[1:]] or [512, 36, 256, 100, 128, 300, 512, 256]
    torch.manual_seed(0)
    model = pkg.MSATransformer(pkg.Vocab(pkg.Alphabet()), num_layers=10, precision="fp16").eval().cuda()
    D, H, N = model.embed_dim, model.num_attention_heads, model.num_layers
    for R, Cc in zip(args[0::2], args[1::2]):
        tok = torch.randint(4, 11, (1, R, Cc), device="cuda")
        code = model._code
        fcode = model._fwd_code
        m = model.c_weights(code)
        nbytes = L.lib.rnamsm_workspace_bytes(R, Cc, D, H, 4 * D, fcode) + ((R * Cc + 255) // 256) * 256
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        x = torch.empty((R, Cc, D), device="cuda")
        att = torch.empty((N, H, Cc, Cc), device="cuda")
        reps = (C.c_void_p * (N + 1))(*[None] * (N + 1))

        def call(stream):
            L.check(L.lib.rnamsm_msa_forward(C.byref(m), L.ptr(tok[0]), R, Cc, 0, fcode, L.ptr(x), L.ptr(att), reps, None,
                                             L.ptr(ws), nbytes, stream), "msa_forward")

        eager = timeit(lambda: call(L.stream_ptr()), 20)
        x_ref = x.clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            call(s.cuda_stream)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g, stream=s):
            call(s.cuda_stream)
        x.zero_()
        replay = timeit(g.replay, 20)
        same = bool(torch.equal(x, x_ref))
        print(f"R={R} C={Cc}: eager {eager:.3f} ms, graph replay {replay:.3f} ms ({eager / replay:.3f}x), identical={same}", flush=True)


if __name__ == "__main__":
    main()
