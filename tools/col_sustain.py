#!/usr/bin/env python
"""Column attention stand-alone, burst vs sustained: the same launch timed over 10 / 100 / 1000 / 3000 back-to-back
iterations, with the SM clock sampled by nvidia-smi in a side thread (is the in-forward rate a clock effect?).
Usage: python tools/col_sustain.py [R C]"""
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnamsm_b200 import _lib as L  # noqa: E402


def clocks(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append((float(r[0]), float(r[1])))
        except Exception:
            pass
        time.sleep(0.05)


def main():
    a = [int(v) for v in sys.argv[1:]] or [512, 256]
    R, C = a[0], a[1]
    D, H = 768, 12
    st = L.stream_ptr()
    qkv = (torch.randn(R * C, 3 * D, device="cuda") * 0.5).half()
    ctx = torch.empty(R * C, D, device="cuda", dtype=torch.float16)
    fn = lambda: L.check(L.lib.rnamsm_col_attn(L.ptr(qkv), R, C, H, L.F16, 1, None, L.ptr(ctx), st))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for iters in (10, 100, 1000, 3000, 10):
        stop, samples = threading.Event(), []
        th = threading.Thread(target=clocks, args=(stop, samples))
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / iters
        mhz = sorted(s[0] for s in samples)
        pw = sorted(s[1] for s in samples)
        print(f"R={R} C={C} iters={iters}: {ms:.3f} ms {4.0 * R * R * C * D / ms / 1e9:.0f} TF/s  clocks {mhz[:1]}..{mhz[-1:]} "
              f"median {mhz[len(mhz) // 2] if mhz else None} power max {pw[-1:] }", flush=True)
        time.sleep(1.0)


if __name__ == "__main__":
    main()
