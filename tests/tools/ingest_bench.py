#!/usr/bin/env python
"""Time the device MSA ingest (clean + greedy diversity select + tokenise) against the CPU oracle's restatement of
the reference (numpy / scipy calls of utils/align.py:128-148) on a synthetic alignment."""
import os, sys, time, tempfile
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rnamsm_b200 as pkg
from rnamsm_b200.ingest import ingest_msa
from oracle import msa_ingest_ref as I

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    Lc = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    num = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    rng = np.random.default_rng(0)
    base = rng.choice(list(b"AGCU-"), size=(N, Lc)).astype(np.uint8)
    path = os.path.join(tempfile.mkdtemp(), "big.a2m_msa2")
    with open(path, "wb") as f:
        for n in range(N):
            f.write(b">s%d\n" % n + bytes(base[n]) + b"\n")
    v = pkg.Vocab(pkg.Alphabet())
    for method in ("first", "diversity-max"):
        ingest_msa(path, v, num, sample_method=method); torch.cuda.synchronize()
        t0 = time.perf_counter(); tok, rows = ingest_msa(path, v, num, sample_method=method); torch.cuda.synchronize()
        print(f"GPU ingest {method}: N={N} L={Lc} -> {tuple(tok.shape)} in {(time.perf_counter() - t0) * 1e3:.1f} ms (incl. file read)", flush=True)
    t0 = time.perf_counter(); idx = I.greedy_select_indices(base, num, "max"); t1 = time.perf_counter()
    print(f"CPU oracle greedy_select: {(t1 - t0):.2f} s; selections identical: {idx == rows.cpu().tolist()}", flush=True)

if __name__ == "__main__":
    main()
