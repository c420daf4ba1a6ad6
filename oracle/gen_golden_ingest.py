#!/usr/bin/env python
"""Generate tests/golden/ingest.npz with the REFERENCE's own utils.align.MSA.greedy_select
(build container only; Bio / tape are stubbed because greedy_select needs numpy + scipy only)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def reference_msa_class():
    for name in ("Bio", "Bio.SeqIO", "Bio.Seq"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["Bio.Seq"].Seq = object
    sys.path.insert(0, REF)
    from utils.align import MSA
    return MSA


def main():
    sys.path.insert(0, ROOT)
    from oracle import msa_ingest_ref as I
    from oracle import msa_ref as O
    MSA = reference_msa_class()
    out = {}
    rng = np.random.default_rng(7)
    cases = {"rand": (60, 33, 16), "dups": (80, 20, 24), "deep": (300, 48, 64)}
    for name, (N, L, num) in cases.items():
        if name == "dups":       # many identical rows -> exact ties in the mean distance
            base = ["".join(rng.choice(list("AGCU-"), L)) for _ in range(12)]
            seqs = [base[i] for i in rng.integers(0, 12, N)]
        else:
            seqs = ["".join(rng.choice(list("AGCU-X"), L, p=[.22, .22, .22, .22, .1, .02])) for _ in range(N)]
        for mode in ("max", "min"):
            sel = MSA.from_sequences(seqs).greedy_select(num, mode=mode)
            # map the selected sequences back to indices exactly as the reference returns them (sorted indices)
            idx = _indices_of(MSA, seqs, num, mode)
            assert [seqs[i] for i in idx] == sel.sequences
            out[f"{name}_{mode}_idx"] = np.array(idx, dtype=np.int32)
        out[f"{name}_chars"] = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(N, L).copy()
        out[f"{name}_num"] = num
    # the shipped MSA of BASELINE config 1: cleaned like from_fasta, diversity-max down to 512 rows
    path = os.path.join(REF, "results", "2DRB_1.a2m_msa2")
    names, seqs = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            names.append(line); seqs.append("")
        elif line:
            seqs[-1] += line
    clean = [I.clean_sequence(s) for s in seqs]
    idx = _indices_of(MSA, clean, 512, "max")
    out["2drb_max_idx"] = np.array(idx, dtype=np.int32)
    out["2drb_depth"] = len(clean)
    out["2drb_chars"] = np.frombuffer("".join(clean).encode(), dtype=np.uint8).reshape(len(clean), -1).copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def _indices_of(MSA, seqs, num, mode):
    """Re-run the reference algorithm but capture the index list it builds (greedy_select returns sequences only):
    a thin subclass whose select() records the indices it is called with."""
    rec = {}

    class Capture(MSA):
        def select(self, indices, axis="seqs"):
            rec["idx"] = list(int(i) for i in indices)
            return super().select(indices, axis)
    Capture.from_sequences(seqs).greedy_select(num, mode=mode)
    return rec.get("idx", list(range(len(seqs))))


if __name__ == "__main__":
    main()
