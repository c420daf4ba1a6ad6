"""CPU oracle for the step in front of the hot path ("next" row 1): MSA cleaning,
diversity sub-sampling and tokenisation.  TEST INFRASTRUCTURE ONLY (same rules as oracle/msa_ref.py).

Restates, citing the reference (paths relative to the reference checkout):
  * ``MSA.from_fasta`` cleaning, utils/align.py:304-316: drop lowercase / '.' / '*', T -> U, IUPAC -> X;
  * ``MSA.greedy_select``, utils/align.py:128-148: start from row 0, repeatedly add the unselected row whose
    MEAN Hamming distance to the rows selected so far is largest ("max") / smallest ("min"), first index on
    ties, return the selection sorted;
  * ``Vocab.encode`` for the RNA alphabet, utils/tokenization.py:107-129 + msm/data.py:166-172.

Parity pinning: ``oracle/gen_golden_ingest.py`` runs the reference's own ``utils.align.MSA.greedy_select``
(importable in the build container once ``Bio`` is stubbed -- greedy_select itself needs only numpy/scipy)
and commits its selections as ``tests/golden/ingest.npz``; tests/test_ingest.py checks this restatement
against them, exact ties included (see greedy_select_indices).
"""
from __future__ import annotations

import re
from typing import List, Sequence

import numpy as np


def clean_sequence(s: str) -> str:
    """utils/align.py:311-313."""
    s = re.sub(r"([a-z]|\.|\*)", "", s)
    s = re.sub(r"[T]", "U", s)
    return re.sub(r"[RYKMSWBDHVN]", "X", s)


def greedy_select_indices(array_u8: np.ndarray, num_seqs: int, mode: str = "max") -> List[int]:
    """utils/align.py:128-148 on a uint8 [N, L] character matrix; returns the sorted selected row indices.

    The reference's own sequence of numpy / scipy calls (cdist "hamming", concatenate, np.delete, .mean(0),
    argmax / argmin).  Exact ties between candidates (duplicate rows, small L) are common and are decided by
    float64 rounding there: np.delete returns a Fortran-ordered array, so .mean(0) runs numpy's 8-accumulator
    pairwise sum along the picks (7.1499999999999995 vs 7.15 for mathematically equal sums).  Keeping the very
    same calls keeps the very same outcome; the CUDA kernel reproduces that summation operation for operation."""
    assert mode in ("max", "min")
    from scipy.spatial.distance import cdist
    N, L = array_u8.shape
    if N <= num_seqs:
        return list(range(N))
    optfunc = np.argmax if mode == "max" else np.argmin
    all_indices = np.arange(N)
    indices = [0]
    pairwise_distances = np.zeros((0, N))
    for _ in range(num_seqs - 1):
        dist = cdist(array_u8[indices[-1:]], array_u8, "hamming")
        pairwise_distances = np.concatenate([pairwise_distances, dist])
        shifted_distance = np.delete(pairwise_distances, indices, axis=1).mean(0)
        shifted_index = optfunc(shifted_distance)
        indices.append(int(np.delete(all_indices, indices)[shifted_index]))
    return sorted(indices)


RNA_TOKS = ("<cls>", "<pad>", "<eos>", "<unk>", "A", "G", "C", "U", "X", "N", "-", "<mask>")


def encode(seqs: Sequence[str]) -> np.ndarray:
    """Vocab.encode: int64 [R, L + 1], column 0 = <cls>, unknown symbols -> <unk>."""
    lut = {t: i for i, t in enumerate(RNA_TOKS) if len(t) == 1}
    L = len(seqs[0])
    out = np.full((len(seqs), L + 1), 3, dtype=np.int64)
    out[:, 0] = 0
    for r, s in enumerate(seqs):
        assert len(s) == L
        for c, ch in enumerate(s):
            out[r, c + 1] = lut.get(ch, 3)
    return out
