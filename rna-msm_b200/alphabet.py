"""RNA alphabet / vocabulary and the minimal MSA ingest the forward path needs.

Mirrors the *interface* the model consumes from the reference -- ``len(vocab)``, ``pad_idx``,
``prepend_bos``, ``append_eos``, ``eos_idx`` (utils/tokenization.py:15-260, built from
``msm.data.Alphabet.from_architecture("rna language")``, msm/data.py:166-172) -- and the token
grid format ``int64 [R, L+1]`` with column 0 = ``<cls>`` (utils/tokenization.py:107-129).
The heavy preprocessing (hhfilter sub-sampling, Biopython parsing) is the reference's CPU data
plane and stays out of scope (SURVEY.md section 8f ranks it 'next').
"""
from __future__ import annotations

import re
from typing import List, Sequence, Tuple

import numpy as np
import torch

RNA_TOKS = ("A", "G", "C", "U", "X", "N", "-")          # msm/constants.py:10-12


class Alphabet:
    """msm/data.py:92-175 for the 'rna language' architecture."""

    def __init__(self, standard_toks: Sequence[str] = RNA_TOKS,
                 prepend_toks: Sequence[str] = ("<cls>", "<pad>", "<eos>", "<unk>"),
                 append_toks: Sequence[str] = ("<mask>",), prepend_bos: bool = True, append_eos: bool = False,
                 use_msa: bool = True):
        self.standard_toks = list(standard_toks)
        self.prepend_toks = list(prepend_toks)
        self.append_toks = list(append_toks)
        self.prepend_bos = prepend_bos
        self.append_eos = append_eos
        self.use_msa = use_msa
        self.all_toks = list(self.prepend_toks) + list(self.standard_toks) + list(self.append_toks)
        self.tok_to_idx = {tok: i for i, tok in enumerate(self.all_toks)}
        self.unk_idx = self.tok_to_idx["<unk>"]
        self.padding_idx = self.get_idx("<pad>")
        self.cls_idx = self.get_idx("<cls>")
        self.mask_idx = self.get_idx("<mask>")
        self.eos_idx = self.get_idx("<eos>")

    def __len__(self):
        return len(self.all_toks)

    def get_idx(self, tok):
        return self.tok_to_idx.get(tok, self.unk_idx)

    def get_tok(self, ind):
        return self.all_toks[ind]

    @classmethod
    def from_architecture(cls, name: str = "rna language") -> "Alphabet":
        if name != "rna language":
            raise ValueError("rnamsm_b200 ships the 'rna language' alphabet only")
        return cls()


class Vocab:
    """The subset of utils/tokenization.Vocab the model and the inference script use."""

    def __init__(self, alphabet: Alphabet):
        self.alphabet = alphabet
        self.tokens = list(alphabet.all_toks)
        self.tokens_to_idx = dict(alphabet.tok_to_idx)
        self.pad_idx = alphabet.padding_idx
        self.bos_idx = alphabet.cls_idx
        self.eos_idx = alphabet.eos_idx
        self.mask_idx = alphabet.mask_idx
        self.unk_idx = alphabet.unk_idx
        self.prepend_bos = alphabet.prepend_bos
        self.append_eos = alphabet.append_eos

    @classmethod
    def from_esm_alphabet(cls, alphabet: Alphabet) -> "Vocab":       # utils/tokenization.py:198-209
        return cls(alphabet)

    def __len__(self):
        return len(self.tokens)

    def encode(self, sequences: Sequence[str]) -> np.ndarray:
        """Aligned sequences -> int64 [R, L + bos + eos]; unknown symbols -> <unk>."""
        L = len(sequences[0])
        if any(len(s) != L for s in sequences):
            raise ValueError("MSA rows must have equal length")
        lut = np.full(256, self.unk_idx, dtype=np.int64)
        for tok, idx in self.tokens_to_idx.items():
            if len(tok) == 1:
                lut[ord(tok)] = idx
        arr = np.frombuffer("".join(sequences).encode("ascii"), dtype=np.uint8).reshape(len(sequences), L)
        out = lut[arr]
        pads = [(0, 0), (int(self.prepend_bos), int(self.append_eos))]
        return np.pad(out, pads, constant_values=[(0, 0), (self.bos_idx, self.eos_idx)])


def read_msa(path: str, max_seqs: int = 512) -> Tuple[List[str], List[str]]:
    """FASTA/.a2m_msa2 reader with the reference's cleaning rules (utils/align.py:311-313: drop
    lowercase/'.'/'*', T -> U, IUPAC ambiguity codes -> X).  Keeps the FIRST ``max_seqs`` rows:
    the reference sub-samples with the external ``hhfilter`` binary first (utils/align.py:165-173),
    which is not reproduced here."""
    names: List[str] = []
    seqs: List[str] = []
    cur: List[str] = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:])
                cur = []
            elif line:
                cur.append(line)
    if names:
        seqs.append("".join(cur))
    out = []
    for s in seqs[:max_seqs]:
        s = re.sub(r"([a-z]|\.|\*)", "", s)
        s = re.sub(r"[T]", "U", s)
        s = re.sub(r"[RYKMSWBDHVN]", "X", s)
        out.append(s)
    return names[:max_seqs], out


def tokenize_msa(path: str, vocab: Vocab, max_seqs: int = 512, max_seqlen: int = 1024) -> torch.Tensor:
    """-> int64 [R, C]; crops to ``max_seqlen`` tokens like RandomCropDataset (dataset.py:142-158;
    the reference crops at a random offset when too long, we keep the leading window)."""
    _, seqs = read_msa(path, max_seqs)
    tok = torch.from_numpy(vocab.encode(seqs))
    return tok[:, :max_seqlen].contiguous()
