"""MSA ingest.  CPU: the oracle restatement of MSA.greedy_select / from_fasta cleaning /
Vocab.encode against golden vectors produced by the REFERENCE's own utils.align.MSA.greedy_select
(oracle/gen_golden_ingest.py).  GPU: the csrc/ingest.cu kernels against the same vectors -- index-exact."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import msa_ingest_ref as I  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "ingest.npz"))
CASES = ["rand", "dups", "deep"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["max", "min"])
def test_oracle_greedy_select_matches_reference(name, mode):
    idx = I.greedy_select_indices(G[f"{name}_chars"], int(G[f"{name}_num"]), mode)
    assert idx == G[f"{name}_{mode}_idx"].tolist()


def test_oracle_greedy_select_2drb_matches_reference():
    idx = I.greedy_select_indices(G["2drb_chars"], 512, "max")
    assert idx == G["2drb_max_idx"].tolist() and len(idx) == 512 and idx[0] == 0


def test_oracle_cleaning_and_encoding():
    assert I.clean_sequence("AGCTagc.N*RY-") == "AGCUXXX-"
    import rnamsm_b200 as pkg
    v = pkg.Vocab(pkg.Alphabet())
    seqs = ["AGCU-XN", "UUZQ-AG"]
    assert np.array_equal(I.encode(seqs), v.encode(seqs))


def test_split_records_follows_seqio_record_rules(tmp_path):
    """Records start at a '>' in column 0 only; blanks inside sequence lines do not count (Bio.SeqIO's FASTA parser,
    which MSA.from_fasta iterates, utils/align.py:303)."""
    from rnamsm_b200.ingest import cleaned_length, split_records
    p = tmp_path / "x.a2m_msa2"
    p.write_bytes(b"# comment before the first record\n>q a->b |x>y\nAC GU\nacg.U \t\n>second\nACGUU\r\n")
    headers, raw, off = split_records(str(p))
    assert headers == ["q a->b |x>y", "second"]
    assert [cleaned_length(raw[off[i]:off[i + 1]]) for i in range(2)] == [5, 5]


def _write_fasta(path, chars, rng):
    """Rows of a cleaned matrix back into a messy a2m file: lowercase insertions, '.', T for U, IUPAC for X."""
    amb = "RYKMSWBDHV"
    with open(path, "w") as f:
        for n, row in enumerate(chars):
            s = ""
            for ch in bytes(row).decode():
                if rng.random() < 0.15:
                    s += rng.choice(list("acgu.")) * int(rng.integers(1, 3))
                if ch == "U" and rng.random() < 0.5:
                    ch = "T"
                elif ch == "X":
                    ch = str(rng.choice(list(amb)))
                s += ch
            f.write(f">seq{n} some description a->b\n")       # a '>' inside a header is text (Bio.SeqIO)
            for i in range(0, len(s), 23):                      # wrapped lines, some with blanks (SeqIO drops them)
                line = s[i:i + 23]
                if rng.random() < 0.3:
                    k = int(rng.integers(0, len(line) + 1))
                    line = line[:k] + " " + line[k:] + "  "
                f.write(line + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES + ["2drb"])
@pytest.mark.parametrize("method", ["first", "diversity-max", "diversity-min"])
def test_gpu_ingest_matches_reference_selection(name, method, tmp_path):
    import rnamsm_b200 as pkg
    from rnamsm_b200.ingest import ingest_msa
    if name == "2drb" and method == "diversity-min":
        pytest.skip("no reference vector")
    chars = G[f"{name}_chars"]
    num = 512 if name == "2drb" else int(G[f"{name}_num"])
    rng = np.random.default_rng(11)
    path = str(tmp_path / "x.a2m_msa2")
    _write_fasta(path, chars, rng)
    v = pkg.Vocab(pkg.Alphabet())
    tokens, rows = ingest_msa(path, v, max_seqs=num, sample_method=method)
    if method == "first":
        want = list(range(min(num, chars.shape[0])))
    else:
        want = G[f"{name}_{method.split('-')[1]}_idx"].tolist()       # the reference's own selection
    assert rows.cpu().tolist() == want                                # index-exact, exact ties included
    seqs = [bytes(chars[i]).decode() for i in want]
    assert torch.equal(tokens.cpu(), torch.from_numpy(I.encode(seqs)))
    assert tokens.dtype == torch.int64 and tokens.is_cuda


@pytest.mark.gpu
def test_gpu_ingest_rejects_ragged_rows(tmp_path):
    import rnamsm_b200 as pkg
    from rnamsm_b200.ingest import ingest_msa
    p = tmp_path / "bad.a2m_msa2"
    p.write_text(">a\nAGCU\n>b\nAGC\n")
    with pytest.raises(AssertionError, match="Seqlen Mismatch"):
        ingest_msa(str(p), pkg.Vocab(pkg.Alphabet()))
    with pytest.raises(ValueError):
        ingest_msa(str(p), pkg.Vocab(pkg.Alphabet()), sample_method="hhfilter")
