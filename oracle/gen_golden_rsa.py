#!/usr/bin/env python
"""tests/golden/rsa_pack.npz from the REFERENCE's own RSA pre-processing (build container only):
`one_hot_encode` of _downstream_tasks/RSA/predict.py:67-74 is imported from the file (Bio / matplotlib stubbed:
the packing needs numpy + sklearn only), and the packing statements of its `__main__` block
(predict.py:131-141, `emb = np.load(...)` ... `x_train.to(device, dtype=torch.float)`) are EXECUTED from the
reference file as they stand, with the shipped statistic_dict_{oh,emb}.pickle.  Nothing of the reference is copied
into the repository; the fixture holds the inputs, the statistics and the resulting tensor."""
import importlib.util
import os
import pickle
import sys
import tempfile
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RSA = "/root/reference/_downstream_tasks/RSA"


def load_predict():
    for name in ("Bio", "Bio.SeqIO", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("rsa_predict", os.path.join(RSA, "predict.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def packing_statements():
    lines = open(os.path.join(RSA, "predict.py")).read().splitlines()
    first = next(i for i, l in enumerate(lines) if l.strip().startswith("emb = np.load("))
    last = next(i for i, l in enumerate(lines) if l.strip().startswith("x_train = x_train.to(device"))
    return textwrap.dedent("\n".join(lines[first:last + 1]))


def main():
    P = load_predict()
    stat = {}
    for n in ("oh", "emb"):
        d = pickle.load(open(os.path.join(RSA, "models", "OH+RNA-MSM_Emb", f"statistic_dict_{n}.pickle"), "rb"))
        stat[n] = (d["mu"], d["std"])
    rng = np.random.default_rng(11)
    seq = "AGCUXN-ACGUUGCAACGNUGGCAUCGAUUAGCNA"               # 35 nt (the 2DRB_1 length), unknown letters included
    L = len(seq)
    emb = (rng.standard_normal((L, 768)) * 0.7).astype(np.float32)      # what *_emb.npy holds
    with tempfile.TemporaryDirectory() as tmp:
        ffeat = os.path.join(tmp, "{pdbid}_emb.npy")
        np.save(ffeat.format(pdbid="g"), emb)
        ns = dict(np=np, torch=torch, one_hot_encode=P.one_hot_encode, ffeat=ffeat, pdbid="g", seq=seq, device="cpu",
                  mu_emb=stat["emb"][0], std_emb=stat["emb"][1], mu_oh=stat["oh"][0], std_oh=stat["oh"][1])
        exec(packing_statements(), ns)
    x = ns["x_train"].contiguous().numpy()
    assert x.shape == (1, 773, L) and x.dtype == np.float32
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rsa_pack.npz"), seq=np.array(seq), emb=emb, x=x,
                        mu_emb=stat["emb"][0], std_emb=stat["emb"][1], mu_oh=stat["oh"][0], std_oh=stat["oh"][1])
    print(x.shape, x.dtype, float(x[0, :4].min()), float(x[0, -1].min()))


if __name__ == "__main__":
    main()
