#!/usr/bin/env python
"""tests/golden/ss_pack.npz from the REFERENCE's own SS pre-processing functions (build container only):
one_hot_encode (sequence_encode.py), outer_concatenation (outer_concatenation.py), the transpose / concatenate of
DataProcess.feature_load (data_processing.py:20-23,42-47) and format_input_shape (data_fomat.py:52-56)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRE = "/root/reference/_downstream_tasks/SS/code/pre_processing"


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(PRE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    enc, outer = load("sequence_encode"), load("outer_concatenation")
    rng = np.random.default_rng(3)
    seq = "AGCUXN-ACGUUGCAACGNU"
    L = len(seq)
    atp = rng.random((120, L, L)).astype(np.float32)            # what *_atp.npy holds
    am = atp.transpose(1, 2, 0)                                 # load_am, data_processing.py:20-23
    oh = enc.one_hot_encode(seq)
    pair = outer.outer_concatenation(oh, oh)                    # [L, L, 8]
    x = np.concatenate((pair, am), axis=2)                      # data_processing.py:46
    x = np.transpose(np.expand_dims(x, 0), (0, 3, 1, 2)).astype(np.float32)   # data_fomat.py:52-56 (.to(float))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ss_pack.npz"), seq=np.array(seq), atp=atp, x=x)
    print(x.shape, x.dtype)


if __name__ == "__main__":
    main()
