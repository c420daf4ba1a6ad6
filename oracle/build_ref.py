"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged where they can travel.

TEST / BASELINE INFRASTRUCTURE ONLY (same status as oracle/msa_ref.py): nothing under ``rna-msm_b200/`` may
import it.  ``oracle/_ref/`` is git-ignored (reference sources never enter this repo's history) but not
gpurun-ignored, so the staged files ride to the GPU box with the snapshot -- ``/root/reference`` does not exist
there.  ``__graft_entry__.build()`` calls :func:`build` in the build container, where the read-only checkout is.

The reference is pure Python (no native code), so "building" it is staging exactly the files its own
``AxialTransformerLayer`` / ``MSATransformer`` import chain needs, byte for byte:

    modules.py               the classes RNA_MSM_Inference.py really runs (modules.py:191-267, 688-945)
    utils/tensor.py          symmetrize / apc, imported by modules.py:6
    utils/__init__.py
    product_key_memory.py    imported by modules.py:7 (unused on this path)
    msm/*.py                 the torch-only twin package (msm.MSATransformer, msm.data.Alphabet)

``python oracle/build_ref.py [--ref /root/reference] [--check]``
"""
from __future__ import annotations

import argparse
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
DEFAULT_REF = "/root/reference"
FILES = ["modules.py", "product_key_memory.py", "utils/__init__.py", "utils/tensor.py",
         "msm/__init__.py", "msm/axial_attention.py", "msm/constants.py", "msm/data.py", "msm/model.py",
         "msm/modules.py", "msm/multihead_attention.py"]


def available() -> bool:
    """True when every staged file is present (on the GPU box: shipped with the snapshot)."""
    return all(os.path.isfile(os.path.join(OUT, f)) for f in FILES)


def build(ref_root: str = DEFAULT_REF, verbose: bool = False) -> bool:
    """Stage the reference files from the read-only checkout.  Returns False (and leaves whatever is already
    staged in place) when the checkout is not there -- the GPU box only uses the prebuilt copy."""
    if not os.path.isdir(os.path.join(ref_root, "msm")):
        return available()
    for f in FILES:
        src, dst = os.path.join(ref_root, f), os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print(f"staged {f}")
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified files of yikunpku/RNA-MSM staged by oracle/build_ref.py for the CPU baseline arm.\n"
                 "Not part of the product; git-ignored.\n" + "\n".join(FILES) + "\n")
    return available()


def check(ref_root: str = DEFAULT_REF) -> bool:
    """Every staged file is byte-identical to the checkout's (build container only)."""
    return all(filecmp.cmp(os.path.join(ref_root, f), os.path.join(OUT, f), shallow=False) for f in FILES)


def import_reference():
    """-> (modules, msm): the reference's own top-level ``modules`` and its ``msm`` package, from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is not staged (run python oracle/build_ref.py where /root/reference exists)")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import modules as ref_modules          # noqa: E402  (the reference's modules.py)
    import msm as ref_msm                  # noqa: E402
    assert os.path.dirname(os.path.abspath(ref_modules.__file__)) == OUT, ref_modules.__file__
    return ref_modules, ref_msm


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    ok = build(a.ref, verbose=True)
    print(f"oracle/_ref staged: {ok}")
    if a.check:
        print(f"byte-identical to {a.ref}: {check(a.ref)}")
    sys.exit(0 if ok else 1)
