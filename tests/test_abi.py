"""CPU: the C-ABI shared library loads and exports every symbol include/rnamsm_b200.h declares
(no compute calls -- there is no GPU here), and the ctypes mirrors agree with the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rnamsm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rnamsm_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("rnamsm_msa_forward", "rnamsm_layer_forward", "rnamsm_linear", "rnamsm_row_attn_logits",
              "rnamsm_row_softmax", "rnamsm_row_attn_av", "rnamsm_col_attn", "rnamsm_embed_layernorm",
              "rnamsm_layernorm", "rnamsm_last_error", "rnamsm_version"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from rnamsm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    raw = C.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.rnamsm_version() == _lib.ABI_VERSION == 3
    assert _lib.lib.rnamsm_launch_count() == 0


def test_ctypes_struct_layout_matches_header():
    from rnamsm_b200 import _lib
    p = C.sizeof(C.c_void_p)
    assert C.sizeof(_lib.AttnWeights) == 7 * p                # 6 pointers + int dtype (padded)
    assert _lib.AttnWeights.dtype.offset == 6 * p
    assert C.sizeof(_lib.LayerWeights) == 2 * 7 * p + 6 * p
    assert C.sizeof(_lib.ModelWeights) == 8 * 4 + 14 * p      # 7 ints + float, then 14 pointers
    assert _lib.ModelWeights.tok_emb.offset == 32


def test_workspace_bytes_is_pure_host_math():
    from rnamsm_b200 import _lib
    # bf16: xn 2 B + qkvh 12 B... per feature; must grow with R*C and be 256-aligned
    a = _lib.lib.rnamsm_workspace_bytes(64, 64, 768, 12, 3072, _lib.F32)
    b = _lib.lib.rnamsm_workspace_bytes(128, 64, 768, 12, 3072, _lib.F32)
    assert 0 < a < b and a % 256 == 0
    assert _lib.lib.rnamsm_workspace_bytes(0, 64, 768, 12, 3072, _lib.BF16) == 0


def test_batch_workspace_bytes_is_pure_host_math():
    """rnamsm_batch_workspace_bytes: token-proportional regions sum over the MSAs, the per-MSA attention scratch is
    the maximum; invalid batches (fp32, a single-row MSA) report 0 with a message instead of a size."""
    import ctypes as C
    from rnamsm_b200 import _lib

    def size(shapes, code=_lib.F16):
        n = len(shapes)
        R = (C.c_int * n)(*[s[0] for s in shapes])
        Cc = (C.c_int * n)(*[s[1] for s in shapes])
        return _lib.lib.rnamsm_batch_workspace_bytes(n, R, Cc, 768, 12, 3072, code)

    one = size([(64, 40)])
    single = _lib.lib.rnamsm_workspace_bytes(64, 40, 768, 12, 3072, _lib.F16)
    assert one % 256 == 0 and single < one <= single + 64 * 40 + 512          # + the pad flags of the batch entry
    two = size([(64, 40), (64, 40)])
    tok = 64 * 40 * (768 * 2 + 3072 * 2 + 1)                                   # xn + qkv|ctx / FFN hidden + pad, per MSA
    assert abs((two - one) - tok) <= 3 * 256                                   # attention scratch is shared, not doubled
    assert size([(64, 40), (32, 90)]) > size([(64, 40), (32, 40)])             # ... and sized by the widest MSA
    assert size([(64, 40)], _lib.F32) == 0
    assert size([(64, 40), (1, 40)]) == 0
    assert b"R >= 2" in _lib.lib.rnamsm_last_error()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The shipped .so really is a tcgen05/TMA build (UTCHMMA / UTMALDG / LDTM in SASS)."""
    import shutil
    import subprocess
    from rnamsm_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM"):
        assert mnemonic in sass, mnemonic
