"""CPU checks of the boundary: the C-ABI library builds, loads and exports every symbol include/dajin2_b200.h
declares; the product path refuses to run without a CUDA device (no CPU fallback)."""
from __future__ import annotations

import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_build_and_symbols():
    import __graft_entry__ as entry

    entry.build()
    from dajin2_b200 import _cabi

    lib = _cabi.load()
    header = (ROOT / "include" / "dajin2_b200.h").read_text()
    declared = set(re.findall(r"\b(dj_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared but not exported"
    assert lib.dj_abi_version() == _cabi.ABI_VERSION
    assert lib.dj_code_pitch(5000) == 5024 and lib.dj_code_pitch(32) == 32
    assert lib.dj_ckpt_pitch(15200) >= (15200 + 15) // _cabi.ENC_CHUNK + 1  # one checkpoint per trip of the encoder
    assert lib.dj_kmer_table_bytes(1024) == 64 + 1024 * 28
    assert lib.dj_row_table_bytes(1024) == 64 + 1024 * 24


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from dajin2_b200 import classification, engine
    from dajin2_b200.ingest import pack_rows

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.EncodedReads(pack_rows(["=A,=C"]), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        classification.calc_match("=A,=C")


def test_product_does_not_import_oracle():
    for path in (ROOT / "dajin2_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    for path in (ROOT / "dajin2_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path
