"""The C-ABI library loads, exports every symbol include/snfb.h declares, and its struct sizes
match the Python mirrors.  No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re

import pytest

from sniffles_b200 import abi, binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "snfb.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(snfb_[a-z_0-9]+)\s*\(", text)))


def test_exports_match_header():
    L = binding.lib()
    names = declared_symbols()
    assert set(names) == set(binding.EXPORTS)
    for n in names:
        assert hasattr(L, n), f"libsnfb200.so does not export {n}"


def test_struct_sizes():
    L = binding.lib()
    assert L.snfb_version() == 3
    want = [abi.REC_DTYPE.itemsize, abi.TASK_DTYPE.itemsize, abi.CONTIG_DTYPE.itemsize, C.sizeof(abi.Records), C.sizeof(abi.Config),
            abi.LEAD_DTYPE.itemsize, abi.CAND_DTYPE.itemsize, C.sizeof(abi.GatherView)]
    assert [L.snfb_sizeof(i) for i in range(8)] == want


def test_hash_name_matches_python():
    L = binding.lib()
    for s in (b"chr1", b"ctg17", b"", b"chrUn_KI270442v1"):
        assert L.snfb_hash_name(s, len(s)) == abi.fnv1a64(s)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(binding.SnfbError):
        binding.Context(0)
