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


def test_fast_exact_sqrt_matches_cpython():
    """The kernels' statistics.stdev arithmetic: a floating-point guess verified exactly (256-bit comparisons against the squared midpoints)
    must give CPython's own correctly rounded sqrt(P / Q) — checked against statistics._float_sqrt_of_frac and the limb-by-limb restatement."""
    import random
    import statistics
    L = binding.lib()
    rnd = random.Random(11)
    cases = []
    for _ in range(4000):
        n = rnd.randrange(2, rnd.choice([5, 40, 120, 5000]))
        v = [rnd.randrange(0, rnd.choice([2, 3, 50, 1000, 10 ** 5, 2 ** 31])) for _ in range(min(n, 60))]
        n = len(v)
        sx, sxx = sum(v), sum(x * x for x in v)
        cases.append((n * sxx - sx * sx, n * (n - 1)))
    # perfect squares, exact midpoints' neighbourhood, powers of two, huge and tiny ratios
    for k in (1, 2, 3, 4, 9, 16, 2 ** 20, 2 ** 52, 2 ** 53 + 1, (2 ** 53 - 1) ** 2, (2 ** 53 + 1) ** 2, 2 ** 100, 2 ** 127 + 12345, 3 * 2 ** 120 + 1):
        for q in (1, 2, 3, 6, 12, 90, 9900, 2 ** 40 + 1, 2 ** 62 - 57):
            cases.append((k, q))
    m = 2 ** 52 + 12345
    for mid in ((2 * m + 1), (2 * m - 1), 2 ** 53 + 1, 2 ** 54 - 1):
        cases += [(mid * mid, 4), (mid * mid + 1, 4), (mid * mid - 1, 4), (mid * mid, 4 * 49)]
    bad = 0
    for P, Q in cases:
        if P <= 0 or P >= 2 ** 128 or Q >= 2 ** 63:
            continue
        want = statistics._float_sqrt_of_frac(P, Q)
        fast = L.snfb_selftest_sqrt_frac(P >> 64, P & (2 ** 64 - 1), Q, 0)
        slow = L.snfb_selftest_sqrt_frac(P >> 64, P & (2 ** 64 - 1), Q, 1)
        if not (fast == want == slow):
            bad += 1
            print(P, Q, want, fast, slow)
    assert bad == 0
