"""Known-answer vectors: behaviours that are easy to "fix" by accident, pinned by outputs of the unmodified
reference (tests/golden/known/known_answers.json, written by tests/golden/make_known_answers.py in the build container) — util.center / trim /
stdev, leadprov.CIGAR_analyze, and the clusters the reference forms on hand-built blocks (cluster.resplit's negative-index wrap,
compute_metrics' over-long sample, merge_inner on equal leads).  CPU: the oracle; GPU (-m gpu): the CUDA path on the same blocks."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import known_blocks
import oracle.oracle as orc
from sniffles_b200 import abi
from sniffles_b200 import config as sconfig

KA = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "known", "known_answers.json")))


def _longs(v):
    return np.asarray(v, dtype=np.int64)


def test_center_is_the_upper_median_of_the_frequent_values():
    L = orc.lib()
    for v, want in KA["center"]:
        a = _longs(v)
        assert L.so_center(a.ctypes.data, len(a)) == want, v


def test_stdev_and_trim_match_python_statistics_bit_for_bit():
    L = orc.lib()
    for v, want in KA["stdev"]:
        a = _longs(v)
        assert L.so_stdev(a.ctypes.data, len(a)) == want, v
    for v, want in KA["stdev_trim"]:
        a = _longs(v)
        assert L.so_stdev_trim(a.ctypes.data, len(a)) == want, v


def test_cigar_analyze_vectors():
    L = orc.lib()
    for text, want in KA["cigar_analyze"]:
        out = np.zeros(4, np.int64)
        rc = L.so_cigar_analyze(text.encode(), len(text), out.ctypes.data)
        if want is None:
            assert rc != 0, text                         # the reference raises (a str is not an exception: TypeError), callers skip the read
        else:
            assert rc == 0 and out.tolist() == want, text


def _cands(res):
    out = []
    for c in res.cand:
        lo, n = int(c["lead_off"]), int(c["lead_n"])
        out.append(dict(svtype=abi.SVTYPE_NAMES[int(c["svtype"])], pos=int(c["pos"]), svlen=int(c["svlen"]), support=int(c["support"]), n_leads=n,
                        stdev_pos=float(c["stdev_pos"]), stdev_len=float(c["stdev_len"]), lead_svlens=[int(x) for x in res.cand_leads[lo:lo + n]["svlen"]]))
    return out


def _check(name, got):
    want = KA["blocks"][name]["cands"]
    assert len(got) == len(want), name
    for g, w in zip(got, want):
        assert (g["svtype"], g["pos"], g["svlen"], g["support"], g["n_leads"]) == (w["svtype"], w["pos"], w["svlen"], w["support"], w["n_leads"]), (name, g, w)
        assert g["lead_svlens"] == w["lead_svlens"], name
        assert g["stdev_pos"] == pytest.approx(float(w["stdev_pos"]), abs=0) and g["stdev_len"] == pytest.approx(float(w["stdev_len"]), abs=0), name


@pytest.mark.parametrize("name", sorted(known_blocks.CASES))
def test_reference_clusters_on_hand_built_blocks_oracle(name):
    blk = known_blocks.ins_block(known_blocks.CASES[name])
    res = orc.run(blk, abi.Config.from_sniffles(sconfig.default_config(*known_blocks.ARGS)), 2, 1)
    _check(name, _cands(res))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(known_blocks.CASES))
def test_reference_clusters_on_hand_built_blocks_cuda(name):
    from sniffles_b200 import binding
    blk = known_blocks.ins_block(known_blocks.CASES[name])
    ctx = binding.Context(0)
    try:
        ctx.set_config(abi.Config.from_sniffles(sconfig.default_config(*known_blocks.ARGS)))
        ctx.load(blk)
        res = ctx.run()
    finally:
        ctx.close()
    _check(name, _cands(res))
