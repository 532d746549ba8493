"""Pins the CPU oracle against the reference: the committed golden fixtures were produced by the
unmodified reference (tests/golden/make_golden.py); the oracle must reproduce them exactly —
lead table, candidates (incl. exact stdev doubles), coverage, INS ALT sequences."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from sniffles_b200 import abi, synth
from sniffles_b200 import config as sconfig
import compare
import oracle.oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "*.json")) if "hg008" not in p and "config" not in p)


def load_fixture(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        fx = json.load(f)
    kw = dict(fx["generator"])
    blk = synth.generate(kw.pop("seed"), kw.pop("contig_len"), kw.pop("coverage"), **kw)
    h = hashlib.sha256()
    for a in (blk.rec, blk.cigar, blk.var, blk.seq, blk.task, blk.tr):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == fx["digest"], "the synthetic generator no longer reproduces the block the fixture was made from"
    if fx.get("n_mask"):
        blk.set_n_mask({int(k): [tuple(x) for x in v] for k, v in fx["n_mask"].items()})
    return fx, blk


_qcache = {}


def qhash(q):
    if q not in _qcache:
        _qcache[q] = orc.qname_hash(q)
    return _qcache[q]


def check_against_golden(fx, blk, res, check_leads=True):
    lo_l = lo_c = 0
    for t, ref in enumerate(fx["tasks"]):
        if check_leads:
            rows = compare.ref_lead_rows(ref["leadtab"], qhash)
            n = len(rows)
            assert (res.leads["task"][lo_l:lo_l + n] == t).all()
            compare.assert_leads_equal(rows, compare.lead_rows(res.leads[lo_l:lo_l + n], blk.contig_names), f"task {t} leads")
            lo_l += n
            assert ref["read_count"] == int(res.task_read_count[t])
        nc = len(ref["cands"])
        compare.assert_cands_equal(ref["cands"], res, blk.contig_names, qhash, lo_c, lo_c + nc)
        assert ref["cov_mean"] == float(res.task_cov_mean[t])
        compare.assert_alts_equal(ref["final"], res, lo_c, lo_c + nc)
        lo_c += nc
    assert lo_c == len(res.cand)
    if check_leads:
        assert lo_l == len(res.leads)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference(name):
    fx, blk = load_fixture(name)
    cfg = abi.Config.from_sniffles(sconfig.default_config(*fx["args"]))
    res = orc.run(blk, cfg, 3, 2)
    check_against_golden(fx, blk, res)
    for t, ref in enumerate(fx["tasks"]):
        assert ref["mean_nm"] == float(res.task_mean_nm[t])      # sequential float sum, same order as the reference


def _bam_block(prefix):
    z = np.load(os.path.join(GOLDEN, "hg008_bnd.npz"))
    rec = z[f"{prefix}_rec"]
    nseq = int(((rec["l_seq"].astype(np.int64) + 1) // 2).sum())
    rec = rec.copy()
    rec["seq_off"] = np.concatenate([[0], np.cumsum((rec["l_seq"].astype(np.int64) + 1) // 2)[:-1]]).astype("<u8")
    return synth.RecordBlock(rec=rec, cigar=z[f"{prefix}_cigar"], var=z[f"{prefix}_var"], seq=np.zeros(nseq, "u1"), task=z[f"{prefix}_task"],
                             contig=z[f"{prefix}_contig"], tr=np.zeros(0, "<i4"), contig_names=[str(x) for x in z[f"{prefix}_names"]])


def bnd_leads_by_record(blk, leads):
    out = {}
    for l in leads:
        if int(l["flags"]) & 7 == abi.BND:
            f = int(l["flags"])
            out[int(l["rec"])] = [int(l["ref_start"]), blk.contig_names[int(l["mate_contig"])], int(l["mate_pos"]), bool(f & abi.LF_BND_FIRST), bool(f & abi.LF_BND_REVERSE)]
    return out


@pytest.mark.parametrize("prefix", ["hg008", "hg002"])
def test_reference_bnd_vectors(prefix):
    """src/tests/test_bnd_leads.py: 8 reads give (23272628, chr5, 52747359, first, reverse) / (21493610, chr20, 25499120, ...);
    same-strand SA reads give no BND lead at HEAD."""
    with open(os.path.join(GOLDEN, "hg008_bnd.json")) as f:
        exp = [e for e in json.load(f)["records"] if e["file"].startswith(prefix)]
    blk = _bam_block(prefix)
    cfg = abi.Config.from_sniffles(sconfig.default_config("--dev-no-qc"))
    res = orc.run(blk, cfg, 1, 1)
    got = bnd_leads_by_record(blk, res.leads)
    assert len(exp) == len(blk.rec)
    for i, e in enumerate(exp):
        assert got.get(i) == e["lead"], (i, e, got.get(i))
    if prefix == "hg008":
        assert sum(e["lead"] is not None for e in exp) == 8
        assert exp[0]["lead"] == [23272628, "chr5", 52747359, True, True] and exp[4]["lead"] == [21493610, "chr20", 25499120, False, False]


def test_sqrt_frac_is_correctly_rounded():
    import fractions
    import statistics
    import random
    rnd = random.Random(5)
    for _ in range(300):
        n = rnd.randrange(2, 120)
        v = [rnd.randrange(0, rnd.choice([3, 100, 10 ** 4, 10 ** 8])) for _ in range(n)]
        sx, sxx = sum(v), sum(x * x for x in v)
        P, Q = n * sxx - sx * sx, n * (n - 1)
        assert orc.lib().so_sqrt_frac(P >> 64, P & (2 ** 64 - 1), Q) == statistics.stdev(v)
