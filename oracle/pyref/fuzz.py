"""Oracle vs the unmodified reference on random shapes and arguments (build container only; CPU).

    python oracle/pyref/fuzz.py [first_seed] [count]

Every case generates a seeded synthetic block, runs the reference task by task through the stub pysam
(harness.run_task) and the C oracle on the whole block, and compares lead table, candidates, coverage and ALT
sequences exactly (tests/compare.py, the comparison the golden tests use)."""
import logging
import os
import random
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "tests")]
logging.disable(logging.CRITICAL)

import harness  # noqa: E402
import compare  # noqa: E402
import oracle.oracle as orc  # noqa: E402
from sniffles_b200 import abi, synth  # noqa: E402
from sniffles_b200 import config as sconfig  # noqa: E402

ARG_POOL = [[], ["--mosaic"], ["--no-qc"], ["--qc-nm"], ["--phase"], ["--repeat"], ["--minsvlen", "30"], ["--minsvlen", "80"], ["--minsupport", "auto"],
            ["--minsupport", "3"], ["--mapq", "30"], ["--min-alignment-length", "2500"], ["--cluster-binsize", "50"], ["--cluster-r", "1.5"],
            ["--cluster-merge-pos", "60"], ["--cluster-merge-len", "0.2"], ["--cluster-merge-bnd", "500"], ["--long-ins-length", "1200"],
            ["--no-consensus"], ["--detect-large-ins", "False"], ["--cluster-repeat-h", "2.0"], ["--max-splits-base", "1"], ["--dev-no-resplit"]]

_q = {}


def qhash(q):
    if q not in _q:
        _q[q] = orc.qname_hash(q)
    return _q[q]


def one(seed):
    rnd = random.Random(seed)
    lens = [rnd.randrange(120000, 420000) for _ in range(rnd.choice([1, 1, 2, 3]))]
    kw = dict(coverage=rnd.choice([8, 15, 30, 50]), len_mean=rnd.choice([3000.0, 8000.0, 20000.0, 60000.0]), len_sd=rnd.choice([300.0, 2000.0, 6000.0]),
              tech=rnd.choice(["ont", "hifi"]), sv_spacing=rnd.choice([800.0, 3000.0, 15000.0]), phased_frac=rnd.choice([0.0, 0.5, 1.0]),
              tr_frac=rnd.choice([0.0, 0.15, 0.6]), ins_only=rnd.random() < 0.15, clip_prob=rnd.choice([0.0, 0.1, 0.5]), lowmapq_prob=rnd.choice([0.05, 0.3]),
              mosaic=rnd.random() < 0.2, sv_min=rnd.choice([30, 50]), sv_max=rnd.choice([2000, 5000, 12000]))
    args = []
    for a in rnd.sample(ARG_POOL, rnd.choice([0, 1, 1, 2, 3])):
        if a and a[0] not in args:
            args += a
    blk = synth.generate(5000 + seed, lens, kw.pop("coverage"), **kw)
    cfg = harness.make_config(*args)
    ref = [harness.run_task(blk, t, cfg) for t in range(len(blk.task))]
    res = orc.run(blk, abi.Config.from_sniffles(sconfig.default_config(*args)), 3, 2)
    lo_l = lo_c = 0
    for t, r in enumerate(ref):
        rows = compare.ref_lead_rows(r["leadtab"], qhash)
        compare.assert_leads_equal(rows, compare.lead_rows(res.leads[lo_l:lo_l + len(rows)], blk.contig_names), f"task {t} leads")
        lo_l += len(rows)
        assert r["read_count"] == int(res.task_read_count[t])
        nc = len(r["cands"])
        compare.assert_cands_equal(r["cands"], res, blk.contig_names, qhash, lo_c, lo_c + nc)
        assert r["cov_mean"] == float(res.task_cov_mean[t])
        compare.assert_alts_equal(r["final"], res, lo_c, lo_c + nc)
        lo_c += nc
    assert lo_c == len(res.cand) and lo_l == len(res.leads)
    return len(blk.rec), len(res.cand), args


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    bad = 0
    for seed in range(first, first + count):
        try:
            n, c, args = one(seed)
            print(f"seed {seed}: ok  {n} records, {c} candidates, args {args}", flush=True)
        except Exception:
            bad += 1
            print(f"seed {seed}: MISMATCH\n{traceback.format_exc()}", flush=True)
    print(f"{count - bad} of {count} cases identical")
    sys.exit(1 if bad else 0)
