"""Generates tests/golden/known/known_answers.json by asking the UNMODIFIED reference (imported from /root/reference through oracle/pyref):
the known-answer vectors of SURVEY.md Appendix A — util.center / trim / stdev, leadprov.CIGAR_analyze — and the clusters the reference
forms on the hand-built blocks of tests/known_blocks.py (cluster.resplit's negative-index wrap, compute_metrics' over-long sample).
Run in the build container:  python tests/golden/make_known_answers.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
import harness            # noqa: E402
import known_blocks       # noqa: E402


def main():
    harness.import_reference()
    from sniffles import leadprov, util
    out = {"center": [], "stdev_trim": [], "stdev": [], "cigar_analyze": [], "blocks": {}}
    for v in ([50, 50, 50, 51, 51, 60], [100, 101, 102, 103], [5, 5, 5, 5, 9, 9, 7], [300, -300], [7], [1, 2], [3, 3, 4, 4], list(range(40)), [10, 10, 11, 11, 12, 12, 12]):
        out["center"].append([v, util.center(v)])
    for v in ([1, 2, 3, 4, 5, 6, 7, 8], [1, 2, 3], [10, 12, 12, 13, 20, 40, 41, 43], [5] * 9, list(range(0, 1000, 7)), [2 ** 31 - 1, -(2 ** 31), 17, 0]):
        out["stdev_trim"].append([v, util.stdev(util.trim(v))])
        out["stdev"].append([v, util.stdev(v)])
    for c in ("100S50M10I40M20S", "5H100M", "30M5D30M", "10S20M3P5M", "50M", "7S3H40M2D8M4S", "12=3X5N20=", "M", ""):
        try:
            out["cigar_analyze"].append([c, list(leadprov.CIGAR_analyze(c))])
        except Exception:
            out["cigar_analyze"].append([c, None])
    for name, svlens in known_blocks.CASES.items():
        blk = known_blocks.ins_block(svlens)
        cfg = harness.make_config(*known_blocks.ARGS)
        res = harness.run_task(blk, 0, cfg, finalize=False)
        out["blocks"][name] = dict(svlens=svlens, read_count=res["read_count"],
                                   cands=[dict(svtype=c["svtype"], pos=c["pos"], svlen=c["svlen"], support=c["support"], n_leads=c["n_leads"], cluster_id=c["cluster_id"],
                                               stdev_pos=c["stdev_pos"], stdev_len=c["stdev_len"], lead_svlens=[ld[2] for ld in c["leads"]]) for c in res["cands"]])
    with open(os.path.join(ROOT, "tests", "golden", "known", "known_answers.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out["blocks"].items():
        print(k, [(c["cluster_id"], c["svlen"], c["support"], c["lead_svlens"][:8]) for c in v["cands"]])


if __name__ == "__main__":
    main()
