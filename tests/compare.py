"""Shared comparison helpers: a reference dump (oracle/pyref/harness.run_task, or a committed
tests/golden/*.json made from it) versus a result in the library's struct layout — either
the CPU oracle's (oracle/oracle.py OracleResult) or the device library's (sniffles_b200.binding)."""
import math

import numpy as np

from sniffles_b200 import abi


def lead_rows(leads, contig_names):
    """struct leads -> comparable rows in the layout of harness.lead_tuple (qname as hash)."""
    rows = []
    for l in leads:
        f = int(l["flags"])
        ty = f & 7
        bnd = None
        if ty == abi.BND:
            mc = int(l["mate_contig"])
            bnd = [contig_names[mc] if mc >= 0 else None, int(l["mate_pos"]), bool(f & abi.LF_BND_FIRST),
                   bool(f & abi.LF_BND_REVERSE)]
        rows.append([abi.SVTYPE_NAMES[ty], int(l["ref_start"]), int(l["ref_end"]), int(l["qry_start"]),
                     int(l["qry_end"]), "-" if f & abi.LF_REVERSE else "+", (f >> 16) & 255,
                     abi.SOURCE_NAMES[(f >> 3) & 3], None if f & abi.LF_SVLEN_NONE else int(l["svlen"]),
                     int(l["seq_len"]) if f & abi.LF_HAS_SEQ else None, int(l["qname_hash"]), (f >> 24) & 3,
                     bool(f & abi.LF_IS_SA), int(l["read_len"]), bnd])
    return rows


def ref_lead_rows(leadtab, qhash):
    rows = []
    for svtype in abi.SVTYPE_NAMES:
        for _bin, leads in leadtab.get(svtype, []):
            for t in leads:
                t = list(t)
                t[10] = qhash(t[10])
                rows.append(t)
    return rows


def assert_leads_equal(ref_rows, got_rows, what="leads"):
    assert len(ref_rows) == len(got_rows), f"{what}: count {len(got_rows)} != reference {len(ref_rows)}"
    for i, (a, b) in enumerate(zip(ref_rows, got_rows)):
        assert a == b, f"{what}[{i}] differs:\n  reference {a}\n  got       {b}"


def _feq(a, b):
    if a is None or (isinstance(a, float) and math.isnan(a)):
        return b is None or (isinstance(b, float) and math.isnan(b))
    return b is not None and float(a) == float(b)


def assert_cands_equal(ref_cands, res, contig_names, qhash, lo=0, hi=None, check_alt=None):
    """ref_cands: list of harness._cand_dict for one task; res.cand[lo:hi] the same task's candidates."""
    cand = res.cand[lo:hi]
    assert len(ref_cands) == len(cand), f"candidate count {len(cand)} != reference {len(ref_cands)}"
    for i, (r, c) in enumerate(zip(ref_cands, cand)):
        tag = f"cand[{i}] {r['svtype']}@{r['pos']}"
        assert r["svtype"] == abi.SVTYPE_NAMES[int(c["svtype"])], tag
        for k_ref, k in (("pos", "pos"), ("end", "end"), ("svlen", "svlen"), ("support", "support"), ("qual", "qual"),
                         ("fwd", "fwd"), ("rev", "rev")):
            assert int(r[k_ref]) == int(c[k]), f"{tag}: {k} {int(c[k])} != reference {r[k_ref]}"
        assert bool(r["precise"]) == bool(c["precise"]), f"{tag}: precise"
        got_cov = [int(c[k]) for k in ("cov_upstream", "cov_start", "cov_center", "cov_end", "cov_downstream")]
        assert r["cov"] == got_cov, f"{tag}: coverage {got_cov} != reference {r['cov']}"
        assert _feq(r["stdev_pos"], float(c["stdev_pos"])), f"{tag}: stdev_pos {float(c['stdev_pos'])!r} != {r['stdev_pos']!r}"
        assert _feq(r["stdev_len"], float(c["stdev_len"])), f"{tag}: stdev_len {float(c['stdev_len'])!r} != {r['stdev_len']!r}"
        if r["support_long"] is not None:
            assert r["support_long"] == int(c["support_long"]), f"{tag}: SUPPORT_LONG"
        if r["support_sa"] is not None:
            assert r["support_sa"] == int(c["support_sa"]), f"{tag}: SUPPORT_SA"
        assert r["hap_counts"] == [int(x) for x in c["hap_counts"]], f"{tag}: hap_counts {c['hap_counts']} != {r['hap_counts']}"
        assert r["sa_counts"][0] == int(c["sa_count"]) and r["sa_counts"][1] == int(c["sa_count"]) / float(int(c["sa_total"])), f"{tag}: sa_counts"
        assert r["n_leads"] == int(c["lead_n"]), f"{tag}: n_leads {int(c['lead_n'])} != {r['n_leads']}"
        assert r["n_long"] == int(c["long_n"]), f"{tag}: n_long {int(c['long_n'])} != {r['n_long']}"
        ll = res.cand_leads[int(c["lead_off"]):int(c["lead_off"]) + int(c["lead_n"])]
        got_leads = [[int(l["qname_hash"]), int(l["ref_start"]), int(l["svlen"]),
                      int(l["seq_len"]) if int(l["flags"]) & abi.LF_HAS_SEQ else None] for l in ll]
        ref_leads = [[qhash(q), rs, sl, sq] for q, rs, sl, sq in r["leads"]]
        assert ref_leads == got_leads, f"{tag}: cluster leads differ\n  reference {ref_leads}\n  got       {got_leads}"
        rn = sorted(int(x) for x in res.rnames[int(res.rn_off[lo + i]):int(res.rn_off[lo + i + 1])])
        assert sorted(qhash(q) for q in r["rnames"]) == rn, f"{tag}: rnames"
        if "bnd" in r:
            mc = int(c["bnd_mate_contig"])
            got = [contig_names[mc] if mc >= 0 else None, int(c["bnd_mate_pos"]), bool(c["bnd_is_first"]), bool(c["bnd_is_reverse"])]
            assert r["bnd"] == got, f"{tag}: bnd {got} != {r['bnd']}"
        if r.get("nm", -1) != -1 or True:
            assert _feq(r["nm"], float(c["nm_mean"])), f"{tag}: nm {float(c['nm_mean'])!r} != {r['nm']!r}"


def assert_alts_equal(ref_final, res, lo=0, hi=None):
    cand = res.cand[lo:hi]
    assert len(ref_final) == len(cand)
    n = 0
    for i, (r, c) in enumerate(zip(ref_final, cand)):
        if r["svtype"] != "INS":
            continue
        if int(c["alt_off"]) < 0:
            assert r["alt"] == "<INS>", f"cand[{i}] INS@{r['pos']}: reference has a sequence, result has none"
            continue
        got = res.alt[int(c["alt_off"]):int(c["alt_off"]) + int(c["alt_len"])].tobytes().decode()
        assert r["alt"] == got, f"cand[{i}] INS@{r['pos']}: ALT differs (len {len(got)} vs {len(r['alt'])})"
        n += 1
    return n
