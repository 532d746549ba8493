"""Field-by-field comparison of the device library's result with the CPU oracle's
(both use the struct layouts of include/snfb.h)."""
import numpy as np

LEAD_FIELDS = ["rec", "ref_start", "ref_end", "qry_start", "qry_end", "svlen", "seq_off", "seq_len", "read_len",
               "mate_pos", "mate_contig", "nm_sa", "flags", "task", "k", "qname_hash"]
CAND_INT_FIELDS = ["task", "svtype", "pos", "end", "svlen", "support", "qual", "precise", "fwd", "rev", "support_long",
                   "support_sa", "cov_upstream", "cov_start", "cov_center", "cov_end", "cov_downstream", "sa_count",
                   "sa_total", "bnd_mate_contig", "bnd_mate_pos", "bnd_is_first", "bnd_is_reverse", "n_strands",
                   "support_inline", "lead_off", "lead_n", "long_off", "long_n", "alt_off", "alt_len", "hp_top",
                   "hp_support", "hp_other", "ps_top", "ps_top_null", "ps_support", "ps_other", "cluster_seed", "resplit_bin"]


def _first_diff(a, b):
    idx = np.nonzero(a != b)[0]
    return int(idx[0]) if len(idx) else -1


def assert_leads_same(want, got, what="leads"):
    assert len(want) == len(got), f"{what}: {len(got)} on device, {len(want)} in the oracle"
    for f in LEAD_FIELDS:
        i = _first_diff(want[f], got[f])
        assert i < 0, f"{what}[{i}].{f}: device {got[f][i]} != oracle {want[f][i]}\n  device {got[i]}\n  oracle {want[i]}"


def _same_f8(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def assert_same(want, got, check_leads=True, check_alt=True, nm_rtol=1e-9):
    """want: oracle.OracleResult, got: binding.Result"""
    if check_leads and got.leads is not None:
        assert_leads_same(want.leads, got.leads)
        assert (want.task_read_count == got.task_read_count).all(), (want.task_read_count, got.task_read_count)
        # the regional NM mean is a float sum: deterministic tree on the device, sequential in the reference
        np.testing.assert_allclose(got.task_mean_nm, want.task_mean_nm, rtol=nm_rtol, atol=0)
    assert len(want.cand) == len(got.cand), f"candidates: {len(got.cand)} on device, {len(want.cand)} in the oracle"
    for f in CAND_INT_FIELDS:
        if f in ("alt_off", "alt_len") and not check_alt:
            continue
        i = _first_diff(want.cand[f], got.cand[f])
        assert i < 0, f"cand[{i}].{f}: device {got.cand[f][i]} != oracle {want.cand[f][i]}\n  device {got.cand[i]}\n  oracle {want.cand[i]}"
    i = _first_diff(want.cand["hap_counts"].reshape(-1), got.cand["hap_counts"].reshape(-1))
    assert i < 0, f"cand[{i // 6}].hap_counts: device {got.cand['hap_counts'][i // 6]} != oracle {want.cand['hap_counts'][i // 6]}"
    for f in ("stdev_pos", "stdev_len", "nm_mean"):
        ok = _same_f8(want.cand[f], got.cand[f])
        assert ok.all(), f"cand[{int(np.nonzero(~ok)[0][0])}].{f}: device {got.cand[f][~ok][0]!r} != oracle {want.cand[f][~ok][0]!r}"
    assert_leads_same(want.cand_leads, got.cand_leads, "cand_leads")
    assert (want.rn_off == got.rn_off).all(), "rnames offsets differ"
    assert (want.rnames == got.rnames).all(), "rnames differ"
    if got.task_cov_mean is not None:
        assert (want.task_cov_mean == got.task_cov_mean).all(), (want.task_cov_mean, got.task_cov_mean)
    if check_alt:
        assert len(want.alt) == len(got.alt), f"ALT arena: {len(got.alt)} bytes on device, {len(want.alt)} in the oracle"
        i = _first_diff(want.alt, got.alt)
        assert i < 0, f"ALT byte {i}: device {chr(got.alt[i])} != oracle {chr(want.alt[i])}"
