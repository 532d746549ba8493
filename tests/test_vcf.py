"""VCF records against the reference's own writer: the golden fixtures hold the lines
/root/reference/src/sniffles/vcf.py emits for the finalized calls, without a reference FASTA ("vcf") and with a deterministic
one that exercises the DEL REF fetch, the INS / BND anchor base and the IUPAC clean-up ("vcf_ref")."""
import copy
import io
import os
import sys

import pytest

from sniffles_b200 import abi, postprocess, tasks, vcf
from sniffles_b200 import config as sconfig
import oracle.oracle as orc
from test_oracle_golden import NAMES, load_fixture

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pyref"))


def final_calls(fx, blk, res, cfg, t):
    ranges = tasks.cand_ranges(res.cand, len(blk.task))
    lo, hi = ranges[t]
    cfg.average_regional_nm = cfg.qc_nm_threshold = float(fx["tasks"][t]["mean_nm"])
    calls = postprocess.calls_from_result(res, t, lo, hi, blk.contig_names, blk.contig_names[int(blk.task[t]["contig"])], int(blk.task[t]["task_id"]), cfg, rec_nm=res.rec_nm, want_leads=True)
    return postprocess.finalize_candidates(calls, False, cfg, float(res.task_cov_mean[t]))


@pytest.mark.parametrize("name", NAMES)
def test_vcf_lines_match_reference_writer(name):
    from harness import FakeFasta          # test helper only (no reference import happens here)
    fx, blk = load_fixture(name)
    if "vcf" not in fx["tasks"][0]:
        pytest.skip("fixture predates the VCF lines")
    cfg = sconfig.default_config(*fx["args"])
    res = orc.run(blk, abi.Config.from_sniffles(cfg), 3, 2, keep_rec_nm=True)
    for t, ref in enumerate(fx["tasks"]):
        calls = final_calls(fx, blk, res, cfg, t)
        for key, fasta in (("vcf", None), ("vcf_ref", FakeFasta())):
            buf = io.StringIO()
            w = vcf.VCFWriter(cfg, buf, reference=fasta)
            for c in calls:
                w.write_call(copy.deepcopy(c))
            got = buf.getvalue().splitlines()
            assert len(got) == len(ref[key]), (key, t)
            for a, b in zip(got, ref[key]):
                assert a == b, f"task {t} {key}:\n  ours      {a[:400]}\n  reference {b[:400]}"


def test_header_has_the_columns():
    cfg = sconfig.default_config()
    buf = io.StringIO()
    vcf.VCFWriter(cfg, buf).write_header([("ctg1", 1000)])
    lines = buf.getvalue().splitlines()
    assert lines[0] == "##fileformat=VCFv4.2" and lines[-1].startswith("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t")
    assert any(l.startswith("##contig=<ID=ctg1,length=1000>") for l in lines)
    assert sum(l.startswith("##FILTER=") for l in lines) == len(vcf.FILTERS)
