"""Host epilogue (QC, genotypes, phasing) against the reference's finalized calls
in the golden fixtures.  The candidates come from the CPU oracle here (no GPU needed); the GPU
variant in test_gpu_parity.py feeds the same epilogue from the device result."""
import numpy as np
import pytest

from sniffles_b200 import abi, postprocess, tasks
from sniffles_b200 import config as sconfig
import oracle.oracle as orc
from test_oracle_golden import NAMES, load_fixture


def check_final(fx, blk, res, rec_nm, cfg):
    ranges = tasks.cand_ranges(res.cand, len(blk.task))
    for t, ref in enumerate(fx["tasks"]):
        lo, hi = ranges[t]
        cfg.average_regional_nm = cfg.qc_nm_threshold = float(ref["mean_nm"])
        calls = postprocess.calls_from_result(res, t, lo, hi, blk.contig_names, blk.contig_names[int(blk.task[t]["contig"])], int(blk.task[t]["task_id"]),
                                              cfg, rec_nm=rec_nm, want_leads=True)
        postprocess.finalize_candidates(calls, False, cfg, float(res.task_cov_mean[t]))
        assert len(calls) == len(ref["final"])
        for c, r in zip(calls, ref["final"]):
            tag = f"{r['id']} {r['svtype']}@{r['pos']}"
            assert (c.svtype, c.pos, c.svlen, c.support, c.id) == (r["svtype"], r["pos"], r["svlen"], r["support"], r["id"]), tag
            assert c.filter == r["filter"], f"{tag}: FILTER {c.filter} != reference {r['filter']}"
            assert bool(c.qc) == r["qc"], f"{tag}: qc"
            assert c.alt == r["alt"], f"{tag}: ALT"
            gt = c.genotypes.get(0)
            got_gt = None if gt is None else [gt[0], gt[1], gt[2], gt[3], gt[4], list(gt[5]) if gt[5] else None]
            assert got_gt == r["gt"], f"{tag}: GT {got_gt} != reference {r['gt']}"
            assert c.info.get("VAF") == r["vaf"], f"{tag}: VAF"
            assert c.info.get("PHASE") == r["phase"], f"{tag}: PHASE {c.info.get('PHASE')} != {r['phase']}"


@pytest.mark.parametrize("name", NAMES)
def test_epilogue_matches_reference(name):
    fx, blk = load_fixture(name)
    cfg = sconfig.default_config(*fx["args"])
    res = orc.run(blk, abi.Config.from_sniffles(cfg), 3, 2, keep_rec_nm=True)
    check_final(fx, blk, res, res.rec_nm, cfg)


def test_load_tandem_repeats_matches_reference(tmp_path):
    """util.load_tandem_repeats (util.py:121-147): padding, file order kept, everything sorted once one contig's starts go backwards
    (compared with the PADDED previous start).  Expected values written down from the unmodified reference (same inputs, build container)."""
    from sniffles_b200 import tasks
    bed1 = tmp_path / "sorted.bed"
    bed1.write_text("chr1\t100\t200\tx\nchr1\t150\t400\nchr2\t10\t20\nshort\tline\nchr1\t9000\t9100\textra\tcols\n")
    assert tasks.load_tandem_repeats(str(bed1), 500) == {"chr1": [(0, 700), (0, 900), (8500, 9600)], "chr2": [(0, 520)]}
    bed2 = tmp_path / "unsorted.bed"
    bed2.write_text("chr1\t5000\t5100\nchr1\t1000\t1100\nchr2\t800\t900\nchr2\t700\t950\nchr2\t10\t20\n")
    # chr2: 700 is not below the previous padded start 300, but chr1 is unsorted -> every contig is sorted
    assert tasks.load_tandem_repeats(str(bed2), 500) == {"chr1": [(500, 1600), (4500, 5600)], "chr2": [(0, 520), (200, 1450), (300, 1400)]}
    bed3 = tmp_path / "kept.bed"
    bed3.write_text("chr2\t800\t900\nchr2\t700\t950\n")                  # 700 >= 300: not flagged, file order kept
    assert tasks.load_tandem_repeats(str(bed3), 500) == {"chr2": [(300, 1400), (200, 1450)]}
