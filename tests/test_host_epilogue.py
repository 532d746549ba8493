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
