"""Host-side mirror of the reference's Task interface for the hot path (parallel.py:42-297):
same method names and meaning — build_leadtab / call_candidates / finalize_candidates /
execute — with the three hot calls forwarded to libsnfb200 through ctypes.

A `Task` works on one contig (one snfb_task); several tasks can share one record block and one
device run (`run_block`), which is how the contigs of a genome are processed per GPU."""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi, binding, postprocess

_CTX = {}      # per-process device contexts, created lazily after fork; never pickled


def device_context(device: int = 0) -> binding.Context:
    if device not in _CTX:
        _CTX[device] = binding.Context(device)
    return _CTX[device]


@dataclass
class BlockRun:
    """Result of the device pass over one record block, shared by the tasks of that block."""
    block: object
    result: object
    cand_range: list          # per task (lo, hi) into result.cand
    rec_nm: Optional[np.ndarray] = None


def run_block(block, config, device: int = 0, ctx=None) -> BlockRun:
    """leadprov -> cluster -> consensus for every task of the block in one device pass."""
    ctx = ctx or device_context(device)
    ctx.set_config(abi.Config.from_sniffles(config))
    ctx.load(block)
    res = ctx.run(want_leads=True, want_cands=True, want_seqs=True)
    rec_nm = abi.view(res._rec_nm_ptr, "<f8", len(block.rec)).copy() if getattr(res, "_rec_nm_ptr", None) else None
    return BlockRun(block, res, cand_ranges(res.cand, len(block.task)), rec_nm)


def cand_ranges(cand, n_task):
    t = cand["task"]
    lo = np.searchsorted(t, np.arange(n_task), side="left")
    hi = np.searchsorted(t, np.arange(n_task), side="right")
    return list(zip(lo.tolist(), hi.tolist()))


@dataclass
class Task:
    id: int
    sv_id: int
    contig: str
    start: int
    end: int
    config: object
    assigned_process_id: Optional[int] = None
    lead_provider: object = None
    bam: object = None
    tandem_repeats: list = None
    genotype_svs: list = None
    regions: list = None
    result: object = None
    # the device pass this task reads from, and its index in that block
    block_run: BlockRun = None
    task_index: int = 0
    coverage_average_total: float = 0.0

    def build_leadtab(self):
        """parallel.py:90-102 — returns (externals, read_count).  Leads outside the region are dropped on the
        device, exactly as the caller discards `externals` (parallel.py:264)."""
        r = self.block_run.result
        self.config.average_regional_nm = float(r.task_mean_nm[self.task_index])      # leadprov.py:577-578
        self.config.qc_nm_threshold = self.config.average_regional_nm
        return [], int(r.task_read_count[self.task_index])

    def call_candidates(self, keep_qc_fails, config):
        """parallel.py:104-127"""
        br = self.block_run
        lo, hi = br.cand_range[self.task_index]
        need_leads = bool(config.mosaic) or bool(config.phase)
        calls = postprocess.calls_from_result(br.result, self.task_index, lo, hi, br.block.contig_names, self.contig, self.id, config,
                                              rec_nm=br.rec_nm, want_leads=need_leads)
        self.sv_id += len(calls)
        self.coverage_average_total = float(br.result.task_cov_mean[self.task_index])
        return calls

    def finalize_candidates(self, candidates, keep_qc_fails, config):
        """parallel.py:129-201"""
        return postprocess.finalize_candidates(candidates, keep_qc_fails, config, self.coverage_average_total)


class CallTask(Task):
    def execute(self, worker=None):
        """parallel.py:256-297 (VCF path; SNF parts are a "next" row)"""
        config = self.config
        qc = not (config.snf is not None or config.no_qc)
        _, read_count = self.build_leadtab()
        cands = self.call_candidates(qc, config)
        calls = self.finalize_candidates(cands, not qc, config)
        if not config.no_qc:
            calls = [c for c in calls if c.qc]
        if config.sort:
            calls = sorted(calls, key=lambda c: c.pos)
        self.result = calls
        return calls, read_count
