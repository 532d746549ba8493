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


_BAM = {}      # per-process open BAM files (path -> bamio.BamFile)


def open_bam(path):
    from . import bamio
    if path not in _BAM:
        _BAM[path] = bamio.BamFile(path)
    return _BAM[path]


def gpu_count():
    """GPUs this process may use: SNFB_N_GPUS, else what the CUDA runtime reports (the worker with id w binds to w % n)"""
    import os
    if os.environ.get("SNFB_N_GPUS"):
        return max(1, int(os.environ["SNFB_N_GPUS"]))
    try:
        import torch
        return max(1, torch.cuda.device_count())
    except Exception:
        return 1


def load_tandem_repeats(filename, padding):
    """util.load_tandem_repeats (util.py:121-147): BED of tandem repeats -> {contig: [(start - padding clipped at 0, end + padding), ...]},
    in file order unless some contig's starts go backwards, in which case every contig is sorted (the reference sorts all of them then).
    The result is what `Task(tandem_repeats=...)` takes per contig (sniffles:313-358)."""
    contigs_tr, unsorted = {}, False
    with open(filename, "r") as handle:
        for line in handle:
            parts = line.split("\t")
            if len(parts) >= 3:
                contig, start, end = parts[0], int(parts[1]), int(parts[2])
                iv = contigs_tr.setdefault(contig, [])
                if iv and start < iv[-1][0]:              # compared with the previous PADDED start, as the reference does (util.py:134-136)
                    unsorted = True
                iv.append((max(0, start - padding), end + padding))
    if unsorted:
        for contig in contigs_tr:
            contigs_tr[contig].sort()
    return contigs_tr


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
    # the device pass this task reads from, and its index in that block.  Either handed in (several tasks sharing one block and one
    # device pass: run_block), or built by the task itself from its BAM region on first use, the way the reference's worker does.
    block_run: BlockRun = None
    task_index: int = 0
    coverage_average_total: float = 0.0
    device: int = 0
    device_ingest: bool = True      # a task that reads its own BAM region ships the COMPRESSED bytes: inflate + record decode on the GPU (snfb_load_bam)

    # ---- the reference's way: Task(id, sv_id, contig, start, end, config, bam=..., tandem_repeats=...) and a worker (parallel.py:47-60, 741-746)
    def _open(self):
        bam = self.bam if self.bam is not None else getattr(self.config, "input", None)
        if isinstance(bam, str):
            bam = open_bam(bam)
        if bam is None:
            raise RuntimeError("Task has neither a block_run nor a BAM to read its region from")
        return bam

    def _tables(self, bam, recs=()):
        from . import bamio
        cidx = bam.name_to_id[self.contig]
        tr = {0: [(int(a), int(b)) for a, b in self.tandem_repeats]} if self.tandem_repeats else None
        return bamio.pack_records(bam.contigs, list(recs), [(cidx, int(self.start), int(self.end), int(self.id))], tandem_repeats=tr)

    def _own_block(self):
        """records of [start, end) from the task's BAM (`self.bam`: an open bamio.BamFile or a path; default config.input), decoded and
        packed on the HOST (device_ingest=False; also what the tests compare the device ingest with)"""
        bam = self._open()
        return self._tables(bam, [(0, r) for r in bam.fetch(self.contig, self.start, self.end)])

    def _ctx(self):
        return device_context(self.device)

    def bind(self, worker=None):
        """device = worker.id % n_gpus: a context per process and device, created lazily after the fork"""
        if worker is not None and getattr(worker, "id", None) is not None:
            self.device = int(worker.id) % gpu_count()
        return self

    def build_leadtab(self):
        """parallel.py:90-102 — returns (externals, read_count).  Leads outside the region are dropped on the
        device, exactly as the caller discards `externals` (parallel.py:264)."""
        if self.block_run is None:
            ctx = self._ctx()
            ctx.set_config(abi.Config.from_sniffles(self.config))
            if self.device_ingest:
                # the reference's `bam.fetch(contig, start, end)` (parallel.py:95-98, leadprov.py:488) with htslib's work on the GPU: the host
                # only resolves the BAI index; BGZF inflate, record decode, region filter and CIGAR16 packing are snfb_load_bam
                bam = self._open()
                block = self._tables(bam)
                bgzf, spans = bam.device_input([(self.contig, int(self.start), int(self.end))])
                n_rec = ctx.load_bam(bgzf, spans, block)["n_rec"]
            else:
                block = self._own_block()
                ctx.load(block, cigar16=False)                   # BAM words: the library converts them (snfb_load_records)
                n_rec = len(block.rec)
            res = ctx.extract_leads()                            # snfb_extract_leads
            rec_nm = abi.view(res._rec_nm_ptr, "<f8", n_rec).copy() if getattr(res, "_rec_nm_ptr", None) else None
            self.block_run = BlockRun(block, res, None, rec_nm)
            self.task_index = 0
            self._staged = True
        r = self.block_run.result
        self.config.average_regional_nm = float(r.task_mean_nm[self.task_index])      # leadprov.py:577-578
        self.config.qc_nm_threshold = self.config.average_regional_nm
        return [], int(r.task_read_count[self.task_index])

    def call_candidates(self, keep_qc_fails, config):
        """parallel.py:104-127"""
        br = self.block_run
        if getattr(self, "_staged", False):
            ctx = self._ctx()
            cv = ctx.cluster_call()                              # snfb_cluster_call: candidate records (ALT offsets already planned)
            sv = ctx.consensus()                                 # snfb_consensus: annotate_sv's INS sequences, which the calls below carry
            r = br.result
            r.cand, r.cand_leads, r.rnames, r.rn_off, r.task_cov_mean, r.alt = cv.cand, cv.cand_leads, cv.rnames, cv.rn_off, cv.task_cov_mean, sv.alt
            br.cand_range = cand_ranges(r.cand, len(br.block.task))
            self._staged = False
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
        self.bind(worker)
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
