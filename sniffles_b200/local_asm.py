"""Local assembly of QC-failed INS / DEL candidates (`--dev-locasm-do`) — the host side of SURVEY.md §8a row C3, mirroring
/root/reference/src/sniffles/local_asm.py: padding and window selection (`select_padding` :115-123, `read_seq_by_name_*` :126-152,
:194-216), the SPOA score classes (:26-73), `solve_ins` / `solve_del` (:154-252) and the `assembly` driver (:254-304, gate
parallel.py:186-196).  The two partial-order alignments the reference hands to pyspoa run on the device (`snfb_poa`).

Parity with pyspoa itself is unpinned (the library is not in this image and no reference test touches local_asm); the stated
tolerance is <= 2 % normalised edit distance of the consensus against the real library and identical solve_ins / solve_del
decisions on the synthetic set; the device kernel equals oracle/poa_oracle.c exactly."""
from dataclasses import dataclass

DEFAULT_SCORES = (5, -4, -8, -6, -10, -4)            # pyspoa's m, n, g, e, q, c (the read pile-up uses the library defaults, local_asm.py:287)
GAP = "-"


def spoa_scores(svlen):
    """score class by SV size (local_asm.py:26-73): (match, mismatch, gap open, gap extend, second open, second extend)"""
    a = abs(svlen)
    if a <= 400:
        return (10, -10, -20, 0, -10, -1)
    if a <= 1200:
        return (13, -5, -25, 0, -25, -1)
    return (10, -10, -30, -1, -15, -1)


def select_padding(svlen, read_type="sv", default_pad=2000):
    """local_asm.py:115-123"""
    svlen_pad = svlen if read_type == "sv" else round(svlen / 2)
    max_padding = default_pad * 3
    if abs(svlen) <= 400:
        return min(max(round(svlen_pad * 0.5), round(default_pad * 0.70)), max_padding)
    if abs(svlen) <= 1200:
        return min(max(round(svlen_pad * 0.5), round(default_pad * 1.0)), max_padding)
    return min(max(round(svlen_pad * 0.75), round(default_pad * 1.2)), max_padding)


def read_windows(svtype, pos, end, svlen, reads):
    """read_seq_by_name_del / _ins (local_asm.py:126-152, 194-216).  reads: iterable of (reference_start, query_sequence) of the supporting
    reads overlapping [pos - 5000, end + 5000).  Returns (windows, region_start, region_stop) — region None when nothing qualified."""
    wins, start, stop = [], [], []
    padding = select_padding(svlen, "sv")
    for ref_start, seq in reads:
        if seq is None:
            continue
        n = len(seq)
        if svtype == "DEL":
            p, e = pos - ref_start - padding, end - ref_start + padding + 1
            w = seq[p:e] if p >= 0 else seq[p:e]                  # python slicing semantics of the reference, negative start included
            if len(w) >= 2 * padding and (p > 0 and 0 < e < n):
                wins.append(w)
                start.append(pos - padding - 100)
                stop.append(end + padding + 100)
        else:
            p, e = max(pos - ref_start - padding, 0), pos - ref_start + svlen + padding
            w = seq[p:e]
            if len(w) >= svlen + 2 * padding and (p > 0 and 0 < e < n):
                wins.append(w)
                start.append(pos - padding)
                stop.append(pos + svlen + padding)
    if not wins:
        return [], None, None
    return wins, min(start), max(stop)


def solve_ins(ref_pos, svlen, sv_aln, ref_aln, eps=0.15, max_gaps_aln=3):
    """local_asm.py:218-252: the first gap run in the reference row whose length is within eps of svlen"""
    gap_size = ins_pos = n_gaps = ref_pos_calc = 0
    count_gap = True
    for ch in ref_aln:
        if ch == GAP:
            gap_size += 1
            if count_gap:
                n_gaps += 1
                count_gap = False
        else:
            count_gap = True
            if abs((gap_size - svlen) / svlen) <= eps and gap_size > 0 and ref_pos > 0:
                ins_pos = ref_pos + ref_pos_calc
                break
            ref_pos_calc += 1
            gap_size = 0
    ins_seq, count_pos = "", 0
    for _ in sv_aln:
        count_pos += 1
        if count_pos == ref_pos_calc:
            ins_seq = sv_aln[count_pos:count_pos + gap_size]
            break
    return ins_pos, ins_seq, len(ins_seq) > 0 and (abs((gap_size - svlen) / svlen) <= eps and gap_size > 0 and n_gaps <= max_gaps_aln)


def solve_del(ref_pos, svlen_signed, sv_aln, ref_aln, eps=0.15, max_gaps_aln=3):
    """local_asm.py:154-192: the first gap run in the consensus row whose length is within eps of |svlen|"""
    svlen = abs(svlen_signed)
    gap_size = del_pos = n_gaps = ref_pos_calc = 0
    count_gap = True
    for ch in sv_aln:
        if ch == GAP:
            gap_size += 1
            if count_gap:
                n_gaps += 1
                count_gap = False
        else:
            count_gap = True
            if abs(gap_size - svlen) / float(svlen) <= eps and gap_size > 0:
                del_pos = ref_pos + ref_pos_calc
                break
            ref_pos_calc += 1
            gap_size = 0
    ref_seq, count_pos = "", 0
    for _ in ref_aln:
        count_pos += 1
        if count_pos == ref_pos_calc:
            ref_seq = ref_aln[count_pos:count_pos + gap_size]
            break
    return del_pos, ref_seq, len(ref_seq) > 0 and (abs((gap_size - svlen) / float(svlen)) <= eps and gap_size > 0 and n_gaps <= max_gaps_aln)


def wants_local_asm(call, config):
    """the gate of Task.finalize_candidates (parallel.py:186-191)"""
    skip = ["PASS", "GT"] if not getattr(config, "dev_locasm_skip_mosaic", False) else ["PASS", "GT", "MOSAIC_VAF"]
    return (call.filter not in skip and call.svtype in ("INS", "DEL") and getattr(config, "dev_locasm_do", False) and not call.qc
            and abs(call.svlen) <= getattr(config, "dev_maxsvlen_extra", 10000)
            and (call.support >= getattr(config, "dev_minreads_extra", 5) or len(call.rnames or []) > getattr(config, "dev_minreads_extra", 5)))


@dataclass
class AsmJob:
    call: object
    windows: list
    region_start: int
    region_stop: int


def assemble(ctx, jobs, fetch_ref, min_reads=5, max_reads=30, band=None):
    """LocalAsm.assembly for a batch (local_asm.py:254-304): device consensus of every job's read windows, then the device two-row alignment of
    each consensus against its reference window, then solve_ins / solve_del and `update_sv_cand` on the host.
    jobs: AsmJob list; fetch_ref(contig, start, stop) -> str (the reference's `fas.fetch(region=f'{contig}:{start}-{stop}')`, 1-based inclusive).
    Returns the list of booleans report_sv."""
    live = [j for j in jobs if len(j.windows) >= min_reads]
    cons_jobs = []
    for j in live:
        wins = j.windows[:max_reads] if len(j.windows) > max_reads else j.windows
        W = band if band is not None else abs(j.call.svlen) + 256
        cons_jobs.append(dict(seqs=[w.encode() for w in wins], mode=0, min_cov=round(len(j.windows) * 0.50), scores=DEFAULT_SCORES, band=W))
    cons = ctx.poa(cons_jobs)
    pair_jobs, keep = [], []
    for j, c in zip(live, cons):
        if c is None:
            continue
        ref = fetch_ref(j.call.contig, j.region_start, j.region_stop)
        W = band if band is not None else abs(j.call.svlen) + 256
        pair_jobs.append(dict(seqs=[c, ref.encode()], mode=1, scores=spoa_scores(abs(j.call.svlen)), band=W))
        keep.append(j)
    msas = ctx.poa(pair_jobs)
    done = {}
    for j, m in zip(keep, msas):
        if m is None:
            continue
        sv_aln, ref_aln = m[0].decode(), m[1].decode()
        if j.call.svtype == "INS":
            sv_pos, sv_seq, ok = solve_ins(j.region_start, j.call.svlen, sv_aln, ref_aln)
        else:
            sv_pos, sv_seq, ok = solve_del(j.region_start, j.call.svlen, sv_aln, ref_aln)
        if ok:
            update_sv_cand(j.call, sv_pos, sv_seq)
        done[id(j)] = ok
    return [done.get(id(j), False) for j in jobs]


def update_sv_cand(call, sv_pos, sv_seq):
    """local_asm.py:83-97"""
    if call.filter == "MOSAIC_VAF":
        call.set_info("MOSAIC", True)
        call.filter = "GT"
    else:
        call.filter = "PASS"
    call.qc = True
    call.pos = sv_pos
    call.set_info("LASM", True)
    if call.svtype == "DEL":
        call.end = sv_pos + len(sv_seq) + 1
    elif call.svtype == "INS":
        call.end = sv_pos + 1
