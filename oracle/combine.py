"""CPU restatement of the multi-sample grouping (TEST INFRASTRUCTURE ONLY — never imported by sniffles_b200/).

Follows cluster.resolve_block_groups (/root/reference/src/sniffles/cluster.py:356-390), SVGroup.from_candidate / add_candidate
(sv.py:265-321) and the chunk loop of CombineTask.execute (parallel.py:518-563) over the flat plan of sniffles_b200.combine.Plan, and
returns the same four arrays snfb_combine_groups returns.  Pinned against the reference itself: tests/golden/combine/ was produced by
oracle/pyref/harness.reference_combine (the unmodified CombineTask) and tests/test_combine.py runs this restatement + the host epilogue
against it.  Plain Python floats = IEEE doubles, operations in the reference's order."""
import math

import numpy as np


def levenshtein(a, b):
    """global edit distance (edlib.align(a, b)['editDistance'] with the defaults), textbook dynamic programme on numpy rows"""
    if len(a) > len(b):
        a, b = b, a
    if len(a) == 0:
        return len(b)
    A = np.frombuffer(bytes(a), np.uint8)
    prev = np.arange(len(a) + 1, dtype=np.int64)
    idx = np.arange(len(a) + 1, dtype=np.int64)
    for j, cb in enumerate(bytes(b), 1):
        sub = prev[:-1] + (A != cb)
        cur = np.empty_like(prev)
        cur[0] = j
        np.minimum(sub, prev[1:] + 1, out=cur[1:])
        # the insertion recurrence cur[i] = min(cur[i], cur[i-1] + 1) is a running minimum of (cur[i] - i)
        cur = np.minimum.accumulate(cur - idx) + idx
        prev = cur
    return int(prev[-1])


def combine_groups(arrays, config):
    a = arrays
    n, S = len(a["pos"]), a["n_samples"]
    cand_group = np.zeros(max(n, 1), "<u4")
    emit_chunk = np.full(max(n, 1), -1, "<i4")
    emit_ord = np.zeros(max(n, 1), "<u4")
    cov_non = np.full((max(n, 1), S), -1, "<i4")
    n_chunk = len(a["chunks"])
    step, per_block = a["cov_binsize"], a["bins_per_block"]
    pctseq = float(getattr(config, "combine_pctseq", 0.0) or 0.0)
    alt_of = lambda i: a["alt"][int(a["alt_off"][i]):int(a["alt_off"][i]) + int(a["alt_len"][i])].tobytes()
    for c0, nc, k0, nk, is_bnd, _ in a["chains"].tolist():
        act, n_groups = [], 0                      # act: list of group dicts
        for k in range(k0, k0 + nk):
            cc0, cn, curr_bin, size, cov_block, _ = a["chunks"][k].tolist()
            for c in range(cc0, cc0 + cn):
                pos, svlen, smp = int(a["pos"][c]), int(a["svlen"][c]), int(a["sample"][c])
                best, best_dist = None, math.inf
                for g in act:
                    if is_bnd:
                        dist = abs(g["pos"] - pos) + abs(g["mate"] - int(a["mate_pos"][c]))
                        ok = dist <= config.cluster_merge_bnd * 2 and g["mc"] == int(a["mate_contig"][c])
                    else:
                        dist = abs(g["pos"] - pos) + abs(abs(g["len"]) - abs(svlen))
                        minlen = float(min(abs(g["len"]), abs(svlen)))
                        ok = minlen > 0 and dist <= config.combine_match * math.sqrt(minlen) and dist <= config.combine_match_max
                    if dist < best_dist and ok and (not config.combine_separate_intra or smp not in g["incl"]):
                        if not is_bnd and pctseq:                    # SVGroup.align_call (sv.py:282-292)
                            d = levenshtein(alt_of(g["first"]), alt_of(c))
                            if not ((g["len"] - d) / g["len"]) > pctseq:
                                continue
                        best, best_dist = g, dist
                if best is None:
                    g = dict(first=c, slot=c0 + n_groups, pos=float(pos), len=float(abs(svlen)), mate=int(a["mate_pos"][c]) if is_bnd else 0,
                             mc=int(a["mate_contig"][c]) if is_bnd else 0, n=1, incl={smp})
                    n_groups += 1
                    act.append(g)
                    emit_chunk[g["slot"]] = n_chunk
                else:
                    g = best
                    m = g["n"]
                    g["pos"] = (g["pos"] * m + pos) / (m + 1)
                    g["len"] = (g["len"] * m + abs(svlen)) / (m + 1)
                    if is_bnd:
                        g["mate"] = (g["mate"] * m + int(a["mate_pos"][c])) / (m + 1)
                    g["n"] = m + 1
                    g["incl"].add(smp)
                cand_group[c] = g["slot"]
            keep, n_call = [], 0
            for g in act:
                cb = int(g["pos"] / step) * step
                kbin = -1
                if cov_block >= 0:
                    off = cb - int(a["block_start"][cov_block])
                    if 0 <= off < per_block * step:
                        kbin = off // step
                for s in range(S):
                    if s in g["incl"]:
                        continue
                    cv = 0
                    if kbin >= 0 and a["cov"][cov_block, s, kbin] >= 0:
                        cv = int(a["cov"][cov_block, s, kbin])
                    cov_non[g["slot"], s] = max(cov_non[g["slot"], s], cv)
                if abs(g["pos"] - curr_bin) < max(size * 0.5, config.combine_overlap_abs):
                    keep.append(g)
                else:
                    emit_chunk[g["slot"]] = k
                    emit_ord[g["slot"]] = n_call
                    n_call += 1
            act = keep
        for i, g in enumerate(act):
            emit_chunk[g["slot"]] = n_chunk
            emit_ord[g["slot"]] = i
    return cand_group, emit_chunk, emit_ord, cov_non
