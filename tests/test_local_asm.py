"""Local assembly.  Host side against vectors the reference's own LocalAsm produced (tests/golden/local_asm/vectors.json:
select_padding, the SPOA score classes, solve_ins / solve_del accept + position + sequence decisions).  The partial-order alignment itself
replaces pyspoa, which is not in this image (parity unpinned): the CPU restatement is sanity-checked here, the CUDA kernel is checked against
it in the GPU tests."""
import json
import os
import random

import pytest

from sniffles_b200 import local_asm
from test_oracle_golden import GOLDEN, NAMES  # noqa: F401  (NAMES: keeps the fixture glob in one place)

with open(os.path.join(GOLDEN, "local_asm", "vectors.json")) as f:
    VEC = json.load(f)


def test_padding_and_score_classes_match_reference():
    for svlen, p_sv, p_half in VEC["padding"]:
        assert (local_asm.select_padding(svlen, "sv"), local_asm.select_padding(svlen, "half")) == (p_sv, p_half), svlen
    for row in VEC["scores"]:
        assert list(local_asm.spoa_scores(row[0])) == row[1:], row


def test_solve_decisions_match_reference():
    assert len(VEC["solve"]) >= 200 and 40 < sum(v["want"][2] for v in VEC["solve"]) < len(VEC["solve"]) - 40
    for v in VEC["solve"]:
        if v["svtype"] == "INS":
            got = local_asm.solve_ins(v["ref_pos"], v["svlen"], v["sv_aln"], v["ref_aln"])
        else:
            got = local_asm.solve_del(v["ref_pos"], v["svlen"], v["sv_aln"], v["ref_aln"])
        assert [got[0], got[1], bool(got[2])] == v["want"], (v["svtype"], v["svlen"])


def _mut(rnd, s, sub=0.03, indel=0.03):
    out = bytearray()
    for ch in s:
        r = rnd.random()
        if r < sub:
            out.append(rnd.choice(b"ACGT"))
        elif r < sub + indel / 2:
            continue
        elif r < sub + indel:
            out.append(ch)
            out.append(rnd.choice(b"ACGT"))
        else:
            out.append(ch)
    return bytes(out)


def _edit(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def test_poa_restatement_recovers_the_truth():
    """the CPU restatement behaves like a POA consensus: noisy copies (3 % substitutions, 3 % indels) give back the template within 2 %
    normalised edit distance, a band wider than the indel drift changes nothing, and consensus-vs-reference shows the planted event as one gap"""
    import re
    from oracle import poa
    rnd = random.Random(17)
    for trial in range(6):
        truth = bytes(rnd.choice(b"ACGT") for _ in range(400 + 150 * trial))
        reads = [_mut(rnd, truth) for _ in range(7 + 2 * trial)]
        cons = poa.consensus(reads, round(len(reads) * 0.5))
        assert _edit(truth, cons) <= 0.02 * len(truth)
        assert poa.consensus(reads, round(len(reads) * 0.5), band=96) == cons
        size = 60 * (trial + 1)
        ref = truth[:150] + truth[150 + size:]
        ra, rb = poa.pair_msa(cons, ref, local_asm.spoa_scores(size))
        gaps = [(m.start(), len(m.group())) for m in re.finditer(b"-+", rb)]
        assert len(gaps) == 1 and abs(gaps[0][1] - size) <= 2 and not re.search(b"-", ra)
        pos, seq, ok = local_asm.solve_ins(1000, size, ra.decode(), rb.decode())
        assert ok and abs(pos - (1000 + 150)) <= 2 and abs(len(seq) - size) <= 2
