"""CIGAR16 (include/snfb.h): snfb_pack_cigar16 is host code, so the format is checked without a GPU:
every op survives the round trip, groups never straddle a 16-byte boundary, records start on one."""
import numpy as np
import pytest

from sniffles_b200 import abi, binding, synth

CLASS = [3, 1, 2, 6, 5, 4, 0, 3, 3]          # M I D N S H P = X


E = [[]]


def decode(words):
    ops = []
    E[0] = []
    for k, w in enumerate(int(x) for x in words):
        if w & 0x8000:
            assert k % 8 != 0, "an extension word starts a 16-byte group"
            assert ops, "extension word without a base word"
            ln, c = ops[-1]
            ops[-1] = (ln + ((w & 0xfff) << (11 + 12 * (((w >> 12) & 7) - 1))), c)
        elif w != 0:
            ln, c = w & 0x7ff, (w >> 11) & 7
            ops.append((ln, c))
            E[0].append((len(ops) - 1, bool(w & 0x4000)))
    return ops


def block_of(cigars):
    rec = np.zeros(len(cigars), abi.REC_DTYPE)
    flat, off = [], 0
    for i, cg in enumerate(cigars):
        rec[i]["cigar_off"], rec[i]["n_cigar"] = off, len(cg)
        flat.extend((ln << 4) | op for ln, op in cg)
        off += len(cg)
    return rec, np.asarray(flat, dtype="<u4")


def check(rec, cigar32):
    rec16, c16 = binding.pack_cigar16(rec, cigar32)
    assert len(c16) % 8 == 0
    for r, r16 in zip(rec, rec16):
        assert int(r16["cigar_off"]) % 8 == 0
        want = [(int(w) >> 4, CLASS[int(w) & 15]) for w in cigar32[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])]]
        want = [(ln, c) for ln, c in want if not (ln == 0 and c == 0)]
        got = decode(c16[int(r16["cigar_off"]):int(r16["cigar_off"]) + int(r16["n_cigar"])])
        assert got == want
        for k, e in E[0]:          # E bit: an I / D / S of at least 11 bases
            assert e == (got[k][1] in (1, 2, 5) and got[k][0] >= 11), (got[k], e)
        for f in ("task", "pos", "flag", "mapq", "l_seq", "seq_off", "var_off", "nm"):
            assert r[f] == r16[f]
    return rec16, c16


def test_round_trip_random_ops():
    rnd = np.random.default_rng(7)
    cigars = []
    for _ in range(300):
        n = int(rnd.integers(1, 60))
        lens = np.where(rnd.random(n) < 0.15, rnd.integers(2048, 1 << 23, n), rnd.integers(0, 300, n))
        lens = np.where(rnd.random(n) < 0.03, rnd.integers(1 << 23, 1 << 28, n), lens)
        cigars.append([(int(l), int(o)) for l, o in zip(lens, rnd.integers(0, 9, n))])
    check(*block_of(cigars))


def test_length_boundaries():
    cigars = [[(2047, 0), (2048, 1), (2049, 2), ((1 << 23) - 1, 3), (1 << 23, 4), ((1 << 28) - 1, 2), (0, 0), (1, 8), (10, 1), (11, 2)],
              [(5000, 4)] * 9, [((1 << 24) + 5, 2)] * 7, [(10, 0)]]
    rec16, c16 = check(*block_of(cigars))
    # [(5000, S)] x 9: two-word groups, four per 16 bytes -> 8, 8, 2 words
    assert int(rec16[1]["n_cigar"]) == 18
    # three-word groups: two per 16 bytes, two pad words each time
    assert int(rec16[2]["n_cigar"]) == 8 * 3 + 3


def test_unknown_op_is_rejected():
    rec, cg = block_of([[(10, 0), (3, 9)]])
    with pytest.raises(binding.SnfbError):
        binding.pack_cigar16(rec, cg)


def test_synthetic_block():
    blk = synth.config_block(2, 0.002)
    check(blk.rec[::53].copy(), blk.cigar)
    blk.pack16()
    assert blk.cigar16.nbytes < 0.52 * blk.cigar.nbytes
