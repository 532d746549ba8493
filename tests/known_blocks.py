"""Hand-built record blocks for the known-answer vectors of SURVEY.md Appendix A (shared by the generator that asks the unmodified
reference, tests/golden/make_known_answers.py, and by tests/test_known_answers.py)."""
import numpy as np

from sniffles_b200 import abi, bamio

CASES = {
    # cluster.resplit's negative-index wrap (cluster.py:125-161): INS lengths of the reads of one cluster, in BAM order
    "resplit_wrap_all": [40, 45, 60, 65, 100, 105],
    "resplit_two_groups": [40, 45, 60, 65, 300, 305],
    "resplit_chain": [50, 70, 90, 110, 130],
    # Cluster.compute_metrics samples every (n // 100)-th lead but divides by n (cluster.py:48-61): many equal leads
    "metrics_150_equal": [100] * 150,
    "same_read_twice": [80, 80, 80, 80],
}
ARGS = ("--minsvlen", "35", "--minsupport", "2", "--mapq", "0")


def ins_block(svlens, pos=20_000, flank=3000, contig_len=100_000):
    """one read per length: <flank>M <len>I <flank>M at the same position, on one contig / one task"""
    rnd = np.random.default_rng(len(svlens) * 7919 + int(sum(svlens)))
    recs = []
    for k, L in enumerate(svlens):
        l_seq = 2 * flank + L
        recs.append((0, dict(pos=pos - flank + (k % 3), flag=0 if k % 2 else 16, mapq=60, l_seq=l_seq, qname=b"read%03d" % k,
                             cigar=np.array([((flank - (k % 3)) << 4) | 0, (L << 4) | 1, ((flank + (k % 3)) << 4) | 0], "<u4"),
                             seq=rnd.integers(0, 256, (l_seq + 1) // 2, dtype=np.uint8) & 0x99 | 0x11, aux={"NM": L})))
    recs.sort(key=lambda tr: tr[1]["pos"])
    return bamio.pack_records([("ctgA", contig_len)], recs, [(0, 0, contig_len, 0)])
