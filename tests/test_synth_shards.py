"""The sharded generator (contig_mask) must produce exactly the records of the full generation that land on the owned
contigs — including the supplementary records of reads that start on a contig another shard owns."""
import numpy as np

from sniffles_b200 import synth
from sniffles_b200 import dist as sdist


def _records(blk, tasks=None):
    out = []
    for r in blk.rec:
        if tasks is not None and int(r["task"]) not in tasks:
            continue
        co, n = int(r["cigar_off"]), int(r["n_cigar"])
        vo, vl = int(r["var_off"]), int(r["l_qname"]) + int(r["sa_len"])
        so, sl = int(r["seq_off"]), (int(r["l_seq"]) + 1) // 2
        out.append((int(r["task"]), int(r["pos"]), int(r["flag"]), int(r["mapq"]), int(r["aux_flags"]), int(r["hp"]), int(r["nm"]), int(r["ps"]), int(r["l_seq"]),
                    blk.cigar[co:co + n].tobytes(), blk.var[vo:vo + vl].tobytes(), blk.seq[so:so + sl].tobytes()))
    return out


def test_shards_reproduce_the_full_block():
    lens = [300_000, 260_000, 220_000, 200_000, 200_000]
    kw = dict(coverage=12.0, len_mean=9000.0, len_sd=2500.0, sv_spacing=4000.0, tr_frac=0.1)
    full = synth.generate(77, lens, **kw)
    n_split = int(((full.rec["flag"] & 2048) != 0).sum())
    assert n_split > 50
    for world in (2, 3):
        owner = sdist.lpt_assign(lens, world)
        total = 0
        for rank in range(world):
            mask = [o == rank for o in owner]
            part = synth.generate(77, lens, contig_mask=mask, **kw)
            mine = {t for t, o in enumerate(owner) if o == rank}
            assert set(np.unique(part.rec["task"]).tolist()) <= mine
            assert _records(part) == _records(full, mine), f"rank {rank} of {world}"
            total += len(part.rec)
        assert total == len(full.rec)
