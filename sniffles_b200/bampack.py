"""BAM -> packed record block (include/snfb.h) with the standard library only.

A small host-side packer: BGZF inflate via zlib, BAM record decode via struct.  It stands in
for pysam/htslib (`bam.fetch` + the AlignedSegment accessors of SURVEY.md §2) for tests and
small inputs; production-scale ingest is a "next" row.  Base qualities are
dropped, SA/NM/HP/PS are the only aux tags kept."""
import gzip
import struct

import numpy as np

from . import abi
from .synth import RecordBlock

_AUX_SIZE = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
_AUX_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}


def _parse_aux(buf):
    tags, i, n = {}, 0, len(buf)
    while i + 3 <= n:
        tag, typ = buf[i:i + 2].decode(), chr(buf[i + 2])
        i += 3
        if typ in _AUX_FMT:
            sz = _AUX_SIZE[typ]
            tags[tag] = struct.unpack(_AUX_FMT[typ], buf[i:i + sz])[0]
            i += sz
        elif typ in ("A", "f"):
            i += _AUX_SIZE[typ]
        elif typ in ("Z", "H"):
            j = buf.index(b"\0", i)
            tags[tag] = buf[i:j]
            i = j + 1
        elif typ == "B":
            sub, cnt = chr(buf[i]), struct.unpack("<I", buf[i + 1:i + 5])[0]
            i += 5 + _AUX_SIZE[sub] * cnt
        else:
            raise ValueError(f"unknown aux type {typ!r}")
    return tags


def read_bam(path):
    """Yield (header_contigs [(name, length)], records [dict]) of an uncompressed-on-the-fly BAM."""
    with gzip.open(path, "rb") as f:     # BGZF is a series of gzip members
        data = f.read()
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file")
    l_text = struct.unpack("<i", data[4:8])[0]
    p = 8 + l_text
    n_ref = struct.unpack("<i", data[p:p + 4])[0]
    p += 4
    contigs = []
    for _ in range(n_ref):
        l_name = struct.unpack("<i", data[p:p + 4])[0]
        name = data[p + 4:p + 4 + l_name - 1].decode()
        length = struct.unpack("<i", data[p + 4 + l_name:p + 8 + l_name])[0]
        contigs.append((name, length))
        p += 8 + l_name
    recs = []
    while p + 4 <= len(data):
        bs = struct.unpack("<i", data[p:p + 4])[0]
        b = data[p + 4:p + 4 + bs]
        p += 4 + bs
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack("<iiBBHHHiiii", b[:32])
        o = 32
        qname = b[o:o + l_rn - 1]
        o += l_rn
        cigar = np.frombuffer(b[o:o + 4 * n_cig], "<u4").copy()
        o += 4 * n_cig
        seq = np.frombuffer(b[o:o + (l_seq + 1) // 2], "u1").copy()
        o += (l_seq + 1) // 2 + l_seq
        recs.append(dict(ref_id=ref_id, pos=pos, mapq=mapq, flag=flag, l_seq=l_seq, qname=qname, cigar=cigar, seq=seq, aux=_parse_aux(b[o:])))
    return contigs, recs


def pack(contigs, recs, with_seq=True, only_contigs=None) -> RecordBlock:
    """One task per contig that has records ([0, len-1], as the reference plans them: sniffles:313-358)."""
    used = sorted({r["ref_id"] for r in recs if r["ref_id"] >= 0 and (only_contigs is None or contigs[r["ref_id"]][0] in only_contigs)})
    task_of = {c: i for i, c in enumerate(used)}
    recs = sorted([r for r in recs if r["ref_id"] in task_of], key=lambda r: (task_of[r["ref_id"]], r["pos"]))
    n = len(recs)
    rec = np.zeros(n, abi.REC_DTYPE)
    cig, var, seq = [], [], []
    co = vo = so = 0
    for i, r in enumerate(recs):
        a = r["aux"]
        sa = a.get("SA", b"")
        flags = (abi.AUX_NM if "NM" in a else 0) | (abi.AUX_HP if "HP" in a else 0) | (abi.AUX_PS if "PS" in a else 0) | (abi.AUX_SA if "SA" in a else 0)
        rec[i] = (task_of[r["ref_id"]], r["pos"], r["flag"], r["mapq"], flags, int(a.get("HP", 0)), len(r["qname"]), 0,
                  int(a.get("NM", 0)), int(a.get("PS", 0)), len(r["cigar"]), r["l_seq"], len(sa), 0, co, so, vo)
        cig.append(r["cigar"])
        var.append(np.frombuffer(r["qname"] + sa, "u1"))
        s = r["seq"] if with_seq else np.zeros((r["l_seq"] + 1) // 2, "u1")
        seq.append(s)
        co += len(r["cigar"])
        vo += len(r["qname"]) + len(sa)
        so += len(s)
    names = [c[0] for c in contigs]
    order = sorted(range(len(names)), key=lambda k: names[k].encode())
    rank = {k: i for i, k in enumerate(order)}
    ctg = np.zeros(len(contigs), abi.CONTIG_DTYPE)
    for k, (nm, ln) in enumerate(contigs):
        ctg[k] = (abi.fnv1a64(nm.encode()), ln, rank[k])
    task = np.zeros(len(used), abi.TASK_DTYPE)
    for c, t in task_of.items():
        task[t] = (c, 0, contigs[c][1] - 1, contigs[c][1], t, 0, 0, 0)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return RecordBlock(rec=rec, cigar=np.concatenate([cat(cig, "<u4"), np.zeros(4, "<u4")])[:co] if co else np.zeros(0, "<u4"),
                       var=cat(var, "u1"), seq=cat(seq, "u1"), task=task, contig=ctg, tr=np.zeros(0, "<i4"),
                       contig_names=names, aligned_bp=0)
