"""BGZF / BAM / BAI with the standard library — the host-side stand-in for htslib behind `bam.fetch(contig, start, end)`
(leadprov.py:488; the accessor contract is SURVEY.md §8a row A0).  Region fetches go through the BAI index (bins + linear
index, SAM spec §5.2), so a task only inflates the BGZF blocks its region touches.  A small writer (BAM + BAI from a packed
record block) exists for the tests and the benchmark inputs; it is not part of the product path.

Base qualities are never decoded; of the aux tags only NM, HP, PS, SA and the CG:B,I long-CIGAR escape are read."""
import struct
import zlib

import numpy as np

from . import abi
from .synth import RecordBlock

_AUX_SIZE = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
_AUX_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# ------------------------------------------------------------------------------------------------ BGZF
class BgzfReader:
    """random access by BGZF virtual offset (coffset << 16 | uoffset)"""

    def __init__(self, path):
        self.f = open(path, "rb")
        self._cache = (None, b"", 0)           # (coffset, data, block length)

    def close(self):
        self.f.close()

    def _block(self, coffset):
        if self._cache[0] == coffset:
            return self._cache[1], self._cache[2]
        self.f.seek(coffset)
        head = self.f.read(18)
        if len(head) < 18:
            return b"", 0
        if head[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block")
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = head[12:] + self.f.read(xlen - 6)
        bsize, i = None, 0
        while i + 4 <= len(extra):
            si1, si2, slen = extra[i], extra[i + 1], struct.unpack("<H", extra[i + 2:i + 4])[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack("<H", extra[i + 4:i + 6])[0] + 1
            i += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without a BC field")
        cdata = self.f.read(bsize - 12 - xlen - 8)
        self.f.read(8)
        data = zlib.decompress(cdata, -15)
        self._cache = (coffset, data, bsize)
        return data, bsize

    def read_from(self, voffset, nbytes):
        """nbytes of uncompressed data starting at a virtual offset; returns (data, virtual offset after it)"""
        coff, uoff = voffset >> 16, voffset & 0xffff
        out = bytearray()
        while len(out) < nbytes:
            data, bsize = self._block(coff)
            if bsize == 0:
                break
            take = data[uoff:uoff + nbytes - len(out)]
            out += take
            uoff += len(take)
            if uoff >= len(data):
                coff, uoff = coff + bsize, 0
        return bytes(out), (coff << 16) | uoff


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
            if tag == "CG" and sub == "I":
                tags["CG"] = np.frombuffer(buf[i + 5:i + 5 + 4 * cnt], "<u4").copy()
            i += 5 + _AUX_SIZE[sub] * cnt
        else:
            raise ValueError(f"unknown aux type {typ!r}")
    return tags


def decode_record(b):
    """one BAM alignment (without its block_size prefix) -> dict of the fields the path reads"""
    ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nr, _np, _tl = struct.unpack("<iiBBHHHiiii", b[:32])
    o = 32
    qname = b[o:o + l_rn - 1]
    o += l_rn
    cigar = np.frombuffer(b[o:o + 4 * n_cig], "<u4").copy()
    o += 4 * n_cig
    seq = np.frombuffer(b[o:o + (l_seq + 1) // 2], "u1").copy()
    o += (l_seq + 1) // 2 + l_seq
    aux = _parse_aux(b[o:])
    # records with more than 65535 CIGAR ops carry the real CIGAR in CG:B,I behind a "<l_seq>S<reflen>N" placeholder (SAM spec §4.2.2;
    # htslib and pysam restore it transparently)
    if n_cig == 2 and "CG" in aux and (int(cigar[0]) & 15) == 4 and (int(cigar[0]) >> 4) == l_seq and (int(cigar[1]) & 15) == 3:
        cigar = aux["CG"]
    elif n_cig == 2 and (int(cigar[0]) & 15) == 4 and (int(cigar[0]) >> 4) == l_seq and (int(cigar[1]) & 15) == 3 and l_seq > 0:
        raise ValueError(f"record {qname!r}: placeholder CIGAR without a CG tag")
    return dict(ref_id=ref_id, pos=pos, mapq=mapq, flag=flag, l_seq=l_seq, qname=qname, cigar=cigar, seq=seq, aux=aux)


def ref_span(cigar):
    ops = cigar & 15
    return int(np.where(np.isin(ops, (0, 2, 3, 7, 8)), cigar >> 4, 0).sum())


# ------------------------------------------------------------------------------------------------ BAI
def reg2bins(beg, end):
    end -= 1
    bins = [0]
    for shift, off in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(off + (beg >> shift), off + (end >> shift) + 1))
    return bins


def reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


class BamFile:
    """`fetch(contig, start, end)` over an indexed BAM; yields decoded records in file (coordinate) order"""

    def __init__(self, path, index_path=None):
        self.path = path
        self.bgzf = BgzfReader(path)
        head, v = self.bgzf.read_from(0, 12)
        if head[:4] != b"BAM\1":
            raise ValueError("not a BAM file")
        l_text = struct.unpack("<i", head[4:8])[0]
        _, v = self.bgzf.read_from(8 << 0, 0)
        data, v = self.bgzf.read_from(0, 12 + l_text)
        n_ref = struct.unpack("<i", data[8 + l_text:12 + l_text])[0]
        self.contigs = []
        for _ in range(n_ref):
            d, v = self.bgzf.read_from(v, 4)
            l_name = struct.unpack("<i", d)[0]
            d, v = self.bgzf.read_from(v, l_name + 4)
            self.contigs.append((d[:l_name - 1].decode(), struct.unpack("<i", d[l_name:l_name + 4])[0]))
        self.first_record = v
        self.name_to_id = {n: i for i, (n, _) in enumerate(self.contigs)}
        import os
        if index_path is None:
            index_path = next((p for p in (path + ".bai", path + ".csi", path[:-4] + ".bai" if path.endswith(".bam") else path + ".bai") if os.path.exists(p)), path + ".bai")
        self.min_shift, self.depth = 14, 5
        self.index = self._load_csi(index_path) if index_path.endswith(".csi") else self._load_bai(index_path)
        self.meta_bin = ((1 << ((self.depth + 1) * 3)) - 1) // 7 + 1          # 37450 for the BAI layout

    def close(self):
        self.bgzf.close()

    @staticmethod
    def _load_bai(path):
        with open(path, "rb") as f:
            d = f.read()
        if d[:4] != b"BAI\1":
            raise ValueError("not a BAI index")
        n_ref = struct.unpack("<i", d[4:8])[0]
        p, refs = 8, []
        for _ in range(n_ref):
            n_bin = struct.unpack("<i", d[p:p + 4])[0]
            p += 4
            bins = {}
            for _ in range(n_bin):
                b, n_chunk = struct.unpack("<Ii", d[p:p + 8])
                p += 8
                bins[b] = [struct.unpack("<QQ", d[p + 16 * k:p + 16 * k + 16]) for k in range(n_chunk)]
                p += 16 * n_chunk
            n_intv = struct.unpack("<i", d[p:p + 4])[0]
            p += 4
            lin = list(struct.unpack(f"<{n_intv}Q", d[p:p + 8 * n_intv]))
            p += 8 * n_intv
            refs.append((bins, lin, None))
        return refs

    def _load_csi(self, path):
        """CSI (SAM spec §5.3 / CSIv1): a BGZF file; bins of a configurable geometry (min_shift, depth), every bin with the smallest virtual
        offset of a record overlapping its first window (`loffset`) instead of BAI's linear index"""
        with open(path, "rb") as f:
            z = f.read()
        d, o = b"", 0
        while o < len(z):                                # concatenated gzip members
            dec = zlib.decompressobj(31)
            d += dec.decompress(z[o:])
            o = len(z) - len(dec.unused_data)
        if d[:4] != b"CSI\1":
            raise ValueError("not a CSI index")
        self.min_shift, self.depth, l_aux = struct.unpack("<iii", d[4:16])
        p = 16 + l_aux
        n_ref = struct.unpack("<i", d[p:p + 4])[0]
        p += 4
        refs = []
        for _ in range(n_ref):
            n_bin = struct.unpack("<i", d[p:p + 4])[0]
            p += 4
            bins, loff = {}, {}
            for _ in range(n_bin):
                b, lo, n_chunk = struct.unpack("<IQi", d[p:p + 16])
                p += 16
                bins[b] = [struct.unpack("<QQ", d[p + 16 * k:p + 16 * k + 16]) for k in range(n_chunk)]
                loff[b] = lo
                p += 16 * n_chunk
            refs.append((bins, None, loff))
        return refs

    # ---- index queries shared by fetch and device_input
    def _reg2bins(self, beg, end):
        end -= 1
        bins, t, s = [], 0, self.min_shift + 3 * self.depth
        for lvl in range(self.depth + 1):
            bins.extend(range(t + (beg >> s), t + (end >> s) + 1))
            t += 1 << (3 * lvl)
            s -= 3
        return bins

    def _min_offset(self, rid, start):
        """smallest virtual offset a record overlapping `start` can have: BAI's linear index, or CSI's per-bin loffset found the way
        htslib looks it up (the leaf bin of start, else the nearest earlier sibling / ancestor that exists)"""
        bins, lin, loff = self.index[rid]
        if lin is not None:
            w = start >> 14
            return lin[w] if w < len(lin) else (lin[-1] if lin else 0)
        first_leaf = ((1 << (3 * self.depth)) - 1) // 7
        b = first_leaf + (start >> self.min_shift)
        while b:
            if b in loff:
                return loff[b]
            parent = (b - 1) >> 3
            b = b - 1 if b > (parent << 3) + 1 else parent
        return loff.get(0, 0)

    def _anchors(self, rid):
        """record-aligned virtual offsets inside a contig's data: where the device ingest may cut its spans"""
        bins, lin, loff = self.index[rid]
        return sorted(set(lin if lin is not None else loff.values()))

    def get_reference_length(self, contig):
        return self.contigs[self.name_to_id[contig]][1]

    def count_mapped(self, contig):
        """mapped-read count of the pseudo-bin 37450 (what get_index_statistics reports), or None"""
        bins = self.index[self.name_to_id[contig]][0]
        ch = bins.get(self.meta_bin)
        return int(ch[1][0]) if ch and len(ch) > 1 else None

    def fetch(self, contig, start, end):
        rid = self.name_to_id[contig]
        bins = self.index[rid][0]
        min_off = self._min_offset(rid, max(start, 0))
        chunks = sorted(c for b in self._reg2bins(max(start, 0), max(end, start + 1)) if b in bins and b != self.meta_bin for c in bins[b] if c[1] > min_off)
        seen_to = 0
        for beg, stop in chunks:
            v = max(beg, seen_to, min_off)
            while v < stop:
                d, v2 = self.bgzf.read_from(v, 4)
                if len(d) < 4:
                    break
                bs = struct.unpack("<i", d)[0]
                b, v = self.bgzf.read_from(v2, bs)
                if len(b) < bs:
                    break
                ref_id, pos = struct.unpack("<ii", b[:8])
                if ref_id != rid or pos >= end:
                    if ref_id > rid or pos >= end:
                        seen_to = stop
                        break
                    continue
                r = decode_record(b)
                if pos + max(ref_span(r["cigar"]), 1) > start:
                    yield r
            seen_to = max(seen_to, v)

    # ---- device ingest (snfb_load_bam): the index work stays on the host, the bytes stay compressed
    def merged_chunks(self, contig, start, end):
        """disjoint, ascending virtual-offset ranges holding every record `fetch(contig, start, end)` would look at (the BAI chunks of
        the region's bins behind the linear-index minimum, merged the way htslib merges them)"""
        rid = self.name_to_id[contig]
        bins = self.index[rid][0]
        min_off = self._min_offset(rid, max(start, 0))
        chunks = sorted(c for b in self._reg2bins(max(start, 0), max(end, start + 1)) if b in bins and b != self.meta_bin for c in bins[b] if c[1] > min_off)
        merged = []
        for beg, stop in chunks:
            beg = max(beg, min_off)
            if merged and beg <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], stop)
            else:
                merged.append([beg, stop])
        return [(a, b) for a, b in merged if b > a]

    def _bsize_at(self, coffset):
        """total size of the BGZF block that starts at a file offset (0 at end of file)"""
        self.bgzf.f.seek(coffset)
        head = self.bgzf.f.read(18)
        if len(head) < 18:
            return 0
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = head[12:] + self.bgzf.f.read(max(xlen - 6, 0))
        i = 0
        while i + 4 <= len(extra):
            slen = struct.unpack("<H", extra[i + 2:i + 4])[0]
            if extra[i] == 66 and extra[i + 1] == 67:
                return struct.unpack("<H", extra[i + 4:i + 6])[0] + 1
            i += 4 + slen
        raise ValueError("BGZF block without a BC field")

    def device_input(self, regions, split=True):
        """regions: [(contig, start, end)] = the tasks, in task order.  Returns (bgzf, spans): the compressed bytes of every BGZF block the
        regions need (file order, each block once) and abi.SPAN_DTYPE rows — the merged index chunks of each task, cut at the linear
        index's record-aligned offsets so that every ~16 kb window is its own parallel walk on the device."""
        pieces = []                                      # (task, vbeg, vend)
        for t, (contig, start, end) in enumerate(regions):
            anchors = self._anchors(self.name_to_id[contig]) if split else []
            for vb, ve in self.merged_chunks(contig, start, end):
                cuts = [vb] + [a for a in anchors if vb < a < ve] + [ve]
                pieces.extend((t, cuts[k], cuts[k + 1]) for k in range(len(cuts) - 1))
        # file intervals [cb, ce) that hold the blocks of the pieces; a piece that ends inside a block needs that block too
        iv, bs_cache = [], {}
        for _, vb, ve in pieces:
            cb, ce = vb >> 16, ve >> 16
            if ve & 0xffff:
                if ce not in bs_cache:
                    bs_cache[ce] = self._bsize_at(ce)
                ce += bs_cache[ce]
            iv.append((cb, ce))
        iv.sort()
        merged = []
        for a, b in iv:
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        starts, base, parts = [], [], []
        off = 0
        for a, b in merged:
            self.bgzf.f.seek(a)
            d = self.bgzf.f.read(b - a)
            if len(d) != b - a:
                raise ValueError("truncated BAM file")
            starts.append(a)
            base.append(off)
            parts.append(d)
            off += len(d)
        bgzf = np.frombuffer(b"".join(parts), "u1") if parts else np.zeros(0, "u1")
        import bisect

        def to_buf(c):                                   # file offset of a block start (or of an interval's end) -> offset in bgzf
            k = bisect.bisect_right(starts, c) - 1
            if k < 0 or c > merged[k][1]:
                raise ValueError("virtual offset outside the loaded intervals")
            return base[k] + (c - starts[k])
        spans = np.zeros(len(pieces), abi.SPAN_DTYPE)
        for i, (t, vb, ve) in enumerate(pieces):
            spans[i] = (to_buf(vb >> 16), to_buf(ve >> 16), vb & 0xffff, ve & 0xffff, t, 0)
        return bgzf, spans


def pack_records(contigs, recs, tasks, with_seq=True, tandem_repeats=None) -> RecordBlock:
    """records (already grouped by task, coordinate sorted inside a task) -> packed block of include/snfb.h.
    tasks: list of (contig index, start, end, task_id); recs: list of (task index, record dict)."""
    n = len(recs)
    rec = np.zeros(n, abi.REC_DTYPE)
    cig, var, seq = [], [], []
    co = vo = so = 0
    for i, (t, r) in enumerate(recs):
        a = r["aux"]
        sa = a.get("SA", b"")
        flags = (abi.AUX_NM if "NM" in a else 0) | (abi.AUX_HP if "HP" in a else 0) | (abi.AUX_PS if "PS" in a else 0) | (abi.AUX_SA if "SA" in a else 0)
        if len(r["qname"]) > 255:
            raise ValueError("query name longer than 255 bytes")
        rec[i] = (t, r["pos"], r["flag"], r["mapq"], flags, int(a.get("HP", 0)) & 255, len(r["qname"]), 0,
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
    task = np.zeros(len(tasks), abi.TASK_DTYPE)
    tr_flat, tr_off = [], 0
    for t, (c, start, end, tid) in enumerate(tasks):
        iv = sorted((tandem_repeats or {}).get(t, []))
        task[t] = (c, start, end, contigs[c][1], tid, tr_off, len(iv), 0)
        tr_flat.extend(x for ab in iv for x in ab)
        tr_off += len(iv)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return RecordBlock(rec=rec, cigar=cat(cig, "<u4"), var=cat(var, "u1"), seq=cat(seq, "u1"), task=task, contig=ctg,
                       tr=np.asarray(tr_flat, dtype="<i4"), contig_names=names, aligned_bp=0)


# ------------------------------------------------------------------------------------------------ writer (tests / benchmark inputs)
def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp
            + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, blk: RecordBlock, block_bytes=0xff00, level=6, qual_seed=None, index="bai"):
    """packed block -> coordinate-sorted BAM + BAI (records keep the block's order; one contig per task).  Returns the paths.
    qual_seed: None writes the "qualities absent" bytes 0xff; an integer writes noisy phred values (what makes a real BAM hard to compress).
    index: "bai" or "csi" (the BAI bin geometry, min_shift 14 / depth 5, with per-bin loffsets derived from the linear index as htslib derives them)."""
    qrng = np.random.default_rng(qual_seed) if qual_seed is not None else None
    names = blk.contig_names
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(f"@SQ\tSN:{n}\tLN:{int(c['length'])}\n".encode() for n, c in zip(names, blk.contig))
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(names))
    for n, c in zip(names, blk.contig):
        head += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", int(c["length"]))
    out = open(path, "wb")
    buf, coff = bytearray(), 0
    index_tabs = [({}, []) for _ in names]
    stats = [[None, None, 0] for _ in names]

    def flush():
        nonlocal buf, coff
        if buf:
            blk_b = _bgzf_block(bytes(buf), level)
            out.write(blk_b)
            coff += len(blk_b)
            buf = bytearray()

    def put(data):
        nonlocal buf
        if len(buf) + len(data) > block_bytes:
            flush()
        v0 = (coff << 16) | len(buf)
        while len(data) > block_bytes:                 # a record larger than one block spans several
            buf += data[:block_bytes - len(buf)]
            data = data[block_bytes - len(buf):] if False else data[len(buf):]
            flush()
        buf += data
        return v0
    put(head)
    flush()
    for r in blk.rec:
        rid = int(blk.task[int(r["task"])]["contig"])
        co, nco = int(r["cigar_off"]), int(r["n_cigar"])
        cig = np.ascontiguousarray(blk.cigar[co:co + nco], dtype="<u4")
        vo, lq, sl = int(r["var_off"]), int(r["l_qname"]), int(r["sa_len"])
        qname = bytes(blk.var[vo:vo + lq]) + b"\0"
        sa = bytes(blk.var[vo + lq:vo + lq + sl])
        l_seq = int(r["l_seq"])
        seq = bytes(blk.seq[int(r["seq_off"]):int(r["seq_off"]) + (l_seq + 1) // 2])
        aux = b""
        af = int(r["aux_flags"])
        if af & abi.AUX_NM:
            aux += b"NMi" + struct.pack("<i", int(r["nm"]))
        if af & abi.AUX_HP:
            aux += b"HPC" + struct.pack("<B", int(r["hp"]))
        if af & abi.AUX_PS:
            aux += b"PSi" + struct.pack("<i", int(r["ps"]))
        if af & abi.AUX_SA:
            aux += b"SAZ" + sa + b"\0"
        pos = int(r["pos"])
        end = pos + max(ref_span(cig), 1)
        n_cig, cig_b = nco, cig.tobytes()
        if nco > 65535:                                # long CIGAR escape
            aux += b"CGBI" + struct.pack("<I", nco) + cig_b
            cig_b = struct.pack("<II", (l_seq << 4) | 4, ((end - pos) << 4) | 3)
            n_cig = 2
        body = struct.pack("<iiBBHHHiiii", rid, pos, len(qname), int(r["mapq"]), reg2bin(pos, end), n_cig, int(r["flag"]), l_seq, -1, -1, 0) \
            + qname + cig_b + seq + (b"\xff" * l_seq if qrng is None else np.clip(qrng.normal(22.0, 9.0, l_seq), 2, 50).astype("u1").tobytes()) + aux
        data = struct.pack("<i", len(body)) + body
        if len(buf) + len(data) > block_bytes:
            flush()
        v0 = (coff << 16) | len(buf)
        if len(data) > block_bytes:
            for k in range(0, len(data), block_bytes):
                buf += data[k:k + block_bytes]
                if len(buf) >= block_bytes:
                    flush()
        else:
            buf += data
        v1 = (coff << 16) | len(buf)
        bins, lin = index_tabs[rid]
        ch = bins.setdefault(reg2bin(pos, end), [])
        if ch and ch[-1][1] == v0:
            ch[-1] = (ch[-1][0], v1)
        else:
            ch.append((v0, v1))
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            while len(lin) <= w:
                lin.append(0)
            if lin[w] == 0:
                lin[w] = v0
        st = stats[rid]
        st[0] = v0 if st[0] is None else st[0]
        st[1] = v1
        st[2] += 1
    flush()
    out.write(_BGZF_EOF)
    out.close()
    for _, lin in index_tabs:
        for w in range(1, len(lin)):                   # empty windows inherit the previous offset
            if lin[w] == 0:
                lin[w] = lin[w - 1]
    if index == "csi":
        body = b"CSI\1" + struct.pack("<iii", 14, 5, 0) + struct.pack("<i", len(names))
        level_first = [((1 << (3 * l)) - 1) // 7 for l in range(6)]
        for (bins, lin), st in zip(index_tabs, stats):
            body += struct.pack("<i", len(bins) + (1 if st[2] else 0))
            for b, ch in bins.items():
                lvl = max(l for l in range(6) if level_first[l] <= b)
                w = (b - level_first[lvl]) << (3 * (5 - lvl))
                body += struct.pack("<IQi", b, lin[w] if w < len(lin) else 0, len(ch)) + b"".join(struct.pack("<QQ", v0, v1) for v0, v1 in ch)
            if st[2]:
                body += struct.pack("<IQi", 37450, 0, 2) + struct.pack("<QQ", st[0], st[1]) + struct.pack("<QQ", st[2], 0)
        with open(path + ".csi", "wb") as f:
            for k in range(0, len(body), 0xff00):
                f.write(_bgzf_block(body[k:k + 0xff00], level))
            f.write(_BGZF_EOF)
        return path, path + ".csi"
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", len(names)))
        for (bins, lin), st in zip(index_tabs, stats):
            nb = len(bins) + (1 if st[2] else 0)
            f.write(struct.pack("<i", nb))
            for b, ch in bins.items():
                f.write(struct.pack("<Ii", b, len(ch)))
                for v0, v1 in ch:
                    f.write(struct.pack("<QQ", v0, v1))
            if st[2]:
                f.write(struct.pack("<Ii", 37450, 2) + struct.pack("<QQ", st[0], st[1]) + struct.pack("<QQ", st[2], 0))
            f.write(struct.pack("<i", len(lin)) + struct.pack(f"<{len(lin)}Q", *lin))
    return path, path + ".bai"
