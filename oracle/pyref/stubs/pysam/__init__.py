"""Minimal stand-in for `pysam` so the UNMODIFIED reference under /root/reference/src
imports in this container (pysam/htslib are not installed here).  Test infrastructure
only: used by oracle/pyref/harness.py to run the reference as the golden-vector source.
Only the names the reference touches at import time and the duck-typed classes the
hot path reads are provided (SURVEY.md §2 accessor list)."""
CMATCH, CINS, CDEL, CREF_SKIP, CSOFT_CLIP, CHARD_CLIP, CPAD, CEQUAL, CDIFF, CBACK = range(10)


class AlignedSegment:  # duck type only; concrete reads come from oracle/pyref/harness.py
    pass


class AlignmentFile:
    def __init__(self, *a, **k):
        raise RuntimeError("stub pysam: AlignmentFile cannot open files")


class FastaFile:
    def __init__(self, *a, **k):
        raise RuntimeError("stub pysam: FastaFile cannot open files")


class VariantFile:
    def __init__(self, *a, **k):
        raise RuntimeError("stub pysam: VariantFile cannot open files")


def tabix_index(*a, **k):
    raise RuntimeError("stub pysam: tabix_index unavailable")
