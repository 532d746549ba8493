"""ctypes front end of oracle/poa_oracle.c (the CPU restatement of the POA behind the reference's LocalAsm).  TEST INFRASTRUCTURE:
only tests/ may import this module.  Sequences are byte strings over any alphabet (compared for equality only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DEFAULT = (5, -4, -8, -6, -10, -4)          # pyspoa's m, n, g, e, q, c


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libpoa_oracle.so")
        src = os.path.join(_HERE, "poa_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libpoa_oracle.so"], env={**os.environ, "CC": ""})
        L = C.CDLL(so)
        L.po_consensus.restype = C.c_int
        L.po_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_int] * 7 + [C.c_void_p, C.c_int]
        L.po_pair_msa.restype = C.c_int
        L.po_pair_msa.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def consensus(seqs, min_cov, scores=DEFAULT, band=1 << 28):
    flat = np.frombuffer(b"".join(seqs), "u1").copy() if seqs else np.zeros(0, "u1")
    offs = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int32)
    out = np.zeros(int(offs[-1]) + 16, "u1")
    n = lib().po_consensus(flat.ctypes.data, offs.ctypes.data, len(seqs), int(min_cov), *[int(x) for x in scores], int(min(band, 1 << 28)), out.ctypes.data, len(out))
    if n < 0:
        raise RuntimeError("po_consensus: graph overflow")
    return out[:n].tobytes()


def pair_msa(a, b, scores, band=1 << 28):
    A, B = np.frombuffer(a, "u1").copy(), np.frombuffer(b, "u1").copy()
    cap = len(a) + len(b) + 16
    ra, rb = np.zeros(cap, "u1"), np.zeros(cap, "u1")
    n = lib().po_pair_msa(A.ctypes.data, len(a), B.ctypes.data, len(b), *[int(x) for x in scores], int(min(band, 1 << 28)), ra.ctypes.data, rb.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("po_pair_msa: overflow")
    dec = lambda r: bytes(45 if x == 255 else x for x in r[:n])
    return dec(ra), dec(rb)
