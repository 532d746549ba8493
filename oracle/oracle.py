"""ctypes front end of the CPU oracle (oracle/libsnf_oracle.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module; the product package sniffles_b200 never does."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from sniffles_b200 import abi  # noqa: E402  (struct mirrors only)

_LIB = None


class _View(C.Structure):
    _fields_ = [("n_leads", C.c_uint64), ("leads", C.c_void_p), ("n_task", C.c_uint32), ("_pad", C.c_uint32),
                ("task_read_count", C.c_void_p), ("task_mean_nm", C.c_void_p), ("rec_nm", C.c_void_p),
                ("task_cov_mean", C.c_void_p), ("n_pass", C.c_uint64), ("soft_errors", C.c_uint64),
                ("n_cand", C.c_uint64), ("cand", C.c_void_p), ("n_cand_leads", C.c_uint64), ("cand_leads", C.c_void_p),
                ("rnames", C.c_void_p), ("rn_off", C.c_void_p), ("n_alt", C.c_uint64), ("alt", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libsnf_oracle.so")
    src = os.path.join(_HERE, "snf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": ""})
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.so_run.restype = C.c_void_p
        L.so_run.argtypes = [C.POINTER(abi.Records), C.POINTER(abi.Config), C.c_int, C.c_int]
        L.so_get.argtypes = [C.c_void_p, C.POINTER(_View)]
        L.so_free.argtypes = [C.c_void_p]
        L.so_qname_hash.restype = C.c_uint64
        L.so_qname_hash.argtypes = [C.c_char_p, C.c_size_t]
        L.so_hash_name.restype = C.c_uint64
        L.so_hash_name.argtypes = [C.c_char_p, C.c_size_t]
        L.so_sqrt_frac.restype = C.c_double
        L.so_sqrt_frac.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.so_center.restype = C.c_long
        L.so_center.argtypes = [C.c_void_p, C.c_long]
        L.so_stdev.restype = C.c_double
        L.so_stdev.argtypes = [C.c_void_p, C.c_long]
        L.so_stdev_trim.restype = C.c_double
        L.so_stdev_trim.argtypes = [C.c_void_p, C.c_long]
        L.so_cigar_analyze.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def qname_hash(name: str) -> int:
    b = name.encode()
    return int(lib().so_qname_hash(b, len(b)))


class OracleResult:
    """Copies of the oracle's outputs as numpy arrays (same struct layouts as the device library)."""

    def __init__(self, v: _View):
        cp = lambda p, dt, n: abi.view(p, dt, n).copy()
        self.leads = cp(v.leads, abi.LEAD_DTYPE, v.n_leads)
        self.task_read_count = cp(v.task_read_count, "<u4", v.n_task)
        self.task_mean_nm = cp(v.task_mean_nm, "<f8", v.n_task)
        self.task_cov_mean = cp(v.task_cov_mean, "<f8", v.n_task)
        self.n_pass, self.soft_errors = int(v.n_pass), int(v.soft_errors)
        self.cand = cp(v.cand, abi.CAND_DTYPE, v.n_cand)
        self.cand_leads = cp(v.cand_leads, abi.LEAD_DTYPE, v.n_cand_leads)
        self.rn_off = cp(v.rn_off, "<u4", v.n_cand + 1)
        self.rnames = cp(v.rnames, "<u8", int(self.rn_off[-1]) if v.n_cand else 0)
        self.alt = cp(v.alt, "u1", v.n_alt)
        self.rec_nm = None

    def alt_of(self, i) -> str:
        c = self.cand[i]
        if c["alt_off"] < 0:
            return None
        return self.alt[int(c["alt_off"]):int(c["alt_off"]) + int(c["alt_len"])].tobytes().decode()


def run(block, config: abi.Config, stages=3, threads=1, keep_rec_nm=False) -> OracleResult:
    L = lib()
    rs = block.as_struct()
    h = L.so_run(C.byref(rs), C.byref(config), int(stages), int(threads))
    if not h:
        raise MemoryError("so_run failed")
    try:
        v = _View()
        L.so_get(h, C.byref(v))
        res = OracleResult(v)
        if keep_rec_nm:
            res.rec_nm = abi.view(v.rec_nm, "<f8", len(block.rec)).copy()
        return res
    finally:
        L.so_free(h)
