"""ctypes binding of oracle/oracle.c (test infrastructure, NOT the product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "libldw_oracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ldwo_aln_param.restype = C.c_int64
        _LIB.ldwo_block_links.restype = C.c_int64
        _LIB.ldwo_num_threads.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def aln_param(aln: np.ndarray, filter: int, gap: float, maf: float):
    aln = np.ascontiguousarray(aln, dtype=np.uint8)
    nseq, L = aln.shape
    counts = np.zeros(5 * L, dtype=np.float64)
    pos = np.zeros(L, dtype=np.int32)
    n = lib().ldwo_aln_param(_p(aln, C.c_uint8), C.c_int64(nseq), C.c_int64(L), C.c_int(filter), C.c_double(gap),
                             C.c_double(maf), _p(counts, C.c_double), _p(pos, C.c_int32))
    return pos[:n].copy(), counts.reshape(L, 5).T.copy()


def extract_snps(aln: np.ndarray, pos: np.ndarray):
    aln = np.ascontiguousarray(aln, dtype=np.uint8)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    nseq, L = aln.shape
    nsnp = len(pos)
    codes = np.zeros((nsnp, nseq), dtype=np.uint8)
    table = np.zeros(5 * nsnp, dtype=np.float64)
    lib().ldwo_extract_snps(_p(aln, C.c_uint8), C.c_int64(nseq), C.c_int64(L), _p(pos, C.c_int32), C.c_int64(nsnp),
                            _p(codes, C.c_uint8), _p(table, C.c_double))
    return codes, table.reshape(nsnp, 5).T.copy()


def acgtn2num(nv: np.ndarray, cv: bytes) -> None:
    """nv: [5, n] float64 Fortran-ordered (R column-major), modified in place."""
    assert nv.flags["F_CONTIGUOUS"] and nv.dtype == np.float64
    lib().ldwo_acgtn2num(_p(nv, C.c_double), C.c_char_p(cv), C.c_int64(nv.shape[1]))


def hdw(codes: np.ndarray, threshold: float, want_dist: bool = False):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    nsnp, nseq = codes.shape
    cnt = np.zeros(nseq, dtype=np.int32)
    w = np.zeros(nseq, dtype=np.float64)
    dist = np.zeros((nseq, nseq), dtype=np.int32, order="F") if want_dist else None
    lib().ldwo_hdw(_p(codes, C.c_uint8), C.c_int64(nsnp), C.c_int64(nseq), C.c_double(threshold), _p(cnt, C.c_int32),
                   _p(w, C.c_double), _p(dist, C.c_int32) if want_dist else None)
    return (w, cnt, dist) if want_dist else (w, cnt)


def block_mi(codes, hdw_, r, uqe, from_idx, to_idx, ncores: int = 0) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    nsnp, nseq = codes.shape
    hdw_ = np.ascontiguousarray(hdw_, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    uqe_f = np.asfortranarray(uqe, dtype=np.float64)
    f = np.ascontiguousarray(from_idx, dtype=np.int32)
    t = np.ascontiguousarray(to_idx, dtype=np.int32)
    MI = np.zeros((len(f), len(t)), dtype=np.float64, order="F")
    if ncores <= 0:
        ncores = lib().ldwo_num_threads()
    lib().ldwo_block_mi(_p(codes, C.c_uint8), C.c_int64(nsnp), C.c_int64(nseq), _p(hdw_, C.c_double),
                        _p(r, C.c_double), _p(uqe_f, C.c_double), _p(f, C.c_int32), C.c_int64(len(f)),
                        _p(t, C.c_int32), C.c_int64(len(t)), C.c_int(ncores), _p(MI, C.c_double))
    return MI


def block_links(MI, POS, from_idx, to_idx, g, sr_dist, lr_retain_links, lr_links_approx, sr_only=False):
    MI = np.asfortranarray(MI, dtype=np.float64)
    POS = np.ascontiguousarray(POS, dtype=np.float64)
    f = np.ascontiguousarray(from_idx, dtype=np.int32)
    t = np.ascontiguousarray(to_idx, dtype=np.int32)
    cap = len(f) * len(t)
    row = np.zeros(cap, dtype=np.int32)
    col = np.zeros(cap, dtype=np.int32)
    ln = np.zeros(cap, dtype=np.float64)
    mi = np.zeros(cap, dtype=np.float64)
    is_sr = np.zeros(cap, dtype=np.uint8)
    keep = np.zeros(cap, dtype=np.uint8)
    thr = C.c_double()
    prob = C.c_double()
    n = lib().ldwo_block_links(_p(MI, C.c_double), _p(POS, C.c_double), _p(f, C.c_int32), C.c_int64(len(f)),
                               _p(t, C.c_int32), C.c_int64(len(t)), C.c_double(g), C.c_double(sr_dist),
                               C.c_double(lr_retain_links), C.c_double(lr_links_approx if lr_links_approx else 1.0),
                               C.c_int(int(sr_only)), _p(row, C.c_int32), _p(col, C.c_int32), _p(ln, C.c_double),
                               _p(mi, C.c_double), _p(is_sr, C.c_uint8), _p(keep, C.c_uint8), C.byref(thr),
                               C.byref(prob))
    return dict(row=row[:n], col=col[:n], len=ln[:n], MI=mi[:n], is_sr=is_sr[:n].astype(bool),
                lr_keep=keep[:n].astype(bool), thr=thr.value, prob=prob.value)
