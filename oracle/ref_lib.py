"""ctypes front of oracle/_ref/libldw_ref.so -- the reference's OWN compiled C++ (test infrastructure, NOT the product).

``libldw_ref.so`` is /root/reference/src/{getACGTNsites.cpp, computeMI.cpp, ACGTN2num_parallel.cpp, fintersect.cpp,
kseq2.h} compiled unmodified against ``oracle/mock_rcpp/Rcpp.h`` (recipe ``oracle/Makefile``, target ``ref``).  Every
function here calls that object code; nothing is restated.  The functions mirror the R-level names of
``R/RcppExports.R:4-46``.  On the GPU box /root/reference does not exist: the prebuilt .so travels with the snapshot,
and ``available()`` tells tests whether it is there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libldw_ref.so")
REFERENCE_ROOT = os.environ.get("LDW_REFERENCE_ROOT", "/root/reference")
_LIB = None


def build() -> Optional[str]:
    """(Re)build from the reference sources when they are present; otherwise return the prebuilt file if any."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "kseq2.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE_ROOT}"])
    return _SO if os.path.exists(_SO) else None


def available() -> bool:
    return build() is not None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libldw_ref.so is missing and /root/reference is absent: cannot build it")
        L = C.CDLL(so)
        vp = C.c_void_p
        for name in ("ldwref_extractAlnParam", "ldwref_extractSNPs", "ldwref_extractRef", "ldwref_kseq_read_all"):
            getattr(L, name).restype = vp
        L.ldwref_extractAlnParam.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double]
        L.ldwref_extractSNPs.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_long]
        L.ldwref_extractRef.argtypes = [C.c_char_p]
        L.ldwref_kseq_read_all.argtypes = [C.c_char_p]
        L.ldwref_list_free.argtypes = [vp]
        L.ldwref_list_has.argtypes = [vp, C.c_char_p]
        L.ldwref_list_int.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
        L.ldwref_list_intvec.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_int))]
        L.ldwref_list_intvec.restype = C.c_long
        L.ldwref_list_matrix.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int)]
        L.ldwref_list_matrix.restype = C.c_long
        L.ldwref_list_strvec_len.argtypes = [vp, C.c_char_p]
        L.ldwref_list_strvec_len.restype = C.c_long
        L.ldwref_list_strvec_get.argtypes = [vp, C.c_char_p, C.c_long]
        L.ldwref_list_strvec_get.restype = C.c_char_p
        L.ldwref_list_str.argtypes = [vp, C.c_char_p, C.POINTER(C.c_long)]
        L.ldwref_list_str.restype = C.c_void_p
        dp = C.POINTER(C.c_double)
        L.ldwref_ACGTN2num.argtypes = [dp, C.c_char_p, C.c_long, C.c_int]
        L.ldwref_fastHadamard.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int, C.c_int, dp, dp, C.c_int]
        L.ldwref_compareToRow.argtypes = [dp, C.c_int, C.c_int, dp, C.c_long, C.POINTER(C.c_int)]
        L.ldwref_vecPosMatch.argtypes = [dp, C.c_long, dp, C.c_long, dp]
        L.ldwref_compareTriplet.argtypes = [dp, dp, C.c_long, C.c_double]
        L.ldwref_compareTriplet.restype = C.c_int
        L.ldwref_fast_intersect.argtypes = [C.POINTER(C.c_int), C.c_long, C.POINTER(C.c_int), C.c_long,
                                            C.POINTER(C.c_int)]
        L.ldwref_fast_intersect.restype = C.c_long
        L.ldwref_kseq_count.argtypes = [vp]
        L.ldwref_kseq_count.restype = C.c_long
        L.ldwref_kseq_last_rc.argtypes = [vp]
        L.ldwref_kseq_name.argtypes = [vp, C.c_long]
        L.ldwref_kseq_name.restype = C.c_char_p
        L.ldwref_kseq_seq.argtypes = [vp, C.c_long, C.POINTER(C.c_long)]
        L.ldwref_kseq_seq.restype = C.c_void_p
        L.ldwref_kseq_free.argtypes = [vp]
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _list_int(h, name: str) -> Optional[int]:
    v = C.c_int()
    return v.value if lib().ldwref_list_int(h, name.encode(), C.byref(v)) == 0 else None


def _list_intvec(h, name: str) -> Optional[np.ndarray]:
    p = C.POINTER(C.c_int)()
    n = lib().ldwref_list_intvec(h, name.encode(), C.byref(p))
    if n < 0:
        return None
    return np.ctypeslib.as_array(p, shape=(n,)).astype(np.int32) if n else np.zeros(0, np.int32)


def _list_matrix(h, name: str) -> Optional[np.ndarray]:
    p = C.POINTER(C.c_double)()
    nr, nc = C.c_int(), C.c_int()
    n = lib().ldwref_list_matrix(h, name.encode(), C.byref(p), C.byref(nr), C.byref(nc))
    if n < 0:
        return None
    flat = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0)
    return flat.reshape((nr.value, nc.value), order="F")


def _list_strvec(h, name: str) -> Optional[List[str]]:
    n = lib().ldwref_list_strvec_len(h, name.encode())
    if n < 0:
        return None
    return [lib().ldwref_list_strvec_get(h, name.encode(), i).decode("latin-1") for i in range(n)]


def kseq_read_all(path: str) -> Tuple[List[bytes], List[bytes], int]:
    """Every record ``kseq_read`` (src/kseq2.h:167-207) returns for the file, as the loop of
    src/getACGTNsites.cpp:50 sees them: (names, sequences of seq.l bytes, the return code that ended the loop)."""
    L = lib()
    h = L.ldwref_kseq_read_all(os.fsencode(path))
    if not h:
        raise OSError(f"gzopen failed: {path}")
    try:
        names, seqs = [], []
        for i in range(L.ldwref_kseq_count(h)):
            names.append(L.ldwref_kseq_name(h, i))
            n = C.c_long()
            p = L.ldwref_kseq_seq(h, i, C.byref(n))
            seqs.append(C.string_at(p, n.value))
        return names, seqs, L.ldwref_kseq_last_rc(h)
    finally:
        L.ldwref_kseq_free(h)


def extractAlnParam(path: str, filter: int, gap_thresh: float, maf_thresh: float) -> Dict:
    """``.extractAlnParam`` (src/getACGTNsites.cpp:13-176).  NB the reference dereferences a NULL sequence buffer when
    the file holds no record at all (``strlen(seq->seq.s)``, :36): callers must not pass such a file."""
    L = lib()
    h = L.ldwref_extractAlnParam(os.fsencode(path), filter, gap_thresh, maf_thresh)
    try:
        out = {"seq.length": _list_int(h, "seq.length")}
        if L.ldwref_list_has(h, b"num.seqs"):
            out.update({"num.seqs": _list_int(h, "num.seqs"), "num.snps": _list_int(h, "num.snps"),
                        "seq.names": _list_strvec(h, "seq.names"), "pos": _list_intvec(h, "pos")})
        return out
    finally:
        L.ldwref_list_free(h)


def extractSNPs(path: str, n_seq: int, n_snp: int, POS) -> Dict:
    """``.extractSNPs`` (src/getACGTNsites.cpp:179-291): seq.names, ACGTN_table [5, n_snp] and the 15 COO vectors."""
    L = lib()
    pos = np.ascontiguousarray(POS, dtype=np.int32)
    h = L.ldwref_extractSNPs(os.fsencode(path), n_seq, n_snp, pos.ctypes.data_as(C.POINTER(C.c_int)), len(pos))
    try:
        out = {"seq.names": _list_strvec(h, "seq.names"), "ACGTN_table": _list_matrix(h, "ACGTN_table")}
        for a in "ACGTN":
            for k in "ijx":
                out[f"{k}_{a}"] = _list_intvec(h, f"{k}_{a}")
        return out
    finally:
        L.ldwref_list_free(h)


def codes_from_coo(coo: Dict, n_snp: int, n_seq: int) -> np.ndarray:
    """The [n_snp, n_seq] class matrix the 15 COO vectors of ``.extractSNPs`` describe (``sparseMatrix(i, j, ...)`` then
    ``t()``, R/extractSNPs.R:100-141).  255 marks a cell no allele vector covers; a cell covered twice raises."""
    codes = np.full((n_snp, n_seq), 255, dtype=np.uint8)
    for a, ch in enumerate("ACGTN"):
        i, j = coo[f"i_{ch}"].astype(np.int64) - 1, coo[f"j_{ch}"].astype(np.int64) - 1
        if np.any(codes[j, i] != 255):
            raise ValueError("cell covered by two allele vectors")
        codes[j, i] = a
        assert np.all(coo[f"x_{ch}"] == a + 1)
    return codes


def ACGTN2num(nv: np.ndarray, cv: bytes, ncores: int = 1) -> None:
    """``.ACGTN2num`` (src/ACGTN2num_parallel.cpp:10-43); nv [5, n] float64 Fortran order, modified in place."""
    assert nv.flags["F_CONTIGUOUS"] and nv.dtype == np.float64 and nv.shape[0] == 5
    lib().ldwref_ACGTN2num(_dp(nv), cv, nv.shape[1], ncores)


def fastHadamard(MI, den, uq, pxy, pxpy, RXY, pXrX, pYrY, ncores: int = 1) -> None:
    """``.fastHadamard`` (src/computeMI.cpp:11-21); all Fortran-ordered float64, MI modified in place, RXY of any
    shape with the same element count (quirk Q1)."""
    for m in (MI, den, uq, pxy, pxpy, RXY, pXrX, pYrY):
        assert m.flags["F_CONTIGUOUS"] and m.dtype == np.float64
    assert RXY.size == MI.size
    lib().ldwref_fastHadamard(_dp(MI), MI.shape[0], MI.shape[1], _dp(den), _dp(uq), _dp(pxy), _dp(pxpy), _dp(RXY),
                              RXY.shape[0], RXY.shape[1], _dp(pXrX), _dp(pYrY), ncores)


def compareToRow(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """``.compareToRow`` (src/computeMI.cpp:25-41)."""
    x = np.asfortranarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(x.shape[0], dtype=np.int32)
    lib().ldwref_compareToRow(_dp(x), x.shape[0], x.shape[1], _dp(y), len(y), out.ctypes.data_as(C.POINTER(C.c_int)))
    return out.astype(bool)


def vecPosMatch(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """``.vecPosMatch`` (src/computeMI.cpp:44-59)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(len(x), dtype=np.float64)
    lib().ldwref_vecPosMatch(_dp(x), len(x), _dp(y), len(y), _dp(out))
    return out


def compareTriplet(MI0X: np.ndarray, MI0Z: np.ndarray, MI0: float) -> bool:
    """``.compareTriplet`` (src/computeMI.cpp:63-77)."""
    a = np.ascontiguousarray(MI0X, dtype=np.float64)
    b = np.ascontiguousarray(MI0Z, dtype=np.float64)
    return bool(lib().ldwref_compareTriplet(_dp(a), _dp(b), len(a), float(MI0)))


def fast_intersect(A, B) -> np.ndarray:
    """``.fast_intersect`` (src/fintersect.cpp:6-32)."""
    a = np.ascontiguousarray(A, dtype=np.int32)
    b = np.ascontiguousarray(B, dtype=np.int32)
    out = np.zeros(max(1, min(len(a), len(b))), dtype=np.int32)
    ip = C.POINTER(C.c_int)
    n = lib().ldwref_fast_intersect(a.ctypes.data_as(ip), len(a), b.ctypes.data_as(ip), len(b), out.ctypes.data_as(ip))
    return out[:n].copy()
