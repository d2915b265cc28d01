"""Host-side mirror of LDWeaver's R interface for the hot path (same names, argument meaning and
error behaviour), driving libldwgpu through its C ABI.  R itself is not available in this image, so
this Python layer stands where the patched R wrappers of ``r_package/`` stand on a machine with R.

Reference functions mirrored (LDWeaver v1.5.2):
  parse_fasta_alignment              R/extractSNPs.R:23-142
  parse_fasta_SNP_alignment          R/extractSNPs.R:168-281
  estimate_Hamming_distance_weights  R/performPopulationStuctureCorrection.R:20-81
  perform_MI_computation             R/computePairwiseMI.R:46-145 (scan + link filter; the model fit /
                                     ARACNE steps after it stay in R and are out of scope here)
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr

ALLELES = "ACGTN"


@dataclass
class SnpDat:
    """``snp.dat`` (R/extractSNPs.R:138-141).  The five sparse ``snp.matrix_*`` slots partition the
    nsnp x nseq grid; they are held as one uint8 class matrix and materialised on request."""
    codes: np.ndarray          # [nsnp, nseq] uint8, 0..4 = A,C,G,T,N
    g: Optional[int]
    nsnp: int
    nseq: int
    seq_names: List[str]
    r: np.ndarray              # rowSums(uqe), float64
    uqe: np.ndarray            # [nsnp, 5] 0/1 float64
    POS: np.ndarray            # int32

    def snp_matrix(self, allele: str):
        """``snp.matrix_<allele>`` as a scipy CSC boolean matrix (nsnp x nseq), like the lgCMatrix."""
        import scipy.sparse as sp
        return sp.csc_matrix(self.codes == ALLELES.index(allele))

    @property
    def snp_matrix_A(self): return self.snp_matrix("A")
    @property
    def snp_matrix_C(self): return self.snp_matrix("C")
    @property
    def snp_matrix_G(self): return self.snp_matrix("G")
    @property
    def snp_matrix_T(self): return self.snp_matrix("T")
    @property
    def snp_matrix_N(self): return self.snp_matrix("N")


def _method_to_filter(method: str) -> int:
    if method == "default":
        return 0
    if method == "relaxed":
        return 1
    warnings.warn("Unkown filtering method, using default...")  # R/extractSNPs.R:34
    return 0


def read_fasta_matrix(aln_path: str):
    """gz/plain multi-FASTA -> (names, uint8 matrix [nseq, seq_len]) via the library's host reader."""
    L = _lib.lib()
    nseq, slen = C.c_int64(), C.c_int64()
    check(L.ldw_read_fasta(aln_path.encode(), C.byref(nseq), C.byref(slen), None, 0, None, 0))
    if slen.value == -1:
        raise ValueError("Error! sequences are of different lengths!")  # R/extractSNPs.R:41
    if nseq.value == 0:
        raise ValueError("File does not contain any sequences!")  # :42
    aln = np.empty((nseq.value, slen.value), dtype=np.uint8)
    names = C.create_string_buffer(1 << 16)
    cap = 1 << 16
    while True:
        n2, l2 = C.c_int64(), C.c_int64(slen.value)
        rc = L.ldw_read_fasta(aln_path.encode(), C.byref(n2), C.byref(l2), ptr(aln), aln.size, names, cap)
        if rc != 0 and b"names buffer" in L.ldw_last_error():
            cap *= 8
            names = C.create_string_buffer(cap)
            continue
        check(rc)
        break
    name_list = names.raw.split(b"\0")[:nseq.value]
    return [n.decode(errors="replace") for n in name_list], aln


def _build_snp_dat(aln: np.ndarray, names: Sequence[str], filt: int, gap_freq: float, maf_freq: float,
                   pos: Optional[np.ndarray], device: int) -> SnpDat:
    L = _lib.lib()
    ctx = _lib.default_context(device)
    aln = np.ascontiguousarray(aln, dtype=np.uint8)
    nseq, slen = aln.shape
    pos_idx = np.empty(slen, dtype=np.int32)
    n_snp = C.c_int64()
    check(L.ldw_aln_param(ctx.handle, ptr(aln), nseq, slen, filt, gap_freq, maf_freq, ptr(pos_idx), C.byref(n_snp), None))
    n = int(n_snp.value)
    if n == 0:
        raise ValueError("File does not contain any SNPs")  # R/extractSNPs.R:43
    if pos is not None and len(pos) != slen:
        raise ValueError("Error! Number of positions do not match the fasta sequence length")  # :194
    pos_idx = np.ascontiguousarray(pos_idx[:n])
    codes = np.empty((n, nseq), dtype=np.uint8)
    table = np.empty(5 * n, dtype=np.float64)
    check(L.ldw_extract_snps(ctx.handle, ptr(aln), nseq, slen, ptr(pos_idx), n, ptr(codes), ptr(table)))
    table = table.reshape(n, 5)  # == t(ACGTN_table)
    uqe = (table > 0).astype(np.float64)  # R/extractSNPs.R:47
    if pos is None:
        POS, g = pos_idx.astype(np.int32), int(slen)  # :140
    else:
        POS, g = np.asarray(pos)[pos_idx.astype(np.int64) - 1].astype(np.int32), None  # :200, :279
    return SnpDat(codes=codes, g=g, nsnp=n, nseq=int(nseq), seq_names=[s.lstrip(">") for s in names],
                  r=uqe.sum(axis=1), uqe=uqe, POS=POS)


def parse_fasta_alignment(aln_path: str, gap_freq: float = 0.15, maf_freq: float = 0.01, method: str = "default",
                          mega_dset: bool = False, device: int = 0) -> SnpDat:
    """R/extractSNPs.R:23-142.  ``mega_dset`` only changed the sparse-matrix package in the reference
    (spam64 vs Matrix); results are identical, so it is accepted and ignored."""
    aln_path = os.path.abspath(os.path.expanduser(aln_path))
    if not os.path.exists(aln_path):
        raise FileNotFoundError(f"Can't locate file {aln_path}")  # :27
    filt = _method_to_filter(method)
    names, aln = read_fasta_matrix(aln_path)
    return _build_snp_dat(aln, names, filt, gap_freq, maf_freq, None, device)


def parse_fasta_SNP_alignment(aln_path: str, pos: Sequence[int], gap_freq: float = 0.15, maf_freq: float = 0.01,
                              method: str = "default", mega_dset: bool = False, device: int = 0) -> SnpDat:
    """R/extractSNPs.R:168-281 (SNP-only alignment + positions; ``g`` is left None, :279)."""
    aln_path = os.path.abspath(os.path.expanduser(aln_path))
    if not os.path.exists(aln_path):
        raise FileNotFoundError(f"Can't locate file {aln_path}")
    filt = _method_to_filter(method)
    names, aln = read_fasta_matrix(aln_path)
    return _build_snp_dat(aln, names, filt, gap_freq, maf_freq, np.asarray(pos), device)


def snp_dat_from_alignment_matrix(aln: np.ndarray, names: Optional[Sequence[str]] = None, pos=None, gap_freq=0.15,
                                  maf_freq=0.01, method="default", device: int = 0) -> SnpDat:
    """Same as the two parsers but from an already tokenised byte matrix (tests, synthetic data)."""
    names = list(names) if names is not None else [f"s{i}" for i in range(aln.shape[0])]
    return _build_snp_dat(aln, names, _method_to_filter(method), gap_freq, maf_freq,
                          None if pos is None else np.asarray(pos), device)


def snp_dat_from_codes(codes: np.ndarray, POS: np.ndarray, g: Optional[int], seq_names=None) -> SnpDat:
    """A ``snp.dat`` loaded without a device handle (e.g. from an RDS in the reference's resume path,
    R/BacGWES.R:281,301-302): uqe / r are re-derived from the class matrix exactly as R/extractSNPs.R:47,141."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    nsnp, nseq = codes.shape
    uqe = np.stack([(codes == a).any(axis=1) for a in range(5)], axis=1).astype(np.float64)
    return SnpDat(codes=codes, g=g, nsnp=nsnp, nseq=nseq, seq_names=list(seq_names) if seq_names is not None else
                  [f"s{i}" for i in range(nseq)], r=uqe.sum(axis=1), uqe=uqe, POS=np.asarray(POS, dtype=np.int32))


def acgtn2num(nv: np.ndarray, cv, ncores: int = 1, device: int = 0) -> None:
    """``.ACGTN2num(nv, cv, ncores)`` (R/RcppExports.R:4-6): in place on a 5 x n column-major matrix."""
    if not (nv.dtype == np.float64 and nv.flags["F_CONTIGUOUS"] and nv.shape[0] == 5):
        raise ValueError("nv must be a Fortran-ordered float64 [5, n] matrix (R column-major)")
    ref = "".join((c[0] if len(c) else " ") for c in cv).encode("latin-1") if not isinstance(cv, (bytes, bytearray)) else bytes(cv)
    if len(ref) != nv.shape[1]:
        raise ValueError("cv length does not match ncol(nv)")
    ctx = _lib.default_context(device)
    check(_lib.lib().ldw_acgtn2num(ctx.handle, ptr(nv), ref, nv.shape[1]))


def estimate_Hamming_distance_weights(snp_dat: SnpDat, threshold: float = 0.1, mega_dset: bool = False,
                                      device: int = 0, return_parts: bool = False):
    """R/performPopulationStuctureCorrection.R:20-81.  Returns the weight vector (``hdw``); with
    ``return_parts`` also the neighbour counts and the integer Hamming-distance matrix."""
    ctx = _lib.default_context(device)
    codes = np.ascontiguousarray(snp_dat.codes, dtype=np.uint8)
    n, S = codes.shape
    cnt = np.empty(S, dtype=np.int32)
    hdw = np.empty(S, dtype=np.float64)
    dist = np.empty((S, S), dtype=np.int32, order="F") if return_parts else None
    check(_lib.lib().ldw_hdw(ctx.handle, ptr(codes), n, S, float(threshold), ptr(cnt), ptr(hdw),
                             ptr(dist) if return_parts else None))
    if return_parts:
        return hdw, cnt, dist
    return hdw
