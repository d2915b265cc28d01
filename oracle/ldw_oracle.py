"""CPU ORACLE (test infrastructure, NOT the product path).

A literal NumPy restatement of LDWeaver's genome-wide pairwise-LD hot path
(reference R package v1.5.2).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product path (``ldweaver_b200``) never does.

PARITY STATUS: "parity unpinned" -- the reference ships no golden vectors, no
known-answer tests and no saved outputs for this path and R
is not installed here, so the reference itself cannot be run.  What pins this
oracle instead: (1) it follows the reference statement by statement (every
function cites the file:line it restates), (2) an independent C restatement
(``oracle/oracle.c``) must agree with it (integers bit-exact, MI <= 1e-12),
(3) the closed form of SURVEY.md section 8a must agree with the literal 8-matrix form,
(4) the survey-time probe values on ``inst/extdata/snp_sample.fa.gz`` are
re-derived in ``tests/test_oracle.py``.

All matrices are dense; this is for small cases (seconds).  Shapes follow R:
``snp.matrix_X`` is nsnp x nseq.
"""
from __future__ import annotations

import gzip
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ALLELES = "ACGTN"


# --------------------------------------------------------------------------------------
# FASTA reading (host side of src/getACGTNsites.cpp: kseq over gzread)
# --------------------------------------------------------------------------------------
def read_fasta(path: str) -> Tuple[List[str], List[bytes]]:
    """Multi-FASTA reader (gz or plain) restating klib's ``kseq_read`` as vendored in src/kseq2.h:167-207, as its
    callers see the records (``strlen(seq->seq.s)`` / C string of ``seq->name.s``, src/getACGTNsites.cpp:36,51-52).

    * :171-174 skip to the first '>' or '@' (anywhere) when no header character is pending;
    * :176-178 name = up to the first isspace() byte; the rest of the line is a comment unless that byte was '\n';
    * :183-186 line by line: first byte '>' / '@' -> next record, '+' -> quality section, ANY other first byte
      (a '\n' of an empty line, '\r', blank) is appended and so is everything up to the next '\n' -- CR of CRLF files
      and inner blanks are sequence bytes, an empty line glues the following line onto the sequence;
    * :196-205 after '+': drop that line, append whole lines to qual until len(qual) >= len(seq); a mismatch or a
      missing quality section returns -2, which ends the callers' ``while ((l = kseq_read(seq)) >= 0)`` loops;
    * :87-98/:176 a header character that is the very last byte of the stream gives no record unless the stream length
      is a multiple of the 16384-byte buffer (EOF not yet known to ks_getuntil).
    Pinned against the compiled reference reader in tests/test_fasta_ref_cpu.py / tests/test_ref_pins_cpu.py."""
    opener = gzip.open if _is_gz(path) else open
    with opener(path, "rb") as fh:
        buf = fh.read()
    n = len(buf)
    names: List[str] = []
    seqs: List[bytes] = []
    space = b" \t\n\v\f\r"
    i = 0
    pending_header = False
    while True:
        if not pending_header:  # :171-174
            while i < n and buf[i] not in b">@":
                i += 1
            if i >= n:
                break
            i += 1
        pending_header = False
        if i >= n and n % 16384 != 0:  # ks_getuntil returns -1 at a known EOF (:88)
            break
        j = i
        while j < n and buf[j] not in space:
            j += 1
        name = buf[i:j]
        delim = buf[j] if j < n else 0
        i = j + 1
        if delim != 0x0A:  # comment up to '\n' (:178)
            k = buf.find(b"\n", i)
            i = n if k < 0 else k + 1
        seq = bytearray()
        c = -1
        while i < n:  # :183-186
            c = buf[i]
            i += 1
            if c in b">+@":
                break
            seq.append(c)
            k = buf.find(b"\n", i)
            if k < 0:
                seq += buf[i:]
                i = n
            else:
                seq += buf[i:k]
                i = k + 1
            c = -1
        if c != -1 and c in b">@":
            pending_header = True
        ok = True
        if c == 0x2B:  # '+'  (:196-205)
            k = buf.find(b"\n", i)
            if k < 0:
                ok = False
                i = n
            else:
                i = k + 1
                qual = 0
                while True:
                    if i >= n:
                        break
                    k = buf.find(b"\n", i)
                    if k < 0:
                        qual += n - i
                        i = n
                    else:
                        qual += k - i
                        i = k + 1
                    if qual >= len(seq):
                        break
                ok = qual == len(seq)
            if not ok:
                break  # return -2: the callers' loops end
        names.append(name.split(b"\0", 1)[0].decode("latin-1"))
        seqs.append(bytes(seq).split(b"\0", 1)[0])
    return names, seqs


def _is_gz(path: str) -> bool:
    with open(path, "rb") as fh:
        return fh.read(2) == b"\x1f\x8b"  # gzread inflates by magic number, not by file name


def classify_bytes(arr: np.ndarray) -> np.ndarray:
    """[Aa]->0 [Cc]->1 [Gg]->2 [Tt]->3 everything else->4
    (reference src/getACGTNsites.cpp:59-69 and :233-263; quirk Q8)."""
    lut = np.full(256, 4, dtype=np.uint8)
    for k, ch in enumerate("ACGT"):
        lut[ord(ch)] = k
        lut[ord(ch.lower())] = k
    return lut[arr]


# --------------------------------------------------------------------------------------
# a1  extractAlnParam  (src/getACGTNsites.cpp:13-176)
# --------------------------------------------------------------------------------------
def extract_aln_param(names: Sequence[str], seqs: Sequence[bytes], filter: int,
                      gap_thresh: float, maf_thresh: float) -> dict:
    """Per-column allele counts and the SNP filter.

    Follows src/getACGTNsites.cpp:33-39 (length of first record), :50-85 (counting,
    unequal length -> seq.length = -1 at :54-56), :104-134 (default filter) and
    :135-166 (relaxed filter).  Returns the same list as :169-174.
    """
    if len(seqs) == 0:
        return {"num.seqs": 0, "num.snps": 0, "seq.length": 0, "seq.names": [], "pos": []}
    seq_length = len(seqs[0])
    n = 0
    allele_counts = np.zeros((5, seq_length), dtype=np.float64)  # NumericMatrix :47
    for s in seqs:
        if len(s) != seq_length:
            return {"seq.length": -1}
        cls = classify_bytes(np.frombuffer(s, dtype=np.uint8))
        for a in range(5):
            allele_counts[a] += (cls == a)
        n += 1
    pos: List[int] = []
    if filter == 0:
        min_maf = int(n * maf_thresh)  # int truncation :105
        for j in range(seq_length):
            col = allele_counts[:, j]
            if int(np.count_nonzero(col[:4] > 0)) > 1:  # chk_flag > 1  :114-117
                if col[4] / n < gap_thresh:  # :118
                    snp = np.sort(col[:4])  # :119-121
                    if snp[2] > min_maf:  # second largest non-gap count :122
                        pos.append(j + 1)
    else:
        min_maf = int(n * (1 - maf_thresh))  # :136
        for j in range(seq_length):
            col = allele_counts[:, j]
            if int(np.count_nonzero(col[:4] > 0)) > 1:  # :145-147
                if col[4] / n < gap_thresh:  # :148
                    if col.max() <= min_maf:  # :153 (max over all 5 rows)
                        pos.append(j + 1)
    return {"num.seqs": n, "num.snps": len(pos), "seq.length": seq_length,
            "seq.names": list(names), "pos": pos, "allele_counts": allele_counts}


# --------------------------------------------------------------------------------------
# a2  extractSNPs  (src/getACGTNsites.cpp:179-291)
# --------------------------------------------------------------------------------------
def extract_snps(seqs: Sequence[bytes], n_seq: int, n_snp: int, POS: Sequence[int]) -> dict:
    """Class of every (sequence, retained column) and the 5 x nsnp ACGTN_table
    (src/getACGTNsites.cpp:222-267).  The 15 COO vectors of the reference partition the
    nseq x nsnp grid, so they are returned here as one uint8 code matrix [nsnp x nseq]
    plus, on request, the explicit triplets (see ``coo_triplets``)."""
    idx = np.asarray(POS, dtype=np.int64) - 1
    codes = np.empty((n_snp, n_seq), dtype=np.uint8)
    for s_i, s in enumerate(seqs):
        arr = np.frombuffer(s, dtype=np.uint8)
        codes[:, s_i] = classify_bytes(arr[idx])
    table = np.zeros((5, n_snp), dtype=np.float64)
    for a in range(5):
        table[a] = (codes == a).sum(axis=1)
    return {"codes": codes, "ACGTN_table": table}


def coo_triplets(codes: np.ndarray) -> Dict[str, np.ndarray]:
    """The i_X/j_X/x_X vectors exactly as src/getACGTNsites.cpp:233-263 pushes them
    (sequence-major, then SNP order; 1-based; x = 1..5 constant per allele)."""
    out: Dict[str, np.ndarray] = {}
    n_snp, n_seq = codes.shape
    ct = codes.T  # [seq, snp] -> iteration order of the reference
    for a, ch in enumerate(ALLELES):
        seq_i, snp_k = np.nonzero(ct == a)
        out["i_" + ch] = (seq_i + 1).astype(np.int32)
        out["j_" + ch] = (snp_k + 1).astype(np.int32)
        out["x_" + ch] = np.full(seq_i.shape, a + 1, dtype=np.int32)
    return out


# --------------------------------------------------------------------------------------
# a3  parse_fasta_alignment / parse_fasta_SNP_alignment  (R/extractSNPs.R:23-142, 168-281)
# --------------------------------------------------------------------------------------
@dataclass
class SnpDat:
    """The ``snp.dat`` list of R/extractSNPs.R:138-141.  ``snp_matrix[a]`` is the
    dense nsnp x nseq 0/1 matrix standing for ``snp.matrix_<a>`` (an lgCMatrix in R)."""
    codes: np.ndarray  # [nsnp, nseq] uint8 0..4  (== the five sparse matrices)
    g: Optional[int]
    nsnp: int
    nseq: int
    seq_names: List[str]
    r: np.ndarray  # rowSums(uqe)
    uqe: np.ndarray  # [nsnp, 5] 0/1 double
    POS: np.ndarray  # int

    def snp_matrix(self, a: int) -> np.ndarray:
        return (self.codes == a).astype(np.float64)


def _method_to_filter(method: str) -> int:
    # R/extractSNPs.R:29-36
    if method == "default":
        return 0
    if method == "relaxed":
        return 1
    return 0  # "Unkown filtering method, using default..."


def parse_fasta_alignment(aln_path: str, gap_freq: float = 0.15, maf_freq: float = 0.01,
                          method: str = "default") -> SnpDat:
    """R/extractSNPs.R:23-142."""
    names, seqs = read_fasta(aln_path)
    filt = _method_to_filter(method)
    param = extract_aln_param(names, seqs, filt, gap_freq, maf_freq)  # :39
    if param["seq.length"] == -1:
        raise ValueError("Error! sequences are of different lengths!")  # :41
    if param["num.seqs"] == 0:
        raise ValueError("File does not contain any sequences!")  # :42
    if param["num.snps"] == 0:
        raise ValueError("File does not contain any SNPs")  # :43
    data = extract_snps(seqs, param["num.seqs"], param["num.snps"], param["pos"])  # :45
    uqe = (data["ACGTN_table"] > 0).T.astype(np.float64)  # :47  -> nsnp x 5
    return SnpDat(codes=data["codes"], g=param["seq.length"], nsnp=param["num.snps"],
                  nseq=param["num.seqs"], seq_names=[n.lstrip(">") for n in names],
                  r=uqe.sum(axis=1), uqe=uqe, POS=np.asarray(param["pos"], dtype=np.int64))


def parse_fasta_SNP_alignment(aln_path: str, pos: Sequence[int], gap_freq: float = 0.15,
                              maf_freq: float = 0.01, method: str = "default") -> SnpDat:
    """R/extractSNPs.R:168-281 (SNP-only input: ``pos`` gives genome coordinates)."""
    names, seqs = read_fasta(aln_path)
    filt = _method_to_filter(method)
    param = extract_aln_param(names, seqs, filt, gap_freq, maf_freq)  # :187
    if param["seq.length"] == -1:
        raise ValueError("Error! sequences are of different lengths!")
    if param["num.seqs"] == 0:
        raise ValueError("File does not contain any sequences!")
    if param["num.snps"] == 0:
        raise ValueError("File does not contain any SNPs")
    if len(pos) != param["seq.length"]:  # :194
        raise ValueError("Error! Number of positions do not match the fasta sequence length")
    data = extract_snps(seqs, param["num.seqs"], param["num.snps"], param["pos"])  # :197
    real_pos = np.asarray(pos)[np.asarray(param["pos"], dtype=np.int64) - 1].astype(np.int64)  # :200
    uqe = (data["ACGTN_table"] > 0).T.astype(np.float64)
    return SnpDat(codes=data["codes"], g=None, nsnp=param["num.snps"], nseq=param["num.seqs"],
                  seq_names=[n.lstrip(">") for n in names], r=uqe.sum(axis=1), uqe=uqe, POS=real_pos)


def snp_dat_from_codes(codes: np.ndarray, POS: np.ndarray, g: Optional[int],
                       seq_names: Optional[List[str]] = None) -> SnpDat:
    """Build a ``snp.dat`` straight from a code matrix (what the five sparse matrices of
    R/extractSNPs.R:100-141 hold); uqe/r as R/extractSNPs.R:47,141."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    nsnp, nseq = codes.shape
    table = np.stack([(codes == a).sum(axis=1) for a in range(5)], axis=0)
    uqe = (table > 0).T.astype(np.float64)
    return SnpDat(codes=codes, g=g, nsnp=nsnp, nseq=nseq,
                  seq_names=seq_names or [f"s{i}" for i in range(nseq)], r=uqe.sum(axis=1), uqe=uqe,
                  POS=np.asarray(POS, dtype=np.int64))


# --------------------------------------------------------------------------------------
# a4  ACGTN2num  (src/ACGTN2num_parallel.cpp:10-43)
# --------------------------------------------------------------------------------------
def acgtn2num(nv: np.ndarray, cv: Sequence[str]) -> None:
    """In place: zero the reference-allele row of a 5 x n (column-major in R) matrix.
    Uppercase A/C/G/T, and 'N' or '-' -> row 4; any other character leaves the column
    untouched (src/ACGTN2num_parallel.cpp:21-38; quirks Q8, Q11).  ``nv`` is [5, n]."""
    rowmap = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4, "-": 4}
    for c, ch in enumerate(cv):
        ch0 = ch[0] if len(ch) else ""
        if ch0 in rowmap:
            nv[rowmap[ch0], c] = 0


# --------------------------------------------------------------------------------------
# a5  estimate_Hamming_distance_weights  (R/performPopulationStuctureCorrection.R:20-81)
# --------------------------------------------------------------------------------------
def estimate_Hamming_distance_weights(snp: SnpDat, threshold: float = 0.1,
                                      return_parts: bool = False):
    thresh = int(snp.nsnp * threshold)  # as.integer truncation :23
    shared = np.zeros((snp.nseq, snp.nseq), dtype=np.float64)
    for a in range(5):  # :49-74   crossprod(as(M_a,"matrix"), M_a)
        m = snp.snp_matrix(a)
        shared += m.T @ m
    flags = (snp.nsnp - shared) < thresh  # strict '<'  :76
    cnt = flags.sum(axis=0)  # colSums (includes self)
    hdw = 1.0 / (cnt + 1)
    if return_parts:
        return hdw, cnt.astype(np.int64), (snp.nsnp - shared).astype(np.int64), thresh
    return hdw


# --------------------------------------------------------------------------------------
# R's Mersenne-Twister + sample()  (base R; used by R/computePairwiseMI.R:95-96)
# --------------------------------------------------------------------------------------
class RMersenne:
    """R's default RNG (Mersenne-Twister, 'Inversion', sample.kind='Rejection') as seeded by
    ``set.seed(seed)``.  Restates base R's src/main/RNG.c (Randomize/RNG_Init/MT_sgenrand/
    MT_genrand/fixup/R_unif_index/rbits).  UNVERIFIED against a live R here (no R in the
    image) -- callers may pass ``lr_links_approx`` explicitly instead."""

    N = 624
    M = 397

    def __init__(self, seed: int):
        seed &= 0xFFFFFFFF
        for _ in range(50):  # initial scrambling
            seed = (69069 * seed + 1) & 0xFFFFFFFF
        iseed = []
        for _ in range(self.N + 1):
            seed = (69069 * seed + 1) & 0xFFFFFFFF
            iseed.append(seed)
        # FixupSeeds: dummy[0] = mti = N
        self.mt = iseed[1:]
        self.mti = self.N

    def _genrand(self) -> float:
        mt, N, M = self.mt, self.N, self.M
        if self.mti >= N:
            for kk in range(N - M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            for kk in range(N - M, N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            y = (mt[N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y * 2.3283064365386963e-10

    def unif_rand(self) -> float:
        v = self._genrand()
        if v <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v

    def _rbits(self, bits: int) -> int:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn: int) -> int:
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample(self, n: int, k: int) -> np.ndarray:
        """``sample(n, k)`` without replacement, n <= 1e7 path (partial Fisher-Yates,
        do_sample in src/main/random.c); returns 1-based indices."""
        x = list(range(n))
        out = np.empty(k, dtype=np.int64)
        nn = n
        for i in range(k):
            j = self.unif_index(nn)
            out[i] = x[j] + 1
            nn -= 1
            x[j] = x[nn]
        return out


def r_round_half_even(x: float) -> float:
    """R's round(x) for the magnitudes used here (IEC 60559 half-to-even)."""
    return float(np.round(x))


def lr_links_approx_reference(POS: np.ndarray, g: float, sr_dist: float, seed: int = 1988) -> float:
    """R/computePairwiseMI.R:94-97."""
    nsnp = len(POS)
    snp_subset = int(min(nsnp, r_round_half_even(nsnp * 0.1)))
    rng = RMersenne(seed)
    idx = rng.sample(nsnp, snp_subset)
    P = POS.astype(np.float64)
    cnt = 0
    for x in P[idx - 1]:
        cnt += int(np.count_nonzero((0.5 * g - np.abs(np.mod(x - P, g) - 0.5 * g)) > sr_dist))
    return cnt / snp_subset * nsnp / 2


# --------------------------------------------------------------------------------------
# a6  make_blocks  (R/computePairwiseMI.R:147-165)
# --------------------------------------------------------------------------------------
def r_round_to_thousands(x: float) -> int:
    """``round(max_blk_sz, -3)`` (R/computePairwiseMI.R:69)."""
    return int(np.round(x / 1000.0) * 1000)


def make_blocks(nsnp: int, max_blk_sz: int) -> List[Tuple[int, int, int, int]]:
    """1-based inclusive (from_s, from_e, to_s, to_e), rows i in 1..p, j in i..p."""
    part1 = int(math.ceil(nsnp / max_blk_sz))
    from_s = [(i - 1) * max_blk_sz + 1 for i in range(1, part1 + 1)]
    from_e = [min(i * max_blk_sz, nsnp) for i in range(1, part1 + 1)]
    out = []
    for i in range(part1):
        for j in range(i, part1):
            out.append((from_s[i], from_e[i], from_s[j], from_e[j]))
    return out


def circ_len(pos1: np.ndarray, pos2: np.ndarray, g: float) -> np.ndarray:
    """R/computePairwiseMI.R:330 with R's floored %%."""
    return 0.5 * g - np.abs(np.mod(pos1 - pos2, g) - 0.5 * g)


def quantile_type7(x: np.ndarray, prob: float) -> float:
    """stats::quantile.default(type = 7) for one probability (base R; quirk Q3)."""
    n = len(x)
    index = 1 + max(n - 1, 0) * prob
    lo = int(math.floor(index))
    hi = int(math.ceil(index))
    xs = np.sort(x)
    qs = xs[lo - 1]
    if index > lo and xs[hi - 1] != qs:
        h = index - lo
        qs = (1 - h) * qs + h * xs[hi - 1]
    return float(qs)


# --------------------------------------------------------------------------------------
# a7-a9  perform_MI_computation_ACGTN / computeMI_Sprase / fastHadamard
# --------------------------------------------------------------------------------------
# When set (tests/test_ref_pins_cpu.py, tests/golden/make_golden_ref.py: ``ref_lib.fastHadamard``), the element-wise
# finish below is executed by the reference's own compiled src/computeMI.cpp instead of the NumPy restatement.
HADAMARD_IMPL = None


def fast_hadamard(MI, den, uq, pxy, pxpy, RXY, pXrX, pYrY) -> None:
    """src/computeMI.cpp:11-21 -- linear (column-major) index over nf*nt; ``RXY`` may have
    a different shape (nt x nf) and is consumed by the same linear index (quirk Q1)."""
    if HADAMARD_IMPL is not None:
        f = [np.asfortranarray(m, dtype=np.float64) for m in (MI, den, uq, pxy, pxpy, RXY, pXrX, pYrY)]
        f[0] = np.array(MI, dtype=np.float64, order="F")  # private copy: modified in place (quirk Q11)
        HADAMARD_IMPL(*f)
        MI[...] = f[0]
        return
    mi = MI.reshape(-1, order="F")
    d = den.reshape(-1, order="F")
    u = uq.reshape(-1, order="F")
    p = pxy.reshape(-1, order="F")
    pp = pxpy.reshape(-1, order="F")
    r = RXY.reshape(-1, order="F")
    a = pXrX.reshape(-1, order="F")
    b = pYrY.reshape(-1, order="F")
    mi += u * p / d * np.log(p / (pp + r + a + b) * d)
    MI[...] = mi.reshape(MI.shape, order="F")


def compute_mi_sparse(MI, tX, tY, pX, pY, rX, rY, RXY, uqX, uqY, den) -> None:
    """R/computePairwiseMI.R:390-398."""
    pxy = tX @ tY.T + 0.5  # :391
    uq = np.outer(uqX, uqY)  # :392
    pXrX = np.outer(pX * rX, np.ones(len(pY)))  # :393
    pYrY = np.outer(np.ones(len(pX)), pY * rY)  # :394
    pxpy = np.outer(pX, pY)  # :395
    fast_hadamard(MI, den, uq, pxy, pxpy, RXY, pXrX, pYrY)  # :396


@dataclass
class BlockLinks:
    """What one call of perform_MI_computation_ACGTN produces, before tsv/rbind."""
    block: int
    # rows of MI_df in reference order (R/computePairwiseMI.R:326-331)
    pos1: np.ndarray
    pos2: np.ndarray
    clust1: np.ndarray
    clust2: np.ndarray
    len: np.ndarray
    MI: np.ndarray
    sr_mask: np.ndarray  # len <= sr_dist  (:333)
    lr_keep: np.ndarray  # indices into the LR subset kept by the quantile filter (:358)
    disc_thresh: Optional[float]
    prob: Optional[float]
    row: np.ndarray = field(default=None)  # 0-based local row (from) index
    col: np.ndarray = field(default=None)  # 0-based local col (to) index
    from_idx: np.ndarray = field(default=None)  # 0-based global SNP indices of the block
    to_idx: np.ndarray = field(default=None)


def block_mi_matrix(snp: SnpDat, hdw: np.ndarray, from_idx: np.ndarray, to_idx: np.ndarray) -> np.ndarray:
    """MI matrix (nf x nt) of one block, literal: R/computePairwiseMI.R:198-298."""
    neff = float(np.sum(hdw))  # :77
    hsq = np.sqrt(hdw)  # diag(sqrt(hdw)) :89
    fromISto = len(from_idx) == len(to_idx) and bool(np.all(from_idx == to_idx))  # :198-202
    rf = snp.r[from_idx].astype(np.float64)
    rt = rf if fromISto else snp.r[to_idx].astype(np.float64)  # :204
    uqf = snp.uqe[from_idx, :]
    uqt = uqf if fromISto else snp.uqe[to_idx, :]  # :205
    tfh, pf, tth, pt = [], [], [], []
    for a in range(5):  # :238-242
        tA = snp.snp_matrix(a)[from_idx, :]
        tAh = tA * hsq[None, :]
        tfh.append(tAh)
        pf.append((tAh ** 2).sum(axis=1))
    if fromISto:  # :244-249
        tth, pt = tfh, pf
    else:
        for a in range(5):  # :252-256
            tA = snp.snp_matrix(a)[to_idx, :]
            tAh = tA * hsq[None, :]
            tth.append(tAh)
            pt.append((tAh ** 2).sum(axis=1))
    den = neff + np.outer(snp.r[from_idx], snp.r[to_idx]) * 0.5  # :260
    rft = np.outer(rf, rt).T * 0.25  # :261  (nt x nf; quirk Q1)
    rf = 0.5 * rf  # :262
    rt = 0.5 * rt  # :263
    MI = np.zeros((len(from_idx), len(to_idx)), dtype=np.float64)  # :268
    for a in range(5):  # :270-298, a-major / b-minor
        for b in range(5):
            compute_mi_sparse(MI, tfh[a], tth[b], pf[a], pt[b], rf, rt, rft, uqf[:, a], uqt[:, b], den)
    return MI


def sr_only_keep(POS_f: np.ndarray, POS_t: np.ndarray, g: float, sr_dist: float):
    """R/computePairwiseMI.R:182-183 (quirk Q12)."""
    kp_f = np.array([np.any(np.abs(0.5 * g - np.abs(np.mod(POS_t - x, g) - 0.5 * g)) < sr_dist) for x in POS_f])
    kp_t = np.array([np.any(np.abs(0.5 * g - np.abs(np.mod(POS_f - x, g) - 0.5 * g)) < sr_dist) for x in POS_t])
    return kp_f, kp_t


def perform_MI_computation_ACGTN(snp: SnpDat, hdw: np.ndarray, paint: np.ndarray, from_idx: np.ndarray,
                                 to_idx: np.ndarray, sr_dist: float, lr_retain_links: float,
                                 lr_links_approx: Optional[float], perform_SR_analysis_only: bool = False,
                                 block: int = 0) -> BlockLinks:
    """R/computePairwiseMI.R:167-386 for one block; indices are 0-based global SNP ids."""
    g = float(snp.g)
    POS_f = snp.POS[from_idx].astype(np.float64)  # :176
    POS_t = snp.POS[to_idx].astype(np.float64)
    if perform_SR_analysis_only:  # :179-189
        kp_f, kp_t = sr_only_keep(POS_f, POS_t, g, sr_dist)
        from_idx = from_idx[kp_f]
        to_idx = to_idx[kp_t]
        POS_f = snp.POS[from_idx].astype(np.float64)
        POS_t = snp.POS[to_idx].astype(np.float64)
    paint_f = paint[from_idx]  # :194
    paint_t = paint[to_idx]
    fromISto = len(from_idx) == len(to_idx) and bool(np.all(from_idx == to_idx))
    MI = block_mi_matrix(snp, hdw, from_idx, to_idx)
    nf, nt = MI.shape
    rr, cc = np.meshgrid(np.arange(nf), np.arange(nt), indexing="ij")
    if fromISto:  # :307  which(lower.tri(t(MI))) -> row > col, column-major order
        m = (rr > cc)
        order = np.argsort((cc[m] * nf + rr[m]), kind="stable")
        row = rr[m][order]
        col = cc[m][order]
    else:  # :309  upper.tri rows (column-major) then lower.tri rows (column-major); diag dropped (Q2)
        mu = rr < cc
        ou = np.argsort(cc[mu] * nf + rr[mu], kind="stable")
        ml = rr > cc
        ol = np.argsort(cc[ml] * nf + rr[ml], kind="stable")
        row = np.concatenate([rr[mu][ou], rr[ml][ol]])
        col = np.concatenate([cc[mu][ou], cc[ml][ol]])
    pos2 = POS_f[row]  # :319
    pos1 = POS_t[col]  # :320
    clust2 = paint_f[row]  # :322
    clust1 = paint_t[col]  # :323
    ln = circ_len(pos1, pos2, g)  # :330
    mi = MI[row, col]  # :331
    sr_mask = ln <= sr_dist  # :333
    lr_keep = np.zeros(0, dtype=np.int64)
    disc = None
    prob = None
    lr_idx = np.nonzero(~sr_mask)[0]
    if len(lr_idx) > 0 and not perform_SR_analysis_only:  # :347
        n_lr = len(lr_idx)
        prob = max(0.0, 1 - ((lr_retain_links * (n_lr / lr_links_approx)) / n_lr))  # :352
        disc = quantile_type7(mi[lr_idx], prob)  # :354
        lr_keep = lr_idx[mi[lr_idx] >= disc]  # :358
    return BlockLinks(block=block, pos1=pos1, pos2=pos2, clust1=clust1, clust2=clust2, len=ln, MI=mi,
                      sr_mask=sr_mask, lr_keep=lr_keep, disc_thresh=disc, prob=prob, row=row, col=col,
                      from_idx=from_idx, to_idx=to_idx)


# --------------------------------------------------------------------------------------
# a6 driver  perform_MI_computation (scan + filter part, R/computePairwiseMI.R:46-116)
# --------------------------------------------------------------------------------------
@dataclass
class ScanResult:
    """The two artefacts of the scan: LR rows as appended to lr_links.tsv
    (R/computePairwiseMI.R:362), SR rows as appended to the per-cluster lists (:372-376),
    kept here as one list in block order plus the cluster routing."""
    lr: Dict[str, np.ndarray]
    sr: Dict[str, np.ndarray]
    sr_by_cluster: List[np.ndarray]  # row indices into ``sr`` for clusters 1..nclust
    blocks: List[BlockLinks]
    lr_links_approx: Optional[float]
    neff: float


def perform_MI_scan(snp: SnpDat, hdw: np.ndarray, paint: np.ndarray, nclust: int, sr_dist: float = 20000,
                    lr_retain_links: float = 1e6, max_blk_sz: float = 10000,
                    perform_SR_analysis_only: bool = False, lr_links_approx: Optional[float] = None,
                    keep_blocks: bool = False) -> ScanResult:
    blk = r_round_to_thousands(max_blk_sz)  # :69
    blocks = make_blocks(snp.nsnp, blk)  # :70
    if not perform_SR_analysis_only and lr_links_approx is None:
        lr_links_approx = lr_links_approx_reference(snp.POS, float(snp.g), sr_dist)  # :94-97
    cols = ["pos1", "pos2", "clust1", "clust2", "len", "MI"]
    lr = {c: [] for c in cols}
    sr = {c: [] for c in cols}
    kept: List[BlockLinks] = []
    for bi, (fs, fe, ts, te) in enumerate(blocks):  # :103-116
        from_idx = np.arange(fs - 1, fe)
        to_idx = np.arange(ts - 1, te)
        bl = perform_MI_computation_ACGTN(snp, hdw, paint, from_idx, to_idx, sr_dist, lr_retain_links,
                                          lr_links_approx, perform_SR_analysis_only, block=bi)
        for c in cols:
            v = getattr(bl, c)
            lr[c].append(v[bl.lr_keep])
            sr[c].append(v[bl.sr_mask])
        if keep_blocks:
            kept.append(bl)
    lr = {c: np.concatenate(v) if v else np.zeros(0) for c, v in lr.items()}
    sr = {c: np.concatenate(v) if v else np.zeros(0) for c, v in sr.items()}
    by_cluster = []
    for c in range(1, nclust + 1):  # :372-376 + src/computeMI.cpp:25-41 (compareToRow)
        by_cluster.append(np.nonzero((sr["clust1"] == c) | (sr["clust2"] == c))[0])
    return ScanResult(lr=lr, sr=sr, sr_by_cluster=by_cluster, blocks=kept,
                      lr_links_approx=lr_links_approx, neff=float(np.sum(hdw)))


# --------------------------------------------------------------------------------------
# Closed form of SURVEY.md section 8a (cross-check of the literal form; also documents Q1)
# --------------------------------------------------------------------------------------
def block_mi_closed_form(snp: SnpDat, hdw: np.ndarray, from_idx: np.ndarray, to_idx: np.ndarray) -> np.ndarray:
    w = hdw.astype(np.float64)
    neff = w.sum()
    nf, nt = len(from_idx), len(to_idx)
    fromISto = nf == nt and bool(np.all(from_idx == to_idx))
    Xf = [(snp.codes[from_idx] == a).astype(np.float64) for a in range(5)]
    Xt = [(snp.codes[to_idx] == a).astype(np.float64) for a in range(5)]
    pf = [x @ w for x in Xf]
    pt = [x @ w for x in Xt]
    rf = snp.r[from_idx].astype(np.float64)
    rt = snp.r[to_idx].astype(np.float64)
    den = neff + 0.5 * np.outer(rf, rt)
    if fromISto:
        Q = 0.25 * np.outer(rf, rt)
    else:
        c = np.arange(nf)[:, None] + np.arange(nt)[None, :] * nf  # linear index il + jl*nf
        Q = 0.25 * rf[c // nt] * rt[c % nt]
    MI = np.zeros((nf, nt))
    for a in range(5):
        for b in range(5):
            cab = (Xf[a] * w[None, :]) @ Xt[b].T + 0.5
            D = np.outer(pf[a], pt[b]) + 0.5 * (pf[a] * rf)[:, None] + 0.5 * (pt[b] * rt)[None, :] + Q
            u = np.outer(snp.uqe[from_idx, a], snp.uqe[to_idx, b])
            MI += u * cab / den * np.log(cab * den / D)
    return MI


def pair_mi_closed_form(snp: SnpDat, hdw: np.ndarray, from_idx: np.ndarray, to_idx: np.ndarray, il: np.ndarray,
                        jl: np.ndarray) -> np.ndarray:
    """MI of selected cells (il[k], jl[k]) of the block (from_idx, to_idx): the same closed form as
    block_mi_closed_form (R/computePairwiseMI.R:238-298, src/computeMI.cpp:19, quirk Q1 by
    the linear index il + jl*nf), evaluated per pair so that full-size blocks can be spot-checked without forming
    the nf x nt matrices."""
    w = hdw.astype(np.float64)
    neff = w.sum()
    from_idx = np.asarray(from_idx); to_idx = np.asarray(to_idx)
    il = np.asarray(il, dtype=np.int64); jl = np.asarray(jl, dtype=np.int64)
    nf, nt = len(from_idx), len(to_idx)
    fromISto = nf == nt and bool(np.all(from_idx == to_idx))
    gi, gj = from_idx[il], to_idx[jl]
    ci, cj = snp.codes[gi], snp.codes[gj]                 # [m, S]
    rf = snp.r[from_idx].astype(np.float64)
    rt = snp.r[to_idx].astype(np.float64)
    ri, rj = rf[il], rt[jl]
    den = neff + 0.5 * ri * rj
    if fromISto:
        Q = 0.25 * ri * rj
    else:
        c = il + jl * nf
        Q = 0.25 * rf[c // nt] * rt[c % nt]
    out = np.zeros(len(il))
    for a in range(5):
        xa = (ci == a)
        pa = xa @ w
        for b in range(5):
            yb = (cj == b)
            pb = yb @ w
            cab = (xa & yb) @ w + 0.5
            D = pa * pb + 0.5 * pa * ri + 0.5 * pb * rj + Q
            u = snp.uqe[gi, a] * snp.uqe[gj, b]
            out += u * cab / den * np.log(cab * den / D)
    return out


def format_r_numeric(x: float) -> str:
    """One double as ``write.table`` encodes a cell (R/computePairwiseMI.R:140,362 -> utils:::writetable ->
    EncodeElement0 -> formatReal with R_print.digits = DBL_DIG = 15, scipen = 0; base R, not in the reference tree):
    the fewest significant digits (<= 15) that reproduce the 15-digit rounding; fixed notation unless scientific
    notation is strictly narrower -- so 20000 stays "20000" but 100000 becomes "1e+05", and MI values below 1e-4 with
    many digits come out as d.ddde-05.  Restated from the published algorithm; not checked against a live R here."""
    x = float(x)
    if x != x:
        return "NA"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0:
        return "0"
    mant, ex = f"{x:.14e}".split("e")
    neg = 1 if mant.startswith("-") else 0
    digits = mant.lstrip("-").replace(".", "").rstrip("0") or "0"
    nsig, kp = len(digits), int(ex)
    if kp >= 0:
        left, rgt = kp + 1, max(0, nsig - kp - 1)
        if 0 < kp <= 22 and abs(x) < 10.0 ** kp:   # formatReal's `roundingwidens`
            left -= 1
    else:
        left, rgt = 1, nsig - kp - 1
    w_fixed = neg + left + (rgt + 1 if rgt else 0)
    w_sci = neg + (nsig + 1 if nsig > 1 else 1) + (5 if abs(kp) >= 100 else 4)
    return f"{x:.{rgt}f}" if w_fixed <= w_sci else f"{x:.{nsig - 1}e}"
