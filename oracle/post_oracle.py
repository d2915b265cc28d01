"""TEST INFRASTRUCTURE -- CPU restatement (NumPy/SciPy) of the post-scan steps of perform_MI_computation:
`mergeNsort_sr_links` (R/computePairwiseMI.R:400-495) and `runARACNE` (R/io_functions.R:101-164 with
.compareToRow / .vecPosMatch / .compareTriplet, src/computeMI.cpp:25-79, and .fast_intersect, src/fintersect.cpp:6-33).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path is the
native code behind include/ldw.h (ldw_sr_postprocess, ldw_run_aracne, ldw_write_sr_tsv).

PARITY UNPINNED: the reference holds no expected values for these steps and R is not installed here.  Arithmetic that
lives in un-vendored CRAN dependencies is restated from their published algorithms:
  * dplyr::group_by(len) + stats::quantile(MI, 0.95)   -> ascending distinct `len`, quantile type 7 (base R
    stats::quantile.default), the same restatement as quirk Q3 of the scan (ldw_oracle.quantile_type7);
  * RcppArmadillo::fastLm(cbind(log(len), 1), log(max)) -> ordinary least squares (arma::solve -> LAPACK dgels, QR);
    here numpy.linalg.lstsq; agreement between solvers is ~1e-13 relative, not bitwise;
  * fitdistrplus::fitdist(x, "beta") (version unpinned, DESCRIPTION Imports) -> maximum likelihood: start values
    shape1 = m*aux, shape2 = (1-m)*aux with m = mean(x), v = (n-1)/n*var(x), aux = m(1-m)/v - 1 (fitdistrplus'
    default start for "beta"), objective -sum(dbeta(x, a, b, log = TRUE)), minimised by stats::optim(method =
    "Nelder-Mead") with optim's defaults (reltol = sqrt(.Machine$double.eps), maxit = 500, alpha 1, beta 0.5, gamma 2);
    `nmmin` below restates R's src/appl/optim.c:nmmin (Nash's Algorithm 19) step by step, so the simplex path -- and
    with it the returned vertex -- is the one R takes as long as no two objective values tie within rounding;
  * stats::pbeta(x, a, b, lower.tail = FALSE, log.p = TRUE) -> log of the regularised incomplete beta complement
    (scipy.special.betaincc; R uses TOMS 708 bratio: same function, ~1e-14 relative).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import special

from ldw_oracle import quantile_type7

SR_COLS = ["pos1", "pos2", "clust1", "clust2", "len", "MI"]


# --------------------------------------------------------------------------------------------------------------------
# stats::optim(method = "Nelder-Mead")  (R: src/appl/optim.c nmmin; called from src/library/stats/src/optim.c)
# --------------------------------------------------------------------------------------------------------------------
def nmmin(fn, start: Sequence[float], abstol: float = -np.inf, intol: float = float(np.sqrt(np.finfo(float).eps)),
          alpha: float = 1.0, bet: float = 0.5, gamm: float = 2.0, maxit: int = 500):
    """Returns (x, fmin, fncount, fail).  Variable names follow nmmin: P is the (n+1) x (n+2) polytope, column C = n+2
    the centroid, L / H the lowest / highest vertex (1-based), row n+1 the function values."""
    big = 1.0e35
    n = len(start)
    B = np.array(start, dtype=np.float64)
    P = np.zeros((n + 1, n + 2))
    fail = 0
    f = fn(B)
    if not np.isfinite(f):
        raise ValueError("function cannot be evaluated at initial parameters")
    funcount = 1
    convtol = intol * (abs(f) + intol)
    n1, C = n + 1, n + 2
    P[n1 - 1, 0] = f
    P[:n, 0] = B
    L = 1
    size = 0.0
    step = 0.0
    for i in range(n):
        if 0.1 * abs(B[i]) > step:
            step = 0.1 * abs(B[i])
    if step == 0.0:
        step = 0.1
    for j in range(2, n1 + 1):
        P[:n, j - 1] = B
        trystep = step
        while P[j - 2, j - 1] == B[j - 2]:
            P[j - 2, j - 1] = B[j - 2] + trystep
            trystep *= 10
        size += trystep
    oldsize = size
    calcvert = True
    while True:
        if calcvert:
            for j in range(n1):
                if j + 1 != L:
                    B = P[:n, j].copy()
                    f = fn(B)
                    if not np.isfinite(f):
                        f = big
                    funcount += 1
                    P[n1 - 1, j] = f
            calcvert = False
        VL = P[n1 - 1, L - 1]
        VH = VL
        H = L
        for j in range(1, n1 + 1):
            if j != L:
                f = P[n1 - 1, j - 1]
                if f < VL:
                    L, VL = j, f
                if f > VH:
                    H, VH = j, f
        if VH <= VL + convtol or VL <= abstol:
            break
        for i in range(n):
            temp = -P[i, H - 1]
            for j in range(n1):
                temp += P[i, j]
            P[i, C - 1] = temp / n
        B = (1.0 + alpha) * P[:n, C - 1] - alpha * P[:n, H - 1]
        f = fn(B)
        if not np.isfinite(f):
            f = big
        funcount += 1
        VR = f
        if VR < VL:
            P[n1 - 1, C - 1] = f
            for i in range(n):
                f = gamm * B[i] + (1 - gamm) * P[i, C - 1]
                P[i, C - 1] = B[i]
                B[i] = f
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < VR:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            else:
                P[:n, H - 1] = P[:n, C - 1]
                P[n1 - 1, H - 1] = VR
        else:
            if VR < VH:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = VR
            B = (1 - bet) * P[:n, H - 1] + bet * P[:n, C - 1]
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < P[n1 - 1, H - 1]:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            elif VR >= VH:
                calcvert = True
                size = 0.0
                for j in range(n1):
                    if j + 1 != L:
                        for i in range(n):
                            P[i, j] = bet * (P[i, j] - P[i, L - 1]) + P[i, L - 1]
                            size += abs(P[i, j] - P[i, L - 1])
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if funcount > maxit:
            break
    if funcount > maxit:
        fail = 1
    return P[:n, L - 1].copy(), float(P[n1 - 1, L - 1]), funcount, fail


def beta_start(x: np.ndarray) -> Tuple[float, float]:
    """fitdistrplus' default start values for "beta" (method-of-moments with the biased variance)."""
    if np.any(x < 0) or np.any(x > 1):
        raise ValueError("values must be in [0-1] to fit a beta distribution")
    n = len(x)
    m = float(np.mean(x))
    v = (n - 1) / n * float(np.var(x, ddof=1))
    aux = m * (1 - m) / v - 1
    return m * aux, (1 - m) * aux


def beta_nll(x: np.ndarray):
    """-sum(dbeta(x, a, b, log = TRUE)) as a function of (a, b); NaN (-> `big` in nmmin) outside a, b > 0, as dbeta."""
    lx, l1x = np.log(x), np.log1p(-x)

    def f(par):
        a, b = float(par[0]), float(par[1])
        if not (a > 0 and b > 0):
            return np.nan
        return -float(np.sum((a - 1) * lx + (b - 1) * l1x - special.betaln(a, b)))
    return f


def fit_beta_mle(x: np.ndarray):
    s = beta_start(x)
    par, fmin, cnt, fail = nmmin(beta_nll(x), s)
    return par, fmin, cnt, fail, s


def neg_log_pbeta_upper(x: np.ndarray, a: float, b: float) -> np.ndarray:
    """-pbeta(x, a, b, lower.tail = FALSE, log.p = TRUE)   (R/computePairwiseMI.R:453)."""
    with np.errstate(divide="ignore"):
        return -np.log(special.betaincc(a, b, x))


# --------------------------------------------------------------------------------------------------------------------
# mergeNsort_sr_links  (R/computePairwiseMI.R:400-495)
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class ClusterFit:
    len: np.ndarray      # maxvls$len  (ascending distinct lengths, :422)
    max: np.ndarray      # maxvls$max  (95th percentile of MI per length)
    fit: np.ndarray      # maxvls$fit = exp(fitted(fastLm))  (:428-429)
    coef: np.ndarray     # slope, intercept of log(max) ~ log(len)
    shape: np.ndarray    # beta MLE (:452)
    start: Tuple[float, float]
    n_pos: int
    nm_evals: int
    nm_fail: int


@dataclass
class SrPost:
    df: Dict[str, np.ndarray]       # sr_links_df: clust_c, row (index into the SR table), srp_max (+ the six link columns)
    red: np.ndarray                 # row indices into df: sr_links_red (:494)
    chk: np.ndarray                 # row indices into df: sr_links_ARACNE_check (:495)
    fits: List[ClusterFit]


def merge_n_sort_sr_links(sr: Dict[str, np.ndarray], nclust: int, sr_dist: float, srp_cutoff: float) -> SrPost:
    pos1, pos2 = np.asarray(sr["pos1"]), np.asarray(sr["pos2"])
    c1, c2 = np.asarray(sr["clust1"]), np.asarray(sr["clust2"])
    ln, MI = np.asarray(sr["len"], dtype=np.float64), np.asarray(sr["MI"], dtype=np.float64)
    df_rows, df_c, df_srp = [], [], []
    dup_rows, dup_c, dup_srp = [], [], []
    fits = []
    for c in range(1, nclust + 1):
        rows = np.nonzero((c1 == c) | (c2 == c))[0]                       # :372-376 (cluster list c, in scan order)
        rows = rows[~np.isnan(ln[rows])]                                  # :417
        rows = rows[ln[rows] < sr_dist]                                   # :418
        rows = rows[ln[rows] > 0]                                         # :419
        l, m = ln[rows], MI[rows]
        o = np.argsort(l, kind="stable")                                  # group_by(len): ascending distinct lengths
        ulen, first = np.unique(l[o], return_index=True)
        bounds = np.append(first, len(l))
        ms = m[o]
        q95 = np.array([quantile_type7(ms[bounds[k]:bounds[k + 1]], 0.95) for k in range(len(ulen))])   # :422
        X = np.stack([np.log(ulen), np.ones(len(ulen))], axis=1)          # :428
        coef = np.linalg.lstsq(X, np.log(q95), rcond=None)[0]
        mean_dist = np.exp(X @ coef)                                      # :429
        # :448 -- `mean_dist[sr_links_t$len]` indexes the fitted values by the VALUE of len (1-based position in
        # maxvls), not by the group of that length; out of range gives NA and the link is dropped by which(NA > 0)
        idx = l.astype(np.int64)                                          # R truncates a double subscript
        md = np.full(len(l), np.nan)
        ok = (idx >= 1) & (idx <= len(mean_dist))
        md[ok] = mean_dist[idx[ok] - 1]
        diff = m - md
        with np.errstate(invalid="ignore"):
            sel = np.nonzero(diff > 0)[0]                                 # :449
        x = diff[sel]
        par, fmin, cnt, fail, start = fit_beta_mle(x)                     # :452
        srp = neg_log_pbeta_upper(x, par[0], par[1])                      # :453
        keep = ~np.isnan(srp)                                             # :458 (is.na is TRUE for NaN too)
        sel, srp = sel[keep], srp[keep]
        r = rows[sel]
        dup = c1[r] != c2[r]                                              # :460
        df_rows.append(r[~dup]); df_c.append(np.full(int((~dup).sum()), c)); df_srp.append(srp[~dup])
        dup_rows.append(r[dup]); dup_c.append(np.full(int(dup.sum()), c)); dup_srp.append(srp[dup])
        fits.append(ClusterFit(ulen, q95, mean_dist, coef, par, start, len(x), cnt, fail))
    cat = lambda v, dt: np.concatenate(v).astype(dt) if v else np.zeros(0, dt)
    rows, cc, srp = cat(df_rows, np.int64), cat(df_c, np.int32), cat(df_srp, np.float64)
    drows, dcc, dsrp = cat(dup_rows, np.int64), cat(dup_c, np.int32), cat(dup_srp, np.float64)
    if len(drows):
        # :474-483 -- data.table `by = keys` (the six link columns): groups in order of first appearance,
        # .I[which.max(srp_max)] = first row holding the group's maximum
        first: Dict[tuple, int] = {}
        best: Dict[tuple, int] = {}
        for k in range(len(drows)):
            r = drows[k]
            key = (pos1[r], pos2[r], c1[r], c2[r], ln[r], MI[r])
            if key not in first:
                first[key] = k
                best[key] = k
            elif dsrp[k] > dsrp[best[key]]:
                best[key] = k
        pick = np.array([best[key] for key in first], dtype=np.int64)
        rows, cc, srp = np.concatenate([rows, drows[pick]]), np.concatenate([cc, dcc[pick]]), np.concatenate([srp, dsrp[pick]])
    red = np.nonzero(srp > srp_cutoff)[0]                                 # :494
    if len(red):
        chk = np.nonzero(MI[rows] >= MI[rows[red]].min())[0]              # :495
    else:
        chk = np.zeros(0, np.int64)                                       # min(numeric(0)) = Inf
    df = {"clust_c": cc, "row": rows, "srp_max": srp}
    for k in SR_COLS:
        df[k] = np.asarray(sr[k])[rows]
    return SrPost(df=df, red=red, chk=chk, fits=fits)


# --------------------------------------------------------------------------------------------------------------------
# runARACNE  (R/io_functions.R:101-164), literal
# --------------------------------------------------------------------------------------------------------------------
def compare_to_row(x: np.ndarray, y: float) -> np.ndarray:
    """src/computeMI.cpp:25-41 with a scalar y: rows of x holding y in any column."""
    return np.any(x == y, axis=1)


def fast_intersect(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """src/fintersect.cpp:6-33: sort both (as int), merge; duplicates pair off one to one."""
    Av, Bv = np.sort(A.astype(np.int64)), np.sort(B.astype(np.int64))
    out = []
    i = j = 0
    while i < len(Av) and j < len(Bv):
        if Av[i] < Bv[j]:
            i += 1
        elif Av[i] > Bv[j]:
            j += 1
        else:
            out.append(Av[i]); i += 1; j += 1
    return np.array(out, dtype=np.int64)


def vec_pos_match(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """src/computeMI.cpp:44-58: 1-based position of the first y equal to each x (0 when absent)."""
    ret = np.zeros(len(x), dtype=np.int64)
    for ii, v in enumerate(x):
        w = np.nonzero(y == v)[0]
        if len(w):
            ret[ii] = w[0] + 1
    return ret


def compare_triplet(MI0X: np.ndarray, MI0Z: np.ndarray, MI0: float) -> bool:
    """src/computeMI.cpp:62-78."""
    for a, b in zip(MI0X, MI0Z):
        if MI0 < a and MI0 < b:
            return False
    return True


def run_aracne(chk_pos1, chk_pos2, chk_MI, full_pos1, full_pos2, full_MI) -> np.ndarray:
    pos_mat = np.stack([np.asarray(full_pos1, dtype=np.float64), np.asarray(full_pos2, dtype=np.float64)], axis=1)
    MIs = np.asarray(full_MI, dtype=np.float64)
    nlinks = len(chk_pos1)
    out = np.ones(nlinks, dtype=bool)
    pX_ = 0
    idX = matX = None
    for i in range(nlinks):
        pX, pZ = float(chk_pos1[i]), float(chk_pos2[i])
        if pX != pX_:
            idX = np.nonzero(compare_to_row(pos_mat, pX))[0]
            matX = pos_mat[idX].reshape(-1)                   # c(rbind(col1, col2)): interleaved
            matX = matX[matX != pX]
            pX_ = pX
        idZ = np.nonzero(compare_to_row(pos_mat, pZ))[0]
        matZ = pos_mat[idZ].reshape(-1)
        matZ = matZ[matZ != pZ]
        com = fast_intersect(matX, matZ)
        if len(com) > 0:
            MI0X = MIs[idX[vec_pos_match(com, matX) - 1]]
            MI0Z = MIs[idZ[vec_pos_match(com, matZ) - 1]]
            out[i] = compare_triplet(MI0X, MI0Z, float(chk_MI[i]))
    return out


def order_links_by_srp(srp_max: np.ndarray) -> np.ndarray:
    """order(srp_max, decreasing = TRUE): R's default radix method is stable, ties keep their original order
    (R/computePairwiseMI.R:134)."""
    return np.argsort(-np.asarray(srp_max), kind="stable")


# --------------------------------------------------------------------------------------------------------------------
# analyse_long_range_links, numerical part  (R/lr_analyser.R:72-116)
# --------------------------------------------------------------------------------------------------------------------
def analyse_long_range_links(lr: Dict[str, np.ndarray], sr: Dict[str, np.ndarray], are_lrlinks_ordered: bool = False):
    mi = np.asarray(lr["MI"], dtype=np.float64)
    q13 = np.array([quantile_type7(mi, 0.25), quantile_type7(mi, 0.75)])          # :73
    iqr = q13[1] - q13[0]                                                            # :74
    thresholds = q13[1] + np.array([1.5, 3.0]) * iqr                                 # :75
    red = mi > thresholds.min()                                                      # :91
    if red.sum() < 5000 and len(mi) >= 5000:                                         # :94-99
        thresholds = np.array([quantile_type7(mi, 1 - (1 / len(mi) * k)) for k in (4000, 5000)])
        red = mi > thresholds.min()
    p1 = np.concatenate([np.asarray(lr["pos1"], float), np.asarray(sr["pos1"], float)])   # :105-106
    p2 = np.concatenate([np.asarray(lr["pos2"], float), np.asarray(sr["pos2"], float)])
    mm = np.concatenate([mi, np.asarray(sr["MI"], float)])
    k = mm > thresholds.min()                                                        # :107
    ar = run_aracne(np.asarray(lr["pos1"], float)[red], np.asarray(lr["pos2"], float)[red], mi[red], p1[k], p2[k], mm[k])  # :108
    idx = np.nonzero(red)[0]
    if not are_lrlinks_ordered:                                                      # :113-116
        o = np.argsort(-mi[idx], kind="stable")
        idx, ar = idx[o], ar[o]
    return idx, ar, thresholds
