"""Shared test helpers."""
import numpy as np


def compare_lr_sets(ref_p1, ref_p2, ref_mi, got_p1, got_p2, got_mi, thr, tol_thr=1e-9, tol_mi=1e-12):
    """Compare two retained long-range link sets of ONE block.

    LR membership is ``MI >= quantile(...)`` (R/computePairwiseMI.R:354-358) and real data holds
    many links with mathematically identical MI (duplicate SNP patterns), so links whose MI lies
    within ``tol_thr`` of the threshold are decided by last-ulp summation noise in *any*
    implementation (including R/MatrixExtra itself).  Those are the "threshold-borderline pairs"
    of the north star: they are returned explicitly, everything else must match exactly.
    Returns (n_common, borderline_only_ref, borderline_only_got).
    """
    ref = {(int(a), int(b)): m for a, b, m in zip(ref_p1, ref_p2, ref_mi)}
    got = {(int(a), int(b)): m for a, b, m in zip(got_p1, got_p2, got_mi)}
    only_ref = [k for k in ref if k not in got]
    only_got = [k for k in got if k not in ref]
    for k in only_ref:
        assert abs(ref[k] - thr) <= tol_thr, f"link {k} MI={ref[k]} missing and not borderline (thr={thr})"
    for k in only_got:
        assert abs(got[k] - thr) <= tol_thr, f"link {k} MI={got[k]} extra and not borderline (thr={thr})"
    common = [k for k in ref if k in got]
    if common:
        d = max(abs(ref[k] - got[k]) for k in common)
        assert d <= tol_mi, f"MI mismatch {d}"
    return len(common), only_ref, only_got
