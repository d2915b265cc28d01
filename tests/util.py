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


def compare_sr_post_with_tolerance(red, post, ref, srp_cutoff, mi_tol):
    """Native mergeNsort_sr_links / runARACNE output computed from MI values that carry the scan's tolerance (`mi_tol`,
    1e-6 absolute) against the oracle chain on exact fp64 MI.  srp_max is a statistic of ALL short-range MI values
    (per-length percentiles, a log-log decay fit, a beta fit whose likelihood weighs residuals near zero by their
    logarithm), so the MI tolerance shows up amplified: +-4e-7 on MI moves the beta shapes by ~1e-3 relative and srp_max
    by up to ~2e-2.  `red` = sr_links_red columns, `post` = api.SrLinks, `ref` = post_oracle.SrPost."""
    import post_oracle as PO
    for h, f in zip(post.fits, ref.fits):
        np.testing.assert_array_equal(h["len"], f.len)
        assert np.abs(h["max"] - f.max).max() < mi_tol              # percentiles of MI values that agree to mi_tol
        assert np.allclose(h["fit"], f.fit, rtol=100 * mi_tol / f.max.min() if f.max.min() > 0 else 1e-2, atol=10 * mi_tol)
        assert np.allclose(h["shape"], f.shape, rtol=5e-2)
    ref_srp = dict(zip(ref.df["row"].tolist(), ref.df["srp_max"].tolist()))
    got_srp = dict(zip(post.df["row"].tolist(), post.df["srp_max"].tolist()))
    common = [k for k in got_srp if k in ref_srp]
    assert len(common) > 0.995 * max(len(ref_srp), len(got_srp))      # residuals within mi_tol of zero may change side
    d = np.array([got_srp[k] - ref_srp[k] for k in common])
    s = np.array([ref_srp[k] for k in common])
    print("srp_max: max abs diff", np.abs(d).max(), "over", len(common), "links")
    assert np.all(np.abs(d) <= 0.1 + 0.1 * s)
    got_red, ref_red = set(red["row"].tolist()), set(ref.df["row"][ref.red].tolist())
    for k in got_red ^ ref_red:                                       # only links at the cut-off may differ
        assert abs(got_srp.get(k, ref_srp.get(k)) - srp_cutoff) < 0.1 + 0.1 * srp_cutoff
    assert len(got_red ^ ref_red) <= 0.02 * len(ref_red) + 2
    assert np.all(np.diff(red["srp_max"]) <= 0)                       # order_links = T
    # ARACNE of the returned links against the literal restatement on the same inputs
    d_, chk = post.df, post.chk
    ar = PO.run_aracne(red["pos1"], red["pos2"], red["MI"], d_["pos1"][chk], d_["pos2"][chk], d_["MI"][chk])
    np.testing.assert_array_equal(red["ARACNE"], ar.astype(float))
