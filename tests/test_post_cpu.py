"""Post-scan steps of perform_MI_computation, native host code against the oracle:
mergeNsort_sr_links (R/computePairwiseMI.R:400-495), runARACNE (R/io_functions.R:101-164), sr_links.tsv
(R/computePairwiseMI.R:140).  No device needed.  What pins the restatements, there being no R here and no expected
values in the reference: R's own printed output of optim() on Rosenbrock's function (?optim), closed forms of the
regularised incomplete beta function, and a literal second implementation (oracle/post_oracle.py)."""
import ctypes as C
import math

import numpy as np
import pytest
from scipy import special

import ldw_oracle as O
import post_oracle as PO
import ldweaver_b200 as ldw
from ldweaver_b200 import _lib


def _fixture_sr(e):
    """Short-range table of the bundled fixture (g = 50 000, max_blk_sz = 1000) from the golden scan output."""
    POS, paint = e["relaxed_POS"], e["paint"]
    p1, p2, MI = e["g50k_b1000_sr_pos1"], e["g50k_b1000_sr_pos2"], e["g50k_b1000_sr_MI"]
    lut = np.zeros(int(POS.max()) + 1, dtype=np.int32)
    lut[POS] = paint
    ln = O.circ_len(p1.astype(float), p2.astype(float), 50000.0)
    return dict(pos1=p1, pos2=p2, clust1=lut[p1], clust2=lut[p2], len=ln, MI=MI), paint


# ---------------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------------
def test_nelder_mead_reproduces_r_optim_example():
    """?optim:  fr <- function(x) 100 * (x[2] - x[1]^2)^2 + (1 - x[1])^2;  optim(c(-1.2, 1), fr)  prints
    $par 1.000260 1.000506, $value 8.825241e-08, $counts function 195."""
    fr = lambda x: 100 * (x[1] - x[0] * x[0]) ** 2 + (1 - x[0]) ** 2
    par, val, cnt, fail = PO.nmmin(fr, [-1.2, 1.0])
    assert cnt == 195 and fail == 0
    assert f"{par[0]:.6f} {par[1]:.6f}" == "1.000260 1.000506" and f"{val:.6e}" == "8.825241e-08"
    start = np.array([-1.2, 1.0])
    out, v, c = np.zeros(2), C.c_double(), C.c_int()
    _lib.check(_lib.lib().ldw_nm_rosenbrock(_lib.ptr(start), _lib.ptr(out), C.byref(v), C.byref(c)))
    assert c.value == 195 and np.array_equal(out, par) and v.value == val  # same path, bit for bit


def test_log_upper_beta_tail_closed_forms_and_scipy():
    L = _lib.lib()

    def native(x, a, b):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros_like(x)
        _lib.check(L.ldw_neg_log_pbeta_upper(_lib.ptr(x), len(x), a, b, _lib.ptr(out)))
        return out
    # pbeta(0.5, 2, 3) = 11/16;  shape (1, b): upper tail (1 - x)^b;  shape (a, 1): 1 - x^a
    assert abs(native([0.5], 2, 3)[0] + math.log(1 - 11 / 16)) < 1e-14
    x = np.array([1e-12, 1e-6, 0.01, 0.3, 0.5, 0.9, 0.999, 1 - 1e-9])
    assert np.allclose(native(x, 1.0, 7.5), -7.5 * np.log1p(-x), rtol=1e-13, atol=1e-300)
    assert np.allclose(native(x, 1.0, 250.0), -250.0 * np.log1p(-x), rtol=1e-13, atol=1e-300)  # up to 5180: far past exp() underflow
    exact = np.where(x < 0.5, -np.log1p(-x ** 3.25), -np.log(-np.expm1(3.25 * np.log(x))))  # each form where it is well conditioned
    assert np.allclose(native(x, 3.25, 1.0), exact, rtol=1e-12, atol=0)
    # shapes of the size the fits produce, deep tails included (srp values of several hundred)
    rng = np.random.default_rng(5)
    for a, b in ((0.95, 23.4), (0.98, 32.4), (0.6, 250.0), (2.5, 9.0), (0.4, 0.7)):
        xs = np.concatenate([rng.uniform(0, 1, 2000), 10.0 ** rng.uniform(-9, 0, 2000), 1 - 10.0 ** rng.uniform(-9, 0, 500)])
        xs = xs[(xs > 0) & (xs < 1)]
        with np.errstate(divide="ignore"):
            ref = -np.log(special.betaincc(a, b, xs))
        ok = np.isfinite(ref) & (ref > 1e-300) & (ref < 690)   # betaincc itself degrades in the subnormal range
        got = native(xs, a, b)
        assert np.allclose(got[ok], ref[ok], rtol=2e-11, atol=1e-15), (a, b, np.abs(got[ok] / ref[ok] - 1).max())
    assert native([0.0, 1.0], 2, 2).tolist() == [0.0, math.inf]


# ---------------------------------------------------------------------------------------------------------------------
# mergeNsort_sr_links
# ---------------------------------------------------------------------------------------------------------------------
def _compare_post(got, ref):
    assert np.array_equal(got.df["row"], ref.df["row"]) and np.array_equal(got.df["clust_c"], ref.df["clust_c"])
    for h, f in zip(got.fits, ref.fits):
        assert np.array_equal(h["len"], f.len)
        assert np.array_equal(h["max"], f.max)                       # type-7 quantiles: same arithmetic, bit-exact
        assert np.allclose(h["fit"], f.fit, rtol=1e-11, atol=0)      # least squares: solver-dependent last digits
        assert np.allclose(h["start"], f.start, rtol=1e-10)
        assert np.allclose(h["shape"], f.shape, rtol=1e-7)           # Nelder-Mead vertex
        assert h["n_pos"] == f.n_pos and h["nm_fail"] == f.nm_fail
    assert np.allclose(got.df["srp_max"], ref.df["srp_max"], rtol=1e-6, atol=1e-10)
    assert np.array_equal(got.red, ref.red) and np.array_equal(got.chk, ref.chk)


def test_merge_n_sort_matches_oracle_on_fixture(fixture_expected):
    sr, paint = _fixture_sr(fixture_expected)
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    got = ldw.mergeNsort_sr_links(ldw.CdsVar(paint, 3), sr, 20000.0, None, 3.0)
    _compare_post(got, ref)
    # the fixture has 19 links with len == sr_dist (quirk Q4): they reach the SR table and are dropped here (:418)
    assert int((sr["len"] == 20000).sum()) == 19 and not np.any(got.df["len"] == 20000)
    # links between two clusters are scored by both and kept once (:474-483)
    cross = got.df["clust1"] != got.df["clust2"]
    assert cross.any()
    keys = set(zip(got.df["pos1"][cross].tolist(), got.df["pos2"][cross].tolist()))
    assert len(keys) == int(cross.sum())
    assert [f["nm_evals"] for f in got.fits] == [f.nm_evals for f in ref.fits]
    assert len(got.red) > 0 and np.all(got.df["srp_max"][got.red] > 3.0)
    assert np.all(got.df["MI"][got.chk] >= got.df["MI"][got.red].min())
    # rows sitting on the two thresholds are listed: the link that sets min(MI) is one of them, and so are its ties
    mi_min = got.df["MI"][got.red].min()
    assert len(got.borderline_chk) >= 1 and np.all(np.abs(got.df["MI"][got.borderline_chk] - mi_min) <= 1e-12)
    assert set(np.nonzero(got.df["MI"] == mi_min)[0]) <= set(got.borderline_chk.tolist())
    assert len(got.borderline_red) == int((np.abs(got.df["srp_max"] - 3.0) <= 3e-9).sum())


def test_missing_lengths_shift_the_fit_lookup_like_the_reference():
    """`mean_dist[sr_links_t$len]` (R/computePairwiseMI.R:448) subscripts the fitted values by the value of len: with
    lengths missing the lookup lands on another length's fit, and lengths beyond the number of groups give NA and drop
    out.  Both implementations must reproduce that, not "fix" it."""
    rng = np.random.default_rng(11)
    n = 40000
    ln = rng.choice(np.arange(2, 400, 2), size=n).astype(np.float64)      # even lengths only: 199 groups, values up to 398
    mi = rng.beta(0.8, 30.0, size=n) * (ln ** -0.3)
    pos1 = rng.integers(1, 10 ** 6, n)
    sr = dict(pos1=pos1, pos2=pos1 + ln.astype(np.int64), clust1=np.ones(n, int), clust2=np.ones(n, int), len=ln, MI=mi)
    ref = PO.merge_n_sort_sr_links(sr, 1, 1000.0, 2.0)
    got = ldw.mergeNsort_sr_links(ldw.CdsVar(None, 1), sr, 1000.0, None, 2.0)
    _compare_post(got, ref)
    assert len(got.fits[0]["len"]) == 199
    assert got.df["len"].max() <= 199                                      # lengths 200..398 index past the fit: NA
    assert got.fits[0]["n_pos"] < n


def test_postprocess_errors_mirror_r():
    one = dict(pos1=np.array([1, 2]), pos2=np.array([5, 9]), clust1=np.array([1, 1]), clust2=np.array([1, 1]),
               len=np.array([4.0, 7.0]), MI=np.array([0.1, 0.2]))
    with pytest.raises(_lib.LdwError, match="fewer than two"):
        ldw.mergeNsort_sr_links(ldw.CdsVar(None, 1), one, 100.0)
    rng = np.random.default_rng(2)
    n = 5000
    ln = rng.integers(1, 50, n).astype(float)
    big = dict(pos1=np.arange(n), pos2=np.arange(n) + ln.astype(int), clust1=np.ones(n, int), clust2=np.ones(n, int), len=ln,
               MI=np.where(rng.random(n) < 0.01, 3.0, rng.beta(0.9, 20.0, n)))   # a few residuals above 1
    with pytest.raises(_lib.LdwError, match=r"values must be in \[0-1\] to fit a beta distribution"):
        ldw.mergeNsort_sr_links(ldw.CdsVar(None, 1), big, 100.0)
    with pytest.raises(ValueError, match=r"values must be in \[0-1\]"):
        PO.merge_n_sort_sr_links(big, 1, 100.0, 3.0)
    big["MI"] = rng.beta(0.9, 20.0, n)
    ldw.mergeNsort_sr_links(ldw.CdsVar(None, 1), big, 100.0)
    with pytest.raises(_lib.LdwError, match="cluster 2 holds no short-range link"):   # R: subscript / quantile errors on an empty frame
        ldw.mergeNsort_sr_links(ldw.CdsVar(None, 2), big, 100.0)


# ---------------------------------------------------------------------------------------------------------------------
# runARACNE
# ---------------------------------------------------------------------------------------------------------------------
def test_aracne_hand_cases():
    # triangle 10-20-30: the weakest edge (10,30) is indirect; ties are NOT indirect (strict <, src/computeMI.cpp:69-70)
    full = dict(pos1=np.array([10, 20, 10, 40, 50.]), pos2=np.array([20, 30, 30, 50, 60.]), MI=np.array([0.9, 0.8, 0.3, 0.5, 0.5]))
    chk = dict(pos1=np.array([10, 10, 20, 40, 40, 70.]), pos2=np.array([30, 20, 30, 60, 50, 80.]),
               MI=np.array([0.3, 0.9, 0.8, 0.5, 0.5, 0.1]))
    want = [False, True, True, True, True, True]  # (40,60): via 50 both 0.5, not strictly larger; (70,80): unknown -> TRUE
    assert ldw.runARACNE(chk, full).tolist() == want
    assert PO.run_aracne(chk["pos1"], chk["pos2"], chk["MI"], full["pos1"], full["pos2"], full["MI"]).tolist() == want
    # a link repeated in `full`: the FIRST row holding the pair is the one compared (.vecPosMatch)
    full2 = dict(pos1=np.array([1, 1, 2, 1.]), pos2=np.array([2, 2, 3, 3.]), MI=np.array([0.2, 0.9, 0.9, 0.5]))
    chk2 = dict(pos1=np.array([1.]), pos2=np.array([3.]), MI=np.array([0.5]))
    assert ldw.runARACNE(chk2, full2).tolist() == [True]     # first (1,2) row has MI 0.2 < 0.5
    assert PO.run_aracne(chk2["pos1"], chk2["pos2"], chk2["MI"], full2["pos1"], full2["pos2"], full2["MI"]).tolist() == [True]
    full2["MI"][:2] = [0.9, 0.2]
    assert ldw.runARACNE(chk2, full2).tolist() == [False]
    assert PO.run_aracne(chk2["pos1"], chk2["pos2"], chk2["MI"], full2["pos1"], full2["pos2"], full2["MI"]).tolist() == [False]
    assert ldw.runARACNE(dict(pos1=[], pos2=[], MI=[]), full).tolist() == []
    assert ldw.runARACNE(chk2, dict(pos1=[], pos2=[], MI=[])).tolist() == [True]


def test_aracne_matches_literal_restatement_on_random_graphs():
    rng = np.random.default_rng(3)
    for nv, ne, nc in ((30, 200, 150), (200, 3000, 400), (50, 60, 60)):
        p1 = rng.integers(1, nv, ne)
        p2 = p1 + rng.integers(1, 8, ne)                      # repeats of the same pair do occur
        mi = np.round(rng.uniform(0, 1, ne), 2)               # ties do occur
        k = rng.integers(0, ne, nc)
        c1, c2, cm = p1[k].astype(float), p2[k].astype(float), mi[k]
        flip = rng.random(nc) < 0.3                           # orientation of a checked link does not matter
        c1[flip], c2[flip] = c2[flip].copy(), c1[flip].copy()
        got = ldw.runARACNE(dict(pos1=c1, pos2=c2, MI=cm), dict(pos1=p1, pos2=p2, MI=mi))
        ref = PO.run_aracne(c1, c2, cm, p1, p2, mi)
        assert np.array_equal(got, ref) and 0 < got.sum() < nc


# ---------------------------------------------------------------------------------------------------------------------
# the whole tail of perform_MI_computation (:118-143) and sr_links.tsv
# ---------------------------------------------------------------------------------------------------------------------
def test_finish_sr_links_and_tsv(fixture_expected, tmp_path):
    sr, paint = _fixture_sr(fixture_expected)
    path = tmp_path / "sr_links.tsv"
    red, post = ldw.finish_sr_links(sr, ldw.CdsVar(paint, 3), 20000.0, 3.0, True, True, str(path))
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    d = ref.df
    ar = PO.run_aracne(d["pos1"][ref.red], d["pos2"][ref.red], d["MI"][ref.red], d["pos1"][ref.chk], d["pos2"][ref.chk], d["MI"][ref.chk])
    o = PO.order_links_by_srp(post.df["srp_max"][post.red])   # order by the native values (they agree to 1e-6 with the oracle's)
    assert np.array_equal(red["row"], d["row"][ref.red][o])
    assert np.array_equal(red["ARACNE"], ar[o].astype(float))
    assert np.all(np.diff(red["srp_max"]) <= 0)
    lines = path.read_text().splitlines()
    assert len(lines) == len(red["row"])
    for k in list(range(5)) + [len(lines) // 2, len(lines) - 1]:
        want = "\t".join([str(int(red["clust_c"][k])), str(int(red["pos1"][k])), str(int(red["pos2"][k]))] +
                         [O.format_r_numeric(float(red[c][k])) for c in ("clust1", "clust2", "len", "MI", "srp_max", "ARACNE")])
        assert lines[k] == want
    # appending (the reference's write.table(append = T))
    ldw.write_sr_tsv(str(path), sr, red["row"][:3], red["clust_c"][:3], red["srp_max"][:3], red["ARACNE"][:3], append=True)
    assert len(path.read_text().splitlines()) == len(lines) + 3
    # runARACNE = FALSE: warning, constant 1 (:128-129); order_links = FALSE keeps sr_links_df order; plt_folder receives
    # the per-cluster maxvls tables (the reference's c<i>_fit_data.rds, as text)
    with pytest.warns(UserWarning, match="ARACNE not run"):
        red2, post2 = ldw.finish_sr_links(sr, ldw.CdsVar(paint, 3), 20000.0, 3.0, False, False, None, str(tmp_path / "PLOTS"))
    for c in (1, 2, 3):
        tab = np.loadtxt(tmp_path / "PLOTS" / f"c{c}_fit_data.tsv", skiprows=1)
        assert np.array_equal(tab[:, 0], post2.fits[c - 1]["len"]) and np.allclose(tab[:, 2], post2.fits[c - 1]["fit"], rtol=1e-14)
    assert np.all(red2["ARACNE"] == 1) and np.array_equal(red2["row"], d["row"][ref.red])


def test_tolerance_of_the_chain_under_scan_sized_mi_errors(fixture_expected):
    """The device hands over fp32-accurate MI (|error| < 1e-6, 2e-7 measured).  The same comparison the GPU test makes
    (tests/util.py), here with the golden fp64 values perturbed by that much: shows how far the MI tolerance travels into
    srp_max and that only links at the srp cut-off can change sides."""
    from util import compare_sr_post_with_tolerance
    sr, paint = _fixture_sr(fixture_expected)
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    rng = np.random.default_rng(0)
    for amp in (2e-7, 8e-7):
        noisy = dict(sr)
        noisy["MI"] = (sr["MI"] + rng.uniform(-amp, amp, len(sr["MI"]))).astype(np.float32).astype(np.float64)
        red, post = ldw.finish_sr_links(noisy, ldw.CdsVar(paint, 3), 20000.0, 3.0, True, True, None)
        compare_sr_post_with_tolerance(red, post, ref, 3.0, 1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# link files back in, and the long-range ARACNE chain (BASELINE config #5: "long-range links fed to runARACNE")
# ---------------------------------------------------------------------------------------------------------------------
def test_link_files_round_trip_and_lr_aracne_chain(fixture_expected, tmp_path):
    e = fixture_expected
    tag = "g2M_b1000"
    lr = {k: e[f"{tag}_lr_{k}"] for k in ("pos1", "pos2", "clust1", "clust2", "len", "MI")}
    lr_path = tmp_path / "lr_links.tsv"
    ldw.write_lr_tsv(str(lr_path), lr, append=False)
    back = ldw.read_LongRangeLinks(str(lr_path))
    assert list(back) == ["pos1", "pos2", "c1", "c2", "len", "MI"] and len(back["MI"]) == len(lr["MI"])
    for a, b in (("pos1", "pos1"), ("pos2", "pos2"), ("c1", "clust1"), ("c2", "clust2"), ("len", "len")):
        np.testing.assert_array_equal(back[a], lr[b].astype(np.float64))
    # write.table keeps 15 significant digits: what comes back is the value R's read.table would see
    assert np.array_equal(back["MI"], np.array([float(O.format_r_numeric(float(v))) for v in lr["MI"]]))
    assert np.abs(back["MI"] / lr["MI"] - 1).max() < 1e-14
    assert len(ldw.read_LongRangeLinks(str(lr_path), sr_dist=5e5)["MI"]) == int((lr["len"] >= 5e5).sum())   # :43
    # spydrpick output: space separated, 5 or 4 columns (R/io_functions.R:36-41), same len filter
    sp = tmp_path / "spydrpick.txt"
    sp.write_text("10 50000 49990 1 0.25\n20 30 10 0 0.5\n")
    got5 = ldw.read_LongRangeLinks(str(sp), links_from_spydrpick=True)
    assert list(got5) == ["pos1", "pos2", "len", "ARACNE", "MI"] and got5["MI"].tolist() == [0.25] and got5["ARACNE"].tolist() == [1.0]
    sp.write_text("10 50000 49990 0.25\n")
    assert list(ldw.read_LongRangeLinks(str(sp), links_from_spydrpick=True)) == ["pos1", "pos2", "len", "MI"]
    # sr_links.tsv
    sr, paint = _fixture_sr(e)
    sr_path = tmp_path / "sr_links.tsv"
    red, _ = ldw.finish_sr_links(sr, ldw.CdsVar(paint, 3), 20000.0, 3.0, True, True, str(sr_path))
    srb = ldw.read_ShortRangeLinks(str(sr_path))
    assert len(srb["MI"]) == len(red["row"])
    np.testing.assert_array_equal(srb["pos1"], red["pos1"].astype(float))
    np.testing.assert_array_equal(srb["ARACNE"], red["ARACNE"])
    np.testing.assert_array_equal(srb["clust_c"], red["clust_c"].astype(float))
    assert np.abs(srb["srp_max"] / red["srp_max"] - 1).max() < 1e-14
    # malformed files fail loudly
    bad = tmp_path / "bad.tsv"
    bad.write_text("1\t2\t1\t1\t5\t0.1\n3\t4\t1\t1\n")
    with pytest.raises(_lib.LdwError, match="line 2 .* did not have 6 numeric fields"):
        ldw.read_LongRangeLinks(str(bad))
    with pytest.raises(_lib.LdwError, match="can't open"):
        ldw.read_ShortRangeLinks(str(tmp_path / "missing.tsv"))
    empty = tmp_path / "empty.tsv"
    empty.write_text("")
    assert len(ldw.read_ShortRangeLinks(str(empty))["MI"]) == 0
    # the long-range chain on what was read back (R/lr_analyser.R:72-116)
    got = ldw.analyse_long_range_links(back, srb)
    idx, ar, thr = PO.analyse_long_range_links(back, srb)
    assert np.array_equal(got["thresholds"], thr)
    np.testing.assert_array_equal(got["pos1"], back["pos1"][idx])
    np.testing.assert_array_equal(got["MI"], back["MI"][idx])
    np.testing.assert_array_equal(got["ARACNE"], ar)
    assert np.all(np.diff(got["MI"]) <= 0) and 0 < len(idx) < len(back["MI"])
    # fewer than 5000 outliers among >= 5000 links: the top-~5000 fallback and its warning (:94-99)
    rng = np.random.default_rng(4)
    many = dict(pos1=rng.integers(1, 10 ** 6, 20000).astype(float), pos2=rng.integers(1, 10 ** 6, 20000).astype(float),
                MI=rng.uniform(0.1, 0.2, 20000))
    with pytest.warns(UserWarning, match="top links were retained"):
        g2 = ldw.analyse_long_range_links(many, dict(pos1=np.zeros(0), pos2=np.zeros(0), MI=np.zeros(0)))
    i2, a2, t2 = PO.analyse_long_range_links(many, dict(pos1=np.zeros(0), pos2=np.zeros(0), MI=np.zeros(0)))
    assert np.array_equal(g2["thresholds"], t2) and np.array_equal(g2["MI"], many["MI"][i2]) and np.array_equal(g2["ARACNE"], a2)
    assert 4990 <= len(i2) <= 5000


def test_sr_pair_indices_address_the_block_matrices(fixture_snp, fixture_expected):
    """MIPlan.sr_exact() (fp64 MI of every short-range link) rests on mapping a link back to its cell of the block's
    MI matrix: checked here without a device -- the oracle's dense block matrices, read at those cells, must reproduce
    the golden short-range MI column exactly (diagonal and off-diagonal blocks, quirks Q1/Q5 included)."""
    import c_oracle as CO
    from ldweaver_b200 import api
    e = fixture_expected
    tag, blk = "g50k_b1000", 1000
    p1, p2, MI = e[f"{tag}_sr_pos1"], e[f"{tag}_sr_pos2"], e[f"{tag}_sr_MI"]
    # block id of every golden link: rows are in make_blocks order; rebuild it from the per-block SR counts
    snp = fixture_snp
    osr = O.perform_MI_scan(snp, e["hdw"], e["paint"], 3, max_blk_sz=blk, lr_links_approx=1e5, lr_retain_links=1e4, keep_blocks=True)
    block = np.concatenate([np.full(int(b.sr_mask.sum()), k) for k, b in enumerate(osr.blocks)])
    assert len(block) == len(MI) and np.array_equal(osr.sr["pos1"], p1)
    sr = dict(pos1=p1, pos2=p2, MI=np.zeros(len(MI)), block=block)
    seen = 0
    for k, idx, il, jl in api.sr_pair_indices(snp.POS, blk, sr):
        fs, fe, ts, te = O.make_blocks(snp.nsnp, blk)[k]
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        dense = CO.block_mi(snp.codes, e["hdw"], snp.r, snp.uqe, f, t)
        assert np.abs(dense[il, jl] - MI[idx]).max() < 1e-12
        seen += len(MI[idx])
    assert seen == len(MI)
    # rows in another order (not what the scan returns, but allowed): same cells
    perm = np.random.default_rng(1).permutation(len(MI))
    srp = {k: v[perm] for k, v in sr.items()}
    for k, idx, il, jl in api.sr_pair_indices(snp.POS, blk, srp):
        fs, fe, ts, te = O.make_blocks(snp.nsnp, blk)[k]
        assert np.array_equal(snp.POS[fs - 1 + il], srp["pos2"][idx]) and np.array_equal(snp.POS[ts - 1 + jl], srp["pos1"][idx])
    with pytest.raises(_lib.LdwError, match="strictly increasing"):
        list(api.sr_pair_indices(np.array([1, 5, 5, 9]), 1000, sr))
    wrong = dict(sr)
    wrong["block"] = np.where(np.arange(len(MI)) == 7, 2, block)      # a link filed under another block
    with pytest.raises(_lib.LdwError, match="link 7 does not belong to the block it names"):
        list(api.sr_pair_indices(snp.POS, blk, wrong))


def test_post_chain_against_committed_golden(fixture_expected):
    """tests/golden/post_expected.npz (tests/golden/make_golden_post.py): the oracle today and the native code must both
    still produce the committed sr_links_red table (rows, clusters, srp_max, ARACNE, order) and the long-range chain."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "post_expected.npz")))
    sr, paint = _fixture_sr(fixture_expected)
    # native
    red, post = ldw.finish_sr_links(sr, ldw.CdsVar(paint, 3), 20000.0, 3.0, True, True, None)
    assert len(post.df["row"]) == int(g["n_df"]) and len(post.chk) == int(g["n_chk"])
    assert int(np.sum(post.df["row"] * (np.arange(len(post.df["row"])) % 1009 + 1))) == int(g["df_row_checksum"])
    # same links; links with (mathematically) equal srp_max -- duplicate SNP patterns -- may swap places in the ordering,
    # their values agreeing to 1e-9 rather than bitwise, so the tables are compared link by link
    a, b = np.argsort(red["row"], kind="stable"), np.argsort(g["red_row"], kind="stable")
    np.testing.assert_array_equal(red["row"][a], g["red_row"][b])
    np.testing.assert_array_equal(red["clust_c"][a], g["red_clust_c"][b])
    np.testing.assert_array_equal(red["ARACNE"][a], g["red_aracne"][b].astype(float))
    assert np.allclose(red["srp_max"][a], g["red_srp_max"][b], rtol=1e-7, atol=0)
    assert np.all(np.diff(red["srp_max"]) <= 0)
    moved = red["row"] != g["red_row"]
    assert moved.mean() < 0.02 and np.allclose(red["srp_max"][moved], g["red_srp_max"][moved], rtol=1e-7)
    assert np.allclose([f["shape"] for f in post.fits], g["shape"], rtol=1e-9)
    assert np.allclose([f["coef"] for f in post.fits], g["coef"], rtol=1e-11)
    assert [f["nm_evals"] for f in post.fits] == g["nm_evals"].tolist() and [f["n_pos"] for f in post.fits] == g["n_pos"].tolist()
    assert np.allclose([f["max"].sum() for f in post.fits], g["q95_sum"], rtol=1e-14)
    # oracle (guards the oracle itself against drift)
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    o = PO.order_links_by_srp(ref.df["srp_max"][ref.red])
    np.testing.assert_array_equal(ref.df["row"][ref.red][o], g["red_row"])
    assert np.array_equal(ref.df["srp_max"][ref.red][o], g["red_srp_max"])
    # long-range chain (g = 2 221 315 tables) through the native ARACNE
    lr = {k: fixture_expected[f"g2M_b1000_lr_{k}"] for k in ("pos1", "pos2", "MI")}
    got = ldw.analyse_long_range_links(lr, {k: red[k] for k in ("pos1", "pos2", "MI")})
    assert np.array_equal(got["thresholds"], g["lr_thresholds"])
    np.testing.assert_array_equal(got["MI"], lr["MI"][g["lr_idx"]])
    np.testing.assert_array_equal(got["ARACNE"], g["lr_aracne"])


def test_sr_exact_writes_every_link_once(fixture_snp, fixture_expected):
    """MIPlan.sr_exact without a device: pairs_exact stubbed by the oracle's dense block matrices.  In place and copying,
    block-ordered and shuffled rows must all give the golden fp64 column."""
    import c_oracle as CO
    from ldweaver_b200 import api
    e = fixture_expected
    snp, blk = fixture_snp, 1000
    MI = e["g50k_b1000_sr_MI"]
    osr = O.perform_MI_scan(snp, e["hdw"], e["paint"], 3, max_blk_sz=blk, lr_links_approx=1e5, lr_retain_links=1e4, keep_blocks=True)
    block = np.concatenate([np.full(int(b.sr_mask.sum()), k) for k, b in enumerate(osr.blocks)]).astype(np.int32)
    dense = {}
    for k, (fs, fe, ts, te) in enumerate(O.make_blocks(snp.nsnp, blk)):
        dense[k] = CO.block_mi(snp.codes, e["hdw"], snp.r, snp.uqe, np.arange(fs - 1, fe), np.arange(ts - 1, te))

    class StubPlan(api.MIPlan):
        def __init__(self):  # no device, no library handle
            self.pos, self.blk, self.handle = np.asarray(snp.POS, dtype=np.int32), blk, None
            self.calls = 0

        def pairs_exact(self, block_index, from_local, to_local, out=None):
            self.calls += 1
            v = dense[block_index][from_local, to_local]
            if out is None:
                return v
            assert out.flags.c_contiguous and out.shape == v.shape
            out[:] = v
            return out

    sr = dict(pos1=e["g50k_b1000_sr_pos1"], pos2=e["g50k_b1000_sr_pos2"], block=block, MI=MI.astype(np.float32).astype(np.float64))
    plan = StubPlan()
    got = plan.sr_exact(sr)                                    # copy: the input column is left alone
    assert np.abs(got - MI).max() < 1e-12 and np.abs(sr["MI"] - MI).max() > 1e-9
    same = plan.sr_exact(sr, inplace=True)
    assert same is sr["MI"] and np.abs(sr["MI"] - MI).max() < 1e-12
    perm = np.random.default_rng(5).permutation(len(MI))
    shuffled = {k: v[perm] for k, v in sr.items()}
    shuffled["MI"] = np.zeros(len(MI))
    assert np.abs(plan.sr_exact(shuffled) - MI[perm]).max() < 1e-12
    StubPlan.__del__ = lambda self: None


def test_post_results_do_not_depend_on_the_thread_count(fixture_expected, tmp_path):
    """ldw_sr_postprocess cuts every pass into a fixed number of pieces, so one host thread and eight must give the same
    bits (LDW_HOST_THREADS overrides the thread count)."""
    import os
    import subprocess
    import sys
    sr, paint = _fixture_sr(fixture_expected)
    np.savez(tmp_path / "sr.npz", paint=paint, **sr)
    code = ("import sys, numpy as np; sys.path[:0] = [%r]; import ldweaver_b200 as ldw; d = dict(np.load(sys.argv[1])); paint = d.pop('paint');"
            "p = ldw.mergeNsort_sr_links(ldw.CdsVar(paint, 3), d, 20000.0, None, 3.0);"
            "np.savez(sys.argv[2], row=p.df['row'], srp=p.df['srp_max'], red=p.red, chk=p.chk, shape=np.array([f['shape'] for f in p.fits]),"
            "fit=np.concatenate([f['fit'] for f in p.fits]))") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for nt in ("1", "3", "8"):
        out = tmp_path / f"o{nt}.npz"
        subprocess.run([sys.executable, "-c", code, str(tmp_path / "sr.npz"), str(out)], check=True, env=dict(os.environ, LDW_HOST_THREADS=nt))
        outs.append(dict(np.load(out)))
    for o in outs[1:]:
        for k in outs[0]:
            assert np.array_equal(o[k], outs[0][k]), k


@pytest.mark.parametrize("seed", range(8))
def test_merge_n_sort_differential_fuzz(seed):
    """Random tables against the literal oracle: 1-5 clusters, cluster ids outside 1..nclust, links listed under two
    clusters, lengths 0 and >= sr_dist, fractional sr_dist, missing lengths (the subscript-by-value quirk), tied MI values."""
    rng = np.random.default_rng(seed)
    nclust = int(rng.integers(1, 6))
    n = int(rng.integers(3000, 30000))
    sr_dist = float(rng.choice([50, 200.5, 1000, 5000]))
    ln = rng.integers(0, int(sr_dist * rng.choice([0.5, 1.0, 1.3])) + 2, n)
    if rng.random() < 0.5:
        ln = ln * 2
    pos1 = rng.integers(1, 10 ** 6, n)
    c1 = rng.integers(0, nclust + 2, n)
    c2 = np.where(rng.random(n) < 0.7, c1, rng.integers(1, nclust + 1, n))
    mi = rng.beta(0.8, rng.choice([20, 60, 200]), n) * (np.maximum(ln, 1) ** -rng.choice([0.0, 0.3]))
    k = rng.integers(0, n, n // 10)
    mi[k] = mi[rng.integers(0, n, n // 10)]
    sr = dict(pos1=pos1, pos2=pos1 + ln, clust1=c1, clust2=c2, len=ln.astype(float), MI=mi)
    cutoff = float(rng.choice([1.0, 3.0]))
    ref = PO.merge_n_sort_sr_links(sr, nclust, sr_dist, cutoff)
    got = ldw.mergeNsort_sr_links(ldw.CdsVar(None, nclust), sr, sr_dist, None, cutoff)
    _compare_post(got, ref)
    assert [f["nm_evals"] for f in got.fits] == [f.nm_evals for f in ref.fits]


def test_links_to_cells_random_blocks():
    """ldw_links_to_cells against a NumPy restatement: random positions, ragged last block, diagonal and off-diagonal
    blocks, rows in block order and shuffled."""
    from ldweaver_b200 import api
    rng = np.random.default_rng(9)
    for n, blk in ((2500, 1000), (4096, 1024), (130, 128), (5000, 3000)):
        pos = np.sort(rng.choice(np.arange(1, 10 * n), n, replace=False)).astype(np.int32)
        blocks = api.make_blocks(n, blk)
        rows = []
        for k, (fs, fe, ts, te) in enumerate(blocks):
            m = int(rng.integers(0, 400))
            i = rng.integers(fs - 1, fe, m)
            j = rng.integers(ts - 1, te, m)
            rows.append(np.stack([np.full(m, k), i, j], axis=1))
        t = np.concatenate(rows)
        for order in (np.arange(len(t)), rng.permutation(len(t))):
            b, gi, gj = t[order, 0], t[order, 1], t[order, 2]
            sr = dict(pos1=pos[gj], pos2=pos[gi], block=b.astype(np.int32))
            seen = np.zeros(len(b), dtype=bool)
            for k, idx, il, jl in api.sr_pair_indices(pos, blk, sr):
                fs, fe, ts, te = blocks[k]
                assert np.array_equal(il, gi[idx] - (fs - 1)) and np.array_equal(jl, gj[idx] - (ts - 1)) and np.all(b[idx] == k)
                seen[idx] = True
            assert seen.all()
