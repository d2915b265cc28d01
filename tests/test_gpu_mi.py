"""GPU parity tests (through the C ABI): weighted pairwise-MI scan vs the oracle and the golden fixture."""
import numpy as np
import pytest

from util import compare_lr_sets, compare_sr_post_with_tolerance

pytestmark = pytest.mark.gpu

MI_TOL = 1e-6  # north star: MI within 1e-6 absolute of the reference double-precision result


def _snp(fixture_snp, g=50000):
    import ldweaver_b200 as ldw
    return ldw.snp_dat_from_codes(fixture_snp.codes, fixture_snp.POS, g)


def test_dense_block_mi_single_block(fixture_snp, fixture_expected):
    """fp32 tensor-core path, one diagonal block (the whole fixture): every cell within 1e-6."""
    import c_oracle as CO
    import ldweaver_b200 as ldw
    e = fixture_expected
    snp = _snp(fixture_snp)
    plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 10000)
    MI = plan.block_dense(0)
    idx = np.arange(snp.nsnp)
    ref = CO.block_mi(fixture_snp.codes, e["hdw"], fixture_snp.r, fixture_snp.uqe, idx, idx)
    err = np.abs(MI - ref)
    print("max abs err (diag block, fp32 path):", err.max())
    assert err.max() < MI_TOL
    plan.close()


def test_dense_block_mi_offdiag_and_ragged(fixture_snp, fixture_expected):
    """Quirk Q1 on a square off-diagonal block (blk=500 -> blocks of 500,500,268) and on ragged ones."""
    import c_oracle as CO
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    e = fixture_expected
    snp = _snp(fixture_snp)
    plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 500)
    blocks = O.make_blocks(snp.nsnp, 500)
    worst = 0.0
    for bi, (fs, fe, ts, te) in enumerate(blocks):
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        ref = CO.block_mi(fixture_snp.codes, e["hdw"], fixture_snp.r, fixture_snp.uqe, f, t)
        MI = plan.block_dense(bi)
        assert MI.shape == ref.shape
        err = np.abs(MI - ref).max()
        worst = max(worst, err)
        assert err < MI_TOL, f"block {bi} ({fs}-{fe} x {ts}-{te}): {err}"
    print("max abs err over blocks:", worst)
    plan.close()


def test_dense_blocks_vs_reference_hadamard_goldens(fixture_input):
    """a7-a9 with quirk Q1 against tests/golden/ref_expected.npz: per-block MI (max_blk_sz 500: diagonal, square
    off-diagonal, ragged) whose element-wise finish was executed by the reference's compiled .fastHadamard
    (src/computeMI.cpp:11-21; tests/golden/make_golden_ref.py).  fp32 epilogue within 1e-6, fp64 refinement 1e-12."""
    import os
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    from conftest import GOLDEN
    e = dict(np.load(os.path.join(GOLDEN, "ref_expected.npz")))
    POS = fixture_input["pos"][e["enc1_pos"].astype(np.int64) - 1]
    snp = ldw.snp_dat_from_codes(e["enc1_codes"], POS, 50000)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    np.testing.assert_array_equal(hdw, e["mi_hdw"])
    plan = ldw.MIPlan(snp, hdw, np.ones(snp.nsnp, dtype=np.int64), int(e["mi_blk"]))
    worst32 = worst64 = 0.0
    for bi, (fs, fe, ts, te) in enumerate(O.make_blocks(snp.nsnp, int(e["mi_blk"]))):
        rows, want = e[f"mi_b{bi}_rows"], e[f"mi_b{bi}"]
        MI = plan.block_dense(bi)
        worst32 = max(worst32, float(np.abs(MI[rows] - want).max()))
        il = np.repeat(rows, want.shape[1])
        jl = np.tile(np.arange(want.shape[1]), len(rows))
        got = plan.pairs_exact(bi, il, jl).reshape(want.shape)
        worst64 = max(worst64, float(np.abs(got - want).max()))
    print("vs reference .fastHadamard goldens: fp32 epilogue", worst32, "fp64 refinement", worst64)
    assert worst32 < MI_TOL and worst64 < 1e-12
    plan.close()


def test_exact_pairs_fp64(fixture_snp, fixture_expected):
    import c_oracle as CO
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    e = fixture_expected
    snp = _snp(fixture_snp)
    plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 1000)
    rng = np.random.default_rng(3)
    for bi, (fs, fe, ts, te) in enumerate(O.make_blocks(snp.nsnp, 1000)):
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        ref = CO.block_mi(fixture_snp.codes, e["hdw"], fixture_snp.r, fixture_snp.uqe, f, t)
        il = rng.integers(0, len(f), 4000)
        jl = rng.integers(0, len(t), 4000)
        got = plan.pairs_exact(bi, il, jl)
        assert np.abs(got - ref[il, jl]).max() < 1e-12
    plan.close()


@pytest.mark.parametrize("tag,g,blk,retain", [("g50k_b10000", 50000, 10000, 1e4), ("g50k_b1000", 50000, 1000, 1e4),
                                               ("g2M_b1000", 2221315, 1000, 2e4)])
def test_scan_vs_golden(fixture_snp, fixture_expected, tag, g, blk, retain):
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    e = fixture_expected
    snp = _snp(fixture_snp, g)
    cds = ldw.CdsVar(paint=e["paint"], nclust=3)
    res = ldw.perform_MI_computation(snp, e["hdw"], cds, ncores=1, sr_dist=20000, lr_retain_links=retain, max_blk_sz=blk,
                                     lr_links_approx=1e5, write_tsv=False)
    # ---- short-range rows: identical set, identical (reference) order, MI within tolerance
    assert len(res.sr["MI"]) == int(e[f"{tag}_sr_n"])
    np.testing.assert_array_equal(res.sr["pos1"], e[f"{tag}_sr_pos1"])
    np.testing.assert_array_equal(res.sr["pos2"], e[f"{tag}_sr_pos2"])
    assert np.abs(res.sr["MI"][::16] - e[f"{tag}_sr_MI_16"]).max() < MI_TOL
    if f"{tag}_sr_MI" in e:
        assert np.abs(res.sr["MI"] - e[f"{tag}_sr_MI"]).max() < MI_TOL
    posmap = {int(p): i for i, p in enumerate(snp.POS)}
    ln = 0.5 * g - np.abs(np.mod(res.sr["pos1"].astype(float) - res.sr["pos2"], g) - 0.5 * g)
    np.testing.assert_array_equal(res.sr["len"], ln.astype(np.int64))
    # ---- thresholds (exact type-7 quantile on fp64-refined values) and long-range rows
    thr_ref = e[f"{tag}_thr"]
    assert np.array_equal(np.isnan(res.thr), np.isnan(thr_ref))
    ok = ~np.isnan(thr_ref)
    assert np.abs(res.thr[ok] - thr_ref[ok]).max() < 1e-12
    assert np.abs(res.prob[ok] - e[f"{tag}_prob"][ok]).max() == 0
    g_p1, g_p2, g_mi = e[f"{tag}_lr_pos1"], e[f"{tag}_lr_pos2"], e[f"{tag}_lr_MI"]
    nb = 0
    blocks = O.make_blocks(snp.nsnp, O.r_round_to_thousands(blk))
    POS = snp.POS.astype(np.float64)
    for bi, (fs, fe, ts, te) in enumerate(blocks):
        if np.isnan(thr_ref[bi]):
            continue
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        in_blk = np.isin(g_p2, POS[f]) & np.isin(g_p1, POS[t])
        if fs != ts:
            in_blk &= (g_p2 < POS[t[0]]) & (g_p1 > POS[f[-1]])
        else:
            in_blk &= (g_p1 <= POS[f[-1]]) & (g_p2 <= POS[f[-1]]) & (g_p1 >= POS[f[0]]) & (g_p2 >= POS[f[0]])
        k = res.lr["block"] == bi
        _, a, b = compare_lr_sets(g_p1[in_blk], g_p2[in_blk], g_mi[in_blk], res.lr["pos1"][k], res.lr["pos2"][k],
                                  res.lr["MI"][k], thr_ref[bi])
        nb += len(a) + len(b)
    print(tag, "LR kept", len(res.lr["MI"]), "golden", len(g_mi), "borderline differences", nb, "stats", res.stats)
    assert nb < 0.02 * max(1, len(g_mi))
    # long-range rows come out in reference order (block, then row order inside the block)
    assert np.all(np.diff(res.lr["block"]) >= 0)
    # cluster routing (R/computePairwiseMI.R:372-376)
    for c in range(3):
        m = (res.sr["clust1"] == c + 1) | (res.sr["clust2"] == c + 1)
        np.testing.assert_array_equal(np.nonzero(m)[0], res.sr_links[c])


def test_scan_lr_order_matches_reference(fixture_snp, fixture_expected):
    """Order of the retained LR rows inside each block equals the reference's row order (non-borderline rows)."""
    import ldweaver_b200 as ldw
    e = fixture_expected
    tag = "g50k_b1000"
    snp = _snp(fixture_snp, 50000)
    res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), lr_retain_links=1e4, max_blk_sz=1000,
                                     lr_links_approx=1e5, write_tsv=False)
    got = list(zip(res.lr["pos1"].tolist(), res.lr["pos2"].tolist()))
    ref = list(zip(e[f"{tag}_lr_pos1"].astype(int).tolist(), e[f"{tag}_lr_pos2"].astype(int).tolist()))
    common = set(got) & set(ref)
    assert [x for x in got if x in common] == [x for x in ref if x in common]


def test_sr_only_mode_vs_oracle(fixture_snp, fixture_expected):
    """Q12: SR-only mode drops far SNPs first (local indices change, so Q1/Q2 act on the reduced lists)."""
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    e = fixture_expected
    g = 2221315
    osnp = O.snp_dat_from_codes(fixture_snp.codes, fixture_snp.POS, g)
    ref = O.perform_MI_scan(osnp, e["hdw"], e["paint"], 3, max_blk_sz=1000, perform_SR_analysis_only=True, sr_dist=2000)
    snp = _snp(fixture_snp, g)
    res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), sr_dist=2000, max_blk_sz=1000,
                                     perform_SR_analysis_only=True, write_tsv=False)
    assert len(res.lr["MI"]) == 0
    np.testing.assert_array_equal(res.sr["pos1"], ref.sr["pos1"].astype(np.int32))
    np.testing.assert_array_equal(res.sr["pos2"], ref.sr["pos2"].astype(np.int32))
    assert np.abs(res.sr["MI"] - ref.sr["MI"]).max() < MI_TOL


def test_multiallelic_nrich_vs_oracle():
    """N-rich, multi-allelic synthetic data: every plane-count class (r = 2..5) and all tile kinds."""
    import c_oracle as CO
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    rng = np.random.default_rng(42)
    S, n = 300, 700
    nall = rng.choice([2, 3, 4], size=n, p=[0.5, 0.3, 0.2])
    founders = np.stack([rng.integers(0, k, size=10) for k in nall]).astype(np.uint8)
    codes = founders[:, rng.integers(0, 10, size=S)]
    flip = rng.random((n, S)) < 0.05
    codes[flip] = (codes[flip] + 1) % np.repeat(nall, S).reshape(n, S)[flip]
    codes[rng.random((n, S)) < 0.08] = 4  # N / gap class
    osnp = O.snp_dat_from_codes(codes, np.sort(rng.choice(np.arange(1, 90000), n, replace=False)), 100000)
    assert set(np.unique(osnp.r)) >= {2.0, 3.0, 4.0, 5.0} or osnp.r.max() >= 4
    keep = osnp.r >= 2
    codes, POS = codes[keep], osnp.POS[keep]
    osnp = O.snp_dat_from_codes(codes, POS, 100000)
    hdw = CO.hdw(codes, 0.1)[0]
    snp = ldw.snp_dat_from_codes(codes, POS, 100000)
    plan = ldw.MIPlan(snp, hdw, np.ones(len(POS), dtype=np.int32), 1000)
    idx = np.arange(len(POS))
    ref = CO.block_mi(codes, hdw, osnp.r, osnp.uqe, idx, idx)
    MI = plan.block_dense(0)
    err = np.abs(MI - ref).max()
    print("multi-allelic max abs err:", err, "r histogram", np.unique(osnp.r, return_counts=True))
    assert err < MI_TOL
    plan.close()


def test_scan_partition_union_and_determinism(fixture_snp, fixture_expected):
    """n_parts/part deal make_blocks rows round-robin: the union of the parts equals the full scan, and a repeated
    scan is bit-identical (SR slots are position-determined; LR rows are sorted into reference order)."""
    import ldweaver_b200 as ldw
    e = fixture_expected
    snp = _snp(fixture_snp, 2221315)
    plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 1000)
    full = plan.scan(2221315, 20000, 2e4, 1e5)
    again = plan.scan(2221315, 20000, 2e4, 1e5)
    for k in ("pos1", "pos2", "MI", "len", "block"):
        np.testing.assert_array_equal(full[0][k], again[0][k])
        np.testing.assert_array_equal(full[1][k], again[1][k])
    parts = [plan.scan(2221315, 20000, 2e4, 1e5, 0, 3, r) for r in range(3)]
    for tbl in (0, 1):
        order = np.argsort(np.concatenate([p[tbl]["block"] for p in parts]), kind="stable")
        for k in ("pos1", "pos2", "MI", "len", "block", "clust1", "clust2"):
            np.testing.assert_array_equal(np.concatenate([p[tbl][k] for p in parts])[order], full[tbl][k])
    assert sum(p[5]["n_pairs"] for p in parts) == full[5]["n_pairs"] == 803010
    plan.close()


def test_scan_synthetic_medium_vs_oracle():
    """Synthetic alignment shaped like the benchmark workload (N at 1 % per cell, so r = 3 dominates), six blocks
    incl. ragged ones: every SR link and every block threshold against the C oracle."""
    import c_oracle as CO
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy = synth.generate(nseq=300, nsnp=2700, seed=11)
    hdw = CO.hdw(sy.codes, 0.1)[0]
    osnp = O.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000)
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(sy.paint, 3), sr_dist=20000, lr_retain_links=5e4, max_blk_sz=1000,
                                     lr_links_approx=lra, write_tsv=False)
    POS = osnp.POS.astype(np.float64)
    worst = 0.0
    for bi, (fs, fe, ts, te) in enumerate(O.make_blocks(osnp.nsnp, 1000)):
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        MI = CO.block_mi(sy.codes, hdw, osnp.r, osnp.uqe, f, t)
        L = CO.block_links(MI, POS, f, t, float(sy.g), 20000.0, 5e4, lra)
        k = res.sr["block"] == bi
        np.testing.assert_array_equal(res.sr["pos1"][k], POS[t[L["col"]]][L["is_sr"]].astype(np.int32))
        np.testing.assert_array_equal(res.sr["pos2"][k], POS[f[L["row"]]][L["is_sr"]].astype(np.int32))
        if k.any():
            worst = max(worst, float(np.abs(res.sr["MI"][k] - L["MI"][L["is_sr"]]).max()))
        if np.isnan(L["thr"]):
            assert np.isnan(res.thr[bi])
        else:
            assert abs(res.thr[bi] - L["thr"]) < 1e-12
            kk = res.lr["block"] == bi
            compare_lr_sets(POS[t[L["col"]]][L["lr_keep"]], POS[f[L["row"]]][L["lr_keep"]], L["MI"][L["lr_keep"]],
                            res.lr["pos1"][kk], res.lr["pos2"][kk], res.lr["MI"][kk], L["thr"])
    print("synthetic medium: max |dMI| over SR links", worst, res.stats)
    assert worst < MI_TOL


def test_rerun_path_after_a_failed_threshold_seed(monkeypatch):
    """A candidate-threshold seed above the true threshold makes the selection fail its completeness check; the block
    is then re-run (first from a zero threshold, then with every long-range pair collected) and must give exactly the
    result of an undisturbed scan."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy = synth.generate(nseq=300, nsnp=6000, seed=21)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000.0)
    plan = ldw.MIPlan(snp, hdw, sy.paint, 2000)
    ref = plan.scan(sy.g, 20000.0, 5e4, lra)   # may itself re-run a block: the seed comes from whichever block was selected last
    monkeypatch.setenv("LDW_DBG_FORCE_SEED", "0.5")          # far above any block threshold of this data set
    got = plan.scan(sy.g, 20000.0, 5e4, lra)
    monkeypatch.delenv("LDW_DBG_FORCE_SEED")
    assert got[5]["n_reruns"] > 0
    for which in (0, 1):
        for col in ("pos1", "pos2", "clust1", "clust2", "len", "MI", "block"):
            assert np.array_equal(ref[which][col], got[which][col]), (which, col)
    assert np.array_equal(ref[3], got[3], equal_nan=True)
    plan.close()


def test_selection_margin_is_verified_and_widened(monkeypatch):
    """The long-range selection measures the fp32 epilogue's error on its refined candidates (BlockResult.eps_obs) and re-runs
    a block with a wider margin when 4 x that exceeds the margin in use.  A tiny initial margin (test hook) must trigger the
    re-run on every block and still give exactly the undisturbed result."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy = synth.generate(nseq=300, nsnp=6000, seed=22)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000.0)
    plan = ldw.MIPlan(snp, hdw, sy.paint, 2000)
    ref = plan.scan(sy.g, 20000.0, 5e4, lra)
    assert 0 < ref[5]["eps_obs_max"] < 1e-6        # the observed error itself stays inside the MI bar
    monkeypatch.setenv("LDW_DBG_SEL_MARGIN", "1e-8")
    got = plan.scan(sy.g, 20000.0, 5e4, lra)
    monkeypatch.delenv("LDW_DBG_SEL_MARGIN")
    assert got[5]["n_reruns"] >= 6                  # every block with a long-range branch was sent back
    for which in (0, 1):
        for col in ("pos1", "pos2", "clust1", "clust2", "len", "MI", "block"):
            assert np.array_equal(ref[which][col], got[which][col]), (which, col)
    assert np.array_equal(ref[3], got[3], equal_nan=True)
    plan.close()


def test_strongly_clonal_weights_at_5000_sequences():
    """one clonal cluster of 5000 near-identical sequences (weights 1/5001, coherently rounded in the 28-bit
    fixed-point unit whose scale is set by the largest weight, 0.5 for the singletons) is where the epilogue error is
    largest.  Thresholds and link sets must still match the oracle, with the observed error reported."""
    import c_oracle as CO
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    rng = np.random.default_rng(77)
    S_big, S_single, n = 5000, 12, 260
    base = rng.integers(0, 2, size=n).astype(np.uint8)
    clone = np.repeat(base[:, None], S_big, axis=1)
    flip = rng.random((n, S_big)) < 0.02                       # within-cluster variation: distances far below 0.1 n
    clone[flip] ^= 1
    singles = rng.integers(0, 2, size=(n, S_single)).astype(np.uint8)
    codes = np.ascontiguousarray(np.concatenate([clone, singles], axis=1))
    codes[rng.random(codes.shape) < 0.01] = 4
    POS = np.sort(rng.choice(np.arange(1, 400000), n, replace=False)).astype(np.int32)
    g = 500000
    snp = ldw.snp_dat_from_codes(codes, POS, g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    hdw_c, _ = CO.hdw(codes, 0.1)
    np.testing.assert_array_equal(hdw, hdw_c)
    assert hdw.max() == 0.5 and np.sum(hdw < 1e-3) >= S_big - 5     # one huge cluster, isolated singletons
    osnp = O.snp_dat_from_codes(codes, POS, g)
    idx = np.arange(n)
    MI = CO.block_mi(codes, hdw, osnp.r, osnp.uqe, idx, idx)
    L = CO.block_links(MI, POS.astype(np.float64), idx, idx, float(g), 20000.0, 3000.0, 2e4)
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(np.ones(n, dtype=np.int32), 1), sr_dist=20000, lr_retain_links=3000,
                                     max_blk_sz=1000, lr_links_approx=2e4, write_tsv=False)
    print("clonal 5000: eps_obs_max", res.stats["eps_obs_max"], "reruns", res.stats["n_reruns"],
          "max |dMI| SR", float(np.abs(res.sr["MI"] - L["MI"][L["is_sr"]]).max()))
    assert abs(res.thr[0] - L["thr"]) < 1e-12
    P = POS.astype(np.float64)
    compare_lr_sets(P[L["col"]][L["lr_keep"]], P[L["row"]][L["lr_keep"]], L["MI"][L["lr_keep"]],
                    res.lr["pos1"], res.lr["pos2"], res.lr["MI"], L["thr"])
    assert np.abs(res.sr["MI"] - L["MI"][L["is_sr"]]).max() < MI_TOL
    assert 4 * res.stats["eps_obs_max"] <= max(4e-6, 8 * res.stats["eps_obs_max"])   # whatever was observed, the margin used covered it


def test_whole_perform_mi_computation_with_tsvs(fixture_snp, fixture_expected, tmp_path):
    """perform_MI_computation as the reference runs it (R/computePairwiseMI.R:46-145): scan on the device, then the
    native mergeNsort_sr_links / runARACNE / ordering / sr_links.tsv + lr_links.tsv, against the oracle chain on the
    golden short-range table.  srp_max is a statistic of ALL short-range MI values (per-length percentiles, a decay fit,
    a beta fit whose likelihood weighs residuals near zero by their logarithm), so the 1e-6 tolerance on MI shows up
    amplified: +-4e-7 on MI moves srp_max by up to ~2e-2, and one (cluster, length) group of this fixture has a 95th
    percentile of +1e-8, which the fp32 epilogue's 3e-7 can turn negative -- log() of it then poisons the decay fit, in R as
    here.  That is why the short-range MI is refined to fp64 whenever the post-processing runs; this test takes the values
    from inside the scan (LDW_SCAN_SR_EXACT), the next one the host-driven way, and tests/test_post_cpu.py covers what
    MI errors of the scan's size do to the chain."""
    import ldw_oracle as O
    import post_oracle as PO
    import ldweaver_b200 as ldw
    e = fixture_expected
    tag = "g50k_b1000"
    snp = _snp(fixture_snp, 50000)
    lr_path, sr_path = tmp_path / "lr_links.tsv", tmp_path / "sr_links.tsv"
    res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), ncores=1, lr_save_path=str(lr_path),
                                     sr_save_path=str(sr_path), plt_folder=str(tmp_path), sr_dist=20000, lr_retain_links=1e4,
                                     max_blk_sz=1000, srp_cutoff=3, runARACNE=True, lr_links_approx=1e5, exact_sr="in_scan")
    red = res.sr_links_red
    assert red is not None and len(red["row"]) > 0
    # oracle chain on the golden (fp64) short-range table
    p1, p2, MI = e[f"{tag}_sr_pos1"], e[f"{tag}_sr_pos2"], e[f"{tag}_sr_MI"]
    np.testing.assert_array_equal(res.sr["pos1"], p1.astype(np.int32))
    lut = np.zeros(int(snp.POS.max()) + 1, dtype=np.int32)
    lut[snp.POS] = e["paint"]
    sr = dict(pos1=p1, pos2=p2, clust1=lut[p1], clust2=lut[p2], len=O.circ_len(p1.astype(float), p2.astype(float), 50000.0), MI=MI)
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    compare_sr_post_with_tolerance(red, res.sr_post, ref, 3.0, MI_TOL)
    # files: one row per returned link, 9 columns; lr_links.tsv 6 columns
    rows = sr_path.read_text().splitlines()
    assert len(rows) == len(red["row"]) and all(len(r.split("\t")) == 9 for r in rows[:50])
    first = rows[0].split("\t")
    assert int(first[1]) == int(red["pos1"][0]) and int(first[2]) == int(red["pos2"][0])
    assert abs(float(first[6]) - red["MI"][0]) < 1e-14 and abs(float(first[7]) - red["srp_max"][0]) < 1e-12 * red["srp_max"][0] + 1e-14
    lrows = lr_path.read_text().splitlines()
    assert len(lrows) == len(res.lr["MI"]) and len(lrows[0].split("\t")) == 6


def test_exact_short_range_mi_gives_tight_post_parity(fixture_snp, fixture_expected):
    """exact_sr=True (the default whenever the post-processing runs): every short-range MI recomputed in fp64 (reference arithmetic) -> the column matches the golden
    fp64 values to 1e-12 and the statistics derived from it (percentiles, decay fit, beta fit, srp_max, the srp cut and
    the ARACNE check set) match the oracle chain without the amplified fp32 tolerance."""
    import ldw_oracle as O
    import post_oracle as PO
    import ldweaver_b200 as ldw
    e = fixture_expected
    tag = "g50k_b1000"
    snp = _snp(fixture_snp, 50000)
    res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), lr_retain_links=1e4, max_blk_sz=1000,
                                     lr_links_approx=1e5, write_tsv=False, postprocess=True)   # exact_sr defaults to True here
    p1, p2, MI = e[f"{tag}_sr_pos1"], e[f"{tag}_sr_pos2"], e[f"{tag}_sr_MI"]
    np.testing.assert_array_equal(res.sr["pos1"], p1.astype(np.int32))
    err = np.abs(res.sr["MI"] - MI).max()
    print("exact short-range MI: max abs err", err)
    assert err < 1e-12
    lut = np.zeros(int(snp.POS.max()) + 1, dtype=np.int32)
    lut[snp.POS] = e["paint"]
    sr = dict(pos1=p1, pos2=p2, clust1=lut[p1], clust2=lut[p2], len=O.circ_len(p1.astype(float), p2.astype(float), 50000.0), MI=MI)
    ref = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    post = res.sr_post
    np.testing.assert_array_equal(post.df["row"], ref.df["row"])
    np.testing.assert_array_equal(post.df["clust_c"], ref.df["clust_c"])
    for h, f in zip(post.fits, ref.fits):
        assert np.abs(h["max"] - f.max).max() < 1e-12 and np.allclose(h["shape"], f.shape, rtol=1e-6)
    assert np.allclose(post.df["srp_max"], ref.df["srp_max"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(post.red, ref.red)
    # sr_links_ARACNE_check is `MI >= min(sr_links_red$MI)` (:495): links that TIE with that minimum (duplicate SNP
    # patterns give mathematically equal MI) are decided by the last ulp in any implementation -- only those may differ
    mi_min = ref.df["MI"][ref.red].min()
    diff = np.setxor1d(post.chk, ref.chk)
    print("ARACNE check set:", len(ref.chk), "links,", len(diff), "ties with the minimum differ")
    assert np.all(np.abs(ref.df["MI"][diff] - mi_min) < 1e-12) and len(diff) < 0.01 * len(ref.chk)
    with pytest.raises(ValueError, match="exact_sr is not available"):
        ldw.perform_MI_computation(_snp(fixture_snp, 2221315), e["hdw"], ldw.CdsVar(e["paint"], 3), sr_dist=2000, max_blk_sz=1000,
                                   perform_SR_analysis_only=True, write_tsv=False, exact_sr=True)


@pytest.mark.parametrize("sr_only", [False, True])
def test_in_scan_exact_short_range_mi(fixture_snp, fixture_expected, sr_only):
    """LDW_SCAN_SR_EXACT (mi_sr_exact_kernel): fp64 short-range MI from inside the scan call -- must equal the golden fp64
    column to 1e-12, and in SR-only mode (quirk Q12: reduced SNP lists) the oracle's SR-only values."""
    import ldw_oracle as O
    import ldweaver_b200 as ldw
    e = fixture_expected
    if not sr_only:
        snp = _snp(fixture_snp, 50000)
        res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), lr_retain_links=1e4, max_blk_sz=1000,
                                         lr_links_approx=1e5, write_tsv=False, exact_sr="in_scan")
        np.testing.assert_array_equal(res.sr["pos1"], e["g50k_b1000_sr_pos1"].astype(np.int32))
        assert np.abs(res.sr["MI"] - e["g50k_b1000_sr_MI"]).max() < 1e-12
    else:
        g = 2221315
        osnp = O.snp_dat_from_codes(fixture_snp.codes, fixture_snp.POS, g)
        ref = O.perform_MI_scan(osnp, e["hdw"], e["paint"], 3, max_blk_sz=1000, perform_SR_analysis_only=True, sr_dist=2000)
        res = ldw.perform_MI_computation(_snp(fixture_snp, g), e["hdw"], ldw.CdsVar(e["paint"], 3), sr_dist=2000, max_blk_sz=1000,
                                         perform_SR_analysis_only=True, write_tsv=False, exact_sr="in_scan")
        np.testing.assert_array_equal(res.sr["pos1"], ref.sr["pos1"].astype(np.int32))
        assert np.abs(res.sr["MI"] - ref.sr["MI"]).max() < 1e-12


def test_device_post_processing_matches_the_host_implementation(fixture_snp, fixture_expected, tmp_path):
    """mergeNsort_sr_links on the device-resident short-range table (ldw_sr_postprocess_dev) against the native host code
    (ldw_sr_postprocess) on the same table: same rows in the same order, same cluster attribution, per-length percentiles
    bit for bit (order statistics), decay fits identical, beta shapes / srp_max to 1e-9 relative (sums formed in another
    order), same sr_links_red / ARACNE check sets apart from listed borderline rows."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api
    e = fixture_expected
    for g, sr_dist in ((50000, 20000.0), (2221315, 7000.0)):
        snp = _snp(fixture_snp, g)
        cds = ldw.CdsVar(e["paint"], 3)
        plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 1000)
        sr, lr, bd, thr, prob, st = plan.scan(float(g), sr_dist, 1e4, 1e5, api.SCAN_SR_EXACT)   # rows on the host AND left on the device
        host = api.mergeNsort_sr_links(cds, sr, sr_dist, None, 3.0)
        dev = api.mergeNsort_sr_links_device(cds, sr_dist, None, 3.0)
        plan.close()
        # the device path returns sr_links_df restricted to the rows of sr_links_red / sr_links_ARACNE_check (all that is used
        # downstream); the host path returns every link above the fit
        need = np.zeros(len(host.df["row"]), dtype=bool)
        need[host.red] = True
        need[host.chk] = True
        np.testing.assert_array_equal(dev.df["row"], host.df["row"][need])
        np.testing.assert_array_equal(dev.df["clust_c"], host.df["clust_c"][need])
        for k in ("pos1", "pos2", "clust1", "clust2", "len", "MI"):
            np.testing.assert_array_equal(dev.df[k], host.df[k][need], err_msg=k)
        for fd, fh in zip(dev.fits, host.fits):
            np.testing.assert_array_equal(fd["len"], fh["len"])
            np.testing.assert_array_equal(fd["max"], fh["max"])          # type-7 percentiles: bit-exact
            np.testing.assert_array_equal(fd["fit"], fh["fit"])
            assert fd["n_pos"] == fh["n_pos"]
            np.testing.assert_allclose(fd["start"], fh["start"], rtol=1e-9)
            np.testing.assert_allclose(fd["shape"], fh["shape"], rtol=1e-7)
        np.testing.assert_allclose(dev.df["srp_max"], host.df["srp_max"][need], rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(dev.df["row"][dev.red], host.df["row"][host.red])
        np.testing.assert_array_equal(dev.df["row"][dev.chk], host.df["row"][host.chk])
        print(f"g={g}: {len(sr['MI'])} links -> df {len(host.df['row'])}, red {len(host.red)}, chk {len(host.chk)}; max rel srp diff",
              float(np.max(np.abs(dev.df["srp_max"] - host.df["srp_max"][need]) / np.maximum(host.df["srp_max"][need], 1e-300))))
    # the whole call with the table kept on the device: same sr_links_red (values to 1e-9), same lr file
    snp = _snp(fixture_snp, 50000)
    outs = []
    for tag, kw in (("host", dict(exact_sr="in_scan")), ("dev", dict(device_post=True))):
        d = tmp_path / tag
        d.mkdir()
        res = ldw.perform_MI_computation(snp, e["hdw"], ldw.CdsVar(e["paint"], 3), lr_save_path=str(d / "lr.tsv"), sr_save_path=str(d / "sr.tsv"),
                                         plt_folder=str(d), lr_retain_links=1e4, max_blk_sz=1000, lr_links_approx=1e5, **kw)
        outs.append((res, (d / "lr.tsv").read_bytes(), api.read_ShortRangeLinks(str(d / "sr.tsv"))))
    (a, lra, sra), (b, lrb, srb) = outs
    assert lra == lrb and len(b.sr["MI"]) == 0 and len(a.sr["MI"]) > 0
    # rows are ordered by srp_max (decreasing, stable): links with mathematically equal srp_max (duplicate SNP patterns) may
    # swap when the two paths' values differ in the 13th digit, so the tables are compared as sets of rows
    def canon(t):
        o = np.lexsort((t["clust_c"], t["pos2"], t["pos1"]))
        return {k: np.asarray(t[k])[o] for k in ("clust_c", "pos1", "pos2", "clust1", "clust2", "len", "MI", "ARACNE", "srp_max")}
    for x, y in ((canon(a.sr_links_red), canon(b.sr_links_red)), (canon(sra), canon(srb))):
        for k in ("clust_c", "pos1", "pos2", "clust1", "clust2", "len", "MI", "ARACNE"):
            np.testing.assert_array_equal(x[k], y[k], err_msg=k)
        np.testing.assert_allclose(x["srp_max"], y["srp_max"], rtol=1e-6)
    assert np.all(np.diff(b.sr_links_red["srp_max"]) <= 0)


def test_repeated_scans_of_multi_kind_ragged_blocks_are_bit_identical():
    """The scan kernel's hand-over protocol (TMA -> expanders of both CTAs of a pair -> cta_group::2 MMAs -> epilogue, ring
    of stages re-cut per tile kind) has no data race: repeated scans of blocks that mix every tile kind, a ragged range
    and 3 K-blocks give bit-identical thresholds, short-range MI and long-range tables (tools/stress_determinism.py runs
    the same check longer; a withdrawn kernel variant failed it in one run of four)."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy = synth.generate(nseq=300, nsnp=2700, seed=11)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000)
    first = None
    for _ in range(6):
        res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(sy.paint, 3), sr_dist=20000, lr_retain_links=5e4, max_blk_sz=1000,
                                         lr_links_approx=lra, write_tsv=False)
        cur = (res.thr.copy(), res.sr["MI"].copy(), res.lr["MI"].copy(), res.lr["pos1"].copy(), res.lr["pos2"].copy())
        if first is None:
            first = cur
            continue
        np.testing.assert_array_equal(first[0], cur[0])
        for a, b in zip(first[1:], cur[1:]):
            np.testing.assert_array_equal(a, b)
