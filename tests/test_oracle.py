"""CPU tests: the two oracle restatements against each other, against the closed form of
SURVEY.md 8a, against the survey-time probe values and against the committed golden fixture."""
import numpy as np
import pytest

import c_oracle as CO
import ldw_oracle as O
from util import compare_lr_sets


def test_probe_values_on_reference_fixture(fixture_input, fixture_expected):
    """SURVEY.md 8(c) probe values, re-derived from the bundled fixture bytes."""
    aln, pos = fixture_input["aln"], fixture_input["pos"]
    assert aln.shape == (400, 1268)
    seqs = [bytes(r) for r in aln]
    for filt in (0, 1):
        par = O.extract_aln_param(fixture_input["names"], seqs, filt, 0.15, 0.01)
        assert par["num.snps"] == 1268 and par["num.seqs"] == 400
    data = O.extract_snps(seqs, 400, 1268, par["pos"])
    snp = O.snp_dat_from_codes(data["codes"], pos[np.array(par["pos"]) - 1], 50000)
    vals, counts = np.unique(snp.r, return_counts=True)
    assert dict(zip(vals.astype(int), counts)) == {2: 1074, 3: 187, 4: 7}
    hdw, cnt, dist, thresh = O.estimate_Hamming_distance_weights(snp, 0.1, True)
    assert thresh == 126
    assert abs(hdw.sum() - 30.1688663309) < 1e-9
    assert hdw.min() == 1 / 86 and hdw.max() == 0.5 and len(np.unique(hdw)) == 30
    assert dist.max() == 616
    np.testing.assert_array_equal(hdw, fixture_expected["hdw"])
    np.testing.assert_array_equal(data["codes"], fixture_expected["codes"])


def test_c_oracle_encoding_matches_numpy(fixture_input):
    aln = fixture_input["aln"].copy()
    rng = np.random.default_rng(5)
    # make the filters bite: inject gaps, monomorphic and rare-allele columns
    aln[:, 10] = ord("A")
    aln[rng.random(400) < 0.3, 20] = ord("-")
    aln[:, 30] = ord("C"); aln[:3, 30] = ord("t")
    aln[:, 40] = ord("g"); aln[:5, 40] = ord("T"); aln[5:9, 40] = ord("N")
    seqs = [bytes(r) for r in aln]
    for filt, gap, maf in ((0, 0.15, 0.01), (1, 0.15, 0.01), (0, 0.05, 0.05), (1, 0.5, 0.2)):
        par = O.extract_aln_param(["x"] * 400, seqs, filt, gap, maf)
        pos_c, counts_c = CO.aln_param(aln, filt, gap, maf)
        np.testing.assert_array_equal(np.array(par["pos"], dtype=np.int32), pos_c)
        np.testing.assert_array_equal(par["allele_counts"], counts_c)
        if len(pos_c):
            data = O.extract_snps(seqs, 400, len(pos_c), par["pos"])
            codes_c, table_c = CO.extract_snps(aln, pos_c)
            np.testing.assert_array_equal(data["codes"], codes_c)
            np.testing.assert_array_equal(data["ACGTN_table"], table_c)
    assert 11 not in par["pos"]


def test_filter_edge_rules():
    """Quirk Q9: integer truncation of min_maf, strict '<' gap test, second-largest rule."""
    n = 100
    def col(a=0, c=0, g=0, t=0, other=0):
        s = "A" * a + "C" * c + "G" * g + "T" * t + "-" * other
        assert len(s) == n
        return np.frombuffer(s.encode(), dtype=np.uint8)
    cols = [col(a=99, c=1), col(a=98, c=2), col(a=85, c=1, other=14), col(a=84, c=1, other=15),
            col(a=50, c=48, g=2), col(a=100), col(a=1, other=99), col(a=97, c=1, g=1, t=1)]
    aln = np.stack(cols, axis=1)
    seqs = [bytes(r) for r in aln]
    # default: min_maf = int(100*0.01) = 1 -> keep iff second largest of ACGT > 1
    par = O.extract_aln_param(["x"] * n, seqs, 0, 0.15, 0.01)
    assert par["pos"] == [2, 5]
    pos_c, _ = CO.aln_param(aln, 0, 0.15, 0.01)
    assert list(pos_c) == [2, 5]
    # relaxed: min_maf = int(100*0.99) = 99 -> keep iff max(all 5) <= 99 and gap/n < 0.15
    par = O.extract_aln_param(["x"] * n, seqs, 1, 0.15, 0.01)
    assert par["pos"] == [1, 2, 3, 5, 8]
    pos_c, _ = CO.aln_param(aln, 1, 0.15, 0.01)
    assert list(pos_c) == [1, 2, 3, 5, 8]


def test_unequal_lengths_and_empty():
    assert O.extract_aln_param(["a", "b"], [b"ACGT", b"ACG"], 0, 0.15, 0.01)["seq.length"] == -1
    assert O.extract_aln_param([], [], 0, 0.15, 0.01)["num.seqs"] == 0


def test_acgtn2num():
    cv = "ACGTN-acgtnXR"
    nv = np.ones((5, len(cv)), order="F")
    O.acgtn2num(nv, list(cv))
    nv2 = np.ones((5, len(cv)), order="F")
    CO.acgtn2num(nv2, cv.encode())
    np.testing.assert_array_equal(nv, nv2)
    assert nv[:, 0].tolist() == [0, 1, 1, 1, 1] and nv[:, 4].tolist() == [1, 1, 1, 1, 0]
    assert nv[:, 5].tolist() == [1, 1, 1, 1, 0] and nv[:, 6:].min() == 1  # lowercase / other untouched (Q8)


def test_hdw_c_vs_numpy_vs_golden(fixture_snp, fixture_expected):
    w, cnt, dist = CO.hdw(fixture_snp.codes, 0.1, want_dist=True)
    np.testing.assert_array_equal(w, fixture_expected["hdw"])
    np.testing.assert_array_equal(cnt, fixture_expected["hdw_cnt"])
    np.testing.assert_array_equal(dist, fixture_expected["hdw_dist"])
    # Q7: thresh = 0 -> all weights 1; isolated sequence -> 0.5
    w0, c0 = CO.hdw(fixture_snp.codes, 0.0)
    assert np.all(w0 == 1.0) and np.all(c0 == 0)
    assert O.estimate_Hamming_distance_weights(fixture_snp, 0.0).min() == 1.0


def test_block_mi_c_vs_numpy_vs_closed_form(fixture_snp, fixture_expected):
    snp, hdw = fixture_snp, fixture_expected["hdw"]
    rng = np.random.default_rng(0)
    cases = [(np.arange(0, 300), np.arange(0, 300)),        # diagonal
             (np.arange(0, 300), np.arange(300, 600)),      # square off-diagonal (Q1 with nf == nt)
             (np.arange(100, 400), np.arange(900, 1077)),   # ragged off-diagonal (Q1 scramble)
             (np.sort(rng.choice(1268, 150, replace=False)), np.sort(rng.choice(1268, 90, replace=False)))]
    for f, t in cases:
        lit = O.block_mi_matrix(snp, hdw, f, t)
        clo = O.block_mi_closed_form(snp, hdw, f, t)
        cc = CO.block_mi(snp.codes, hdw, snp.r, snp.uqe, f, t)
        assert np.abs(lit - clo).max() < 1e-12
        assert np.abs(lit - cc).max() < 1e-12


def test_pair_closed_form_matches_block_form(fixture_snp, fixture_expected):
    # the per-pair evaluator used by the full-size GPU spot checks is the block closed form cell by cell
    snp, hdw = fixture_snp, fixture_expected["hdw"]
    rng = np.random.default_rng(5)
    for f, t in [(np.arange(0, 300), np.arange(0, 300)), (np.arange(100, 400), np.arange(900, 1077)),
                 (np.arange(0, 300), np.arange(300, 600))]:
        blk = O.block_mi_closed_form(snp, hdw, f, t)
        il = rng.integers(0, len(f), 500)
        jl = rng.integers(0, len(t), 500)
        got = O.pair_mi_closed_form(snp, hdw, f, t, il, jl)
        assert np.abs(got - blk[il, jl]).max() < 1e-13


def test_golden_single_block_mi(fixture_snp, fixture_expected):
    e = fixture_expected
    idx = np.arange(fixture_snp.nsnp)
    MI = CO.block_mi(fixture_snp.codes, e["hdw"], fixture_snp.r, fixture_snp.uqe, idx, idx)
    assert np.abs(MI[e["MI_rows"], :] - e["MI_sub"]).max() < 1e-12
    off = ~np.eye(len(idx), dtype=bool)
    assert abs(MI[off].max() - 0.682443935150819) < 1e-12
    assert abs(MI[off].min() - 1.796e-13) < 1e-15
    assert np.abs(MI - MI.T).max() < 1e-14


def test_q1_quirk_is_material(fixture_snp, fixture_expected):
    """Off-diagonal non-square block: transposed-rft linear indexing (Q1) must change results."""
    snp, hdw = fixture_snp, fixture_expected["hdw"]
    f, t = np.arange(0, 1000), np.arange(1000, 1268)
    lit = O.block_mi_matrix(snp, hdw, f, t)
    # "ideal" Q = 0.25 r_i r_j
    w = hdw
    ideal = np.zeros_like(lit)
    neff = w.sum()
    rf, rt = snp.r[f], snp.r[t]
    den = neff + 0.5 * np.outer(rf, rt)
    for a in range(5):
        for b in range(5):
            Xa = (snp.codes[f] == a).astype(float); Xb = (snp.codes[t] == b).astype(float)
            pa, pb = Xa @ w, Xb @ w
            c = (Xa * w) @ Xb.T + 0.5
            D = np.outer(pa + 0.5 * rt.mean() * 0, pb) + 0.5 * (pa * rf)[:, None] + 0.5 * (pb * rt)[None, :] + 0.25 * np.outer(rf, rt)
            ideal += np.outer(snp.uqe[f, a], snp.uqe[t, b]) * c / den * np.log(c * den / D)
    dev = np.abs(lit - ideal)
    assert dev.max() > 1e-2 and (dev > 1e-6).mean() > 0.2


@pytest.mark.parametrize("tag,g,blk,retain", [("g50k_b10000", 50000, 10000, 1e4), ("g50k_b1000", 50000, 1000, 1e4),
                                               ("g2M_b1000", 2221315, 1000, 2e4)])
def test_links_c_vs_golden(fixture_snp, fixture_expected, tag, g, blk, retain):
    """Full scan through the C oracle (block MI + link enumeration/filter) equals the NumPy golden."""
    e = fixture_expected
    snp, hdw = fixture_snp, e["hdw"]
    POS = snp.POS.astype(np.float64)
    sr_p1, sr_p2, sr_mi, thr, npairs = [], [], [], [], []
    # golden LR rows are concatenated in block order; split them back per block through pos ranges
    g_p1, g_p2, g_mi = e[f"{tag}_lr_pos1"], e[f"{tag}_lr_pos2"], e[f"{tag}_lr_MI"]
    n_border = 0
    for bi, (fs, fe, ts, te) in enumerate(O.make_blocks(snp.nsnp, O.r_round_to_thousands(blk))):
        f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
        MI = CO.block_mi(snp.codes, hdw, snp.r, snp.uqe, f, t)
        L = CO.block_links(MI, POS, f, t, float(g), 20000.0, retain, 1e5)
        p2, p1 = POS[f[L["row"]]], POS[t[L["col"]]]
        sr_p1.append(p1[L["is_sr"]]); sr_p2.append(p2[L["is_sr"]]); sr_mi.append(L["MI"][L["is_sr"]])
        thr.append(L["thr"]); npairs.append(len(L["MI"]))
        if np.isnan(L["thr"]):
            assert np.isnan(e[f"{tag}_thr"][bi])
            continue
        assert abs(L["thr"] - e[f"{tag}_thr"][bi]) < 1e-12 and L["prob"] == e[f"{tag}_prob"][bi]
        in_blk = np.isin(g_p2, POS[f]) & np.isin(g_p1, POS[t])
        if fs != ts:  # off-diagonal block: from-range positions are all smaller than to-range positions
            in_blk &= (g_p2 < POS[t[0]]) & (g_p1 > POS[f[-1]])
        else:
            in_blk &= (g_p1 <= POS[f[-1]]) & (g_p2 <= POS[f[-1]]) & (g_p1 >= POS[f[0]]) & (g_p2 >= POS[f[0]])
        k = L["lr_keep"]
        _, a, b = compare_lr_sets(g_p1[in_blk], g_p2[in_blk], g_mi[in_blk], p1[k], p2[k], L["MI"][k], L["thr"])
        n_border += len(a) + len(b)
    cat = np.concatenate
    assert n_border < 0.02 * max(1, len(g_mi))
    np.testing.assert_array_equal(np.array(npairs), e[f"{tag}_npairs"])
    np.testing.assert_array_equal(cat(sr_p1).astype(np.int32), e[f"{tag}_sr_pos1"])
    np.testing.assert_array_equal(cat(sr_p2).astype(np.int32), e[f"{tag}_sr_pos2"])
    assert np.abs(cat(sr_mi)[::16] - e[f"{tag}_sr_MI_16"]).max() < 1e-12
    if tag == "g50k_b1000":  # Q2: 268 equal-local-index pairs dropped from the ragged block
        assert int(e[f"{tag}_npairs"].sum()) == 803010


def test_quantile_type7_and_blocks():
    x = np.array([3.0, 1.0, 2.0, 10.0])
    assert O.quantile_type7(x, 0.0) == 1.0 and O.quantile_type7(x, 1.0) == 10.0
    assert O.quantile_type7(x, 0.5) == 2.5
    assert abs(O.quantile_type7(x, 0.9) - np.quantile(x, 0.9)) < 1e-15
    assert O.make_blocks(2500, 1000) == [(1, 1000, 1, 1000), (1, 1000, 1001, 2000), (1, 1000, 2001, 2500),
                                         (1001, 2000, 1001, 2000), (1001, 2000, 2001, 2500), (2001, 2500, 2001, 2500)]
    assert O.r_round_to_thousands(10499) == 10000 and O.r_round_to_thousands(2500) == 2000  # half-to-even


def test_sr_only_mode_drops_far_snps(fixture_snp, fixture_expected):
    """Q12: SR-only mode first drops SNPs with no partner < sr_dist, which changes local indices."""
    snp, hdw = fixture_snp, fixture_expected["hdw"]
    snp_g = O.snp_dat_from_codes(snp.codes, snp.POS, 2221315)
    paint = fixture_expected["paint"]
    full = O.perform_MI_scan(snp_g, hdw, paint, 3, max_blk_sz=1000, lr_links_approx=1e5, sr_dist=2000)
    sro = O.perform_MI_scan(snp_g, hdw, paint, 3, max_blk_sz=1000, perform_SR_analysis_only=True, sr_dist=2000)
    assert len(sro.lr["MI"]) == 0
    assert 0 < len(sro.sr["MI"]) <= len(full.sr["MI"])


def test_r_rng_emulation_shape():
    rng = O.RMersenne(1988)
    s = rng.sample(1268, 127)
    assert len(np.unique(s)) == 127 and s.min() >= 1 and s.max() <= 1268
