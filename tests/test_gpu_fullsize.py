"""Full-size runs of the BASELINE.json configurations on the GPU, checked through size-independent properties and
oracle spot checks (the oracle cannot finish a 10^4 x 10^4 block in test time, let alone 5 x 10^9 pairs).

 * C2 (616 x 100k, the benchmark workload): pair / short-range / long-range bookkeeping against an independent
   NumPy count from the positions, per-block kept counts against the type-7 rank arithmetic, every kept long-range
   link >= its block threshold, sampled links against the per-pair closed form of the oracle (MI within 1e-6,
   integer columns exact), two-rank partition == single run bit for bit (also a determinism check).
 * C4-shaped (5000 sequences: 40 K-blocks per tile, the tensor-heavy regime) and C5-shaped (2000 sequences,
   N-rich, multi-allelic, SNP-only input with lowercase letters through the encoder): sampled links vs the oracle.
 * C3 (10 000 x 50 000, weights only): neighbour counts / distances of sampled sequences against a NumPy count,
   symmetry of the distance matrix.
"""
import numpy as np
import pytest

import ldw_oracle as O

pytestmark = pytest.mark.gpu

SR_DIST = 20000.0


def _sample_check(snp_o, hdw, links, blocks, blk_ids, n_per_block, rng, tol=1e-6):
    """links: dict of columns; for a few blocks compare sampled rows with the oracle's per-pair closed form."""
    worst = 0.0
    pos = snp_o.POS
    for b in blk_ids:
        rows = np.nonzero(links["block"] == b)[0]
        if len(rows) == 0:
            continue
        pick = rng.choice(rows, size=min(n_per_block, len(rows)), replace=False)
        fs, fe, ts, te = blocks[b]
        f = np.arange(fs - 1, fe)
        t = np.arange(ts - 1, te)
        # pos2 = row ("from") SNP, pos1 = column ("to") SNP (R/computePairwiseMI.R:319-320); diagonal: pos1 < pos2
        il = np.searchsorted(pos[f], links["pos2"][pick])
        jl = np.searchsorted(pos[t], links["pos1"][pick])
        assert np.array_equal(pos[f][il], links["pos2"][pick]) and np.array_equal(pos[t][jl], links["pos1"][pick])
        ref = O.pair_mi_closed_form(snp_o, hdw, f, t, il, jl)
        worst = max(worst, float(np.abs(ref - links["MI"][pick]).max()))
        ln = O.circ_len(links["pos1"][pick].astype(np.float64), links["pos2"][pick].astype(np.float64), float(snp_o.g))
        assert np.array_equal(ln.astype(np.int64), links["len"][pick].astype(np.int64))
    assert worst < tol, f"max |MI - oracle| = {worst}"
    return worst


def _expected_sr_pairs(POS, g, blk, sr):
    """Independent count: emitted pairs (quirk Q2 drops equal-local-index pairs of off-diagonal blocks) and
    short-range pairs (len <= sr_dist, quirk Q4) from the positions alone."""
    P = np.asarray(POS, dtype=np.int64)
    n = len(P)
    total = n * (n - 1) // 2
    hi = np.searchsorted(P, P + int(np.floor(sr)), side="right")
    sr_lin = int((hi - np.arange(n) - 1).sum())
    wrap = np.searchsorted(P, P + int(np.ceil(g - sr)), side="left")
    n_sr = sr_lin + int((n - wrap).sum())
    # Q2 losses: off-diagonal block (i, j): pairs (from[k], to[k]), k < min(nf, nt)
    nr = -(-n // blk)
    lost, lost_sr = 0, 0
    for i in range(nr):
        for j in range(i + 1, nr):
            m = min(min(n, (i + 1) * blk) - i * blk, min(n, (j + 1) * blk) - j * blk)
            a = P[i * blk:i * blk + m]
            b = P[j * blk:j * blk + m]
            d = np.abs(a - b) % g
            ln = np.minimum(d, g - d)
            lost += m
            lost_sr += int((ln <= sr).sum())
    return total - lost, n_sr - lost_sr


@pytest.fixture(scope="module")
def c2():
    from ldweaver_b200 import synth
    import ldweaver_b200 as ldw
    S, n, seed, probs, nrate = synth.CONFIGS["C2"]
    sy = synth.generate(S, n, seed, probs, nrate)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    return sy, snp, hdw


def test_c2_full_scan_properties_and_spot_checks(c2):
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api, synth
    sy, snp, hdw = c2
    blk = 10000
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, SR_DIST)
    plan = ldw.MIPlan(snp, hdw, sy.paint, blk)
    sr, lr, bd, thr, prob, st = plan.scan(sy.g, SR_DIST, 1e6, lra)
    # ---- bookkeeping against an independent count from the positions
    n_pairs, n_sr = _expected_sr_pairs(sy.POS, sy.g, blk, SR_DIST)
    assert st["n_pairs"] == n_pairs and st["n_sr"] == n_sr == len(sr["MI"])
    assert st["n_lr_total"] == n_pairs - n_sr
    assert st["n_lr_kept"] == len(lr["MI"])
    assert st["n_reruns"] <= 5   # a re-run costs time, never correctness; the seed comes from whichever block was selected last
    blocks = api.make_blocks(snp.nsnp, blk)
    # ---- per block: kept count follows the type-7 rank arithmetic, every kept link clears its threshold
    kept_by_block = np.bincount(lr["block"], minlength=len(blocks))
    for b, (fs, fe, ts, te) in enumerate(blocks):
        nf, nt = fe - fs + 1, te - ts + 1
        m = (nf * (nf - 1) // 2 if fs == ts else nf * nt - min(nf, nt)) - int((sr["block"] == b).sum())
        p = max(0.0, 1 - ((1e6 * (m / lra)) / m))
        assert abs(prob[b] - p) < 1e-15
        h = 1 + (m - 1) * p
        lo, hi = int(np.floor(h)), int(np.ceil(h))
        # values >= thr: at least those from rank hi upwards, at most those from rank lo upwards plus ties
        assert m - hi + 1 <= kept_by_block[b] <= m - lo + 1 + int((bd["block"] == b).sum())
        sel = lr["block"] == b
        assert lr["MI"][sel].min() >= thr[b]
        assert np.all(lr["len"][sel] > SR_DIST)
    assert np.all(sr["len"] <= SR_DIST)
    # reference row order: blocks ascending
    assert np.all(np.diff(sr["block"]) >= 0) and np.all(np.diff(lr["block"]) >= 0)
    # ---- clusters and lengths of a random sample of rows, MI of sampled rows vs the oracle
    rng = np.random.default_rng(11)
    snp_o = O.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    idx = rng.choice(len(sr["MI"]), 20000, replace=False)
    pos_to_i = {int(p): i for i, p in enumerate(sy.POS)}
    i1 = np.array([pos_to_i[int(p)] for p in sr["pos1"][idx]])
    i2 = np.array([pos_to_i[int(p)] for p in sr["pos2"][idx]])
    assert np.array_equal(sy.paint[i1], sr["clust1"][idx]) and np.array_equal(sy.paint[i2], sr["clust2"][idx])
    w_sr = _sample_check(snp_o, hdw, sr, blocks, [0, 1, 10, 54], 400, rng)
    w_lr = _sample_check(snp_o, hdw, lr, blocks, [0, 1, 7, 30, 53, 54], 300, rng, tol=1e-9)  # LR rows carry fp64 MI
    print(f"C2 full: sr max err {w_sr:.2e}, lr max err {w_lr:.2e}, stats {st}")
    plan.close()


def test_c2_partition_union_is_bitwise_single_run(c2):
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    sy, snp, hdw = c2
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, SR_DIST)
    plan = ldw.MIPlan(snp, hdw, sy.paint, 10000)
    whole = plan.scan(sy.g, SR_DIST, 1e6, lra)
    parts = [plan.scan(sy.g, SR_DIST, 1e6, lra, 0, 2, r) for r in range(2)]
    for which in (0, 1):  # sr, lr
        for col in ("pos1", "pos2", "clust1", "clust2", "len", "MI", "block"):
            cat = np.concatenate([p[which][col] for p in parts])
            blkcol = np.concatenate([p[which]["block"] for p in parts])
            order = np.argsort(blkcol, kind="stable")
            assert np.array_equal(cat[order], whole[which][col]), (which, col)
    thr = np.where(np.isnan(parts[0][3]), parts[1][3], parts[0][3])
    assert np.array_equal(thr, whole[3])
    plan.close()


from ldweaver_b200.synth import cheap_codes as _cheap_codes  # noqa: E402


def test_c4_shaped_5000_sequences_spot_check():
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api, synth
    S, n = 5000, 12000
    codes = _cheap_codes(S, n, 5000300, 0.01, (0.847, 0.147, 0.006))
    rng = np.random.default_rng(1)
    POS = np.sort(rng.choice(np.arange(1, synth.G_DEFAULT + 1), n, replace=False)).astype(np.int32)
    paint = (1 + np.arange(n) * 3 // n).astype(np.int32)
    snp = ldw.snp_dat_from_codes(codes, POS, synth.G_DEFAULT)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    assert len(np.unique(hdw)) > 3
    lra = synth.exact_lr_links_approx(POS, synth.G_DEFAULT, SR_DIST)
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(paint, 3), sr_dist=SR_DIST, lr_retain_links=1e5, max_blk_sz=5000,
                                     lr_links_approx=lra, write_tsv=False)
    snp_o = O.snp_dat_from_codes(codes, POS, synth.G_DEFAULT)
    blocks = api.make_blocks(n, 5000)
    w_sr = _sample_check(snp_o, hdw, res.sr, blocks, range(len(blocks)), 150, rng)
    w_lr = _sample_check(snp_o, hdw, res.lr, blocks, range(len(blocks)), 150, rng, tol=1e-9)
    print(f"C4-shaped: sr max err {w_sr:.2e}, lr max err {w_lr:.2e}, reruns {res.stats['n_reruns']}")


def test_c5_shaped_snp_only_nrich_multiallelic_spot_check():
    import c_oracle as CO
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api, synth
    S, n = 2000, 30000
    codes = _cheap_codes(S, n, 2000500, 0.10, (0.60, 0.30, 0.10))
    rng = np.random.default_rng(2)
    aln = synth.codes_to_alignment(codes, lowercase_frac=0.3)            # [S, n] ASCII, gaps and lowercase
    pos_in = np.sort(rng.choice(np.arange(1, synth.G_DEFAULT + 1), n, replace=False)).astype(np.int32)
    snp = ldw.snp_dat_from_alignment_matrix(aln, pos=pos_in, method="relaxed")   # SNP-only style: positions given
    snp.g = synth.G_DEFAULT
    keep_c, _ = CO.aln_param(aln, 1, 0.15, 0.01)
    assert np.array_equal(snp.POS, pos_in[keep_c - 1])
    codes_c, _ = CO.extract_snps(aln, keep_c)
    assert np.array_equal(snp.codes, codes_c)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    paint = (1 + np.arange(snp.nsnp) * 3 // snp.nsnp).astype(np.int32)
    lra = synth.exact_lr_links_approx(snp.POS, snp.g, SR_DIST)
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(paint, 3), sr_dist=SR_DIST, lr_retain_links=2e5, max_blk_sz=10000,
                                     lr_links_approx=lra, write_tsv=False)
    snp_o = O.snp_dat_from_codes(codes_c, snp.POS, snp.g)
    assert np.bincount(snp_o.r.astype(int))[4:].sum() > 100          # plenty of 4- and 5-class sites
    blocks = api.make_blocks(snp.nsnp, 10000)
    w_sr = _sample_check(snp_o, hdw, res.sr, blocks, range(len(blocks)), 150, rng)
    w_lr = _sample_check(snp_o, hdw, res.lr, blocks, range(len(blocks)), 150, rng, tol=1e-9)
    print(f"C5-shaped: nsnp {snp.nsnp}, sr max err {w_sr:.2e}, lr max err {w_lr:.2e}, reruns {res.stats['n_reruns']}")


def test_c3_full_hamming_weights_10000_by_50000():
    import ldweaver_b200 as ldw
    S, n = 10000, 50000
    codes = _cheap_codes(S, n, 1000050, 0.01, (0.847, 0.147, 0.006))
    snp = ldw.snp_dat_from_codes(codes, np.arange(1, n + 1, dtype=np.int32), n)
    hdw, cnt, dist = ldw.estimate_Hamming_distance_weights(snp, 0.1, return_parts=True)
    thresh = int(n * 0.1)
    assert np.array_equal(dist, dist.T) and np.all(np.diag(dist) == 0)
    assert np.array_equal(cnt, (dist < thresh).sum(axis=0))                       # strict <, self included (Q7)
    assert np.array_equal(hdw, 1.0 / (cnt + 1.0))
    # the fused path (upper-triangle tiles, neighbours counted straight from TMEM, no distance matrix) gives the same weights
    assert np.array_equal(ldw.estimate_Hamming_distance_weights(snp, 0.1), hdw)
    rng = np.random.default_rng(3)
    for s in rng.choice(S, 12, replace=False):                                     # exact distances of sampled rows
        ref = (codes != codes[:, s][:, None]).sum(axis=0)
        assert np.array_equal(ref, dist[:, s])
    assert len(np.unique(cnt)) > 10


def _dense_cells_check(plan, snp_o, hdw, blocks, b, rng, n_cells=3000, tol=1e-6):
    """Sampled cells of one block's dense MI matrix (the fp32 epilogue path that feeds the short-range slots) against the
    oracle's per-pair closed form."""
    fs, fe, ts, te = blocks[b]
    f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
    MI = plan.block_dense(b)
    il = rng.integers(0, len(f), n_cells)
    jl = rng.integers(0, len(t), n_cells)
    ref = O.pair_mi_closed_form(snp_o, hdw, f, t, il, jl)
    err = float(np.abs(MI[il, jl] - ref).max())
    assert err < tol, f"block {b}: max |MI - oracle| = {err}"
    return err


def _full_size_scan_checks(tag, snp, hdw, paint, codes, retain):
    """Whole scan at a BASELINE shape whose short-range table is too large to bring to the host (C4: 8.2e8 rows, C5: 2.3e9):
    LDW_SCAN_LR_ONLY.  Checked: pair / short-range / long-range counts against an independent count from the positions,
    per-block kept counts against the type-7 rank arithmetic, every kept link >= its threshold and longer than sr_dist,
    sampled long-range rows (fp64) and sampled cells of dense blocks (fp32 epilogue) against the oracle's closed form."""
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api, synth
    blk = 10000
    lra = synth.exact_lr_links_approx(snp.POS, snp.g, SR_DIST)
    plan = ldw.MIPlan(snp, hdw, paint, blk)
    sr, lr, bd, thr, prob, st = plan.scan(snp.g, SR_DIST, retain, lra, api.SCAN_LR_ONLY)
    n_pairs, n_sr = _expected_sr_pairs(snp.POS, snp.g, blk, SR_DIST)
    assert st["n_pairs"] == n_pairs and st["n_sr"] == n_sr and len(sr["MI"]) == 0
    assert st["n_lr_total"] == n_pairs - n_sr and st["n_lr_kept"] == len(lr["MI"])
    blocks = api.make_blocks(snp.nsnp, blk)
    assert st["n_blocks"] == len(blocks)
    kept = np.bincount(lr["block"], minlength=len(blocks))
    has_lr = ~np.isnan(thr)
    assert np.all(kept[~has_lr] == 0) and np.all(kept[has_lr] > 0)
    thr_of_row = thr[lr["block"]]
    assert np.all(lr["MI"] >= thr_of_row) and np.all(lr["len"] > SR_DIST)
    assert np.all(np.diff(lr["block"]) >= 0)
    assert abs(len(lr["MI"]) - retain) < 0.02 * retain + len(bd["MI"]) + 2 * len(blocks)   # ~lr_retain_links rows in total
    rng = np.random.default_rng(21)
    snp_o = O.snp_dat_from_codes(codes, snp.POS, snp.g)
    pick = sorted(set([0, 1, len(blocks) // 3, len(blocks) // 2, len(blocks) - 2, len(blocks) - 1]))
    w_lr = _sample_check(snp_o, hdw, lr, blocks, pick, 200, rng, tol=1e-9)
    nr = -(-snp.nsnp // blk)
    w_d = max(_dense_cells_check(plan, snp_o, hdw, blocks, b, rng) for b in (0, 1, nr))   # diagonal, off-diagonal, next diagonal
    print(f"{tag} full: {snp.nseq} x {snp.nsnp}, {st['n_pairs']:.3e} pairs, {len(blocks)} blocks, kept {len(lr['MI'])}, lr max err {w_lr:.2e}, "
          f"dense (fp32) max err {w_d:.2e}, reruns {st['n_reruns']}, scan {st['t_scan_ms'] + st['t_select_ms']:.0f} ms")
    plan.close()


def test_c4_full_5000_by_300000():
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    S, n, seed, probs, nrate = synth.CONFIGS["C4"]
    codes = _cheap_codes(S, n, seed, nrate, probs)
    rng = np.random.default_rng(seed)
    POS = np.sort(rng.choice(np.arange(1, synth.G_DEFAULT + 1), n, replace=False)).astype(np.int32)
    paint = (1 + np.arange(n) * 3 // n).astype(np.int32)
    snp = ldw.snp_dat_from_codes(codes, POS, synth.G_DEFAULT)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    assert len(np.unique(hdw)) > 3
    _full_size_scan_checks("C4", snp, hdw, paint, codes, 1e6)


def test_c5_full_2000_by_500000_snp_only_through_the_encoder():
    import ldweaver_b200 as ldw
    from ldweaver_b200 import api, synth
    S, n, seed, probs, nrate = synth.CONFIGS["C5"]
    codes = _cheap_codes(S, n, seed, nrate, probs)
    rng = np.random.default_rng(seed)
    pos_in = np.sort(rng.choice(np.arange(1, synth.G_DEFAULT + 1), n, replace=False)).astype(np.int32)
    aln = synth.codes_to_alignment(codes, lowercase_frac=0.3)             # [S, n] ASCII: gaps, N, lowercase
    snp = ldw.snp_dat_from_alignment_matrix(aln, pos=pos_in, method="relaxed")   # SNP-only input: positions given, g from the annotation
    del aln
    snp.g = synth.G_DEFAULT
    assert snp.nsnp > 0.9 * n
    keep = np.searchsorted(pos_in, snp.POS)
    assert np.array_equal(pos_in[keep], snp.POS) and np.array_equal(snp.codes, codes[keep])   # the encoder reproduces the generator's classes
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    paint = (1 + np.arange(snp.nsnp) * 3 // snp.nsnp).astype(np.int32)
    _full_size_scan_checks("C5", snp, hdw, paint, snp.codes, 1e6)
    # "long-range links fed to runARACNE" (R/lr_analyser.R:72-108) on the full-size output
    lra = synth.exact_lr_links_approx(snp.POS, snp.g, SR_DIST)
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(paint, 3), sr_dist=SR_DIST, lr_retain_links=1e6, lr_links_approx=lra,
                                     write_tsv=False, scan_flags=api.SCAN_LR_ONLY)
    lr = {"pos1": res.lr["pos1"].astype(float), "pos2": res.lr["pos2"].astype(float), "c1": res.lr["clust1"], "c2": res.lr["clust2"],
          "len": res.lr["len"].astype(float), "MI": res.lr["MI"]}
    out = api.analyse_long_range_links(lr, {k: np.zeros(0) for k in ("pos1", "pos2", "MI")})
    assert len(out["MI"]) >= 4000 and np.all(np.diff(out["MI"]) <= 0) and set(np.unique(out["ARACNE"])) <= {0, 1, True, False}
    print(f"C5 full: {len(res.lr['MI'])} long-range links -> {len(out['MI'])} above the Tukey threshold, ARACNE keeps {int(np.sum(out['ARACNE']))}")
