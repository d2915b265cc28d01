"""GPU parity tests (through the C ABI): encoding and Hamming-distance weights vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    import ldweaver_b200 as ldw
    return ldw


_lib_ = _lib


def test_encode_fixture_bit_exact(fixture_input, fixture_expected):
    import ldw_oracle as O
    ldw = _lib()
    aln, pos = fixture_input["aln"], fixture_input["pos"]
    for method in ("default", "relaxed"):
        snp = ldw.snp_dat_from_alignment_matrix(aln, fixture_input["names"], pos=pos, method=method)
        assert snp.nsnp == 1268 and snp.nseq == 400 and snp.g is None
        np.testing.assert_array_equal(snp.POS, fixture_expected[f"{method}_POS"])
        np.testing.assert_array_equal(snp.r, fixture_expected[f"{method}_r"])
        np.testing.assert_array_equal(snp.uqe, fixture_expected[f"{method}_uqe"])
        np.testing.assert_array_equal(snp.codes, fixture_expected["codes"])


@pytest.fixture(scope="module")
def ref_expected():
    import os
    from conftest import GOLDEN
    return dict(np.load(os.path.join(GOLDEN, "ref_expected.npz")))


def test_encode_vs_reference_object_code_goldens(fixture_input, ref_expected):
    """a1/a2 against tests/golden/ref_expected.npz = outputs of the reference's OWN compiled .extractAlnParam /
    .extractSNPs (oracle/_ref, tests/golden/make_golden_ref.py) on the bundled fixture, six (filter, gap, maf)."""
    ldw = _lib()
    e = ref_expected
    aln = fixture_input["aln"]
    k = 0
    while f"enc{k}_setting" in e:
        filt, gap, maf = e[f"enc{k}_setting"]
        method = "default" if filt == 0 else "relaxed"
        snp = ldw.snp_dat_from_alignment_matrix(aln, fixture_input["names"], method=method, gap_freq=float(gap), maf_freq=float(maf))
        assert snp.nseq == int(e[f"enc{k}_num_seqs"]) and snp.g == int(e[f"enc{k}_seq_length"])
        np.testing.assert_array_equal(snp.POS, e[f"enc{k}_pos"])
        np.testing.assert_array_equal(snp.codes, e[f"enc{k}_codes"])
        np.testing.assert_array_equal(snp.uqe, (e[f"enc{k}_table"] > 0).T.astype(float))
        k += 1
    assert k == 6


def test_whole_parse_chain_on_the_crlf_edge_file_vs_reference_goldens(ref_expected, tmp_path):
    """File bytes -> ldw_read_fasta_alloc (kseq grammar: CR columns kept) -> device count/filter/gather, against what the
    reference's compiled code returned for the very same file."""
    ldw = _lib()
    e = ref_expected
    p = tmp_path / "edge.fa"
    p.write_bytes(e["edge_file"].tobytes())
    for k, (method, gap, maf) in enumerate([("default", 0.15, 0.01), ("relaxed", 0.15, 0.01), ("default", 0.05, 0.10),
                                            ("relaxed", 0.05, 0.10)]):
        if int(e[f"edge{k}_num_snps"]) == 0:
            with pytest.raises(ValueError, match="any SNPs"):
                ldw.parse_fasta_alignment(str(p), gap_freq=gap, maf_freq=maf, method=method)
            continue
        snp = ldw.parse_fasta_alignment(str(p), gap_freq=gap, maf_freq=maf, method=method)
        assert snp.g == int(e[f"edge{k}_seq_length"]) == 426 and snp.nseq == 60
        assert snp.seq_names == [str(x) for x in e["edge_names"]]
        np.testing.assert_array_equal(snp.POS, e[f"edge{k}_pos"])
        np.testing.assert_array_equal(snp.codes, e[f"edge{k}_codes"])
        np.testing.assert_array_equal(snp.uqe, (e[f"edge{k}_table"] > 0).T.astype(float))


def test_acgtn2num_vs_reference_object_code_golden(ref_expected):
    ldw = _lib()
    e = ref_expected
    cv = [chr(c) for c in e["a2n_cv"]]
    nv = np.ones((5, 256), order="F")
    ldw.acgtn2num(nv, cv)
    np.testing.assert_array_equal(nv, e["a2n_nv"])


def test_encode_filters_vs_oracle_ragged():
    """Filters that bite, lower case, gaps, IUPAC, odd sizes (unaligned rows)."""
    import c_oracle as CO
    import ldw_oracle as O
    ldw = _lib()
    rng = np.random.default_rng(11)
    for (S, L) in ((37, 1001), (128, 4096), (301, 777)):
        alphabet = np.frombuffer(b"ACGTacgtNn-RYK", dtype=np.uint8)
        base = rng.integers(0, 4, size=L)
        aln = alphabet[base][None, :].repeat(S, axis=0).copy()
        mut = rng.random((S, L)) < rng.random(L)[None, :] * 0.3
        aln[mut] = alphabet[rng.integers(0, len(alphabet), size=int(mut.sum()))]
        for filt, method in ((0, "default"), (1, "relaxed")):
            for gap, maf in ((0.15, 0.01), (0.05, 0.1)):
                pos_c, counts_c = CO.aln_param(aln, filt, gap, maf)
                if len(pos_c) == 0:
                    with pytest.raises(ValueError):
                        ldw.snp_dat_from_alignment_matrix(aln, method=method, gap_freq=gap, maf_freq=maf)
                    continue
                snp = ldw.snp_dat_from_alignment_matrix(aln, method=method, gap_freq=gap, maf_freq=maf)
                np.testing.assert_array_equal(snp.POS, pos_c)
                codes_c, table_c = CO.extract_snps(aln, pos_c)
                np.testing.assert_array_equal(snp.codes, codes_c)
                np.testing.assert_array_equal(snp.uqe, (table_c > 0).T.astype(float))
                assert snp.g == L


def test_encode_streams_in_row_chunks_and_separate_entry_points_agree(monkeypatch):
    """The alignment goes to the device in row chunks (8 GiB by default; a few kilobytes here): counts accumulate across
    chunks, every chunk's classes land in its own sequence range.  The two reference-shaped entry points (ldw_aln_param,
    ldw_extract_snps: what `.extractAlnParam` / `.extractSNPs` map to) must agree with the one-call encoder, for row
    lengths that are and are not multiples of 16 (rows are padded to 16 bytes on the device)."""
    import ctypes as C
    import c_oracle as CO
    from ldweaver_b200 import _lib
    ldw = _lib_()
    rng = np.random.default_rng(5)
    for (S, L) in ((300, 1000), (517, 1601), (64, 16), (1, 33)):
        alphabet = np.frombuffer(b"ACGTacgtNn-RYK", dtype=np.uint8)
        aln = alphabet[rng.integers(0, 4, size=L)][None, :].repeat(S, axis=0).copy()
        mut = rng.random((S, L)) < rng.random(L)[None, :] * 0.4
        aln[mut] = alphabet[rng.integers(0, len(alphabet), size=int(mut.sum()))]
        pos_c, counts_c = CO.aln_param(aln, 1, 0.3, 0.0)
        if len(pos_c) == 0:
            continue
        codes_c, table_c = CO.extract_snps(aln, pos_c)
        for chunk in (None, 7 * ((L + 15) // 16 * 16), 1):     # one chunk / 7 rows per chunk / one row per chunk
            if chunk is None:
                monkeypatch.delenv("LDW_ENCODE_CHUNK_BYTES", raising=False)
            else:
                monkeypatch.setenv("LDW_ENCODE_CHUNK_BYTES", str(chunk))
            snp = ldw.snp_dat_from_alignment_matrix(aln, method="relaxed", gap_freq=0.3, maf_freq=0.0)
            np.testing.assert_array_equal(snp.POS, pos_c)
            np.testing.assert_array_equal(snp.codes, codes_c)
            np.testing.assert_array_equal(snp.uqe, (table_c > 0).T.astype(float))
            # the separate entry points
            Lb = _lib.lib()
            ctx = _lib.default_context(0)
            pos = np.empty(L, dtype=np.int32)
            n = C.c_int64()
            counts = np.empty(5 * L, dtype=np.float64)
            _lib.check(Lb.ldw_aln_param(ctx.handle, _lib.ptr(aln), S, L, 1, 0.3, 0.0, _lib.ptr(pos), C.byref(n), _lib.ptr(counts)))
            np.testing.assert_array_equal(pos[:n.value], pos_c)
            np.testing.assert_array_equal(counts.reshape(L, 5).T, counts_c)
            codes = np.empty((n.value, S), dtype=np.uint8)
            table = np.empty(5 * n.value, dtype=np.float64)
            _lib.check(Lb.ldw_extract_snps(ctx.handle, _lib.ptr(aln), S, L, _lib.ptr(pos_c), n.value, _lib.ptr(codes), _lib.ptr(table)))
            np.testing.assert_array_equal(codes, codes_c)
            np.testing.assert_array_equal(table.reshape(n.value, 5).T, table_c)
    monkeypatch.delenv("LDW_ENCODE_CHUNK_BYTES", raising=False)


def test_acgtn2num_vs_oracle():
    import c_oracle as CO
    ldw = _lib()
    cv = ("ACGTN-acgtnXR" * 50)[:601]
    nv = np.ones((5, len(cv)), order="F")
    nv2 = nv.copy(order="F")
    ldw.acgtn2num(nv, list(cv))
    CO.acgtn2num(nv2, cv.encode())
    np.testing.assert_array_equal(nv, nv2)


def test_hdw_fixture_bit_exact(fixture_snp, fixture_expected):
    ldw = _lib()
    snp = ldw.snp_dat_from_codes(fixture_snp.codes, fixture_snp.POS, 50000)
    hdw, cnt, dist = ldw.estimate_Hamming_distance_weights(snp, 0.1, return_parts=True)
    np.testing.assert_array_equal(dist, fixture_expected["hdw_dist"])
    np.testing.assert_array_equal(cnt, fixture_expected["hdw_cnt"])
    np.testing.assert_array_equal(hdw, fixture_expected["hdw"])
    # fused path (no distance matrix requested) must give the same weights
    hdw2 = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    np.testing.assert_array_equal(hdw2, fixture_expected["hdw"])
    # Q7: threshold 0 -> all weights 1
    assert np.all(ldw.estimate_Hamming_distance_weights(snp, 0.0) == 1.0)


@pytest.mark.parametrize("S,n,seed", [(130, 900, 1), (616, 3000, 2), (1000, 5000, 3), (2500, 1200, 4)])
def test_hdw_synthetic_vs_c_oracle(S, n, seed):
    """Multi-allelic, N-rich synthetic codes; sizes that exercise split-K, the fused epilogue (several
    column tiles) and padding."""
    import c_oracle as CO
    ldw = _lib()
    rng = np.random.default_rng(seed)
    founders = rng.integers(0, 5, size=(n, 12)).astype(np.uint8)
    assign = rng.integers(0, 12, size=S)
    codes = founders[:, assign]
    flip = rng.random((n, S)) < 0.04
    codes[flip] = rng.integers(0, 5, size=int(flip.sum()))
    codes[: n // 10] = codes[: n // 10] % 2          # biallelic block
    codes[n // 10: n // 10 + 7] = 3                  # monomorphic sites (own no plane)
    w_ref, cnt_ref, dist_ref = CO.hdw(codes, 0.1, want_dist=True)
    snp = ldw.snp_dat_from_codes(codes, np.arange(1, n + 1), n)
    w, cnt, dist = ldw.estimate_Hamming_distance_weights(snp, 0.1, return_parts=True)
    np.testing.assert_array_equal(dist, dist_ref)
    np.testing.assert_array_equal(cnt, cnt_ref)
    np.testing.assert_array_equal(w, w_ref)
    w2 = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    np.testing.assert_array_equal(w2, w_ref)
