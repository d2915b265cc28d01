"""lr_links.tsv writer (host only, no GPU): ldw_write_lr_tsv / ldw_format_r_real against the oracle's restatement of
write.table's cell encoding (R/computePairwiseMI.R:362) and against known R outputs."""
import numpy as np

import ldw_oracle as O
from ldweaver_b200 import api

# print(x, digits = 15) / write.table cells as R produces them
KNOWN = {100000.0: "1e+05", 123456.0: "123456", 0.1: "0.1", 1e-5: "1e-05", 0.0001: "1e-04", 1234567.1: "1234567.1",
         1e15: "1e+15", 20000.0: "20000", 0.5: "0.5", 1 / 3: "0.333333333333333", 0.1 + 0.2: "0.3", 123456789.0: "123456789",
         0.0001234: "0.0001234", 0.00001234: "1.234e-05", -2.5: "-2.5", 1.0: "1", 2221315.0: "2221315", 3e5: "3e+05",
         1110657.5: "1110657.5", 1.23456789012345e-5: "1.23456789012345e-05", 100001.0: "100001", 1e6: "1e+06", 0.0: "0"}


def test_cell_encoder_known_values():
    for x, want in KNOWN.items():
        assert api.format_r_real_native(x) == want, (x, api.format_r_real_native(x), want)
        assert O.format_r_numeric(x) == want
        assert api._format_r(x) == want


def test_cell_encoder_native_equals_oracle_random():
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.random(30000) * 10.0 ** rng.integers(-9, 9, 30000), rng.integers(0, 3_000_000, 5000).astype(float),
                         -rng.random(100), [1e-300, 1e300, 5e-324, 999999999999999.9, 99999.5]])
    for x in xs:
        assert api.format_r_real_native(float(x)) == O.format_r_numeric(float(x)), repr(x)
        assert float(api.format_r_real_native(float(x))) == float(f"{float(x):.15g}")      # 15 significant digits survive


def test_write_lr_tsv_rows_and_append(tmp_path):
    rng = np.random.default_rng(4)
    n = 200_000
    lr = {"pos1": rng.integers(1, 2_221_315, n).astype(np.int32), "pos2": rng.integers(1, 2_221_315, n).astype(np.int32),
          "clust1": rng.integers(1, 4, n).astype(np.int32), "clust2": rng.integers(1, 4, n).astype(np.int32),
          "len": (rng.integers(2, 111, n) * 10000).astype(np.int32), "MI": rng.random(n) * 10.0 ** rng.integers(-7, 0, n),
          "block": np.zeros(n, np.int32)}
    lr["pos1"][:3] = (100000, 2000000, 1200000)    # round positions take the scientific form, as in R
    p = str(tmp_path / "lr_links.tsv")
    api.write_lr_tsv(p, lr, append=True)          # the reference appends (file may not exist yet)
    api.write_lr_tsv(p, {k: v[:5] for k, v in lr.items()}, append=True)
    rows = open(p).read().split("\n")
    assert rows[-1] == "" and len(rows) == n + 5 + 1
    for i in list(rng.integers(0, n, 300)) + [0, n - 1]:
        # pos1 / pos2 are doubles in the reference's data.frame (as.numeric(POS), R/computePairwiseMI.R:176-177): 100000 -> 1e+05
        want = "\t".join([O.format_r_numeric(lr["pos1"][i]), O.format_r_numeric(lr["pos2"][i]), O.format_r_numeric(lr["clust1"][i]),
                          O.format_r_numeric(lr["clust2"][i]), O.format_r_numeric(lr["len"][i]), O.format_r_numeric(lr["MI"][i])])
        assert rows[i] == want
    assert rows[n] == rows[0]
    assert [r.split("\t")[0] for r in rows[:3]] == ["1e+05", "2e+06", "1200000"]
    back = np.loadtxt(p, delimiter="\t")           # what read.table / read_LongRangeLinks would parse (R/io_functions.R:34-35)
    assert back.shape == (n + 5, 6)
    assert np.array_equal(back[:n, 4], lr["len"]) and np.allclose(back[:n, 5], lr["MI"], rtol=1e-14, atol=0)
