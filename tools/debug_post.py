import os, sys
import numpy as np
ROOT = "/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ldweaver_b200 as ldw
from ldweaver_b200 import api
import ldw_oracle as O
e = dict(np.load(os.path.join(ROOT, "tests/golden/fixture_expected.npz")))
fs = O.snp_dat_from_codes(e["codes"], e["relaxed_POS"], 50000)
snp = ldw.snp_dat_from_codes(fs.codes, fs.POS, 50000)
cds = ldw.CdsVar(e["paint"], 3)
plan = ldw.MIPlan(snp, e["hdw"], e["paint"], 1000)
for flags in (0, api.SCAN_SR_EXACT):
    sr, lr, bd, thr, prob, st = plan.scan(50000.0, 20000.0, 1e4, 1e5, flags)
    mi = sr["MI"]
    print("flags", flags, "n", len(mi), "zeros", int((mi == 0).sum()), "nan", int(np.isnan(mi).sum()), "neg", int((mi < 0).sum()), "min", mi.min(), "max", mi.max())
    ref = e["g50k_b1000_sr_MI"]
    print("  max |dMI| vs golden", np.abs(mi - ref).max(), "clusters", np.unique(sr["clust1"]), np.unique(sr["clust2"]))
    for c in (1, 2, 3):
        m = ((sr["clust1"] == c) | (sr["clust2"] == c)) & (sr["len"] > 0) & (sr["len"] < 20000)
        print("  cluster", c, "links", int(m.sum()), "MI mean", mi[m].mean(), "min", mi[m].min())
    try:
        host = api.mergeNsort_sr_links(cds, sr, 20000.0, None, 3.0)
        print("  host post OK: df", len(host.df["row"]), "red", len(host.red), [f["n_pos"] for f in host.fits], [f["coef"].tolist() for f in host.fits])
    except Exception as ex:
        print("  host post FAILED:", ex)
        # which groups have non-positive q95?
        for c in (3,):
            m = ((sr["clust1"] == c) | (sr["clust2"] == c)) & (sr["len"] > 0) & (sr["len"] < 20000)
            L = sr["len"][m]; M = mi[m]
            import collections
            q = {}
            for l in np.unique(L)[:20000]:
                q[l] = np.quantile(M[L == l], 0.95)
            qs = np.array(list(q.values()))
            print("   cluster 3 groups", len(qs), "q95 <= 0:", int((qs <= 0).sum()), "min q95", qs.min())
plan.close()
