"""Times the native post-scan steps (ldw_sr_postprocess, ldw_run_aracne, ldw_write_sr_tsv) on a synthetic short-range
table of the size the 616 x 100k scan produces (9.0e7 links, 3 clusters, sr_dist 20000).  Host only; prints one JSON
line.  Usage: python tools/bench_post.py [n_links]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldweaver_b200 as ldw  # noqa: E402
from ldweaver_b200 import _lib  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 90_000_000
    g, sr_dist = 2_221_315, 20000
    rng = np.random.default_rng(1)
    ln = rng.integers(1, sr_dist + 1, n, dtype=np.int32)
    pos1 = rng.integers(1, g - sr_dist, n, dtype=np.int32)
    pos1.sort()
    pos2 = pos1 + ln
    third = g // 3
    c1 = (np.minimum(pos1 // third, 2) + 1).astype(np.int32)
    c2 = (np.minimum(pos2 // third, 2) + 1).astype(np.int32)
    mi = (rng.beta(0.9, 25.0, n) * (ln.astype(np.float64) ** -0.25)).astype(np.float32).astype(np.float64)  # fp32-valued like the scan's
    sr = _lib.Links.from_dict(dict(pos1=pos1, pos2=pos2, clust1=c1, clust2=c2, len=ln, MI=mi))
    t0 = time.perf_counter()
    post = ldw.mergeNsort_sr_links(ldw.CdsVar(None, 3), sr, float(sr_dist), None, 3.0)
    t1 = time.perf_counter()
    red = {k: post.df[k][post.red] for k in ("pos1", "pos2", "MI")}
    chk = {k: post.df[k][post.chk] for k in ("pos1", "pos2", "MI")}
    ar = ldw.runARACNE(red, chk)
    t2 = time.perf_counter()
    with tempfile.TemporaryDirectory() as d:
        ldw.write_sr_tsv(os.path.join(d, "sr_links.tsv"), sr, post.df["row"][post.red], post.df["clust_c"][post.red],
                         post.df["srp_max"][post.red], ar.astype(np.float64), append=False)
        t3 = time.perf_counter()
        size = os.path.getsize(os.path.join(d, "sr_links.tsv"))
    print(json.dumps({"n_sr_links": n, "host_threads": os.cpu_count(), "mergeNsort_sr_links_s": round(t1 - t0, 3),
                      "links_per_s": round(n / (t1 - t0)), "n_df": int(len(post.df["row"])), "n_red": int(len(post.red)),
                      "n_aracne_check": int(len(post.chk)), "runARACNE_s": round(t2 - t1, 3), "aracne_kept": int(ar.sum()),
                      "write_sr_tsv_s": round(t3 - t2, 3), "tsv_bytes": size,
                      "nm_evals": [f["nm_evals"] for f in post.fits], "shape": [f["shape"].tolist() for f in post.fits]}))


if __name__ == "__main__":
    main()
