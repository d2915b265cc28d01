"""Generates tests/golden/post_expected.npz: the post-scan chain (mergeNsort_sr_links, runARACNE, ordering; and the
long-range chain of analyse_long_range_links) of the NumPy/SciPy oracle (oracle/post_oracle.py) on the golden short-range
/ long-range tables of the reference's fixture (fixture_expected.npz, g = 50 000 and 2 221 315, max_blk_sz = 1000).

    python tests/golden/make_golden_post.py

Needs nothing outside the repository.  "Parity unpinned": the reference holds no expected values for these steps; the
file pins OUR oracle's output so that the native code, the oracle itself and later rounds are held to the same numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import ldw_oracle as O  # noqa: E402
import post_oracle as PO  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def fixture_sr(e):
    POS, paint = e["relaxed_POS"], e["paint"]
    p1, p2, MI = e["g50k_b1000_sr_pos1"], e["g50k_b1000_sr_pos2"], e["g50k_b1000_sr_MI"]
    lut = np.zeros(int(POS.max()) + 1, dtype=np.int32)
    lut[POS] = paint
    return dict(pos1=p1, pos2=p2, clust1=lut[p1], clust2=lut[p2], len=O.circ_len(p1.astype(float), p2.astype(float), 50000.0), MI=MI)


def main():
    e = dict(np.load(os.path.join(OUT, "fixture_expected.npz")))
    sr = fixture_sr(e)
    r = PO.merge_n_sort_sr_links(sr, 3, 20000.0, 3.0)
    d = r.df
    ar = PO.run_aracne(d["pos1"][r.red], d["pos2"][r.red], d["MI"][r.red], d["pos1"][r.chk], d["pos2"][r.chk], d["MI"][r.chk])
    order = PO.order_links_by_srp(d["srp_max"][r.red])
    out = dict(n_df=len(d["row"]), df_row_checksum=int(np.sum(d["row"] * (np.arange(len(d["row"])) % 1009 + 1))),
               red_row=d["row"][r.red][order], red_clust_c=d["clust_c"][r.red][order], red_srp_max=d["srp_max"][r.red][order],
               red_aracne=ar[order], n_chk=len(r.chk), coef=np.array([f.coef for f in r.fits]), shape=np.array([f.shape for f in r.fits]),
               start=np.array([f.start for f in r.fits]), n_pos=np.array([f.n_pos for f in r.fits]),
               nm_evals=np.array([f.nm_evals for f in r.fits]), n_len=np.array([len(f.len) for f in r.fits]),
               q95_sum=np.array([f.max.sum() for f in r.fits]))
    # long-range chain on the g = 2 221 315 tables
    lr = {k: e[f"g2M_b1000_lr_{k}"] for k in ("pos1", "pos2", "MI")}
    srl = dict(pos1=d["pos1"][r.red], pos2=d["pos2"][r.red], MI=d["MI"][r.red])
    idx, lar, thr = PO.analyse_long_range_links(lr, srl)
    out.update(lr_idx=idx, lr_aracne=lar, lr_thresholds=thr)
    np.savez_compressed(os.path.join(OUT, "post_expected.npz"), **out)
    print({k: (np.asarray(v).shape, np.asarray(v).dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
