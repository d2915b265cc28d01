"""Generates tests/golden/*.npz from the reference's bundled fixture with the NumPy oracle.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference ships no expected outputs for this path ("parity unpinned"); these
files pin OUR oracle's outputs on the reference's own input fixture so that the C oracle, the
CUDA path and later rounds are all held to the same numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ldw_oracle as O  # noqa: E402

REF = "/root/reference/inst/extdata"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    names, seqs = O.read_fasta(os.path.join(REF, "snp_sample.fa.gz"))
    pos_all = np.loadtxt(os.path.join(REF, "snp_sample.pos"), dtype=np.int64)
    aln = np.stack([np.frombuffer(s, dtype=np.uint8) for s in seqs])
    np.savez_compressed(os.path.join(OUT, "fixture_input.npz"), aln=aln, pos=pos_all,
                        names=np.array(names))
    out = {}
    for method in ("default", "relaxed"):
        snp = O.parse_fasta_SNP_alignment(os.path.join(REF, "snp_sample.fa.gz"), pos_all, method=method)
        out[f"{method}_POS"] = snp.POS
        out[f"{method}_r"] = snp.r
        out[f"{method}_uqe"] = snp.uqe
    out["codes"] = snp.codes
    snp.g = 50000
    hdw, cnt, dist, thresh = O.estimate_Hamming_distance_weights(snp, 0.1, True)
    out.update(hdw=hdw, hdw_cnt=cnt, hdw_dist=dist.astype(np.int32), hdw_thresh=thresh)
    idx = np.arange(snp.nsnp)
    MI = O.block_mi_matrix(snp, hdw, idx, idx)
    out["MI_rows"] = np.arange(0, snp.nsnp, 6)
    out["MI_sub"] = MI[out["MI_rows"], :]  # every 6th row of the single-block MI matrix (212 x 1268 doubles)
    paint = np.ones(snp.nsnp, dtype=np.int64)
    paint[snp.nsnp // 3:] = 2
    paint[2 * snp.nsnp // 3:] = 3
    out["paint"] = paint
    lra = O.lr_links_approx_reference(snp.POS, 50000.0, 20000.0)
    out["lr_links_approx_g50000"] = lra
    for tag, g, blk, retain in (("g50k_b10000", 50000, 10000, 1e4), ("g50k_b1000", 50000, 1000, 1e4),
                                ("g2M_b1000", 2221315, 1000, 2e4)):
        snp.g = g
        res = O.perform_MI_scan(snp, hdw, paint, 3, max_blk_sz=blk, lr_retain_links=retain,
                                lr_links_approx=1e5, keep_blocks=True)
        out[f"{tag}_sr_n"] = len(res.sr["MI"])
        out[f"{tag}_lr_n"] = len(res.lr["MI"])
        for c in ("pos1", "pos2", "MI", "len", "clust1", "clust2"):
            out[f"{tag}_lr_{c}"] = res.lr[c]
        # SR is large (7e5 rows): keep MI + positions as float32/int32-friendly arrays
        # SR is large (5-7e5 rows): positions in full (they compress well), MI in full for the
        # multi-block case that exercises quirks Q1/Q2, every 16th value + checksums otherwise
        out[f"{tag}_sr_pos1"] = res.sr["pos1"].astype(np.int32)
        out[f"{tag}_sr_pos2"] = res.sr["pos2"].astype(np.int32)
        if tag == "g50k_b1000":
            out[f"{tag}_sr_MI"] = res.sr["MI"]
        out[f"{tag}_sr_MI_16"] = res.sr["MI"][::16]
        out[f"{tag}_sr_MI_sum"] = float(np.sum(res.sr["MI"]))
        out[f"{tag}_thr"] = np.array([b.disc_thresh if b.disc_thresh is not None else np.nan for b in res.blocks])
        out[f"{tag}_prob"] = np.array([b.prob if b.prob is not None else np.nan for b in res.blocks])
        out[f"{tag}_npairs"] = np.array([len(b.MI) for b in res.blocks])
    np.savez_compressed(os.path.join(OUT, "fixture_expected.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
