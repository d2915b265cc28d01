#!/usr/bin/env python
"""Compares this repo's outputs with a directory written by baseline/run_reference.R (real R package) on the same FASTA:
    python tools/compare_with_reference.py <aln.fa[.gz]> <ref_out_dir> [pos_file] [g] [method]
Needs a GPU (it runs the product path).  Integers/positions must be identical, hdw bit-exact, MI within 1e-6, link sets
identical up to threshold-borderline pairs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import ldweaver_b200 as ldw
    aln, ref = sys.argv[1], sys.argv[2]
    pos_file = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    g = float(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] != "-" else None
    method = sys.argv[5] if len(sys.argv) > 5 else "default"
    if pos_file:
        snp = ldw.parse_fasta_SNP_alignment(aln, np.loadtxt(pos_file), method=method)
        snp.g = int(g)
    else:
        snp = ldw.parse_fasta_alignment(aln, method=method)
    POS = np.loadtxt(os.path.join(ref, "POS.txt"))
    assert np.array_equal(snp.POS, POS), "POS differs"
    assert np.array_equal(snp.r, np.loadtxt(os.path.join(ref, "r.txt"))), "r differs"
    assert np.array_equal(snp.uqe, np.loadtxt(os.path.join(ref, "uqe.tsv")).reshape(snp.uqe.shape)), "uqe differs"
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    hdw_ref = np.array([float(x) for x in open(os.path.join(ref, "hdw.txt"))])
    assert np.array_equal(hdw, hdw_ref), "hdw differs"
    n = snp.nsnp
    paint = np.ones(n, dtype=np.int32)
    paint[n // 3:] = 2
    paint[2 * n // 3:] = 3
    res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(paint, 3), write_tsv=False)
    lr = ldw.read_LongRangeLinks(os.path.join(ref, "lr_links.tsv"), sr_dist=0)
    key = lambda a, b: {(int(x), int(y)) for x, y in zip(a, b)}
    mine, theirs = key(res.lr["pos1"], res.lr["pos2"]), key(lr["pos1"], lr["pos2"])
    print("LR links: common", len(mine & theirs), "only here", len(mine - theirs), "only reference", len(theirs - mine),
          "(listed borderline pairs:", len(res.borderline["MI"]) if hasattr(res, "borderline") else "n/a", ")")
    ref_mi = {(int(a), int(b)): m for a, b, m in zip(lr["pos1"], lr["pos2"], lr["MI"])}
    d = [abs(m - ref_mi[(int(a), int(b))]) for a, b, m in zip(res.lr["pos1"], res.lr["pos2"], res.lr["MI"]) if (int(a), int(b)) in ref_mi]
    print("LR MI max abs diff over common links:", max(d) if d else None)
    print("OK: POS, r, uqe identical; hdw bit-exact")


if __name__ == "__main__":
    main()
