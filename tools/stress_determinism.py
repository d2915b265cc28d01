#!/usr/bin/env python
"""Developer tool: repeat small multi-kind scans (ragged blocks, every tile kind, 3 K-blocks) and compare every run with
the first one bit for bit -- a pipeline race in the scan kernel shows up as a run-to-run difference."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import ldweaver_b200 as ldw
from ldweaver_b200 import synth

def one(nseq, nsnp, seed, blk, reps, probs=(0.847, 0.147, 0.006), nrate=0.01):
    sy = synth.generate(nseq=nseq, nsnp=nsnp, seed=seed, allele_probs=probs, n_rate=nrate)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000)
    first = None
    bad = 0
    for r in range(reps):
        res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(sy.paint, 3), sr_dist=20000, lr_retain_links=5e4, max_blk_sz=blk,
                                         lr_links_approx=lra, write_tsv=False)
        cur = (res.thr.copy(), res.sr["MI"].copy(), res.lr["MI"].copy(), res.lr["pos1"].copy(), res.lr["pos2"].copy())
        if first is None:
            first = cur
            continue
        same = all(np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b) for a, b in zip(first, cur))
        if not same:
            bad += 1
            d_thr = np.nanmax(np.abs(first[0] - cur[0])) if len(cur[0]) else 0
            nsr = int((first[1] != cur[1]).sum()) if first[1].shape == cur[1].shape else -1
            print(f"  run {r}: DIFFERS  max|dthr|={d_thr:.3g}  sr MI differing={nsr}  n_lr {len(first[2])} vs {len(cur[2])}", flush=True)
    print(f"nseq={nseq} nsnp={nsnp} blk={blk} probs={probs} nrate={nrate}: {bad} of {reps - 1} repeats differ", flush=True)
    return bad

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    bad = 0
    bad += one(300, 2700, 11, 1000, reps)
    bad += one(300, 2700, 12, 1000, reps, probs=(0.4, 0.3, 0.3), nrate=0.05)
    bad += one(616, 6000, 13, 2000, max(4, reps // 3))
    print("TOTAL differing runs:", bad)
    sys.exit(1 if bad else 0)
