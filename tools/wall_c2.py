#!/usr/bin/env python
"""Wall time "to sr/lr links" at BASELINE config #2 (616 x 100k synthetic), through the public API exactly as
a user of the reference would call it: estimate_Hamming_distance_weights(), then perform_MI_computation() with both
TSV files written.  `--host-post`: the short-range table comes to the host and mergeNsort_sr_links runs on host threads;
default: the table stays on the device (ldw_sr_postprocess_dev).  LDW_DBG_TIMING=1 prints the library's phase times on
stderr.  Prints one JSON line; the second of two calls is reported (the first one allocates buffers)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    import ldweaver_b200 as ldw
    from ldweaver_b200 import synth
    S, n, seed, probs, nrate = synth.CONFIGS["C2"]
    sy = synth.generate(S, n, seed, probs, nrate)
    snp = ldw.snp_dat_from_codes(sy.codes, sy.POS, sy.g)
    lra = synth.exact_lr_links_approx(sy.POS, sy.g, 20000.0)
    kw = dict(exact_sr="in_scan") if "--host-post" in sys.argv else dict(device_post=True)
    out = None
    for it in range(2):
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            hdw = ldw.estimate_Hamming_distance_weights(snp, 0.1)
            t1 = time.perf_counter()
            res = ldw.perform_MI_computation(snp, hdw, ldw.CdsVar(sy.paint, 3), ncores=1, lr_save_path=os.path.join(d, "lr_links.tsv"),
                                             sr_save_path=os.path.join(d, "sr_links.tsv"), plt_folder=d, sr_dist=20000,
                                             lr_retain_links=1e6, max_blk_sz=10000, srp_cutoff=3, runARACNE=True, lr_links_approx=lra, **kw)
            t2 = time.perf_counter()
            out = {"workload": "C2: synthetic 616 x 100000, sr_dist 20000, lr_retain_links 1e6, max_blk_sz 10000, srp_cutoff 3, ARACNE on",
                   "variant": "host_post" if "--host-post" in sys.argv else "device_post", "call": it, "hdw_s": t1 - t0,
                   "perform_MI_computation_s": t2 - t1, "total_s": t2 - t0, "n_sr_links": int(res.stats["n_sr"]), "n_lr_links": int(len(res.lr["MI"])),
                   "n_sr_links_red": int(len(res.sr_links_red["row"])), "aracne_kept": int(res.sr_links_red["ARACNE"].sum()),
                   "phases_s": res.stats.get("phases"), "host_threads": os.cpu_count()}
        del res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
