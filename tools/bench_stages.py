#!/usr/bin/env python
"""Stage timings beside the headline scan (SURVEY.md 8d asks for HDW time and encode GB/s separately).

    python tools/bench_stages.py            # prints one JSON line per stage

 * encode : ldw_aln_param + ldw_extract_snps (column counts, site filter, class matrix) on a synthetic
            616 x 2.2 Mb alignment from pinned host memory -- wall clock through the C ABI, host->device copy included;
            algorithmic bytes = S*L read + S*nsnp written (HBM-bound kernels behind a PCIe-bound upload).
 * hdw    : ldw_hdw at C2 (616 x 100k) and C3 (10 000 x 50 000); algorithmic work 5*nsnp*S^2 int8 op.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    import torch
    import ldweaver_b200 as ldw
    from ldweaver_b200.synth import cheap_codes as _cheap_codes
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    # ---------------------------------------------------------------- encode
    S, L, nvar = 616, 2_221_315, 100_000
    rng = np.random.default_rng(0)
    aln = np.full((S, L), ord("A"), dtype=np.uint8)
    cols = np.sort(rng.choice(L, nvar, replace=False))
    codes = _cheap_codes(S, nvar, 7, 0.01, (0.847, 0.147, 0.006))
    aln[:, cols] = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes.T]
    aln = torch.from_numpy(aln).pin_memory().numpy()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        snp = ldw.snp_dat_from_alignment_matrix(aln, method="default")
        best = min(best, time.perf_counter() - t0)
    alg = S * L + S * snp.nsnp
    print(json.dumps({"stage": "encode (aln_param + extract_snps, host buffers)", "nseq": S, "seq_len": L, "nsnp": int(snp.nsnp),
                      "ms": 1e3 * best, "algorithmic_GB": alg / 1e9, "GB_per_s": alg / best / 1e9,
                      "bound": "PCIe upload of the alignment, then HBM", "hbm_peak_GB_per_s": peaks.get("hbm_gbs")}))
    del aln
    # ---------------------------------------------------------------- hdw
    for name, S, n in (("C2", 616, 100_000), ("C3", 10_000, 50_000)):
        codes = _cheap_codes(S, n, 11, 0.01, (0.847, 0.147, 0.006))
        snp = ldw.snp_dat_from_codes(codes, np.arange(1, n + 1, dtype=np.int32), n)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ldw.estimate_Hamming_distance_weights(snp, 0.1)
            best = min(best, time.perf_counter() - t0)
        ops = 5.0 * n * S * S
        print(json.dumps({"stage": f"hdw {name} (ldw_hdw, host buffers)", "nseq": S, "nsnp": n, "ms": 1e3 * best,
                          "algorithmic_int8_TOPS": ops / best / 1e12, "h2d_MB": codes.nbytes / 1e6,
                          "int8_peak_TOPS (2 x measured bf16)": 2 * peaks.get("bf16_tflops", 0)}))


if __name__ == "__main__":
    main()
