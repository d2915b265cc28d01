#!/usr/bin/env python
"""Writes a BASELINE-shaped synthetic alignment as the FASTA (+ .pos) file the reference package reads, so that
baseline/run_reference.R and this repo see the same bytes:  python tools/make_synthetic_fasta.py C2 out.fa.gz [nsnp]
SNP-only style (one column per SNP, genome positions in out.fa.gz.pos, genome length printed)."""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]


def main():
    from ldweaver_b200 import synth
    cfg, out = sys.argv[1], sys.argv[2]
    sy = synth.generate_config(cfg, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    aln = synth.codes_to_alignment(sy.codes, lowercase_frac=0.3 if cfg == "C5" else 0.0)
    opener = gzip.open if out.endswith(".gz") else open
    with opener(out, "wb") as fh:
        for k in range(aln.shape[0]):
            fh.write(b">seq%d\n" % k + aln[k].tobytes() + b"\n")
    with open(out + ".pos", "w") as fh:
        fh.write("\n".join(str(int(p)) for p in sy.POS) + "\n")
    print(f"{out}: {aln.shape[0]} x {aln.shape[1]}; positions {out}.pos; g = {sy.g}")


if __name__ == "__main__":
    main()
