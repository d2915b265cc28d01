"""Generates tests/golden/ref_expected.npz by RUNNING THE REFERENCE'S OWN COMPILED CODE (oracle/_ref/libldw_ref.so =
/root/reference/src/{getACGTNsites.cpp, computeMI.cpp, ACGTN2num_parallel.cpp, kseq2.h} built unmodified against
oracle/mock_rcpp/Rcpp.h; recipe oracle/Makefile target `ref`).

Run in the build container only (it needs /root/reference):
    python tests/golden/make_golden_ref.py
Unlike fixture_expected.npz (outputs of OUR NumPy oracle), every array written here comes out of reference object code:
  * enc_*      .extractAlnParam + .extractSNPs on inst/extdata/snp_sample.fa.gz for several (filter, gap, maf)
  * edge_*     the same two calls on a small CRLF / lower-case / IUPAC / gap-rich alignment whose FILE BYTES are stored
               too, so that the GPU box can push the identical file through reader + device kernels
  * a2n_*      .ACGTN2num on all 256 byte values
  * mi_*       per-block MI matrices (max_blk_sz 500: diagonal, square off-diagonal and ragged blocks, quirk Q1) whose
               element-wise finish was executed by the reference's .fastHadamard; the R-level matrix algebra around it
               (R/computePairwiseMI.R:238-298,390-396) is oracle/ldw_oracle.py's restatement -- R itself is not here.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ldw_oracle as O  # noqa: E402
import ref_lib as R  # noqa: E402

REF = "/root/reference/inst/extdata"
OUT = os.path.dirname(os.path.abspath(__file__))

ENC_SETTINGS = [(0, 0.15, 0.01), (1, 0.15, 0.01), (0, 0.05, 0.10), (1, 0.05, 0.10), (0, 0.15, 0.30), (1, 0.02, 0.40)]
MI_ROW_STEP = 16
MI_BLK = 500


def edge_alignment_bytes() -> bytes:
    """60 records x 420 columns in lines of 70 with CRLF line ends (=> 6 CR columns per record, class 'other'),
    mixed case, IUPAC codes, gaps; deterministic."""
    rng = np.random.default_rng(20260101)
    nseq, L = 60, 420
    alphabet = np.frombuffer(b"ACGTacgtNn-RYKM", dtype=np.uint8)
    base = alphabet[rng.integers(0, 4, size=L)]
    rows = np.repeat(base[None, :], nseq, axis=0)
    rate = rng.random(L) * 0.5
    mut = rng.random((nseq, L)) < rate[None, :]
    rows[mut] = alphabet[rng.integers(0, len(alphabet), size=int(mut.sum()))]
    out = bytearray()
    for k in range(nseq):
        out += b">seq_%d sample %d\r\n" % (k, k)
        for o in range(0, L, 70):
            out += rows[k, o:o + 70].tobytes() + b"\r\n"
    return bytes(out)


def encode_with_reference(path, filt, gap, maf):
    par = R.extractAlnParam(path, filt, gap, maf)
    out = {"num_seqs": par["num.seqs"], "num_snps": par["num.snps"], "seq_length": par["seq.length"], "pos": par["pos"]}
    if par["num.snps"] > 0:
        sn = R.extractSNPs(path, par["num.seqs"], par["num.snps"], par["pos"])
        out["table"] = sn["ACGTN_table"]
        out["codes"] = R.codes_from_coo(sn, par["num.snps"], par["num.seqs"])
        assert not np.any(out["codes"] == 255)
    return out, par["seq.names"]


def main():
    out = {}
    fa = os.path.join(REF, "snp_sample.fa.gz")
    for k, (filt, gap, maf) in enumerate(ENC_SETTINGS):
        enc, names = encode_with_reference(fa, filt, gap, maf)
        out[f"enc{k}_setting"] = np.array([filt, gap, maf])
        for key, v in enc.items():
            out[f"enc{k}_{key}"] = np.asarray(v)
    out["enc_names"] = np.array(names)

    edge = edge_alignment_bytes()
    tmp = os.path.join(OUT, "_edge_tmp.fa")
    with open(tmp, "wb") as fh:
        fh.write(edge)
    try:
        out["edge_file"] = np.frombuffer(edge, dtype=np.uint8)
        for k, (filt, gap, maf) in enumerate(ENC_SETTINGS[:4]):
            enc, names = encode_with_reference(tmp, filt, gap, maf)
            for key, v in enc.items():
                out[f"edge{k}_{key}"] = np.asarray(v)
        out["edge_names"] = np.array(names)
    finally:
        os.remove(tmp)

    cv = bytes(range(256))
    nv = np.ones((5, 256), order="F")
    R.ACGTN2num(nv, cv, 2)
    out["a2n_cv"] = np.frombuffer(cv, dtype=np.uint8)
    out["a2n_nv"] = nv

    # MI with the reference's .fastHadamard doing the element-wise finish
    pos_all = np.loadtxt(os.path.join(REF, "snp_sample.pos"), dtype=np.int64)
    snp = O.snp_dat_from_codes(out["enc1_codes"], pos_all[out["enc1_pos"].astype(np.int64) - 1], 50000)
    hdw = O.estimate_Hamming_distance_weights(snp, 0.1)
    O.HADAMARD_IMPL = R.fastHadamard
    try:
        for bi, (fs, fe, ts, te) in enumerate(O.make_blocks(snp.nsnp, MI_BLK)):
            f, t = np.arange(fs - 1, fe), np.arange(ts - 1, te)
            MI = O.block_mi_matrix(snp, hdw, f, t)
            out[f"mi_b{bi}_rows"] = np.arange(0, len(f), MI_ROW_STEP)
            out[f"mi_b{bi}"] = MI[::MI_ROW_STEP, :]
            out[f"mi_b{bi}_sum"] = float(MI.sum())
    finally:
        O.HADAMARD_IMPL = None
    out["mi_blk"] = MI_BLK
    out["mi_hdw"] = hdw
    np.savez_compressed(os.path.join(OUT, "ref_expected.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
