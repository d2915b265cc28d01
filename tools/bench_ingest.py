#!/usr/bin/env python
"""FASTA/gz ingestion beside zlib's own inflate time for the same file: the single-pass reader
(ldw_read_fasta_alloc) against the two-pass character loop it replaced (ldw_read_fasta).  Host only; one JSON line.

    python tools/bench_ingest.py [nseq seq_len]        # default 616 x 500000 (308 MB of sequence)
"""
import ctypes as C
import gzip
import json
import os
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from ldweaver_b200 import _lib, api  # noqa: E402


def main():
    S, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (616, 500_000)
    rng = np.random.default_rng(0)
    ref = rng.integers(0, 4, L).astype(np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "aln.fa.gz")
        with gzip.open(path, "wb", compresslevel=1) as fh:
            for s in range(S):
                seq = ref.copy()
                m = rng.random(L) < 0.01
                seq[m] = rng.integers(0, 4, int(m.sum()))
                fh.write(b">seq%d\n" % s + lut[seq].tobytes() + b"\n")
        t0 = time.perf_counter()
        dec = zlib.decompressobj(31)
        n = 0
        with open(path, "rb") as fh:
            while True:
                b = fh.read(1 << 20)
                if not b:
                    break
                n += len(dec.decompress(b))
        t_inflate = time.perf_counter() - t0
        t0 = time.perf_counter()
        names, aln = api.read_fasta_matrix(path)
        t_new = time.perf_counter() - t0
        lib = _lib.lib()
        nseq, slen = C.c_int64(), C.c_int64()
        t0 = time.perf_counter()
        _lib.check(lib.ldw_read_fasta(path.encode(), C.byref(nseq), C.byref(slen), None, 0, None, 0))
        old = np.empty((nseq.value, slen.value), dtype=np.uint8)
        nb = C.create_string_buffer(1 << 20)
        n2, l2 = C.c_int64(), C.c_int64(slen.value)
        _lib.check(lib.ldw_read_fasta(path.encode(), C.byref(n2), C.byref(l2), old.ctypes.data_as(C.c_void_p), old.size, nb, 1 << 20))
        t_old = time.perf_counter() - t0
        assert np.array_equal(old, aln) and len(names) == S
        print(json.dumps({"nseq": S, "seq_len": L, "sequence_MB": S * L / 1e6, "gz_MB": os.path.getsize(path) / 1e6,
                          "zlib_inflate_s": round(t_inflate, 3), "single_pass_reader_s": round(t_new, 3),
                          "two_pass_character_loop_s": round(t_old, 3), "single_pass_MB_per_s": round(S * L / t_new / 1e6, 1),
                          "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
