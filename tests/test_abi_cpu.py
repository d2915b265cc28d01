"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/ldw.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ldw.h")).read()
    return sorted(set(re.findall(r"LDW_API [a-z0-9_ \*]+?\b(ldw_[a-z0-9_]+)\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ldw_create", "ldw_destroy", "ldw_aln_param", "ldw_extract_snps", "ldw_acgtn2num", "ldw_hdw",
              "ldw_mi_plan_create", "ldw_mi_scan", "ldw_last_error", "ldw_read_fasta"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from ldweaver_b200 import _lib
    L = _lib.lib()
    for s in _declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/ldw.h but not exported"
    assert L.ldw_abi_version() == 2


def test_no_cpu_fallback_without_device():
    import torch
    from ldweaver_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.LdwError) as ei:
        _lib.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_host_fasta_reader(tmp_path):
    import gzip
    import numpy as np
    from ldweaver_b200 import api
    p = tmp_path / "x.fa.gz"
    with gzip.open(p, "wt") as fh:
        fh.write(">s1 some description\nACGT\nac-n\n>s2\nTTTTGGGG\n>s3\nAC\nGTAC\nGT\n")
    names, aln = api.read_fasta_matrix(str(p))
    assert names == ["s1", "s2", "s3"]
    assert [bytes(r) for r in aln] == [b"ACGTac-n", b"TTTTGGGG", b"ACGTACGT"]
    q = tmp_path / "bad.fa"
    q.write_text(">a\nACGT\n>b\nACG\n")
    with pytest.raises(ValueError, match="different lengths"):
        api.read_fasta_matrix(str(q))
    e = tmp_path / "empty.fa"
    e.write_text("")
    with pytest.raises(ValueError, match="any sequences"):
        api.read_fasta_matrix(str(e))


def test_r_shim_type_checks_against_the_abi():
    """r_package/src/r_shim.cpp (the .Call glue a maintainer adds to the R package) cannot be built here -- no R -- but
    it must at least type-check against include/ldw.h; tests/mock_r/ declares the handful of R C API functions it uses."""
    import subprocess
    cmd = ["g++", "-fsyntax-only", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "mock_r"), os.path.join(ROOT, "r_package", "src", "r_shim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    shim = open(os.path.join(ROOT, "r_package", "src", "r_shim.cpp")).read()
    rcode = open(os.path.join(ROOT, "r_package", "R", "gpu_hotpath.R")).read()
    entries = re.findall(r'\{"(_LDWeaver_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', shim)
    assert {n for n, _ in entries} >= {"_LDWeaver_gpu_encode", "_LDWeaver_gpu_hdw", "_LDWeaver_gpu_mi_scan", "_LDWeaver_ACGTN2num"}
    for name, nargs in entries:
        calls = re.findall(r'\.Call\("%s",(.*?)PACKAGE = "LDWeaver"\)' % name, rcode, flags=re.S)
        assert calls, f"{name} registered but never called from gpu_hotpath.R"
        for body in calls:  # the registered argument count must be the number of arguments the R side passes
            body = re.sub(r"#[^\n]*", "", body)
            depth, nargs_r = 0, 0
            for ch in body:
                depth += ch in "([{"
                depth -= ch in ")]}"
                nargs_r += ch == "," and depth == 0
            assert nargs_r == int(nargs), f"{name}: R passes {nargs_r} arguments, the shim registers {nargs}"
    # Rf_error is a longjmp: the shim must hold no object with a destructor
    code = re.sub(r"//[^\n]*", "", shim)
    assert "std::" not in code and "#include <vector>" not in code and "#include <string>" not in code
    # the reference's own stub name is bound (R/RcppExports.R:4-6) and devices come from options()/LDW_GPUS
    assert re.search(r"^\.ACGTN2num <- function\(nv, cv, ncores\)", rcode, flags=re.M)
    assert "LDWeaver.gpus" in rcode and "LDW_GPUS" in rcode and "ldw_group_mi_scan" in shim


def test_links_copy_matches_the_columns():
    """Links.to_dict() (ldw_links_copy, host threads) is how every scan result reaches NumPy: exact copies, right dtypes,
    NULL columns (LDW_SCAN_NO_D2H results) come back empty."""
    import numpy as np
    from ldweaver_b200 import _lib
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 1 << 20, (1 << 20) + 7, 2_500_001):
        d = dict(pos1=rng.integers(0, 1e6, n), pos2=rng.integers(0, 1e6, n), clust1=rng.integers(1, 4, n), clust2=rng.integers(1, 4, n),
                 len=rng.integers(1, 20000, n), MI=rng.random(n), block=rng.integers(0, 55, n))
        a = _lib.Links.from_dict(d)
        b = _lib.Links()
        C.memmove(C.byref(b), C.byref(a), C.sizeof(a))      # a table that only holds pointers, like one returned by ldw_mi_scan
        o = b.to_dict()
        for k in d:
            assert o[k].dtype == (np.float64 if k == "MI" else np.int32) and np.array_equal(o[k], np.asarray(d[k]).astype(o[k].dtype))
    e = _lib.Links()
    e.n = 10
    assert all(len(v) == 0 for v in e.to_dict().values())


def test_single_pass_fasta_reader_on_records_that_cross_its_buffers(tmp_path):
    """Grammar parity with the reference's kseq reader lives in tests/test_fasta_ref_cpu.py (checked against the compiled
    src/kseq2.h); this is the large-input case."""
    import gzip
    import numpy as np
    from ldweaver_b200 import api
    # records far longer than the reader's 16 MB hand-over buffers, plain and gz, lines of 70 characters: every state of
    # the tokeniser crosses a buffer boundary somewhere
    rng = np.random.default_rng(0)
    L = 9_000_001
    seqs = [np.frombuffer(b"ACGTN-acgt", dtype=np.uint8)[rng.integers(0, 10, L)] for _ in range(4)]
    body = bytearray()
    for k, s in enumerate(seqs):
        body += b">sequence_number_%d some description\n" % k
        lines = np.split(s, np.arange(70, L, 70))
        body += b"\n".join(x.tobytes() for x in lines) + b"\n"
    for path, opener in ((tmp_path / "big.fa", open), (tmp_path / "big.fa.gz", lambda p, m: gzip.open(p, m, compresslevel=1))):
        with opener(path, "wb") as fh:
            fh.write(bytes(body))
        names, aln = api.read_fasta_matrix(str(path))
        assert names == ["sequence_number_%d" % k for k in range(4)] and aln.shape == (4, L)
        assert all(np.array_equal(aln[k], seqs[k]) for k in range(4))
    with pytest.raises(Exception, match="can't open"):
        api.read_fasta_matrix(str(tmp_path / "missing.fa"))
