"""ctypes loader for libldwgpu.so (the C-ABI of include/ldw.h).

The product path has NO CPU fallback: if the shared library is missing, or no CUDA device is
present, loading / context creation raises loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDW_LIBRARY_PATH: developer override (same-box A/B of kernel variants, tools/kernel_ab.py); the product ships one library
LIB_PATH = os.environ.get("LDW_LIBRARY_PATH") or os.path.join(_HERE, "libldwgpu.so")

i64 = C.c_int64
f64 = C.c_double
P = C.POINTER


class LdwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libldwgpu error {code}: {msg}")
        self.code = code


class Links(C.Structure):
    """struct ldw_links (include/ldw.h)."""
    _fields_ = [("n", i64), ("pos1", P(C.c_int32)), ("pos2", P(C.c_int32)), ("clust1", P(C.c_int32)),
                ("clust2", P(C.c_int32)), ("len", P(C.c_int32)), ("MI", P(f64)), ("block", P(C.c_int32))]

    @classmethod
    def from_dict(cls, d: dict):
        """Borrow the columns of a link table held as NumPy arrays (kept alive on the returned object)."""
        self = cls()
        keep = {}
        for name, dt, ct in (("pos1", np.int32, C.c_int32), ("pos2", np.int32, C.c_int32), ("clust1", np.int32, C.c_int32),
                             ("clust2", np.int32, C.c_int32), ("len", np.int32, C.c_int32), ("MI", np.float64, f64),
                             ("block", np.int32, C.c_int32)):
            a = np.ascontiguousarray(d[name] if name in d else np.zeros(len(d["MI"]), dtype=dt), dtype=dt)
            keep[name] = a
            setattr(self, name, a.ctypes.data_as(P(ct)))
        self.n = len(keep["MI"])
        self._keep = keep
        return self

    def views(self) -> dict:
        """Zero-copy NumPy views of the columns (valid as long as the owner keeps the buffers: the context until its next
        scan, or the arrays passed to from_dict)."""
        if hasattr(self, "_keep"):
            return self._keep
        n = int(self.n)
        out = {}
        for name, _ in self._fields_[1:]:
            ptr = getattr(self, name)
            if n and ptr:
                out[name] = np.ctypeslib.as_array(ptr, shape=(n,))
        return out

    def to_dict(self) -> dict:
        """Copies of the columns as NumPy arrays (``ldw_links_copy``: host threads, 2.9 GB at 616 x 100k)."""
        n = int(self.n)
        cols = (("pos1", np.int32), ("pos2", np.int32), ("clust1", np.int32), ("clust2", np.int32), ("len", np.int32),
                ("MI", np.float64), ("block", np.int32))
        out = {name: np.empty(n if (n and getattr(self, name)) else 0, dtype=dt) for name, dt in cols}
        if n:
            dst = [out[name].ctypes.data_as(C.c_void_p) if len(out[name]) else None for name, _ in cols]
            check(lib().ldw_links_copy(C.byref(self), *dst))
        return out


class ScanStats(C.Structure):
    """struct ldw_scan_stats (include/ldw.h)."""
    _fields_ = [("n_blocks", i64), ("n_pairs", i64), ("n_sr", i64), ("n_lr_total", i64), ("n_lr_kept", i64),
                ("n_borderline", i64), ("n_reruns", i64), ("n_candidates", i64), ("t_pack_ms", f64), ("t_scan_ms", f64), ("t_select_ms", f64),
                ("t_d2h_ms", f64), ("t_kernel_ms", f64), ("n_scan_launches", i64), ("n_launches", i64), ("n_tiles", i64),
                ("exec_int8_ops", f64), ("t_host_prep_ms", f64), ("exec_mufu_ops", f64), ("eps_obs_max", f64)]

    def to_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class SrPost(C.Structure):
    """struct ldw_sr_post (include/ldw.h)."""
    _fields_ = [("n_df", i64), ("clust_c", P(C.c_int32)), ("row", P(i64)), ("srp_max", P(f64)),
                ("n_red", i64), ("red", P(i64)), ("n_chk", i64), ("chk", P(i64)), ("nclust", C.c_int32),
                ("fit_off", P(i64)), ("fit_len", P(C.c_int32)), ("fit_q95", P(f64)), ("fit_val", P(f64)),
                ("coef", P(f64)), ("shape", P(f64)), ("start", P(f64)), ("n_pos", P(i64)), ("nm_evals", P(C.c_int32)),
                ("nm_fail", P(C.c_int32)), ("priv", C.c_void_p)]


def copy_array(ptr, n: int, dtype) -> np.ndarray:
    """Copy n elements out of a library-owned buffer."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


_lib = None


def lib():
    """Load libldwgpu.so (built by ``python __graft_entry__.py`` / ``make -C ldweaver_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "ldweaver_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ldw_last_error.restype = C.c_char_p
    L.ldw_create.argtypes = [C.c_int, P(C.c_void_p)]
    L.ldw_destroy.argtypes = [C.c_void_p]
    L.ldw_destroy.restype = None
    L.ldw_aln_param.argtypes = [C.c_void_p, C.c_void_p, i64, i64, C.c_int, f64, f64, C.c_void_p, P(i64), C.c_void_p]
    L.ldw_extract_snps.argtypes = [C.c_void_p, C.c_void_p, i64, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    L.ldw_encode_alignment.argtypes = [C.c_void_p, C.c_void_p, i64, i64, C.c_int, f64, f64, P(i64), P(C.c_void_p), P(C.c_void_p),
                                       P(C.c_void_p)]
    L.ldw_read_fasta.argtypes = [C.c_char_p, P(i64), P(i64), C.c_void_p, i64, C.c_void_p, i64]
    L.ldw_read_fasta_alloc.argtypes = [C.c_char_p, P(i64), P(i64), P(C.c_void_p), P(C.c_void_p), P(i64)]
    L.ldw_buffer_free.argtypes = [C.c_void_p]
    L.ldw_buffer_free.restype = None
    L.ldw_acgtn2num.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, i64]
    L.ldw_hdw.argtypes = [C.c_void_p, C.c_void_p, i64, i64, f64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ldw_mi_plan_create.argtypes = [C.c_void_p, C.c_void_p, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, i64,
                                     P(C.c_void_p)]
    L.ldw_mi_plan_destroy.argtypes = [C.c_void_p]
    L.ldw_mi_plan_destroy.restype = None
    L.ldw_mi_scan.argtypes = [C.c_void_p, f64, f64, f64, f64, C.c_int, C.c_int, C.c_int, P(Links), P(Links), P(Links),
                              C.c_void_p, C.c_void_p, P(ScanStats)]
    L.ldw_mi_block_dense.argtypes = [C.c_void_p, i64, C.c_void_p, P(i64), P(i64)]
    L.ldw_write_lr_tsv.argtypes = [C.c_char_p, P(Links), C.c_int]
    L.ldw_format_r_real.argtypes = [f64, C.c_char_p, C.c_int]
    L.ldw_sr_postprocess.argtypes = [P(Links), C.c_int32, f64, f64, P(SrPost)]
    L.ldw_sr_postprocess_dev.argtypes = [C.c_void_p, C.c_int32, f64, f64, P(SrPost), P(Links)]
    L.ldw_sr_post_free.argtypes = [P(SrPost)]
    L.ldw_sr_post_free.restype = None
    L.ldw_run_aracne.argtypes = [i64, C.c_void_p, C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ldw_write_sr_tsv.argtypes = [C.c_char_p, P(Links), i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ldw_read_numeric_tsv.argtypes = [C.c_char_p, C.c_int, P(i64), P(P(f64))]
    L.ldw_table_free.argtypes = [P(f64)]
    L.ldw_table_free.restype = None
    L.ldw_links_to_cells.argtypes = [C.c_void_p, i64, i64, P(Links), C.c_void_p, C.c_void_p]
    L.ldw_links_copy.argtypes = [P(Links)] + [C.c_void_p] * 7
    L.ldw_nm_rosenbrock.argtypes = [C.c_void_p, C.c_void_p, P(f64), P(C.c_int)]
    L.ldw_neg_log_pbeta_upper.argtypes = [C.c_void_p, i64, f64, f64, C.c_void_p]
    L.ldw_mi_pairs_exact.argtypes = [C.c_void_p, i64, C.c_void_p, C.c_void_p, i64, C.c_void_p]
    L.ldw_group_create.argtypes = [C.c_void_p, C.c_int, P(C.c_void_p)]
    L.ldw_group_unique_id.argtypes = [C.c_char_p]
    L.ldw_group_create_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, P(C.c_void_p)]
    L.ldw_group_destroy.argtypes = [C.c_void_p]
    L.ldw_group_destroy.restype = None
    L.ldw_group_info.argtypes = [C.c_void_p, P(C.c_int), P(C.c_int), P(C.c_int)]
    L.ldw_group_ctx.argtypes = [C.c_void_p, C.c_int]
    L.ldw_group_ctx.restype = C.c_void_p
    L.ldw_group_load_codes.argtypes = [C.c_void_p, C.c_void_p, i64, i64]
    L.ldw_group_hdw.argtypes = [C.c_void_p, f64, C.c_int, C.c_void_p, C.c_void_p, P(C.c_int)]
    L.ldw_group_mi_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i64, f64, f64, f64, f64, C.c_int,
                                    P(Links), P(Links), P(Links), C.c_void_p, C.c_void_p, C.c_void_p, P(f64)]
    _lib = L
    return L


_nccl_preloaded = False


def preload_nccl() -> None:
    """libldwgpu binds NCCL at run time by soname (csrc/group.cu); a later ``import torch`` needs its own bundled,
    newer libnccl.so.2.  Loading that copy first (when the nvidia-nccl wheel is installed) keeps a process on ONE copy
    whichever order the imports come in.  No wheel: the system library is found by the loader."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            so = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(so):
                C.CDLL(so, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def check(rc: int) -> None:
    if rc != 0:
        raise LdwError(rc, lib().ldw_last_error().decode(errors="replace"))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one ldw_ctx (one CUDA device)."""

    def __init__(self, device: int = 0):
        self.handle = C.c_void_p()
        check(lib().ldw_create(device, C.byref(self.handle)))
        self.device = device

    def close(self):
        if self.handle:
            lib().ldw_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
