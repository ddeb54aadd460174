"""ctypes binding of libcrgpu.so — the C ABI declared in include/crgpu.h.

The product has no CPU path: if the shared library is missing, or there is no
CUDA device when a context is created, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

STAT_NAMES = ["reads", "valid_before", "corrected", "invalid", "keys", "distinct_keys", "umi_corrected_keys",
              "low_support_keys", "molecules", "nnz", "barcodes", "kernel_launches", "umi_corrected_reads",
              "low_support_reads", "sort_violations", "rle_violations", "filtered_target_umis"]
NO_FEATURE = 0xFFFFFFFF
NO_RANK = 0x3FFFFFFF


class CrgpuError(RuntimeError):
    pass


class LibraryDef(C.Structure):
    _fields_ = [("whitelist", C.c_int32), ("bc_offset", C.c_int32), ("bc_length", C.c_int32),
                ("umi_offset", C.c_int32), ("umi_length", C.c_int32), ("umi_correction", C.c_int32),
                ("is_feature_barcode", C.c_int32), ("feature_type", C.c_int32), ("fb_offset", C.c_int32),
                ("fb_length", C.c_int32)]


class ReadBatch(C.Structure):
    _fields_ = [("n", C.c_uint64), ("r1_len", C.c_int32), ("r1_seq", C.c_void_p), ("r1_qual", C.c_void_p),
                ("feature", C.c_void_p), ("r2_len", C.c_int32), ("r2_seq", C.c_void_p), ("r2_qual", C.c_void_p),
                ("on_device", C.c_int32), ("select_key", C.c_void_p)]


class SynthParams(C.Structure):
    _fields_ = [("seed_mix", C.c_uint64), ("seed_mol", C.c_uint64),
                ("n_whitelist", C.c_uint32), ("n_cells", C.c_uint32), ("n_genes", C.c_uint32), ("n_fb", C.c_uint32),
                ("bc_len", C.c_int32), ("umi_len", C.c_int32), ("fb_offset", C.c_int32), ("fb_len", C.c_int32),
                ("is_fb", C.c_int32),
                ("ambient_thr", C.c_uint32), ("unmapped_thr", C.c_uint32), ("bc_err_thr", C.c_uint32),
                ("umi_err_thr", C.c_uint32), ("n_thr", C.c_uint32), ("fb_err_thr", C.c_uint32),
                ("homopolymer_thr", C.c_uint32),
                ("n_qual_ascii", C.c_int32), ("n_qual_classes", C.c_int32), ("n_equal_classes", C.c_int32),
                ("qual_thr", C.c_uint32 * 8), ("equal_thr", C.c_uint32 * 8),
                ("qual_val", C.c_uint8 * 8), ("equal_val", C.c_uint8 * 8),
                ("wl_packed", C.c_void_p), ("cell_rank", C.c_void_p), ("cell_cdf", C.c_void_p),
                ("n_mol", C.c_void_p), ("gene_cdf", C.c_void_p), ("fb_cdf", C.c_void_p), ("fb_packed", C.c_void_p)]


_lib = None


def lib_path() -> str:
    return _build.SO


def load(build_if_missing: bool = True):
    """Load libcrgpu.so. Builds it with nvcc when absent (cross-compiles without a GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if not os.path.exists(so):
        if not build_if_missing:
            raise CrgpuError(f"{so} is missing: build it with `python -m cellranger_b200.build` "
                             "(the product has no CPU fallback)")
        _build.build_library()
    L = C.CDLL(so)
    L.crgpu_last_error.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().crgpu_last_error().decode(errors="replace")
        raise CrgpuError(f"{what or 'crgpu call'} failed ({rc}): {msg}")


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C contiguous"
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))
