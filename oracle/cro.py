"""ctypes binding of the CPU oracle (oracle/cr_oracle.cpp).

TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and the CPU legs of
bench.py (cpu_baseline, --impl reference) may import this module. Nothing in
cellranger_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcr_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cr_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcr_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.cro_probability.restype = C.c_double
        L.cro_probability.argtypes = [C.c_uint8]
        L.cro_ctx_new.restype = C.c_void_p
        L.cro_n_reads.restype = C.c_uint64
        L.cro_matrix_n_barcodes.restype = C.c_uint64
        L.cro_matrix_nnz.restype = C.c_uint64
        L.cro_n_molecules.restype = C.c_uint64
        L.cro_total_barcode_counts.restype = C.c_uint64
        L.cro_kat_encode_2bit.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def ascii_mat(seqs, L=None) -> np.ndarray:
    """list of str/bytes (equal length) or (n, L) uint8 array -> (n, L) uint8."""
    if isinstance(seqs, np.ndarray):
        return np.ascontiguousarray(seqs, dtype=np.uint8)
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    if L is None:
        L = len(bs[0]) if bs else 0
    assert all(len(b) == L for b in bs)
    return np.frombuffer(b"".join(bs), dtype=np.uint8).reshape(len(bs), L).copy()


class Oracle:
    """One GEM well worth of reads through MAKE_SHARD → BARCODE_CORRECTION →
    ALIGN_AND_COUNT (dedup part) → matrix, on the CPU."""

    def __init__(self, threshold: float = 0.975, max_expected_errors: float = 1.7976931348623157e308,
                 filter_umis: bool = True):
        self.L = lib()
        self.ctx = C.c_void_p(self.L.cro_ctx_new())
        self.L.cro_set_params(self.ctx, C.c_double(threshold), C.c_double(max_expected_errors), int(filter_umis))
        self._keep = []
        self.bc_len = None
        self.umi_len = None
        self.n_features = 0

    def close(self):
        if self.ctx:
            self.L.cro_ctx_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def total_barcode_counts(self, min_reads_to_report_bc: int = 1):
        """bc_counts_total of BARCODE_CORRECTION (barcode_correction.rs:327-362) -> (seqs (n, L), valid, counts)."""
        n = int(self.L.cro_total_barcode_counts(self.ctx, C.c_uint64(min_reads_to_report_bc), self.bc_len, None, None, None))
        seqs = np.zeros((n, self.bc_len), dtype=np.uint8)
        valid = np.zeros(n, dtype=np.uint8)
        counts = np.zeros(n, dtype=np.uint64)
        self.L.cro_total_barcode_counts(self.ctx, C.c_uint64(min_reads_to_report_bc), self.bc_len, _p(seqs), _p(valid),
                                        _p(counts))
        return seqs, valid, counts

    def set_target_filter(self, on_target, min_read_count: int):
        """DupBuilder::build(.., targeted_umi_min_read_count) with the panel's target set (mark_dups.rs:311-320)."""
        t = np.ascontiguousarray(on_target, dtype=np.uint8)
        self.L.cro_set_target_filter(self.ctx, _p(t), C.c_int32(t.shape[0]), C.c_uint64(int(min_read_count)))

    def add_whitelist(self, seqs, trans=None) -> int:
        s = ascii_mat(seqs)
        t = ascii_mat(trans) if trans is not None else None
        return self.L.cro_add_whitelist(self.ctx, _p(s), C.c_uint64(s.shape[0]), int(s.shape[1]), _p(t))

    def add_library(self, wl: int, bc_off=0, bc_len=16, umi_off=16, umi_len=12, umi_correction=True,
                    is_fb=False, ftype=0, fb_offset=0, fb_len=0) -> int:
        self.bc_len = bc_len if self.bc_len is None else self.bc_len
        self.umi_len = umi_len if self.umi_len is None else self.umi_len
        assert self.bc_len == bc_len and self.umi_len == umi_len
        return self.L.cro_add_library(self.ctx, wl, bc_off, bc_len, umi_off, umi_len, int(umi_correction),
                                      int(is_fb), ftype, fb_offset, fb_len)

    def set_features(self, feature_type, fb_seqs=None):
        ft = np.ascontiguousarray(feature_type, dtype=np.int32)
        self.n_features = len(ft)
        if fb_seqs is None:
            fb = np.zeros((len(ft), 1), dtype=np.uint8)
        else:
            fb = ascii_mat(fb_seqs)
        self.L.cro_set_features(self.ctx, len(ft), _p(ft), _p(fb), int(fb.shape[1]))

    def add_reads(self, lib: int, r1_seq, r1_qual, feature=None, r2_seq=None, r2_qual=None):
        r1_seq = np.ascontiguousarray(r1_seq, dtype=np.uint8)
        r1_qual = np.ascontiguousarray(r1_qual, dtype=np.uint8)
        n, r1_len = r1_seq.shape
        feat = None if feature is None else np.ascontiguousarray(feature, dtype=np.uint32)
        r2_len = 0
        if r2_seq is not None:
            r2_seq = np.ascontiguousarray(r2_seq, dtype=np.uint8)
            r2_qual = np.ascontiguousarray(r2_qual, dtype=np.uint8)
            r2_len = r2_seq.shape[1]
        self._keep += [r1_seq, r1_qual, feat, r2_seq, r2_qual]
        self.L.cro_add_reads(self.ctx, lib, C.c_uint64(n), r1_len, _p(r1_seq), _p(r1_qual), _p(feat),
                             r2_len, _p(r2_seq), _p(r2_qual))

    def reset_reads(self):
        self.L.cro_reset_reads(self.ctx)
        self._keep = []

    def pass1(self, threads=1):
        self.L.cro_pass1(self.ctx, threads)

    def pass2(self, threads=1):
        self.L.cro_pass2(self.ctx, threads)

    def count(self, threads=1):
        self.L.cro_count(self.ctx, threads)

    def run(self, threads=1):
        self.L.cro_run(self.ctx, threads)

    # ---- cross-chunk state (priors are global per library type) ----
    def prior_add(self, lib, seqs, counts):
        s = ascii_mat(seqs)
        c = np.ascontiguousarray(counts, dtype=np.int64)
        self.L.cro_prior_add(self.ctx, lib, _p(s), C.c_uint64(len(c)), int(s.shape[1]), _p(c))

    def prior_clear(self, lib):
        self.L.cro_prior_clear(self.ctx, lib)

    def fb_counts_set(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        self.L.cro_fb_counts_set(self.ctx, _p(c))

    # ---- results ----
    def counts(self, lib, which, seqs) -> np.ndarray:
        s = ascii_mat(seqs)
        out = np.zeros(s.shape[0], dtype=np.int64)
        self.L.cro_get_counts(self.ctx, lib, which, _p(s), C.c_uint64(s.shape[0]), int(s.shape[1]), _p(out))
        return out

    def fb_counts(self) -> np.ndarray:
        out = np.zeros(self.n_features, dtype=np.int64)
        self.L.cro_get_fb_counts(self.ctx, _p(out))
        return out

    def feat_dist(self) -> np.ndarray:
        out = np.zeros(self.n_features, dtype=np.float64)
        self.L.cro_get_feat_dist(self.ctx, _p(out))
        return out

    def reads(self) -> dict:
        n = int(self.L.cro_n_reads(self.ctx))
        bc = np.zeros((n, self.bc_len), dtype=np.uint8)
        umi = np.zeros((n, self.umi_len), dtype=np.uint8)
        state = np.zeros(n, dtype=np.uint8)
        flags = np.zeros(n, dtype=np.uint8)
        feature = np.zeros(n, dtype=np.uint32)
        rc = np.zeros(n, dtype=np.uint32)
        self.L.cro_get_reads(self.ctx, self.bc_len, self.umi_len, _p(bc), _p(state), _p(umi), _p(flags),
                             _p(feature), _p(rc))
        return dict(bc=bc, state=state, umi=umi, flags=flags, feature=feature, read_count=rc)

    def stats(self) -> dict:
        out = np.zeros(8, dtype=np.uint64)
        self.L.cro_get_stats(self.ctx, _p(out))
        keys = ["valid_before", "corrected", "invalid", "dup_reads", "umi_corrected_reads",
                "low_support_reads", "umis", "molecules"]
        return {k: int(v) for k, v in zip(keys, out)}

    def matrix(self) -> dict:
        nb = int(self.L.cro_matrix_n_barcodes(self.ctx))
        nnz = int(self.L.cro_matrix_nnz(self.ctx))
        barcodes = np.zeros((nb, self.bc_len), dtype=np.uint8)
        indptr = np.zeros(nb + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.uint32)
        data = np.zeros(nnz, dtype=np.int32)
        self.L.cro_matrix_get(self.ctx, self.bc_len, _p(barcodes), _p(indptr), _p(indices), _p(data))
        return dict(barcodes=barcodes, indptr=indptr, indices=indices, data=data)

    def molecules(self) -> np.ndarray:
        """UmiCount rows in the order ALIGN_AND_COUNT emits them (by barcode, then umi_counts.sort()):
        (barcode column, library_idx, feature_idx, umi 2-bit, read_count, umi_type as in molecule_info)."""
        n = int(self.L.cro_n_molecules(self.ctx))
        out = np.zeros((n, 6), dtype=np.uint32)
        self.L.cro_molecules_get(self.ctx, _p(out))
        return out

    def set_select_keys(self, keys):
        """UmiSelectKey per read as one word: bit 63 = 1 for NonTxomic, low bits = rank of the qname."""
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        self.L.cro_set_select_keys(self.ctx, _p(k), C.c_uint64(k.shape[0]))


# ---- single-function known-answer entry points ----

def probability(q: int) -> float:
    return lib().cro_probability(q)


def kat_correct_barcode(wl, counts: dict, observed, qual, max_expected_errors, threshold, trans=None):
    L_ = lib()
    w = ascii_mat(wl)
    t = ascii_mat(trans) if trans is not None else None
    Ln = w.shape[1]
    cs = ascii_mat(list(counts.keys()), Ln) if counts else np.zeros((0, Ln), dtype=np.uint8)
    cv = np.array(list(counts.values()), dtype=np.int64)
    obs = ascii_mat([observed])
    q = None if qual is None else np.ascontiguousarray(np.frombuffer(bytes(qual), dtype=np.uint8))
    out = np.zeros(Ln, dtype=np.uint8)
    ok = L_.cro_kat_correct_barcode(_p(w), C.c_uint64(w.shape[0]), Ln, _p(t), _p(cs), _p(cv), C.c_uint64(len(cv)),
                                    _p(obs), _p(q), C.c_double(max_expected_errors), C.c_double(threshold), _p(out))
    return out.tobytes() if ok else None


def kat_match_to_whitelist(wl, seq):
    w = ascii_mat(wl)
    s = ascii_mat([seq])
    out = np.zeros(w.shape[1], dtype=np.uint8)
    ok = lib().cro_kat_match_to_whitelist(_p(w), C.c_uint64(w.shape[0]), int(w.shape[1]), _p(s), _p(out))
    return out.tobytes() if ok else None


def kat_correct_umis(table):
    """table: list of (umi, gene, count). Returns dict {(umi, gene): corrected_umi} of recorded corrections."""
    um = ascii_mat([t[0] for t in table])
    g = np.array([t[1] for t in table], dtype=np.uint32)
    c = np.array([t[2] for t in table], dtype=np.uint64)
    out = np.zeros_like(um)
    n = lib().cro_kat_correct_umis(_p(um), _p(g), _p(c), C.c_uint64(len(table)), int(um.shape[1]), _p(out))
    res = {}
    for i, t in enumerate(table):
        o = out[i].tobytes()
        k = t[0].encode() if isinstance(t[0], str) else bytes(t[0])
        if o != k:
            res[(k, t[1])] = o
    assert len(res) == n
    return res


def kat_low_support(table):
    um = ascii_mat([t[0] for t in table])
    g = np.array([t[1] for t in table], dtype=np.uint32)
    c = np.array([t[2] for t in table], dtype=np.uint64)
    out = np.zeros(len(table), dtype=np.uint8)
    lib().cro_kat_low_support(_p(um), _p(g), _p(c), C.c_uint64(len(table)), int(um.shape[1]), _p(out))
    return out.astype(bool)


def kat_umi_is_valid(seq: bytes, qual: bytes) -> bool:
    s = ascii_mat([seq])
    q = ascii_mat([qual])
    return bool(lib().cro_kat_umi_is_valid(_p(s), _p(q), len(seq)))


def kat_encode_2bit(seq: bytes) -> int:
    s = ascii_mat([seq])
    return int(lib().cro_kat_encode_2bit(_p(s), len(seq)))


def kat_feature_dist(raw, ftype) -> np.ndarray:
    r = np.ascontiguousarray(raw, dtype=np.int64)
    f = np.ascontiguousarray(ftype, dtype=np.int32)
    out = np.zeros(len(r), dtype=np.float64)
    lib().cro_kat_feature_dist(_p(r), _p(f), len(r), _p(out))
    return out


def kat_feature_match(feat_seqs, feat_idx, feat_dist, seq, qual) -> int:
    fs = ascii_mat(feat_seqs)
    fi = np.ascontiguousarray(feat_idx, dtype=np.int32)
    fd = None if feat_dist is None else np.ascontiguousarray(feat_dist, dtype=np.float64)
    s = ascii_mat([seq])
    q = ascii_mat([qual])
    return int(lib().cro_kat_feature_match(_p(fs), _p(fi), len(fi), int(fs.shape[1]), _p(fd),
                                           0 if fd is None else len(fd), _p(s), _p(q)))
