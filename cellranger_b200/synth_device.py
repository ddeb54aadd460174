"""Runs the synthetic read generator on the GPU (csrc/synth.cu), straight into HBM.

Same function as synth.generate_reads (numpy); tables are built once on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import synth
from ._lib import SynthParams, check, ptr


def _params(t: synth.SynthTables, library: str) -> tuple:
    cfg = t.cfg
    fb = library == "fb"
    p = SynthParams()
    p.seed_mix = int(synth._mix_seed(cfg.seed, 11 if fb else 7))
    p.seed_mol = int(synth._mix_seed(cfg.seed, 13))
    p.n_whitelist = cfg.n_whitelist
    p.n_cells = int(t.cell_rank.shape[0])
    p.n_genes = cfg.n_genes
    p.n_fb = int(t.fb_cdf.shape[0])
    p.bc_len, p.umi_len = cfg.bc_len, cfg.umi_len
    p.fb_offset, p.fb_len, p.is_fb = cfg.fb_offset, cfg.fb_len, int(fb)
    p.ambient_thr = synth._thr32(cfg.ambient_frac)
    p.unmapped_thr = synth._thr32(cfg.unmapped_frac)
    p.bc_err_thr = synth._thr32(cfg.bc_err)
    p.umi_err_thr = synth._thr32(cfg.umi_err)
    p.n_thr = synth._thr32(cfg.n_frac)
    p.fb_err_thr = synth._thr32(cfg.fb_err)
    p.homopolymer_thr = synth._thr32(cfg.umi_homopolymer_frac) if cfg.umi_homopolymer_frac > 0 else 0
    p.n_qual_ascii = cfg.n_qual + 33
    p.n_qual_classes = len(t.qual_val)
    p.n_equal_classes = len(t.equal_val)
    for i in range(len(t.qual_val)):
        p.qual_thr[i] = int(t.qual_thr[i])
        p.qual_val[i] = int(t.qual_val[i])
    for i in range(len(t.equal_val)):
        p.equal_thr[i] = int(t.equal_thr[i])
        p.equal_val[i] = int(t.equal_val[i])
    wl = t.trans_packed if (fb and t.trans_packed is not None) else t.wl_packed
    keep = [np.ascontiguousarray(wl, dtype=np.uint32), np.ascontiguousarray(t.cell_rank, dtype=np.uint32),
            np.ascontiguousarray(t.cell_cdf, dtype=np.uint32), np.ascontiguousarray(t.n_mol, dtype=np.uint32),
            np.ascontiguousarray(t.gene_cdf, dtype=np.uint32), np.ascontiguousarray(t.fb_cdf, dtype=np.uint32),
            np.ascontiguousarray(synth.pack_2bit(t.fb_seqs).astype(np.uint32) if len(t.fb_seqs) else
                                 np.zeros(0, dtype=np.uint32))]
    p.wl_packed, p.cell_rank, p.cell_cdf, p.n_mol, p.gene_cdf, p.fb_cdf, p.fb_packed = [a.ctypes.data for a in keep]
    return p, keep


class DeviceReads:
    """Device buffers of one synthetic batch (freed on close())."""

    def __init__(self, gw, n, r1_len, r2_len, is_fb):
        self.gw, self.n, self.r1_len, self.r2_len, self.is_fb = gw, n, r1_len, r2_len, is_fb
        self.r1_seq = self._alloc(n * r1_len + 64)
        self.r1_qual = self._alloc(n * r1_len + 64)
        self.feature = 0 if is_fb else self._alloc(n * 4 + 64)
        self.r2_seq = self._alloc(n * r2_len + 64) if is_fb else 0
        self.r2_qual = self._alloc(n * r2_len + 64) if is_fb else 0

    def _alloc(self, nbytes):
        p = C.c_void_p()
        check(self.gw.L.crgpu_dev_alloc(self.gw.ctx, C.c_uint64(nbytes), C.byref(p)), "crgpu_dev_alloc")
        return p.value

    @property
    def input_bytes(self) -> int:
        return self.n * (2 * self.r1_len + (2 * self.r2_len if self.is_fb else 4))

    def to_host(self) -> dict:
        out = {}

        def pull(dev, shape, dtype):
            a = np.zeros(shape, dtype=dtype)
            check(self.gw.L.crgpu_memcpy_d2h(self.gw.ctx, ptr(a), C.c_void_p(dev), C.c_uint64(a.nbytes)))
            return a

        out["r1_seq"] = pull(self.r1_seq, (self.n, self.r1_len), np.uint8)
        out["r1_qual"] = pull(self.r1_qual, (self.n, self.r1_len), np.uint8)
        if self.is_fb:
            out["r2_seq"] = pull(self.r2_seq, (self.n, self.r2_len), np.uint8)
            out["r2_qual"] = pull(self.r2_qual, (self.n, self.r2_len), np.uint8)
        else:
            out["feature"] = pull(self.feature, (self.n,), np.uint32)
        return out

    def close(self):
        for p in (self.r1_seq, self.r1_qual, self.feature, self.r2_seq, self.r2_qual):
            if p:
                self.gw.L.crgpu_dev_free(self.gw.ctx, C.c_void_p(p))
        self.r1_seq = self.r1_qual = self.feature = self.r2_seq = self.r2_qual = 0


def generate_device(gw, t: synth.SynthTables, start: int, n: int, library: str = "gex") -> DeviceReads:
    fb = library == "fb"
    p, keep = _params(t, library)
    d = DeviceReads(gw, n, t.cfg.r1_len, t.cfg.r2_len if fb else 0, fb)
    check(gw.L.crgpu_synth_generate(gw.ctx, C.byref(p), C.c_uint64(start), C.c_uint64(n), C.c_void_p(d.r1_seq),
                                    C.c_void_p(d.r1_qual), C.c_void_p(d.feature or None),
                                    C.c_void_p(d.r2_seq or None), C.c_void_p(d.r2_qual or None)),
          "crgpu_synth_generate")
    del keep
    return d
