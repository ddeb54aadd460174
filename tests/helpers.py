"""Shared helpers of the parity tests: one synthetic problem fed to the CPU oracle and to the GPU product."""
from __future__ import annotations

import numpy as np

from cellranger_b200 import synth


def make_problem(name: str, n_reads: int, fb_frac=None, **overrides) -> dict:
    cfg = synth.preset(name, n_reads)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    if fb_frac is not None:
        cfg.fb_frac = fb_frac
    tables = synth.make_tables(cfg, n_reads)
    n_fb = int(round(n_reads * cfg.fb_frac)) if cfg.n_fb_features else 0
    n_gex = n_reads - n_fb
    prob = dict(cfg=cfg, tables=tables, n_gex=n_gex, n_fb=n_fb)
    prob["gex"] = synth.generate_reads(tables, 0, n_gex, "gex")
    if n_fb:
        prob["fb"] = synth.generate_reads(tables, 0, n_fb, "fb")
    return prob


def feature_tables(prob):
    cfg, t = prob["cfg"], prob["tables"]
    n_feat = cfg.n_genes + cfg.n_fb_features
    ftype = np.zeros(n_feat, dtype=np.int32)
    ftype[cfg.n_genes:] = 1
    fb_seqs = np.full((n_feat, max(cfg.fb_len, 1)), ord("A"), dtype=np.uint8)
    if cfg.n_fb_features:
        fb_seqs[cfg.n_genes:] = t.fb_seqs
    return ftype, fb_seqs


def run_oracle(prob, threads: int = 4, stages: bool = True):
    from oracle import cro

    cfg, t = prob["cfg"], prob["tables"]
    o = cro.Oracle()
    wl = o.add_whitelist(t.whitelist)
    lib = o.add_library(wl, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    fb_lib = None
    if prob["n_fb"]:
        wl2 = o.add_whitelist(t.trans, trans=t.whitelist)
        fb_lib = o.add_library(wl2, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len, umi_correction=True, is_fb=True, ftype=1,
                               fb_offset=cfg.fb_offset, fb_len=cfg.fb_len)
    ftype, fb_seqs = feature_tables(prob)
    o.set_features(ftype, fb_seqs)
    g = prob["gex"]
    o.add_reads(lib, g["r1_seq"], g["r1_qual"], g["feature"])
    if prob["n_fb"]:
        f = prob["fb"]
        o.add_reads(fb_lib, f["r1_seq"], f["r1_qual"], None, f["r2_seq"], f["r2_qual"])
    if stages:
        o.run(threads)
    return o


def run_gpu(prob, annotate: bool = True, run: bool = True):
    import cellranger_b200 as cb

    cfg, t = prob["cfg"], prob["tables"]
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    chem = cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    lib = gw.add_library(wl, chem)
    fb_lib = None
    if prob["n_fb"]:
        wl2 = gw.add_whitelist(cb.Whitelist.trans(t.trans, t.whitelist))
        fb_lib = gw.add_library(wl2, chem, umi_correction=True, feature_type=1, fb_offset=cfg.fb_offset,
                                fb_length=cfg.fb_len)
    fr = cb.FeatureReference(cfg.n_genes)
    for i in range(cfg.n_fb_features):
        fr.add_feature_barcode(f"FB{i}", bytes(t.fb_seqs[i]).decode(), 1, "5P" + "N" * cfg.fb_offset + "(BC)")
    gw.set_feature_reference(fr)
    g = prob["gex"]
    gw.add_reads(lib, g["r1_seq"], g["r1_qual"], g["feature"])
    if prob["n_fb"]:
        f = prob["fb"]
        gw.add_reads(fb_lib, f["r1_seq"], f["r1_qual"], None, f["r2_seq"], f["r2_qual"])
    if run:
        gw.run(annotate_reads=annotate)
    return gw


def compare_all(o, gw, prob, check_reads: bool = True):
    """Bit-exact comparison of every output the two sides share. Returns a dict of sizes."""
    import cellranger_b200 as cb

    cfg = prob["cfg"]
    Lb, Lu = cfg.bc_len, cfg.umi_len
    wl = prob["tables"].whitelist
    # priors / corrected counts per library
    n_libs = 2 if prob["n_fb"] else 1
    for lib in range(n_libs):
        assert np.array_equal(o.counts(lib, 0, wl), gw.prior(lib).astype(np.int64)), f"prior of library {lib}"
        assert np.array_equal(o.counts(lib, 1, wl), gw.corrected_counts(lib).astype(np.int64)), f"corrected counts {lib}"
    if prob["n_fb"]:
        assert np.array_equal(o.fb_counts(), gw.fb_exact_counts())
    # per-read
    if check_reads:
        ro = o.reads()
        base = 0
        for batch in range(n_libs):
            rg = gw.reads(batch)
            n = rg["state"].shape[0]
            so = ro["state"][base:base + n]
            assert np.array_equal(so, rg["state"]), f"barcode state, batch {batch}"
            valid = (so == 1) | (so == 2)
            bc_g = gw.barcode_seqs(rg["bc_rank"][valid])
            assert np.array_equal(ro["bc"][base:base + n][valid], bc_g), f"corrected barcode, batch {batch}"
            assert np.all(rg["bc_rank"][~valid] == cb.NO_RANK)
            assert np.array_equal(ro["feature"][base:base + n], rg["feature"]), f"feature, batch {batch}"
            fo = ro["flags"][base:base + n]
            assert np.array_equal(fo, rg["flags"]), f"read flags, batch {batch}"
            has_dup = (fo & 2) != 0
            umi_g = cb.unpack_2bit(rg["umi"], Lu)
            assert np.array_equal(ro["umi"][base:base + n][has_dup], umi_g[has_dup]), f"processed UMI, batch {batch}"
            # BarcodeSummary rows of this library (aligner.rs:33-68), from the oracle's per-read DupInfo
            from oracle import pyref
            eb, er, eu, ec, ek = pyref.barcode_summary(ro["bc"][base:base + n], so, fo)
            sm = gw.barcode_summary(batch)
            assert np.array_equal(eb, gw.barcode_seqs(sm["barcode_rank"])), f"summary barcodes, library {batch}"
            assert np.array_equal(er, sm["reads"].astype(np.int64)), f"summary reads, library {batch}"
            assert np.array_equal(eu, sm["umis"].astype(np.int64)), f"summary umis, library {batch}"
            assert np.array_equal(ec, sm["candidate_dup_reads"].astype(np.int64)), f"candidate_dup_reads, library {batch}"
            assert np.array_equal(ek, sm["umi_corrected_reads"].astype(np.int64)), f"umi_corrected_reads, library {batch}"
            base += n
    # matrix
    mo = o.matrix()
    mg = gw.count_matrix()
    assert np.array_equal(mo["barcodes"], mg.barcodes), "barcode index"
    assert np.array_equal(mo["indptr"], mg.indptr), "indptr"
    assert np.array_equal(mo["indices"], mg.indices), "indices"
    assert np.array_equal(mo["data"], mg.data), "data"
    # molecules (UmiCount rows): the same rows in the same order - by barcode, then umi_counts.sort()
    # (library_idx, feature_idx, umi, read_count; cr_types/src/types.rs:152-160, align_and_count.rs:314)
    a = o.molecules()
    b = gw.molecules()
    assert a.shape == b.shape
    assert np.array_equal(a, b), "molecule rows (values or order)"
    so_, sg = o.stats(), gw.stats()
    assert so_["valid_before"] == sg["valid_before"]
    assert so_["corrected"] == sg["corrected"]
    assert so_["invalid"] == sg["invalid"]
    assert so_["umis"] == sg["molecules"]
    assert so_["dup_reads"] == sg["keys"]
    assert so_["umi_corrected_reads"] == sg["umi_corrected_reads"]
    assert so_["low_support_reads"] == sg["low_support_reads"]
    return dict(n_barcodes=len(mg.barcodes), nnz=len(mg.data), molecules=len(b), stats=sg)
