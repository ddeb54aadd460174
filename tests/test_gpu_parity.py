"""Parity of the CUDA path (through the C ABI) against the CPU oracle. Needs a B200: -m gpu."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

F64_MAX = 1.7976931348623157e308


# ---- the reference's known-answer tests, through the C ABI on the GPU ----

def test_kat_barcode_correction_gpu(kats):
    import cellranger_b200 as cb

    for name in ("barcode_correction", "barcode_correction_no_counts"):
        k = kats[name]
        corr = cb.BarcodeCorrector(cb.Whitelist.plain(k["whitelist"]), k["counts"],
                                   cb.Posterior(k["max_expected_errors"], k["threshold"]))
        seqs = [c["seq"] for c in k["cases"]]
        quals = [bytes(c["qual"]) for c in k["cases"]]
        out, state = corr.correct_barcodes(seqs, quals)
        for i, c in enumerate(k["cases"]):
            if c["expect"] is None:
                # "AAAAA" is itself on the whitelist: the batch seam reports the exact hit
                if c["seq"] in k["whitelist"]:
                    assert state[i] == cb.api.VALID_BEFORE_CORRECTION
                else:
                    assert state[i] == cb.api.INVALID, (name, c["name"])
            else:
                assert state[i] == cb.api.VALID_AFTER_CORRECTION, (name, c["name"])
                assert bytes(out[i]).decode() == c["expect"], (name, c["name"])
        corr.close()


def test_kat_n_rescue_gpu(kats):
    import cellranger_b200 as cb

    k = kats["barcode_n_rescue"]
    bc = k["whitelist"][0]
    corr = cb.BarcodeCorrector(cb.Whitelist.plain(k["whitelist"]), {}, cb.Posterior(1.0, k["threshold"]))
    seqs, quals = [], []
    for p in range(16):
        seqs.append(bc[:p] + "N" + bc[p + 1:])
        q = [k["qual_else"]] * 16
        q[p] = k["qual_at_n"]
        quals.append(bytes(q))
    seqs.append("NN" + bc[2:])
    quals.append(bytes([53] * 16))
    out, state = corr.correct_barcodes(seqs, quals)
    for p in range(16):
        assert state[p] == cb.api.VALID_AFTER_CORRECTION and bytes(out[p]).decode() == bc
    assert state[16] == cb.api.INVALID
    corr.close()


def test_corrector_batch_matches_oracle_random():
    """Random invalid segments with random priors and qualities: every accept/reject decision and every
    corrected sequence must equal the oracle's (f64 posterior in the reference's operation order)."""
    import cellranger_b200 as cb
    from oracle import cro

    rng = np.random.default_rng(7)
    for L, W in ((16, 5000), (8, 3000), (5, 200)):
        wl = np.unique(rng.integers(0, 4, size=(W, L)), axis=0)
        wl_ascii = np.frombuffer(b"ACGT", dtype=np.uint8)[wl]
        counts = {bytes(s): int(c) for s, c in zip(wl_ascii, rng.integers(0, 500, size=len(wl_ascii))) if c % 3}
        corr = cb.BarcodeCorrector(cb.Whitelist.plain(wl_ascii), counts)
        n = 3000
        src = wl_ascii[rng.integers(0, len(wl_ascii), size=n)].copy()
        for i in range(n):  # 1-2 substitutions, sometimes an N
            for _ in range(rng.integers(1, 3)):
                src[i, rng.integers(0, L)] = b"ACGT"[rng.integers(0, 4)]
            if rng.random() < 0.1:
                src[i, rng.integers(0, L)] = ord("N")
        quals = rng.integers(33 + 2, 33 + 41, size=(n, L)).astype(np.uint8)
        out, state = corr.correct_barcodes(src, quals)
        wl_set = {bytes(s) for s in wl_ascii}
        for i in range(n):
            s = bytes(src[i])
            if s in wl_set:
                assert state[i] == 1
                continue
            exp = cro.kat_correct_barcode(wl_ascii, counts, s, bytes(quals[i]), F64_MAX, 0.975)
            if exp is None:
                assert state[i] == 3, (L, i, s)
            else:
                assert state[i] == 2 and bytes(out[i]) == exp, (L, i, s, exp, bytes(out[i]))
        corr.close()


def test_synth_device_matches_numpy():
    import cellranger_b200 as cb
    from cellranger_b200 import synth, synth_device

    n = 50_000
    for name in ("cfg1", "cfg5", "cfg4"):
        cfg = synth.preset(name, n)
        cfg.n_whitelist = 20_000
        cfg.n_cells = 50
        t = synth.make_tables(cfg, n)
        gw = cb.GemWell()
        for libname in (("gex", "fb") if cfg.n_fb_features else ("gex",)):
            d = synth_device.generate_device(gw, t, 1000, n, libname)
            h = d.to_host()
            ref = synth.generate_reads(t, 1000, n, libname)
            for key in h:
                assert np.array_equal(h[key], ref[key]), (name, libname, key)
            d.close()
        gw.close()


# ---- whole-path parity on the BASELINE.json configs, at sizes the oracle finishes in seconds ----

@pytest.mark.parametrize("name,n,kw", [
    ("cfg1", 200_000, {}),                                        # 3' v2, 737K whitelist
    ("cfg2", 300_000, {"n_whitelist": 200_000, "n_cells": 400}),   # 3' v3 (12 bp UMI)
    ("cfg5", 200_000, {"n_whitelist": 100_000, "n_cells": 20}),    # high error, saturated UMIs
    ("cfg1", 1000, {"n_whitelist": 5000, "n_cells": 10}),          # ragged: below one tile
    ("cfg1", 513, {"n_whitelist": 5000, "n_cells": 3}),            # one tile + 1
])
def test_full_path_matches_oracle(name, n, kw):
    prob = helpers.make_problem(name, n, **kw)
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    assert info["nnz"] > 0
    gw.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4", "cfg5"])
def test_full_path_matches_oracle_1m_full_whitelist(name):
    """BASELINE.json's configs as written except for the read count: 1 M reads against the FULL whitelists
    (737 280 / 6 794 880 entries; cfg4 with its translated feature-barcode whitelist and Antibody Capture
    library, cfg5 with 5 % barcode errors and saturated UMIs), every output bit-exact against the oracle:
    priors, per-read barcode / state / UMI / flags, barcode index, CSC arrays, molecule rows, summaries."""
    prob = helpers.make_problem(name, 1_000_000)
    cfg = prob["cfg"]
    assert cfg.n_whitelist == (737_280 if name == "cfg1" else 6_794_880)
    if name == "cfg4":
        assert prob["n_fb"] > 0
    o = helpers.run_oracle(prob, threads=16)
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    assert info["nnz"] > 0 and info["stats"]["reads"] == 1_000_000
    gw.close()
    o.close()


def test_feature_index_out_of_range_is_an_error():
    """A feature index the matrix has no row for must not reach a key (it would spill into the barcode-rank
    bits): the batch is refused with CRGPU_E_INVALID, and a GEX batch without crgpu_features_set is refused too."""
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 5000, n_whitelist=3000, n_cells=5)
    g = prob["gex"]
    bad = g["feature"].copy()
    mapped = np.flatnonzero(bad != 0xFFFFFFFF)
    bad[mapped[7]] = prob["cfg"].n_genes          # first index past the matrix
    bad[mapped[11]] = 0x7FFFFFFF
    prob["gex"]["feature"] = bad
    gw = helpers.run_gpu(prob, run=False)
    gw.make_shard()
    gw.barcode_correction()
    with pytest.raises(cb.CrgpuError, match="feature index"):
        gw.align_and_count()
    gw.close()
    t = prob["tables"]
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    lib = gw.add_library(wl, cb.ChemistryDef.SC3Pv2())
    gw.add_reads(lib, g["r1_seq"], g["r1_qual"], g["feature"])
    with pytest.raises(cb.CrgpuError, match="crgpu_features_set"):
        gw.make_shard()
    gw.close()


def test_barcode_correction_twice_is_idempotent():
    """crgpu_pass2 called again (a retry) must not add the corrections, or the corrected reads' keys, twice."""
    prob = helpers.make_problem("cfg1", 60_000)
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob, run=False)
    gw.make_shard()
    gw.barcode_correction()
    gw.barcode_correction()
    gw.align_and_count(annotate_reads=True)
    helpers.compare_all(o, gw, prob)
    # after the count stage the key buffer is sorted: a further pass 2 is refused, a whole new run is fine
    import cellranger_b200 as cb
    with pytest.raises(cb.CrgpuError, match="crgpu_pass1 first"):
        gw.barcode_correction()
    gw.run(annotate_reads=True)
    helpers.compare_all(o, gw, prob)
    gw.close()


def test_feature_barcode_library_matches_oracle():
    prob = helpers.make_problem("cfg4", 200_000, n_whitelist=100_000, n_cells=200)
    assert prob["n_fb"] > 0
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    assert info["nnz"] > 0
    gw.close()


def test_full_whitelist_sizes():
    """The real whitelist sizes (737 280 and 6 794 880 entries) with their 2- and 3-ordering tables."""
    for name, n in (("cfg1", 100_000), ("cfg2", 100_000)):
        prob = helpers.make_problem(name, n, n_cells=100)
        o = helpers.run_oracle(prob)
        gw = helpers.run_gpu(prob)
        helpers.compare_all(o, gw, prob)
        gw.close()


def test_edge_cases():
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 4000, n_whitelist=3000, n_cells=5)
    g = prob["gex"]
    # all reads unmapped: barcodes still enter the barcode index, matrix has empty columns only
    g["feature"][:] = 0xFFFFFFFF
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    assert info["nnz"] == 0 and info["n_barcodes"] > 0
    gw.close()
    # every barcode invalid and uncorrectable (poly-N)
    prob = helpers.make_problem("cfg1", 2000, n_whitelist=3000, n_cells=5)
    prob["gex"]["r1_seq"][:, :16] = ord("N")
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob)
    info = helpers.compare_all(o, gw, prob)
    assert info["n_barcodes"] == 0
    gw.close()
    # empty batch
    prob = helpers.make_problem("cfg1", 1000, n_whitelist=3000, n_cells=5)
    for k in ("r1_seq", "r1_qual", "feature", "true_rank"):
        prob["gex"][k] = prob["gex"][k][:0]
    gw = helpers.run_gpu(prob)
    assert gw.stats()["reads"] == 0 and gw.count_matrix().data.shape[0] == 0
    gw.close()


def test_umi_chain_and_low_support_hand_case():
    """A hand-built barcode: chain A->B->C (single hop), tie broken towards the larger UMI, a UMI seen
    with two genes (low support), homopolymer and low-quality UMIs dropped."""
    import cellranger_b200 as cb
    from oracle import cro

    wl = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT"]
    rows = []  # (barcode, umi, gene, copies, qual char)

    def add(bc, umi, gene, copies, q="I"):
        for _ in range(copies):
            rows.append((bc, umi, gene, q))

    b = wl[0]
    add(b, "AAAAAAAAAC", 5, 1)   # A -> B (count 2) ...
    add(b, "AAAAAAAACC", 5, 2)   # B -> C (count 3): single hop, B keeps A's read
    add(b, "AAAAAAACCC", 5, 3)   # C
    add(b, "CCCCCCCCCA", 7, 1)   # tie 1:1 -> lexicographically larger CCCCCCCCCG
    add(b, "CCCCCCCCCG", 7, 1)
    add(b, "GATTACAGAT", 3, 4)   # same UMI on two genes: gene 9 is sub-maximal -> low support
    add(b, "GATTACAGAT", 9, 1)
    add(b, "TTTTGGGGCC", 3, 2)   # tie across genes -> both low support
    add(b, "TTTTGGGGCC", 4, 2)
    add(b, "GGGGGGGGGG", 3, 5)   # homopolymer: invalid UMI
    add(b, "ACGTTGCAAC", 3, 2, q="*")  # Q9 < 10: invalid UMI
    add(wl[1], "ACGTTGCAAC", 3, 2)
    n = len(rows)
    r1 = np.zeros((n, 26), dtype=np.uint8)
    q1 = np.zeros((n, 26), dtype=np.uint8)
    feat = np.zeros(n, dtype=np.uint32)
    for i, (bc, umi, gene, q) in enumerate(rows):
        r1[i] = np.frombuffer((bc + umi).encode(), dtype=np.uint8)
        q1[i] = ord("I")
        q1[i, 16:] = ord(q)
        feat[i] = gene
    o = cro.Oracle()
    w = o.add_whitelist(wl)
    lib = o.add_library(w, 0, 16, 16, 10)
    o.set_features(np.zeros(16, dtype=np.int32))
    o.add_reads(lib, r1, q1, feat)
    o.run()
    gw = cb.GemWell()
    w2 = gw.add_whitelist(cb.Whitelist.plain(wl))
    lib2 = gw.add_library(w2, cb.ChemistryDef.SC3Pv2())
    gw.set_feature_reference(cb.FeatureReference(16))
    gw.add_reads(lib2, r1, q1, feat)
    gw.run(annotate_reads=True)
    mo, mg = o.matrix(), gw.count_matrix()
    assert np.array_equal(mo["indptr"], mg.indptr) and np.array_equal(mo["indices"], mg.indices)
    assert np.array_equal(mo["data"], mg.data)
    # spelled out: barcode 0 → gene 3: GATTACAGAT (1), gene 5: B and C (2), gene 7: CCCCCCCCCG (1)
    assert mg.indices.tolist() == [3, 5, 7, 3] and mg.data.tolist() == [1, 2, 1, 1]
    ro, rg = o.reads(), gw.reads(0)
    assert np.array_equal(ro["flags"], rg["flags"])
    has = (ro["flags"] & 2) != 0
    assert np.array_equal(ro["umi"][has], cb.unpack_2bit(rg["umi"], 10)[has])
    gw.close()


def test_multiplexing_capture_disables_umi_correction():
    import cellranger_b200 as cb
    from oracle import cro

    prob = helpers.make_problem("cfg1", 50_000, n_whitelist=5000, n_cells=20)
    g = prob["gex"]
    t = prob["tables"]
    o = cro.Oracle()
    w = o.add_whitelist(t.whitelist)
    lib = o.add_library(w, 0, 16, 16, 10, umi_correction=False)
    o.set_features(np.zeros(prob["cfg"].n_genes, dtype=np.int32))
    o.add_reads(lib, g["r1_seq"], g["r1_qual"], g["feature"])
    o.run(4)
    gw = cb.GemWell()
    w2 = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    lib2 = gw.add_library(w2, cb.ChemistryDef.SC3Pv2(), umi_correction=False)
    gw.set_feature_reference(cb.FeatureReference(prob["cfg"].n_genes))
    gw.add_reads(lib2, g["r1_seq"], g["r1_qual"], g["feature"])
    gw.run(annotate_reads=True)
    assert gw.stats()["umi_corrected_keys"] == 0
    mo, mg = o.matrix(), gw.count_matrix()
    assert np.array_equal(mo["data"], mg.data) and np.array_equal(mo["indices"], mg.indices)
    gw.close()


def test_size_independent_properties_large():
    """2 M reads (the oracle is not run): sortedness, conservation and idempotence."""
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 2_000_000)
    gw = helpers.run_gpu(prob, annotate=True)
    st = gw.stats()
    m = gw.count_matrix()
    assert st["valid_before"] + st["corrected"] + st["invalid"] == st["reads"] == 2_000_000
    assert np.all(np.diff(m.barcode_rank.astype(np.int64)) > 0)  # barcode index strictly sorted
    assert m.indptr[0] == 0 and m.indptr[-1] == len(m.data) and np.all(np.diff(m.indptr) >= 0)
    for c in np.random.default_rng(0).integers(0, len(m.indptr) - 1, size=2000):  # features sorted per column
        seg = m.indices[m.indptr[c]:m.indptr[c + 1]]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
    assert int(m.data.sum()) == st["molecules"] and np.all(m.data > 0)
    mol = gw.molecules()
    assert int(mol[:, 4].sum()) + st["low_support_reads"] == st["keys"]  # every deduped read is accounted for
    r = gw.reads(0)
    assert int(((r["flags"] & 16) != 0).sum()) == st["molecules"]  # one representative read per UMI
    # idempotence: running the stages again gives the same matrix
    gw.run(annotate_reads=False)
    m2 = gw.count_matrix()
    assert np.array_equal(m.indptr, m2.indptr) and np.array_equal(m.indices, m2.indices)
    assert np.array_equal(m.data, m2.data)
    gw.close()


def test_repeated_runs_are_deterministic(monkeypatch):
    """The staged pass-1 kernel refills its shared-memory stages while other warps are still working; a
    missing ordering there shows up as run-to-run differences (it did once: stale feature words). Ten
    back-to-back runs must give identical key sets and matrices, with the sort / RLE self-checks on."""
    import ctypes as C

    from cellranger_b200._lib import check, ptr

    monkeypatch.setenv("CRGPU_VERIFY", "1")
    prob = helpers.make_problem("cfg1", 1_500_000)
    gw = helpers.run_gpu(prob, annotate=False, run=False)
    ref_keys = ref_m = None
    for it in range(10):
        gw.make_shard()
        gw.barcode_correction()
        p, nk = gw.keys_dev()
        k = np.zeros(nk, dtype=np.uint64)
        check(gw.L.crgpu_memcpy_d2h(gw.ctx, ptr(k), C.c_void_p(p), C.c_uint64(nk * 8)))
        k.sort()
        gw.align_and_count()
        st = gw.stats()
        assert st["sort_violations"] == 0 and st["rle_violations"] == 0
        m = gw.count_matrix()
        if ref_keys is None:
            ref_keys, ref_m = k, m
        else:
            assert np.array_equal(k, ref_keys), f"key set differs in run {it}"
            assert np.array_equal(m.indptr, ref_m.indptr) and np.array_equal(m.indices, ref_m.indices)
            assert np.array_equal(m.data, ref_m.data)
    gw.close()


def test_long_umi_segments_span_several_windows():
    """(barcode, gene) segments far longer than the shared-memory window of the UMI-correction kernel
    (one cell, one or two genes, up to ~40 k distinct UMIs each): every key must still see its whole segment."""
    import cellranger_b200 as cb
    from oracle import cro

    rng = np.random.default_rng(5)
    wl = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT"]
    n = 120_000
    r1 = np.zeros((n, 26), dtype=np.uint8)
    r1[:, :16] = np.frombuffer(wl[0].encode(), dtype=np.uint8)
    r1[n // 2:, :16] = np.frombuffer(wl[1].encode(), dtype=np.uint8)
    umis = rng.integers(0, 4, size=(n, 10))
    umis[:, :2] = 0  # 4^8 = 65536 UMIs: dense Hamming-1 neighbourhoods, chains and ties
    r1[:, 16:] = np.frombuffer(b"ACGT", dtype=np.uint8)[umis]
    q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
    feat = (rng.random(n) < 0.7).astype(np.uint32)  # gene 0: ~30 %, gene 1: ~70 %
    o = cro.Oracle()
    w = o.add_whitelist(wl)
    lib = o.add_library(w, 0, 16, 16, 10)
    o.set_features(np.zeros(2, dtype=np.int32))
    o.add_reads(lib, r1, q1, feat)
    o.run(4)
    gw = cb.GemWell()
    w2 = gw.add_whitelist(cb.Whitelist.plain(wl))
    lib2 = gw.add_library(w2, cb.ChemistryDef.SC3Pv2())
    gw.set_feature_reference(cb.FeatureReference(2))
    gw.add_reads(lib2, r1, q1, feat)
    gw.run(annotate_reads=True)
    assert gw.stats()["distinct_keys"] > 50_000
    mo, mg = o.matrix(), gw.count_matrix()
    assert np.array_equal(mo["indptr"], mg.indptr) and np.array_equal(mo["indices"], mg.indices)
    assert np.array_equal(mo["data"], mg.data)
    ro, rg = o.reads(), gw.reads(0)
    assert np.array_equal(ro["flags"], rg["flags"])
    has = (ro["flags"] & 2) != 0
    assert np.array_equal(ro["umi"][has], cb.unpack_2bit(rg["umi"], 10)[has])
    so, sg = o.stats(), gw.stats()
    assert so["umi_corrected_reads"] == sg["umi_corrected_reads"] and so["low_support_reads"] == sg["low_support_reads"]
    gw.close()


def test_barcode_correction_metrics_match_oracle_counts():
    """corrected_bc / good_bc of InnerBarcodeCorrectionMetrics (barcode_correction_metrics.rs:16-39,62-86)."""
    prob = helpers.make_problem("cfg1", 60_000)
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob, annotate=False)
    so = o.stats()
    m = gw.barcode_correction_metrics(0)
    n = 60_000
    assert m["total_reads"] == n and m["valid_before"] == so["valid_before"] and m["corrected"] == so["corrected"]
    assert m["corrected_bc"] == so["corrected"] / n
    assert m["good_bc"] == (so["valid_before"] + so["corrected"]) / n
    gw.close()


def test_select_keys_choose_the_representative_read_and_the_umi_type():
    """UmiSelectKey{utype, qname} (tx_annotation/src/mark_dups.rs:110-152,248-268,300-326) through
    crgpu_read_batch.select_key: random qname ranks and a random Txomic / NonTxomic bit per read. The representative
    read of every molecule (is_umi_count) and UmiCount::utype must follow the smallest (utype, qname), not the
    read index."""
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 120_000, n_whitelist=20_000, n_cells=30)
    g = prob["gex"]
    n = prob["n_gex"]
    rng = np.random.default_rng(42)
    keys = rng.permutation(n).astype(np.uint64) | (rng.integers(0, 2, size=n).astype(np.uint64) << np.uint64(63))
    o = helpers.run_oracle(prob, stages=False)
    o.set_select_keys(keys)
    o.run(4)
    cfg, t = prob["cfg"], prob["tables"]
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    lib = gw.add_library(wl, cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len))
    gw.set_feature_reference(cb.FeatureReference(cfg.n_genes))
    gw.add_reads(lib, g["r1_seq"], g["r1_qual"], g["feature"], select_key=keys)
    gw.make_shard()
    gw.barcode_correction()
    gw.align_and_count()
    with pytest.raises(cb.CrgpuError, match="crgpu_annotate_reads"):
        gw.molecules()  # umi_type needs the representative reads
    gw.align_and_count(annotate_reads=True)
    info = helpers.compare_all(o, gw, prob)
    mol = gw.molecules()
    assert 0 < int((mol[:, 5] == 0).sum()) < len(mol)  # both types occur
    # and it differs from the default choice (lowest read index) somewhere
    ref = helpers.run_gpu(prob)
    assert not np.array_equal(ref.reads(0)["flags"], gw.reads(0)["flags"])
    assert np.array_equal(ref.count_matrix().data, gw.count_matrix().data)  # the matrix does not depend on it
    assert info["molecules"] == len(mol)
    ref.close()
    gw.close()


def test_molecule_rows_are_library_major_when_libraries_are_not_in_feature_order():
    """UmiCount sorts by library_idx before feature_idx (cr_types/src/types.rs:152-160). Here the Antibody Capture
    library is library 0 although its features are the LAST rows of the matrix, so the key order (feature above
    library) is not the row order: the rows must come out re-sorted, exactly as the oracle's umi_counts.sort()."""
    import cellranger_b200 as cb
    from oracle import cro

    prob = helpers.make_problem("cfg4", 120_000, n_whitelist=50_000, n_cells=60)
    cfg, t = prob["cfg"], prob["tables"]
    ftype, fb_seqs = helpers.feature_tables(prob)
    g, f = prob["gex"], prob["fb"]
    o = cro.Oracle()
    wl_fb = o.add_whitelist(t.trans, trans=t.whitelist)
    lib_fb = o.add_library(wl_fb, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len, umi_correction=True, is_fb=True, ftype=1,
                           fb_offset=cfg.fb_offset, fb_len=cfg.fb_len)
    wl_gex = o.add_whitelist(t.whitelist)
    lib_gex = o.add_library(wl_gex, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    o.set_features(ftype, fb_seqs)
    o.add_reads(lib_fb, f["r1_seq"], f["r1_qual"], None, f["r2_seq"], f["r2_qual"])
    o.add_reads(lib_gex, g["r1_seq"], g["r1_qual"], g["feature"])
    o.run(4)
    gw = cb.GemWell()
    # the first whitelist defines the content space: the translation whitelist maps onto the plain one's sequences
    w_fb = gw.add_whitelist(cb.Whitelist.trans(t.trans, t.whitelist))
    chem = cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    l_fb = gw.add_library(w_fb, chem, umi_correction=True, feature_type=1, fb_offset=cfg.fb_offset, fb_length=cfg.fb_len)
    w_gex = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    l_gex = gw.add_library(w_gex, chem)
    fr = cb.FeatureReference(cfg.n_genes)
    for i in range(cfg.n_fb_features):
        fr.add_feature_barcode(f"FB{i}", bytes(t.fb_seqs[i]).decode(), 1, "5P" + "N" * cfg.fb_offset + "(BC)")
    gw.set_feature_reference(fr)
    gw.add_reads(l_fb, f["r1_seq"], f["r1_qual"], None, f["r2_seq"], f["r2_qual"])
    gw.add_reads(l_gex, g["r1_seq"], g["r1_qual"], g["feature"])
    gw.run(annotate_reads=True)
    a, b = o.molecules(), gw.molecules()
    assert a.shape == b.shape and len(a) > 1000
    assert np.array_equal(a, b), "UmiCount rows: values or order"
    # library-major inside a barcode: somewhere a row of library 1 with a LOWER feature follows a row of library 0
    same_bc = b[1:, 0] == b[:-1, 0]
    assert np.any(same_bc & (b[1:, 1] > b[:-1, 1]) & (b[1:, 2] < b[:-1, 2]))
    mo, mg = o.matrix(), gw.count_matrix()
    assert np.array_equal(mo["indptr"], mg.indptr) and np.array_equal(mo["indices"], mg.indices)
    assert np.array_equal(mo["data"], mg.data)
    gw.close()


def test_finish_sort_path_matches_oracle(monkeypatch):
    """CRGPU_FINISH_SORT=1: radix sort on the bits above the UMI only, UMI bits sorted per segment in shared memory
    (pairwise / bucketed / pre-sorted long segments) fused with the run-length encoding. Same results as the default
    path on a case with short, medium (hundreds of reads) and long (tens of thousands) segments."""
    monkeypatch.setenv("CRGPU_FINISH_SORT", "1")
    monkeypatch.setenv("CRGPU_VERIFY", "1")
    for name, n, kw in (("cfg5", 600_000, {"n_whitelist": 100_000, "n_cells": 6}), ("cfg2", 400_000, {"n_whitelist": 200_000, "n_cells": 40}),
                        ("cfg1", 3000, {"n_whitelist": 5000, "n_cells": 10})):
        prob = helpers.make_problem(name, n, **kw)
        o = helpers.run_oracle(prob, threads=8)
        gw = helpers.run_gpu(prob)
        helpers.compare_all(o, gw, prob)
        st = gw.stats()
        assert st["sort_violations"] == 0 and st["rle_violations"] == 0
        gw.close()
        o.close()


def test_barcode_diversity_matches_histogram_formula():
    """barcodes_detected / effective_barcode_diversity of BARCODE_CORRECTION's join (barcode_correction.rs:428-441):
    (sum c)^2 / sum c^2 over the corrected barcode histogram, here from the oracle's raw-valid + corrected counts
    with the reference's own f64 expression (histogram.rs:161-171)."""
    prob = helpers.make_problem("cfg1", 80_000)
    o = helpers.run_oracle(prob)
    gw = helpers.run_gpu(prob, annotate=False)
    wl = prob["tables"].whitelist
    c = (o.counts(0, 0, wl) + o.counts(0, 1, wl)).astype(np.float64)
    c = c[c > 0]
    s, s2 = 0.0, 0.0
    for x in c:  # literally the loop of effective_diversity()
        s += x
        s2 += x ** 2
    d = gw.barcode_diversity(0)
    assert d["barcodes_detected"] == len(c)
    assert d["effective_barcode_diversity"] == s ** 2 / s2
    gw.close()


def _fastq_text(seq, qual, rng, trailing_newline=True):
    """4-line FASTQ records with headers of varying length (what bcl2fastq writes, roughly)."""
    parts = []
    for i in range(seq.shape[0]):
        head = f"@A00123:45:HXXXXXXXX:{1 + i % 4}:{1101 + int(rng.integers(0, 500))}:{i}:{int(rng.integers(1000, 99999))} 1:N:0:SI-GA-A1"
        parts.append(head.encode() + b"\n" + bytes(seq[i]) + b"\n+\n" + bytes(qual[i]) + b"\n")
    text = b"".join(parts)
    return text if trailing_newline else text[:-1]


@pytest.mark.parametrize("trailing_newline", [True, False])
def test_fastq_ingestion_matches_array_ingestion(trailing_newline):
    """FASTQ text -> device read arrays (crgpu_fastq_extract) gives the same arrays, and the same
    matrix, as handing the arrays over directly."""
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 50_000)
    g = prob["gex"]
    rng = np.random.default_rng(5)
    text = _fastq_text(g["r1_seq"], g["r1_qual"], rng, trailing_newline)
    ref = helpers.run_gpu(prob, annotate=False)
    m_ref = ref.count_matrix()

    cfg, t = prob["cfg"], prob["tables"]
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    lib = gw.add_library(wl, cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len))
    gw.set_feature_reference(cb.FeatureReference(cfg.n_genes))
    info = gw.add_fastq(lib, text, g["feature"])
    assert info["n_records"] == 50_000 and info["n_short"] == 0 and info["n_malformed"] == 0
    rl = info["read_len"]
    assert rl == cfg.bc_len + cfg.umi_len
    assert np.array_equal(gw.read_device(info["dev_seq"], (50_000, rl)), g["r1_seq"][:, :rl])
    assert np.array_equal(gw.read_device(info["dev_qual"], (50_000, rl)), g["r1_qual"][:, :rl])
    gw.run()
    m = gw.count_matrix()
    assert np.array_equal(m.barcode_rank, m_ref.barcode_rank) and np.array_equal(m.indptr, m_ref.indptr)
    assert np.array_equal(m.indices, m_ref.indices) and np.array_equal(m.data, m_ref.data)
    assert gw.stats()["molecules"] == ref.stats()["molecules"]
    gw.close()
    ref.close()


def test_paired_fastq_gz_ingestion_of_a_feature_barcode_library():
    """a GEX + Antibody Capture run fed from FASTQ text - R1 (gzip) of both libraries and the R2 of
    the feature-barcode library (the capture sequence comes out of R2's first fb_offset + fb_length cycles on the
    device) - gives the matrix, and the UmiCount rows, of the same reads handed over as arrays."""
    import gzip

    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg4", 60_000, n_whitelist=50_000, n_cells=50)
    cfg, t = prob["cfg"], prob["tables"]
    g, f = prob["gex"], prob["fb"]
    ref = helpers.run_gpu(prob, annotate=False)
    rng = np.random.default_rng(9)
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    chem = cb.ChemistryDef(cfg.name, 0, cfg.bc_len, cfg.bc_len, cfg.umi_len)
    lib = gw.add_library(wl, chem)
    wl2 = gw.add_whitelist(cb.Whitelist.trans(t.trans, t.whitelist))
    fb_lib = gw.add_library(wl2, chem, umi_correction=True, feature_type=1, fb_offset=cfg.fb_offset, fb_length=cfg.fb_len)
    fr = cb.FeatureReference(cfg.n_genes)
    for i in range(cfg.n_fb_features):
        fr.add_feature_barcode(f"FB{i}", bytes(t.fb_seqs[i]).decode(), 1, "5P" + "N" * cfg.fb_offset + "(BC)")
    gw.set_feature_reference(fr)
    i1 = gw.add_fastq(lib, gzip.compress(_fastq_text(g["r1_seq"], g["r1_qual"], rng)), g["feature"])
    # R2 as the sequencer writes it: longer than the capture region (90 cycles), the rest is cDNA
    r2_full = np.concatenate([f["r2_seq"], np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(prob["n_fb"], 65))]], axis=1)
    q2_full = np.concatenate([f["r2_qual"], np.full((prob["n_fb"], 65), ord("F"), dtype=np.uint8)], axis=1)
    i2 = gw.add_fastq(fb_lib, gzip.compress(_fastq_text(f["r1_seq"], f["r1_qual"], rng)),
                      r2_fastq=_fastq_text(r2_full, q2_full, rng, trailing_newline=False))
    assert (i1["n_records"], i2["n_records"]) == (prob["n_gex"], prob["n_fb"]) and i2["r2_len"] == cfg.fb_offset + cfg.fb_len
    assert i2["n_short_r2"] == 0 and i2["n_malformed_r2"] == 0
    assert np.array_equal(gw.read_device(i2["dev_r2_seq"], (prob["n_fb"], i2["r2_len"])), f["r2_seq"][:, :i2["r2_len"]])
    gw.run()
    m, m_ref = gw.count_matrix(), ref.count_matrix()
    assert np.array_equal(m.barcode_rank, m_ref.barcode_rank) and np.array_equal(m.indptr, m_ref.indptr)
    assert np.array_equal(m.indices, m_ref.indices) and np.array_equal(m.data, m_ref.data)
    assert np.array_equal(gw.molecules(), ref.molecules())
    assert int((m.indices >= cfg.n_genes).sum()) > 0  # feature-barcode rows are there
    with pytest.raises(ValueError, match="R2 FASTQ"):
        gw.add_fastq(fb_lib, _fastq_text(f["r1_seq"][:10], f["r1_qual"][:10], rng))
    gw.close()
    ref.close()


def test_fastq_short_and_malformed_records():
    import cellranger_b200 as cb

    recs = [b"@r0\nACGTACGTACGTACGTAAAACCCCGG\n+\nIIIIIIIIIIIIIIIIIIIIIIIIII\n",
            b"@r1\nACGTACGT\n+\nIIIIIIII\n",                       # shorter than the 26 cycles asked for
            b"r2\nTTTTACGTACGTACGTAAAACCCCGG\n-\nIIIIIIIIIIIIIIIIIIIIIIIIII\n",   # bad header and separator
            b"@r3\nGGGGACGTACGTACGTAAAACCCCGGTTTTTTTT\r\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\r\n"]  # longer, CRLF
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(np.frombuffer(b"ACGTACGTACGTACGT", dtype=np.uint8).reshape(1, 16)))
    lib = gw.add_library(wl, cb.ChemistryDef("SC3Pv2", 0, 16, 16, 10))
    gw.set_feature_reference(cb.FeatureReference(10))
    info = gw.add_fastq(lib, b"".join(recs), np.zeros(4, dtype=np.uint32))
    assert (info["n_records"], info["n_short"], info["n_malformed"]) == (4, 1, 2)
    seq = gw.read_device(info["dev_seq"], (4, 26))
    qual = gw.read_device(info["dev_qual"], (4, 26))
    assert bytes(seq[0]) == b"ACGTACGTACGTACGTAAAACCCCGG" and bytes(qual[0]) == b"I" * 26
    assert bytes(seq[1]) == b"ACGTACGT" + b"N" * 18 and bytes(qual[1]) == b"I" * 8 + b"#" * 18
    assert bytes(seq[2]) == b"TTTTACGTACGTACGTAAAACCCCGG"
    assert bytes(seq[3]) == b"GGGGACGTACGTACGTAAAACCCCGG"
    with pytest.raises(cb.CrgpuError):
        gw.add_fastq(lib, b"@r0\nACGT\n+\n", np.zeros(0, dtype=np.uint32))   # three lines: not whole records
    gw.close()


def test_write_mex_matches_matrix(tmp_path):
    """MtxWriter (write_matrix_market.rs:41-120): header lines, 1-based triplets in (barcode, feature) order,
    SEQ-gem_group barcodes."""
    import gzip

    prob = helpers.make_problem("cfg1", 40_000)
    gw = helpers.run_gpu(prob, annotate=False)
    m = gw.count_matrix()
    out = tmp_path / "raw_feature_bc_matrix"
    gw.write_mex(str(out), software_version="Cell Ranger test-1.0", gem_group=1)
    lines = gzip.open(out / "matrix.mtx.gz", "rt").read().split("\n")
    assert lines[0] == "%%MatrixMarket matrix coordinate integer general"
    assert lines[1] == '%metadata_json: {"software_version": "Cell Ranger test-1.0", "format_version": 2}'
    assert lines[2] == f"{m.n_features} {len(m.barcode_rank)} {len(m.data)}"
    assert lines[3:-1] == m.mtx_lines() and lines[-1] == ""
    bcs = gzip.open(out / "barcodes.tsv.gz", "rt").read().split("\n")
    assert bcs[:-1] == m.barcode_strings(1) and bcs[-1] == ""
    feats = gzip.open(out / "features.tsv.gz", "rt").read().split("\n")
    assert len(feats) - 1 == m.n_features and feats[0].split("\t")[2] == "Gene Expression"
    # read back the way the reference's own loader does (lib/python/cellranger/mtx_to_matrix_converter.py:70-92,
    # from_v3_mtx: pandas for the two TSVs, scipy.io.mmread for the matrix): the same CSC comes out
    import pandas as pd
    import scipy.io as sp_io
    import scipy.sparse as sp_sparse

    barcodes = pd.read_csv(str(out / "barcodes.tsv.gz"), delimiter="\t", header=None, usecols=[0], dtype=bytes).values.squeeze()
    features = pd.read_csv(str(out / "features.tsv.gz"), delimiter="\t", header=None)
    mat = sp_sparse.csc_matrix(sp_io.mmread(str(out / "matrix.mtx.gz")))
    mat.sort_indices()
    assert mat.shape == (m.n_features, len(m.barcode_rank)) and len(features) == m.n_features
    assert [b.decode() if isinstance(b, bytes) else b for b in barcodes.tolist()] == m.barcode_strings(1)
    assert np.array_equal(mat.indptr, m.indptr) and np.array_equal(mat.indices, m.indices)
    assert np.array_equal(mat.data, m.data)
    gw.close()


def test_bench_parity_at_scale_check_works_and_detects_a_wrong_count():
    """bench.py's sampled-barcode oracle check (run at the 200 M bench size by `bench.py`): green on a correct run
    of 1.5 M device-generated reads, and red when the GPU's matrix is tampered with."""
    import bench
    import cellranger_b200 as cb
    from cellranger_b200 import synth, synth_device

    n = 1_500_000
    cfg = synth.preset("cfg2", n)
    tables = synth.make_tables(cfg, n)
    gw = cb.GemWell()
    libs = bench.setup_problem(gw, cfg, tables)
    reads = synth_device.generate_device(gw, tables, 0, n, "gex")
    gw.add_reads_device(libs[0], n, cfg.r1_len, reads.r1_seq, reads.r1_qual, reads.feature)
    gw.run()
    res = bench.parity_at_scale(gw, libs[0], cfg, tables, reads, n, n_targets=60)
    assert res["columns_equal"] and res["priors_equal"] and res["molecule_rows_equal"], res
    assert res["barcodes_checked"] >= 50 and res["entries_checked"] > 1000 and res["reads_through_oracle"] < n // 2
    # a checker that cannot fail proves nothing: hand it a matrix with one count changed
    real = gw.count_matrix

    def tampered(pinned=False):
        m = real()
        cells = np.unique(np.asarray(tables.cell_rank))
        col = int(np.searchsorted(m.barcode_rank, cells[0]))
        m.data[m.indptr[col]] += 1
        return m

    gw.count_matrix = tampered
    bad = bench.parity_at_scale(gw, libs[0], cfg, tables, reads, n, n_targets=60, check_molecules=False)
    assert not bad["columns_equal"] and bad["columns_ok"] == bad["barcodes_checked"] - 1
    gw.count_matrix = real
    reads.close()
    gw.close()


def test_segmented_barcode_corrector_matches_oracle_per_segment():
    """GelBeadAndProbe: a 16-base gel-bead segment and an 8-base probe segment, each checked and corrected against
    its own whitelist with its own segment counts (correct_barcode_in_read walks the segments,
    barcode_correction.rs:88-99); the barcode is valid when both segments are (barcode/src/lib.rs:818-823)."""
    import cellranger_b200 as cb
    from oracle import cro

    rng = np.random.default_rng(23)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    wls, counts = [], []
    for L, W in ((16, 4000), (8, 16)):
        wl = acgt[np.unique(rng.integers(0, 4, size=(W, L)), axis=0)]
        wls.append(wl)
        counts.append({bytes(s): int(c) for s, c in zip(wl, rng.integers(1, 400, size=len(wl)))})
    n = 2500
    seqs, quals = [], []
    for wl in wls:
        L = wl.shape[1]
        src = wl[rng.integers(0, len(wl), size=n)].copy()
        for i in range(n):
            if rng.random() < 0.6:  # 40 % of the segments arrive exact
                for _ in range(rng.integers(1, 3)):
                    src[i, rng.integers(0, L)] = acgt[rng.integers(0, 4)]
        seqs.append(src)
        quals.append(rng.integers(33 + 2, 33 + 41, size=(n, L)).astype(np.uint8))
    corr = cb.SegmentedBarcodeCorrector([cb.Whitelist.plain(w) for w in wls], counts)
    outs, states, valid = corr.correct_barcodes(seqs, quals)
    exp_valid = np.ones(n, dtype=bool)
    for k, wl in enumerate(wls):
        members = {bytes(s) for s in wl}
        for i in range(n):
            s = bytes(seqs[k][i])
            if s in members:
                assert states[k][i] == cb.api.VALID_BEFORE_CORRECTION and bytes(outs[k][i]) == s
                continue
            exp = cro.kat_correct_barcode(wl, counts[k], s, bytes(quals[k][i]), F64_MAX, 0.975)
            if exp is None:
                assert states[k][i] == cb.api.INVALID, (k, i)
                exp_valid[i] = False
            else:
                assert states[k][i] == cb.api.VALID_AFTER_CORRECTION and bytes(outs[k][i]) == exp, (k, i)
    assert np.array_equal(valid, exp_valid)
    assert 0 < int(valid.sum()) < n
    corr.close()


def test_barcode_diversity_against_reference_python_vectors(kats):
    """effective_barcode_diversity on read sets whose valid-barcode histogram is a given count vector, against the
    values the reference's Python effective_diversity() gives for that vector (tests/golden, generated)."""
    import cellranger_b200 as cb

    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for case in kats["effective_diversity"]["cases"]:
        counts = case["counts"]
        wl = acgt[np.unique(rng.integers(0, 4, size=(len(counts) + 5, 16)), axis=0)]
        assert len(wl) >= len(counts)
        rows = np.repeat(np.arange(len(counts)), counts)
        n = len(rows)
        r1 = np.concatenate([wl[rows], np.tile(np.frombuffer(b"ACGTTGCAAC", dtype=np.uint8), (n, 1))], axis=1)
        q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
        gw = cb.GemWell()
        lib = gw.add_library(gw.add_whitelist(cb.Whitelist.plain(wl)), cb.ChemistryDef.SC3Pv2())
        gw.set_feature_reference(cb.FeatureReference(4))
        gw.add_reads(lib, r1, q1, np.zeros(n, dtype=np.uint32))
        gw.run()
        d = gw.barcode_diversity(0)
        assert d["barcodes_detected"] == len(counts)
        assert d["effective_barcode_diversity"] == case["expect"], case
        gw.close()
