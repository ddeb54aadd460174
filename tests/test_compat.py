"""the other users of the resident whitelist - whitelist match rate and CHECK_BARCODES_COMPATIBILITY."""
import numpy as np
import pytest

from oracle import compat_ref


def test_nx_known_answers_from_the_reference():
    # lib/rust/stats/src/nx.rs:113-131 (tests) and :60-101 (doc examples)
    assert compat_ref.nx([100, 70, 60, 50, 50, 40, 30], 0.5) == 60
    assert compat_ref.nx([70, 60, 50, 40, 30, 100, 50], 0.5) == 60
    assert compat_ref.nx([68, 90, 11, 50, 15, 57, 27, 67, 24, 45], 0.5) == 57
    assert compat_ref.nx([], 0.5) is None
    assert compat_ref.nx([2, 3, 4, 5, 6, 7, 8, 9, 10], 0.5) == 8
    assert compat_ref.nx([2, 3, 4, 5, 6, 7, 8, 9, 10], 0.9) == 4
    with pytest.raises(AssertionError):
        compat_ref.nx([68, 90, 0, 50], 0.5)


def test_match_to_whitelist_known_answers(kats):
    # lib/rust/barcode/src/whitelist.rs:554-567
    k = kats["match_to_whitelist"]
    wl = {s.encode() for s in k["whitelist"]}
    for case in k["cases"]:
        got = compat_ref.match_to_whitelist(wl, case["seq"].encode())
        assert (got.decode() if got else None) == case["expect"], case


def test_robust_cosine_similarity_hand_case():
    a = {b"AA": 10, b"AC": 10, b"AG": 1000}   # N92.5 caps the outlier: descending 1000 reaches 0.925 * 1020 alone
    b = {b"AA": 5, b"AC": 5}
    assert compat_ref.nx(a.values(), 0.925) == 1000 and compat_ref.nx(b.values(), 0.925) == 5
    # no capping applies here: dot = 10*5 + 10*5, |a| = sqrt(100 + 100 + 10^6), |b| = sqrt(50)
    assert compat_ref.robust_cosine_similarity(a, b) == 100.0 / (np.sqrt(1000200.0) * np.sqrt(50.0))
    assert compat_ref.robust_cosine_similarity({}, b) == 0.0 and compat_ref.robust_cosine_similarity(a, {}) == 0.0
    assert compat_ref.robust_cosine_similarity(a, a) == pytest.approx(1.0)


def _upload_hist(gw, hist, counts):
    import ctypes as C

    from cellranger_b200._lib import check, ptr

    c = np.ascontiguousarray(counts, dtype=np.uint32)
    assert c.shape[0] == hist.n
    check(gw.L.crgpu_memcpy_h2d(gw.ctx, C.c_void_p(hist.dev), ptr(c), C.c_uint64(c.nbytes)))


@pytest.mark.gpu
def test_gpu_nx_and_similarity_known_answers():
    import cellranger_b200 as cb

    rng = np.random.default_rng(3)
    wl = np.unique(rng.integers(0, 4, size=(64, 8)), axis=0)[:40]
    wl_ascii = np.frombuffer(b"ACGT", dtype=np.uint8)[wl]
    gw = cb.GemWell()
    w = gw.add_whitelist(cb.Whitelist.plain(wl_ascii))
    n = wl_ascii.shape[0]
    for items, frac, exp in (([100, 70, 60, 50, 50, 40, 30], 0.5, 60), ([68, 90, 11, 50, 15, 57, 27, 67, 24, 45], 0.5, 57),
                             ([2, 3, 4, 5, 6, 7, 8, 9, 10], 0.5, 8), ([2, 3, 4, 5, 6, 7, 8, 9, 10], 0.9, 4), ([], 0.5, 0)):
        h = cb.WhitelistHistogram(gw, w)
        counts = np.zeros(n, dtype=np.uint32)
        counts[rng.permutation(n)[:len(items)]] = items
        _upload_hist(gw, h, counts)
        assert h.nx(frac) == exp, items
        h.close()
    # random histograms: similarity bit-identical to the restatement
    keys = [bytes(s) for s in wl_ascii[np.lexsort(wl_ascii.T[::-1])]]
    for trial in range(20):
        a = (rng.random(n) < 0.7) * rng.integers(1, 5000 if trial % 2 else 20, size=n)
        b = (rng.random(n) < 0.6) * rng.integers(1, 300, size=n)
        if trial == 7:
            b[:] = 0
        ha, hb = cb.WhitelistHistogram(gw, w), cb.WhitelistHistogram(gw, w)
        _upload_hist(gw, ha, a)
        _upload_hist(gw, hb, b)
        da = {k: int(v) for k, v in zip(keys, a) if v}
        db = {k: int(v) for k, v in zip(keys, b) if v}
        assert ha.robust_cosine_similarity(hb) == compat_ref.robust_cosine_similarity(da, db), trial
        ha.close()
        hb.close()
    gw.close()


@pytest.mark.gpu
def test_gpu_sample_valid_barcodes_and_translation_decision():
    """GEX reads against Antibody reads of the same cells whose gel-bead barcodes are the TRANSLATED partners (the
    3' v3 feature-barcode oligos): similarity without translation is near 0, with translation near 1 -> the library
    is translated; an untranslated library of the same cells is not; a library of other cells fails the check.
    Histograms, match counts and both similarities are compared with the restatement, bit for bit."""
    import cellranger_b200 as cb
    from cellranger_b200 import synth

    n = 150_000
    cfg = synth.preset("cfg4", n)
    # 2 % of the reads get an N (the rescue path matters); little ambient RNA, so that the N92.5 cap sits at cell level
    cfg.n_whitelist, cfg.n_cells, cfg.n_frac, cfg.ambient_frac = 60_000, 80, 0.02, 0.02
    t = synth.make_tables(cfg, n)
    gex = synth.generate_reads(t, 0, n, "gex")["r1_seq"]
    fb = synth.generate_reads(t, 0, n // 2, "fb")["r1_seq"]          # barcodes from t.trans (the raw partners)
    plain_set = {bytes(s) for s in t.whitelist}
    translate = {bytes(r): bytes(c) for r, c in zip(t.trans, t.whitelist)}   # raw -> translated (content)
    assert set(translate) == plain_set  # the partners are a permutation of the same whitelist
    gw = cb.GemWell()
    w_plain = gw.add_whitelist(cb.Whitelist.plain(t.whitelist))
    w_trans = gw.add_whitelist(cb.Whitelist.trans(t.trans, t.whitelist))
    # histograms and match counts
    L = cfg.bc_len
    h_gex = cb.WhitelistHistogram(gw, w_plain)
    m_gex = h_gex.observe(gex)
    ref_gex, n_ref, m_ref = compat_ref.sample_valid_barcodes(plain_set, (bytes(r[:L]) for r in gex))
    assert (m_gex, h_gex.reads) == (m_ref, n_ref) and 0.5 < h_gex.fraction() < 1.0
    keys = [bytes(s) for s in t.whitelist]  # sorted: entry index = position
    got = h_gex.counts()
    assert {k: int(v) for k, v in zip(keys, got) if v} == ref_gex
    assert any(b"N" in bytes(r[:L]) and compat_ref.match_to_whitelist(plain_set, bytes(r[:L])) for r in gex[:20000])
    h_fb = cb.WhitelistHistogram(gw, w_plain)
    h_fb.observe(fb)
    ref_fb, _, _ = compat_ref.sample_valid_barcodes(plain_set, (bytes(r[:L]) for r in fb))
    s_plain = h_gex.robust_cosine_similarity(h_fb)
    s_trans = h_gex.robust_cosine_similarity(h_fb, w_trans)
    assert s_plain == compat_ref.robust_cosine_similarity(ref_gex, ref_fb)
    assert s_trans == compat_ref.robust_cosine_similarity(ref_gex, compat_ref.map_key(ref_fb, translate))
    assert s_plain < 0.01 and s_trans > 0.99
    h_fb.close()
    h_gex.close()
    # the stage's decisions
    same_cells_untranslated = synth.generate_reads(t, n, n // 2, "gex")["r1_seq"]
    cfg2 = synth.preset("cfg4", n)
    cfg2.n_whitelist, cfg2.n_cells, cfg2.n_frac, cfg2.ambient_frac = 60_000, 80, 0.02, 0.02
    cfg2.seed = cfg.seed + 99                                                     # other cells, same whitelist
    t2 = synth.make_tables(cfg2, n)
    res = cb.check_barcodes_compatibility(gw, w_plain, w_trans, gex,
                                          {"Antibody Capture": fb, "CRISPR Guide Capture": same_cells_untranslated})
    exp_tr, exp_sims = compat_ref.libraries_to_translate(
        ref_gex, {"Antibody Capture": ref_fb,
                  "CRISPR Guide Capture": compat_ref.sample_valid_barcodes(plain_set, (bytes(r[:L]) for r in same_cells_untranslated))[0]},
        translate)
    assert set(res["libraries_to_translate"]) == exp_tr == {"Antibody Capture"}
    for name, (s0, s1) in exp_sims.items():
        assert res["similarity"][name]["without_translation"] == s0 and res["similarity"][name]["with_translation"] == s1
    other = synth.generate_reads(t2, 0, n // 2, "gex")["r1_seq"]
    assert not np.array_equal(np.sort(t2.cell_rank), np.sort(t.cell_rank))
    with pytest.raises(ValueError, match="insufficient overlap"):
        cb.check_barcodes_compatibility(gw, w_plain, w_trans, gex, {"Multiplexing Capture": other})
    gw.close()
