"""The C++ oracle against the structurally different Python restatement (oracle/pyref.py) on seeded
random inputs: this is what pins the parts of the path for which the reference ships no test
(one-read pre-move, low-support filter, representative read, feature counts, barcode index, CSC)."""
import numpy as np
import pytest

from cellranger_b200 import synth
from oracle import cro, pyref
from tests import helpers


def _pyref_from_problem(prob):
    cfg, t = prob["cfg"], prob["tables"]
    wls = [(t.whitelist, None)]
    libs = [dict(wl=0, bc_off=0, bc_len=cfg.bc_len, umi_off=cfg.bc_len, umi_len=cfg.umi_len, umi_correction=True,
                 is_fb=False, ftype=0, fb_offset=0, fb_len=0)]
    batches = [dict(lib=0, r1_seq=prob["gex"]["r1_seq"], r1_qual=prob["gex"]["r1_qual"], feature=prob["gex"]["feature"])]
    ftype, fb_seqs = helpers.feature_tables(prob)
    if prob["n_fb"]:
        wls.append((t.trans, t.whitelist))
        libs.append(dict(wl=1, bc_off=0, bc_len=cfg.bc_len, umi_off=cfg.bc_len, umi_len=cfg.umi_len,
                         umi_correction=True, is_fb=True, ftype=1, fb_offset=cfg.fb_offset, fb_len=cfg.fb_len))
        f = prob["fb"]
        batches.append(dict(lib=1, r1_seq=f["r1_seq"], r1_qual=f["r1_qual"], r2_seq=f["r2_seq"], r2_qual=f["r2_qual"]))
    return pyref.run_pipeline(wls, libs, ftype.tolist(), fb_seqs, batches)


def _compare(o, p, prob):
    cfg = prob["cfg"]
    ro = o.reads()
    n = len(p["reads"])
    assert ro["state"].shape[0] == n
    content_ascii = synth.unpack_2bit(np.array(p["content"], dtype=np.uint64), cfg.bc_len)
    for gi, rd in enumerate(p["reads"]):
        assert ro["state"][gi] == rd["state"], gi
        if rd["rank"] is not None:
            assert bytes(ro["bc"][gi]) == bytes(content_ascii[rd["rank"]]), gi
        assert ro["feature"][gi] == rd["feature"], gi
        assert ro["flags"][gi] == rd["flags"], (gi, ro["flags"][gi], rd["flags"])
        if rd["proc_umi"] is not None:
            assert cro.kat_encode_2bit(bytes(ro["umi"][gi])) == rd["proc_umi"], gi
    m = o.matrix()
    assert [bytes(b) for b in m["barcodes"]] == [bytes(content_ascii[r]) for r in p["barcode_ranks"]]
    assert m["indptr"].tolist() == p["indptr"]
    assert m["indices"].tolist() == p["indices"]
    assert m["data"].tolist() == p["data"]
    mol = o.molecules()
    assert sorted(map(tuple, mol[:, :5].tolist())) == sorted(p["molecules"])
    assert np.all(mol[:, 5] == 1)  # every read Txomic without select keys
    wl = prob["tables"].whitelist
    for lib in range(2 if prob["n_fb"] else 1):
        assert o.counts(lib, 0, wl).tolist() == p["prior"][lib].tolist()
        assert o.counts(lib, 1, wl).tolist() == p["corrected"][lib].tolist()
    if prob["n_fb"]:
        assert o.fb_counts().tolist() == p["fb_counts"]


@pytest.mark.parametrize("name,n,kw", [
    ("cfg1", 6000, dict(n_whitelist=3000, n_cells=6)),
    ("cfg2", 6000, dict(n_whitelist=3000, n_cells=6)),
    ("cfg5", 8000, dict(n_whitelist=2000, n_cells=3, top_genes=5)),   # saturated UMIs: chains, ties, low support
    ("cfg4", 6000, dict(n_whitelist=3000, n_cells=6)),                # GEX + feature barcodes, translation whitelist
])
def test_oracle_matches_python_restatement(name, n, kw):
    prob = helpers.make_problem(name, n, **kw)
    o = helpers.run_oracle(prob, threads=2)
    p = _pyref_from_problem(prob)
    _compare(o, p, prob)
    st = o.stats()
    assert st["umis"] > 0
    if name == "cfg5":
        assert st["umi_corrected_reads"] > 0


def test_umi_collision_stress_exercises_low_support():
    """4-base UMIs over 3 genes in 2 cells: every rule of A.4 fires many times."""
    rng = np.random.default_rng(11)
    wl = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT", "TTTTGGGGCCCCAAAA"]
    n = 4000
    bcs = rng.integers(0, 2, size=n)
    umis = rng.integers(0, 4, size=(n, 10))
    umis[:, :6] = 0  # only the last 4 bases vary -> 256 UMIs
    genes = rng.integers(0, 3, size=n)
    r1 = np.zeros((n, 26), dtype=np.uint8)
    for i in range(n):
        r1[i, :16] = np.frombuffer(wl[bcs[i]].encode(), dtype=np.uint8)
    r1[:, 16:] = np.frombuffer(b"ACGT", dtype=np.uint8)[umis]
    q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
    feat = genes.astype(np.uint32)
    o = cro.Oracle()
    w = o.add_whitelist(wl)
    lib = o.add_library(w, 0, 16, 16, 10)
    o.set_features(np.zeros(3, dtype=np.int32))
    o.add_reads(lib, r1, q1, feat)
    o.run(2)
    p = pyref.run_pipeline([(cro.ascii_mat(wl), None)],
                           [dict(wl=0, bc_off=0, bc_len=16, umi_off=16, umi_len=10, umi_correction=True, is_fb=False,
                                 ftype=0, fb_offset=0, fb_len=0)], [0, 0, 0], np.zeros((3, 1), dtype=np.uint8),
                           [dict(lib=0, r1_seq=r1, r1_qual=q1, feature=feat)])
    ro = o.reads()
    assert ro["flags"].tolist() == [rd["flags"] for rd in p["reads"]]
    m = o.matrix()
    assert m["data"].tolist() == p["data"] and m["indices"].tolist() == p["indices"]
    st = o.stats()
    assert st["low_support_reads"] > 100 and st["umi_corrected_reads"] > 100


def test_barcode_summary_hand_case():
    """BarcodeSummary::observe (aligner.rs:53-67) on six reads written out by hand."""
    from oracle import pyref

    bc = np.array([list(b"AAAA"), list(b"AAAA"), list(b"CCCC"), list(b"AAAA"), list(b"GGGG"), list(b"CCCC")], dtype=np.uint8)
    state = np.array([1, 2, 1, 1, 3, 1], dtype=np.uint8)      # the GGGG read stays invalid: no row
    #        has_dup | corrected | low support | umi_count
    flags = np.array([2 | 16, 2 | 4, 0, 2 | 8, 2 | 16, 2 | 4 | 8], dtype=np.uint8)
    b, reads, umis, cand, corr = pyref.barcode_summary(bc, state, flags)
    assert [bytes(x) for x in b] == [b"AAAA", b"CCCC"]
    assert reads.tolist() == [3, 2]
    assert umis.tolist() == [1, 0]
    assert cand.tolist() == [2, 0]      # AAAA: reads 0 and 1 (read 3 is low support); CCCC: read 5 is low support
    assert corr.tolist() == [1, 1]
