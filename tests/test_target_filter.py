"""Targeted-panel UMI filter: BarcodeDupMarker::process marks a molecule is_filtered_target_umi when its feature is
in the panel's target set, its read count (after UMI correction) is below targeted_umi_min_read_count and it is not
low support; such a molecule is no UMI count (lib/rust/tx_annotation/src/mark_dups.rs:189-191,311-323). The
reference ships no test for it: the expectations of the hand case below are derived by hand from those lines and
must be met by the C++ oracle, by the Python restatement and (under -m gpu) by the library."""
import numpy as np
import pytest

from oracle import cro, pyref
from tests import helpers

WL = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT"]
# (barcode, umi, gene, copies)
ROWS = [
    (0, "ACGTACGTAC", 3, 1),  # X: on target, 1 read < 2                          -> filtered target UMI
    (0, "CAGTCAGTCA", 3, 3),  # Y: on target, 3 reads                             -> counted
    (0, "GGATCCGGAA", 3, 1),  # Z: corrected onto Z' (2 reads): read count 3      -> counted (as Z')
    (0, "GGATCCGGAC", 3, 2),  # Z'
    (0, "TTGACCATGA", 5, 1),  # W: gene 5 is off target                           -> counted
    (0, "CCATGGTTAA", 3, 1),  # V on gene 3 (1 read) and gene 5 (2 reads): (V, 3) is low support - NOT "filtered"
    (0, "CCATGGTTAA", 5, 2),  #                                                    (V, 5) counted
    (1, "ACGTACGTAC", 3, 2),  # another barcode: 2 reads, not below the threshold  -> counted
]
ON_TARGET = np.zeros(16, dtype=np.uint8)
ON_TARGET[3] = 1
MIN_READS = 2


def _reads():
    rows = [(WL[b], u, g) for b, u, g, c in ROWS for _ in range(c)]
    n = len(rows)
    r1 = np.zeros((n, 26), dtype=np.uint8)
    q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
    feat = np.zeros(n, dtype=np.uint32)
    for i, (bc, umi, gene) in enumerate(rows):
        r1[i] = np.frombuffer((bc + umi).encode(), dtype=np.uint8)
        feat[i] = gene
    return rows, r1, q1, feat


def _oracle(filter_on: bool):
    rows, r1, q1, feat = _reads()
    o = cro.Oracle()
    lib = o.add_library(o.add_whitelist(WL), 0, 16, 16, 10)
    o.set_features(np.zeros(16, dtype=np.int32))
    if filter_on:
        o.set_target_filter(ON_TARGET, MIN_READS)
    o.add_reads(lib, r1, q1, feat)
    o.run()
    return o, rows


def test_hand_case_oracle():
    o, rows = _oracle(True)
    m = o.matrix()
    # barcode 0: gene 3 -> Y and Z' (X filtered, (V, 3) low support); gene 5 -> W and V. barcode 1: gene 3 -> 1
    assert m["indptr"].tolist() == [0, 2, 3]
    assert m["indices"].tolist() == [3, 5, 3] and m["data"].tolist() == [2, 2, 1]
    fl = o.reads()["flags"]
    filtered = [(rows[i][1], rows[i][2]) for i in range(len(rows)) if fl[i] & 32]
    assert filtered == [("ACGTACGTAC", 3)]  # the one read of X in barcode 0, nothing else
    low = sorted({(rows[i][1], rows[i][2]) for i in range(len(rows)) if fl[i] & 8})
    assert low == [("CCATGGTTAA", 3)]
    assert not any(fl[i] & 16 for i in range(len(rows)) if fl[i] & 32)  # a filtered read is never the UMI count
    mol = o.molecules()
    assert mol.shape[0] == 5  # Y, Z', W, (V, 5) and the molecule of barcode 1
    o.close()
    # without the filter X counts: gene 3 of barcode 0 has three UMIs
    o2, _ = _oracle(False)
    assert o2.matrix()["data"].tolist() == [3, 2, 1]
    assert not np.any(o2.reads()["flags"] & 32)
    o2.close()


def test_hand_case_python_restatement():
    keys = []
    for b, u, g, c in ROWS:
        keys += [(b, 0, g, cro.kat_encode_2bit(u.encode()))] * c
    d = pyref.dedup_count(keys, None, {0: True}, True, on_target=ON_TARGET, target_min_reads=MIN_READS)
    assert d["entries"] == [(0, 3, 2), (0, 5, 2), (1, 3, 1)]
    d0 = pyref.dedup_count(keys, None, {0: True}, True)
    assert d0["entries"] == [(0, 3, 3), (0, 5, 2), (1, 3, 1)]


def test_seeded_oracle_against_python_restatement():
    """Random tiny barcodes (4-base UMIs over two letters, three genes): the oracle's matrix with the filter on
    equals the restatement's entries."""
    rng = np.random.default_rng(5)
    wl = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT", "TTTTGGGGCCCCAAAA"]
    on_target = np.array([1, 0, 1, 0], dtype=np.uint8)
    for trial in range(30):
        n = int(rng.integers(5, 120))
        bcs = rng.integers(0, 3, size=n)
        umis = ["".join("AC"[x] for x in rng.integers(0, 2, size=4)) + "GTACGT" for _ in range(n)]
        genes = rng.integers(0, 3, size=n).astype(np.uint32)
        r1 = np.zeros((n, 26), dtype=np.uint8)
        for i in range(n):
            r1[i] = np.frombuffer((wl[bcs[i]] + umis[i]).encode(), dtype=np.uint8)
        q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
        thr = int(rng.integers(2, 5))
        o = cro.Oracle()
        lib = o.add_library(o.add_whitelist(wl), 0, 16, 16, 10)
        o.set_features(np.zeros(4, dtype=np.int32))
        o.set_target_filter(on_target, thr)
        o.add_reads(lib, r1, q1, genes)
        o.run()
        m = o.matrix()
        order = {s: i for i, s in enumerate(sorted(wl))}
        keys = [(order[wl[bcs[i]]], 0, int(genes[i]), cro.kat_encode_2bit(umis[i].encode())) for i in range(n)]
        d = pyref.dedup_count(keys, None, {0: True}, True, on_target=on_target, target_min_reads=thr)
        got = []
        present = sorted({order[wl[b]] for b in bcs})
        for col, r in enumerate(present):
            for e in range(m["indptr"][col], m["indptr"][col + 1]):
                got.append((r, int(m["indices"][e]), int(m["data"][e])))
        assert got == d["entries"], trial
        o.close()


@pytest.mark.gpu
def test_hand_case_gpu():
    import cellranger_b200 as cb

    o, rows = _oracle(True)
    _, r1, q1, feat = _reads()
    gw = cb.GemWell()
    lib = gw.add_library(gw.add_whitelist(cb.Whitelist.plain(WL)), cb.ChemistryDef.SC3Pv2())
    gw.set_feature_reference(cb.FeatureReference(16))
    gw.set_target_filter(ON_TARGET, MIN_READS)
    gw.add_reads(lib, r1, q1, feat)
    gw.run(annotate_reads=True)
    mo, mg = o.matrix(), gw.count_matrix()
    assert mg.indptr.tolist() == [0, 2, 3] and mg.indices.tolist() == [3, 5, 3] and mg.data.tolist() == [2, 2, 1]
    assert np.array_equal(mo["data"], mg.data)
    assert np.array_equal(o.reads()["flags"], gw.reads(0)["flags"])
    assert np.array_equal(o.molecules(), gw.molecules())
    assert gw.stats()["filtered_target_umis"] == 1
    # switching the filter off again restores the unfiltered counts
    gw.set_target_filter(ON_TARGET, None)
    gw.run(annotate_reads=True)
    assert gw.count_matrix().data.tolist() == [3, 2, 1]
    assert gw.stats()["filtered_target_umis"] == 0
    gw.close()
    o.close()


@pytest.mark.gpu
def test_random_problem_with_target_filter_gpu():
    """cfg1 at 150 k reads with half of the genes on target and a threshold of 3 reads: every output equal."""
    import cellranger_b200 as cb

    prob = helpers.make_problem("cfg1", 150_000, n_whitelist=20_000, n_cells=60)
    cfg, t = prob["cfg"], prob["tables"]
    on_target = (np.arange(cfg.n_genes) % 2 == 0).astype(np.uint8)
    o = helpers.run_oracle(prob, stages=False)
    o.set_target_filter(on_target, 3)
    o.run(4)
    gw = helpers.run_gpu(prob, run=False)
    gw.set_target_filter(on_target, 3)
    gw.run(annotate_reads=True)
    info = helpers.compare_all(o, gw, prob)
    n_filtered = info["stats"]["filtered_target_umis"]
    assert n_filtered > 0
    # every filtered molecule has exactly one representative-less read group: count molecules via the flags
    fl = o.reads()["flags"]
    assert int(np.count_nonzero(fl & 32)) >= n_filtered
    gw.close()
    o.close()
