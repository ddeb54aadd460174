"""total_barcode_counts of BARCODE_CORRECTION (lib/rust/cr_lib/src/stages/barcode_correction.rs:327-362): the reads
the stage reads back (not valid before correction) counted under their barcode after correction, entries below
min_reads_to_report_bc dropped. No reference test exists: a hand case with the expectation written out, the
oracle against an independent recount from its own per-read outputs, and (under -m gpu) the library against both."""
from collections import Counter

import numpy as np
import pytest

from oracle import cro
from tests import helpers

WL = ["AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT", "TTTTGGGGCCCCAAAA"]


def _hand_reads():
    # (barcode as sequenced, copies): exact hits never reach the stage; one substitution is corrected (high quality
    # elsewhere, the prior of the target > 0); two substitutions stay invalid; an N in an otherwise exact barcode
    # is rescued; two Ns stay invalid
    rows = [("AAAACCCCGGGGTTTT", 5), ("AAAACCCCGGGGTTTA", 3), ("AAAACCCCGGGGTTAA", 2), ("NAAACCCCGGGGTTTT", 1),
            ("NNAACCCCGGGGTTTT", 2), ("ACGTACGTACGTACGT", 4), ("ACGTACGTACGTACGA", 1), ("GGGGGGGGGGGGGGGG", 3)]
    seqs = [bc for bc, c in rows for _ in range(c)]
    n = len(seqs)
    r1 = np.zeros((n, 26), dtype=np.uint8)
    for i, bc in enumerate(seqs):
        r1[i] = np.frombuffer((bc + "ACGTTGCAAC").encode(), dtype=np.uint8)
    q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
    feat = np.zeros(n, dtype=np.uint32)
    return r1, q1, feat


HAND_EXPECT_MIN1 = [  # Barcode order: invalid sequences first, each group ascending (A < C < G < N < T)
    ("AAAACCCCGGGGTTAA", 0, 2), ("GGGGGGGGGGGGGGGG", 0, 3), ("NNAACCCCGGGGTTTT", 0, 2),
    ("AAAACCCCGGGGTTTT", 1, 4),  # 3 corrected substitutions + the rescued N
    ("ACGTACGTACGTACGT", 1, 1),
]


def _as_list(t):
    seqs, valid, counts = t
    return [(bytes(s).decode(), int(v), int(c)) for s, v, c in zip(seqs, valid, counts)]


def test_hand_case_oracle():
    r1, q1, feat = _hand_reads()
    o = cro.Oracle()
    lib = o.add_library(o.add_whitelist(WL), 0, 16, 16, 10)
    o.set_features(np.zeros(4, dtype=np.int32))
    o.add_reads(lib, r1, q1, feat)
    o.run()
    assert _as_list(o.total_barcode_counts(1)) == HAND_EXPECT_MIN1
    assert _as_list(o.total_barcode_counts(3)) == [e for e in HAND_EXPECT_MIN1 if e[2] >= 3]
    o.close()


def _recount(o, min_reads):
    rd = o.reads()
    hist = Counter()
    for s, st in zip(rd["bc"], rd["state"]):
        if st != 1:
            hist[(st == 2, bytes(s).decode())] += 1
    return [(k[1], int(k[0]), c) for k, c in sorted(hist.items()) if c >= min_reads]


def test_oracle_against_recount_of_its_reads():
    prob = helpers.make_problem("cfg1", 30_000, n_whitelist=3000, n_cells=10)
    o = helpers.run_oracle(prob)
    for m in (1, 2, 5):
        assert _as_list(o.total_barcode_counts(m)) == _recount(o, m)
    o.close()


@pytest.mark.gpu
def test_gpu_hand_case_and_random_problem():
    import cellranger_b200 as cb

    r1, q1, feat = _hand_reads()
    gw = cb.GemWell()
    lib = gw.add_library(gw.add_whitelist(cb.Whitelist.plain(WL)), cb.ChemistryDef.SC3Pv2())
    gw.set_feature_reference(cb.FeatureReference(4))
    gw.add_reads(lib, r1, q1, feat)
    gw.run()
    assert _as_list(gw.total_barcode_counts(1)) == HAND_EXPECT_MIN1
    assert _as_list(gw.total_barcode_counts(3)) == [e for e in HAND_EXPECT_MIN1 if e[2] >= 3]
    gw.close()
    for name, n, kw in (("cfg1", 200_000, dict(n_whitelist=20_000, n_cells=50)),
                        ("cfg4", 100_000, dict(n_whitelist=20_000, n_cells=50))):
        prob = helpers.make_problem(name, n, **kw)
        o = helpers.run_oracle(prob)
        gw = helpers.run_gpu(prob, annotate=False)
        for m in (1, 2, 4):
            assert _as_list(gw.total_barcode_counts(m)) == _as_list(o.total_barcode_counts(m)), (name, m)
        gw.close()
        o.close()


def test_oracle_recount_with_many_non_acgt_bases():
    """Barcodes with one or several N (and exact hits, one- and two-base substitutions) over a tiny whitelist: the
    histogram equals the recount, its order is (valid, sequence) with N between G and T, and the thresholds nest."""
    rng = np.random.default_rng(17)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    wl = acgt[np.unique(rng.integers(0, 4, size=(40, 16)), axis=0)]
    n = 6000
    src = wl[rng.integers(0, len(wl), size=n)].copy()
    for i in range(n):
        r = rng.random()
        if r < 0.5:
            for _ in range(rng.integers(1, 3)):
                src[i, rng.integers(0, 16)] = acgt[rng.integers(0, 4)]
        if r > 0.7:
            for _ in range(rng.integers(1, 3)):
                src[i, rng.integers(0, 16)] = ord("N")
    r1 = np.concatenate([src, np.tile(np.frombuffer(b"ACGTTGCAAC", dtype=np.uint8), (n, 1))], axis=1)
    q1 = np.full((n, 26), ord("I"), dtype=np.uint8)
    o = cro.Oracle()
    lib = o.add_library(o.add_whitelist(wl), 0, 16, 16, 10)
    o.set_features(np.zeros(4, dtype=np.int32))
    o.add_reads(lib, r1, q1, np.zeros(n, dtype=np.uint32))
    o.run()
    prev = None
    for m in (1, 2, 3, 10):
        got = _as_list(o.total_barcode_counts(m))
        assert got == _recount(o, m)
        keys = [(v, s) for s, v, _ in got]
        assert keys == sorted(keys)
        if prev is not None:
            assert set(got) <= set(prev)
        prev = got
    assert any("N" in s for s, v, _ in _as_list(o.total_barcode_counts(1)) if v == 0)
    o.close()
