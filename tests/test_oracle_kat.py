"""The CPU oracle against every self-contained known-answer test the reference
holds for this path (tests/golden/reference_kats.json; each block cites its
reference test)."""
import math

import numpy as np

from oracle import cro

F64_MAX = 1.7976931348623157e308


def test_probability_is_libm_pow():
    # lib/rust/barcode/src/corrector.rs:167-171
    for q in range(0, 256):
        assert cro.probability(q) == math.pow(10.0, -(float(q) - 33.0) / 10.0)


def test_barcode_correction(kats):
    for name in ("barcode_correction", "barcode_correction_no_counts"):
        k = kats[name]
        for case in k["cases"]:
            got = cro.kat_correct_barcode(k["whitelist"], k["counts"], case["seq"], bytes(case["qual"]),
                                          k["max_expected_errors"], k["threshold"])
            exp = None if case["expect"] is None else case["expect"].encode()
            assert got == exp, (name, case["name"])


def test_barcode_n_rescue(kats):
    k = kats["barcode_n_rescue"]
    bc = k["whitelist"][0]
    for n_pos in range(len(bc)):
        seq = bc[:n_pos] + "N" + bc[n_pos + 1:]
        qual = [k["qual_else"]] * len(bc)
        qual[n_pos] = k["qual_at_n"]
        got = cro.kat_correct_barcode(k["whitelist"], k["counts"], seq, bytes(qual), k["max_expected_errors"],
                                      k["threshold"])
        assert got == k["expect"].encode(), n_pos


def test_two_n_is_uncorrectable(kats):
    k = kats["barcode_n_rescue"]
    bc = k["whitelist"][0]
    seq = "NN" + bc[2:]
    assert cro.kat_correct_barcode(k["whitelist"], {}, seq, bytes([53] * 16), F64_MAX, 0.975) is None


def test_match_to_whitelist(kats):
    k = kats["match_to_whitelist"]
    for case in k["cases"]:
        got = cro.kat_match_to_whitelist(k["whitelist"], case["seq"])
        assert got == (None if case["expect"] is None else case["expect"].encode())


def test_correct_umis(kats):
    for case in kats["correct_umis"]["cases"]:
        table = [(u, g, c) for u, g, c in case["table"]]
        got = cro.kat_correct_umis(table)
        exp = {(u.encode(), g): d.encode() for u, g, d in case["expect"]}
        assert got == exp


def test_umi_extraction(kats):
    # UMI = R1[offset : offset+length], or min_length when R1 is too short
    # (lib/rust/cr_types/src/rna_read.rs:103-138); the oracle takes fixed (offset, length)
    for case in kats["umi_extraction"]["cases"]:
        r1 = case["r1"]
        length = case["length"] if len(r1) >= case["offset"] + case["length"] else case["min_length"]
        assert r1[case["offset"]:case["offset"] + length] == case["expect"]


def test_encode_2bit(kats):
    for case in kats["encode_2bit"]["cases"]:
        assert cro.kat_encode_2bit(case["seq"].encode()) == case["expect"]


def test_feature_dist(kats):
    k = kats["feature_dist"]
    got = cro.kat_feature_dist(k["raw"], k["ftype"])
    assert got.tolist() == [float(np.float64(x)) for x in (0.5, 0.5, 0.0, 9.0 / 10.0, 1.0 / 10.0, 0.0)]
    # all-zero counts → uniform over every feature (feature_checker.rs:36-47)
    assert cro.kat_feature_dist([0, 0, 0, 0], [0, 0, 1, 1]).tolist() == [0.25] * 4


def _feature_cases(k):
    feats = k["features"]
    ftype = [t for _, t in feats]
    dist = cro.kat_feature_dist(k["raw_counts"], ftype)
    for case in k["cases"]:
        sel = [i for i, (_, t) in enumerate(feats) if t == case["ftype"]]
        got = cro.kat_feature_match([feats[i][0] for i in sel], sel, dist, case["seq"].encode(),
                                    case["qual"].encode())
        exp = -1 if case["expect"] is None else [s for s, _ in feats].index(case["expect"])
        yield case, got, exp


def test_correct_feature(kats):
    for name in ("correct_feature", "correct_bare_feature_fixed"):
        for case, got, exp in _feature_cases(kats[name]):
            assert got == exp, (name, case)


def test_feature_exact_only_without_dist(kats):
    # MAKE_SHARD's extractor has feat_dist=None: only exact captures match
    # (lib/rust/cr_lib/src/make_shard_metrics.rs:210-213)
    k = kats["correct_bare_feature_fixed"]
    seqs = [s for s, _ in k["features"]]
    assert cro.kat_feature_match(seqs, [0, 1, 2], None, b"TTTA", b"IIII") == -1
    assert cro.kat_feature_match(seqs, [0, 1, 2], None, b"TTTT", b"IIII") == 2


def test_umi_validity():
    # lib/rust/umi/src/info.rs:20-74
    good_q = b"I" * 10
    assert cro.kat_umi_is_valid(b"ACGTACGTAC", good_q)
    assert not cro.kat_umi_is_valid(b"ACGTNCGTAC", good_q)  # has N
    assert not cro.kat_umi_is_valid(b"AAAAAAAAAA", good_q)  # homopolymer
    assert not cro.kat_umi_is_valid(b"ACGTACGTAC", b"IIII*IIIII")  # '*' = Q9 < 10
    assert cro.kat_umi_is_valid(b"ACGTACGTAC", b"IIII+IIIII")  # '+' = Q10


def test_low_support_rule():
    # lib/rust/tx_annotation/src/mark_dups.rs:87-108 (no reference test: pinned by code reading)
    low = cro.kat_low_support([("AAAA", 0, 5), ("AAAA", 1, 2), ("CCCC", 0, 3), ("CCCC", 1, 3), ("GGGG", 2, 1),
                               ("TTTT", 0, 0), ("TTTT", 1, 0)])
    assert low.tolist() == [False, True, True, True, False, True, True]


def test_effective_diversity_formula_against_the_reference_python(kats):
    """SimpleHistogram::effective_diversity (metric/src/histogram.rs:161-171) restated as the loop the GPU test uses,
    against values GENERATED by the reference's own Python twin (cellranger/stats.py:17-21)."""
    for case in kats["effective_diversity"]["cases"]:
        s, s2 = 0.0, 0.0
        for x in case["counts"]:
            s += float(x)
            s2 += float(x) ** 2
        assert s ** 2 / s2 == case["expect"], case
