"""A third, literal restatement of the dedup core, pinned by hand-derived expectations.

`lit_correct_umis`, `lit_determine_low_support_umigenes`, `lit_new` and `lit_process` below are line-for-line
transliterations of lib/rust/tx_annotation/src/mark_dups.rs:19-59, 87-108, 201-277 and 280-363 (the line numbers are
in the comments), written with dicts keyed by (umi, gene) exactly as the reference's HashMaps are. They are test
infrastructure only. The reference ships no test for BarcodeDupMarker::new / ::process; here

  1. the reference's own `test_correct_umis` table (mark_dups.rs:371-392), joined into one barcode, is taken through
     the transliteration and every intermediate (corrections, one-read pre-move counts c1, low-support set, final
     counts c2, min-raw-UMI representatives, per-read DupInfo, UmiCount rows) is compared with values worked out BY HAND
     and written out below;
  2. the C++ oracle (oracle/cr_oracle.cpp) must give the same per-read flags, molecule rows and matrix;
  3. hypothesis drives thousands of tiny random barcodes (3-4-base UMIs, 2-3 genes, 1-2 library types, select keys)
     through transliteration and oracle and demands equality: a differential test with an independent third party;
  4. (-m gpu) the CUDA path gets the same hand case and the same kind of random tiny barcodes, in one launch.
"""
import itertools

import numpy as np
import pytest

from oracle import cro

NUCS = b"ACGT"


# ---------------------------------------------------------------------------
# literal transliteration (test infrastructure)
# ---------------------------------------------------------------------------
def lit_correct_umis(umigene_counts):  # mark_dups.rs:19-59
    corrections = {}
    for (umi, gene), orig_count in umigene_counts.items():  # :24
        test_umi = bytearray(umi)  # :25
        best_dest_count = orig_count  # :27
        best_dest_umi = umi  # :28
        for pos in range(len(umi)):  # :30
            for test_char in NUCS:  # :32
                if test_char == umi[pos]:  # :33
                    continue
                test_umi[pos] = test_char  # :37
                test_count = umigene_counts.get((bytes(test_umi), gene), 0)  # :40
                # :44-46 greater count, or equal count and lexicographically larger UMI
                if test_count > best_dest_count or (test_count == best_dest_count and bytes(test_umi) > best_dest_umi):
                    best_dest_umi = bytes(test_umi)  # :48
                    best_dest_count = test_count  # :49
            test_umi[pos] = umi[pos]  # :53
        if umi != best_dest_umi:  # :55
            corrections[(umi, gene)] = best_dest_umi  # :56
    return corrections


def lit_determine_low_support_umigenes(umigene_counts):  # mark_dups.rs:87-108
    low = set()
    vec = sorted((umi, gene, count) for (umi, gene), count in umigene_counts.items())  # :91-96
    for umi, grouped in itertools.groupby(vec, key=lambda x: x[0]):  # :97
        gene_counts = list(grouped)  # :98
        max_count = max(x[2] for x in gene_counts)  # :99
        max_is_tied = sum(1 for x in gene_counts if x[2] == max_count) >= 2  # :100
        for _umi, gene, count in gene_counts:  # :101
            if max_is_tied or count < max_count:  # :102
                low.add((umi, gene))  # :103
    return low


def lit_new(umigene_counts, umigene_min_key, filter_umis=True, umi_correction=True):  # mark_dups.rs:201-277
    umigene_counts = dict(umigene_counts)
    umigene_min_key = dict(umigene_min_key)
    umi_corrections = lit_correct_umis(umigene_counts) if umi_correction else {}  # :209-212
    umi_correction_counts = [(raw_key, (corrected_umi, raw_key[1]), umigene_counts[raw_key])  # :214-223
                             for raw_key, corrected_umi in umi_corrections.items()]
    for raw_key, corrected_key, _raw_count in umi_correction_counts:  # :227-231 one read first
        umigene_counts[raw_key] -= 1
        umigene_counts[corrected_key] += 1
    c1 = dict(umigene_counts)
    low_support_umigenes = lit_determine_low_support_umigenes(umigene_counts) if filter_umis else set()  # :234-238
    for raw_key, corrected_key, raw_count in umi_correction_counts:  # :241-245 the remaining reads
        umigene_counts[raw_key] -= raw_count - 1
        umigene_counts[corrected_key] += raw_count - 1
    min_raw_umis = {}  # :249-259
    for (raw_seq, gene), corr_seq in umi_corrections.items():
        if raw_seq < corr_seq or (corr_seq, gene) in umi_corrections:  # :252
            prev = min_raw_umis.pop((corr_seq, gene), None)  # :253
            min_raw_umis[(corr_seq, gene)] = raw_seq if prev is None else min(prev, raw_seq)  # :254-257
    min_umi_key_corrections = {}  # :261-265
    for (corr_seq, gene), raw_seq in min_raw_umis.items():
        min_umi_key_corrections[(corr_seq, gene)] = umigene_min_key[(raw_seq, gene)]
    for key, umi_key in min_umi_key_corrections.items():  # :266-268
        umigene_min_key[key] = umi_key
    return dict(counts=umigene_counts, corrections=umi_corrections, low=low_support_umigenes, min_key=umigene_min_key,
                c1=c1, min_raw=min_raw_umis)


def lit_process(marker, umi, gene, select_key):  # mark_dups.rs:280-363 for one read with a valid UMI and a feature
    raw_key = (umi, gene)  # :293
    if raw_key in marker["corrections"]:  # :294-297
        corrected_umi, is_corrected = marker["corrections"][raw_key], True
    else:
        corrected_umi, is_corrected = umi, False
    corrected_key = (corrected_umi, gene)  # :299
    is_low_support_umi = corrected_key in marker["low"]  # :300
    min_key = marker["min_key"][corrected_key]  # :302-305: header == min_key.qname
    is_min_qname = (select_key[1] == min_key[1])
    read_count = marker["counts"][corrected_key]  # :307
    is_umi_count = (not is_low_support_umi) and is_min_qname  # :323-324 (no target filter, sampling factor 1)
    return dict(processed_umi=corrected_umi, is_corrected=is_corrected, is_low_support=is_low_support_umi,
                is_umi_count=is_umi_count, read_count=read_count,
                utype=select_key[0])  # :326-329 the type of the read that carries the count


def lit_barcode(reads, filter_umis=True, umi_correction=True):
    """reads: list of (umi bytes, gene, select_key = (utype_ord, qname_rank)) of ONE barcode and library, every
    UMI valid, every read mapped. Returns per-read DupInfo dicts and the sorted UmiCount rows."""
    counts, min_key = {}, {}
    for umi, gene, sk in reads:  # DupBuilder::observe :128-155
        key = (umi, gene)
        counts[key] = counts.get(key, 0) + 1
        min_key[key] = sk if key not in min_key else min(min_key[key], sk)
    marker = lit_new(counts, min_key, filter_umis, umi_correction)
    infos = [lit_process(marker, umi, gene, sk) for umi, gene, sk in reads]
    rows = sorted((gene, encode_2bit(i["processed_umi"]), i["read_count"], i["utype"])
                  for (umi, gene, sk), i in zip(reads, infos) if i["is_umi_count"])
    return marker, infos, rows


def encode_2bit(umi: bytes) -> int:
    v = 0
    for ch in umi:
        v = (v << 2) | NUCS.index(ch)
    return v


# ---------------------------------------------------------------------------
# the reference's test table, joined into one barcode, with hand-derived expectations
# ---------------------------------------------------------------------------
G0, G1 = 0, 1
TABLE = [(b"AAAA", G0, 3), (b"AAAT", G0, 2), (b"AAAA", G1, 1), (b"AATT", G1, 1),   # mark_dups.rs:375-379
         (b"CCCC", G0, 1), (b"CGCC", G0, 1)]                                        # mark_dups.rs:385-387


def table_reads(prefix=b""):
    """The table as reads. Through the whole pipeline AAAA and CCCC would be dropped as homopolymer UMIs
    (umi/src/info.rs:20-37), so the pipeline tests put a G in front of every UMI: same Hamming distances, same
    lexicographic order, no homopolymer."""
    reads = []
    for umi, gene, count in TABLE:
        for _ in range(count):
            reads.append((prefix + umi, gene, (0, len(reads))))  # Txomic, qname ordered like the read index
    return reads  # reads 0-2 AAAA/g0, 3-4 AAAT/g0, 5 AAAA/g1, 6 AATT/g1, 7 CCCC/g0, 8 CGCC/g0


def test_hand_derived_values_of_the_reference_table():
    marker, infos, rows = lit_barcode(table_reads())
    # correct_umis: AAAT(2) moves to AAAA(3); CCCC ties with CGCC at 1 and moves to the larger sequence; AAAA/g1 and
    # AATT/g1 are two substitutions apart; nothing crosses genes (the reference asserts the first and the third)
    assert marker["corrections"] == {(b"AAAT", G0): b"AAAA", (b"CCCC", G0): b"CGCC"}
    # :226-232 one read of each corrected key moved before the low-support test
    assert marker["c1"] == {(b"AAAA", G0): 4, (b"AAAT", G0): 1, (b"AAAA", G1): 1, (b"AATT", G1): 1,
                            (b"CCCC", G0): 0, (b"CGCC", G0): 2}
    # :87-108 UMI AAAA is seen with g0 (4) and g1 (1): the sub-maximal gene is low support; the zero-count CCCC entry
    # is alone in its group and not low support
    assert marker["low"] == {(b"AAAA", G1)}
    # :241-246 the remaining reads follow
    assert marker["counts"] == {(b"AAAA", G0): 5, (b"AAAT", G0): 0, (b"AAAA", G1): 1, (b"AATT", G1): 1,
                                (b"CCCC", G0): 0, (b"CGCC", G0): 2}
    # :248-268 CCCC < CGCC, so CGCC's representative becomes CCCC's read (read 7); AAAT > AAAA and AAAA is not itself
    # corrected, so AAAA keeps its own (read 0)
    assert marker["min_raw"] == {(b"CGCC", G0): b"CCCC"}
    assert marker["min_key"][(b"CGCC", G0)] == (0, 7) and marker["min_key"][(b"AAAA", G0)] == (0, 0)
    flags = [(i["is_corrected"], i["is_low_support"], i["is_umi_count"]) for i in infos]
    assert flags == [(False, False, True), (False, False, False), (False, False, False),   # AAAA/g0: read 0 counts
                     (True, False, False), (True, False, False),                            # AAAT -> AAAA
                     (False, True, False),                                                  # AAAA/g1 low support
                     (False, False, True),                                                  # AATT/g1
                     (True, False, True),                                                   # CCCC -> CGCC, carries it
                     (False, False, False)]                                                 # CGCC's own read does not
    assert [i["processed_umi"] for i in infos] == [b"AAAA"] * 5 + [b"AAAA", b"AATT", b"CGCC", b"CGCC"]
    # UmiCount rows (feature, umi 2-bit, read_count, utype order): AAAA = 0, CGCC = 0b01100101, AATT = 0b00001111
    assert rows == [(G0, 0, 5, 0), (G0, 0b01100101, 2, 0), (G1, 0b00001111, 1, 0)]
    # the G-prefixed table of the pipeline tests is the same case: GAAAA = 0b10_00000000 ...
    marker_g, infos_g, rows_g = lit_barcode(table_reads(b"G"))
    assert [(i["is_corrected"], i["is_low_support"], i["is_umi_count"]) for i in infos_g] == flags
    assert rows_g == [(G0, 0b1000000000, 5, 0), (G0, 0b1001100101, 2, 0), (G1, 0b1000001111, 1, 0)]


def _arrays(barcodes, reads_per_barcode, umi_len):
    """Read arrays for a list of barcodes (ASCII 16-mers) with their (umi, gene, select key) read lists."""
    n = sum(len(r) for r in reads_per_barcode)
    r1 = np.zeros((n, 16 + umi_len), dtype=np.uint8)
    q1 = np.full((n, 16 + umi_len), ord("I"), dtype=np.uint8)
    feat = np.zeros(n, dtype=np.uint32)
    sel = np.zeros(n, dtype=np.uint64)
    k = 0
    for bc, reads in zip(barcodes, reads_per_barcode):
        for umi, gene, (ut, qn) in reads:
            r1[k, :16] = np.frombuffer(bc, dtype=np.uint8)
            r1[k, 16:] = np.frombuffer(umi, dtype=np.uint8)
            feat[k] = gene
            sel[k] = (ut << 63) | qn
            k += 1
    return r1, q1, feat, sel


def _expected(barcodes, reads_per_barcode, umi_correction=True):
    """Per-read flag bytes (CRGPU_F_*), molecule rows and matrix triplets from the transliteration."""
    order = np.argsort([bytes(b) for b in barcodes], kind="stable")
    col_of = {int(b): c for c, b in enumerate(order)}
    flags, mols, trip = [], [], []
    per_bc_rows = {}
    for b, reads in enumerate(reads_per_barcode):
        _, infos, rows = lit_barcode(reads, umi_correction=umi_correction)
        per_bc_rows[b] = rows
        for i in infos:
            flags.append(1 | 2 | (4 if i["is_corrected"] else 0) | (8 if i["is_low_support"] else 0) |
                         (16 if i["is_umi_count"] else 0))
    for c, b in enumerate(order):
        rows = per_bc_rows[int(b)]
        for gene, umi, rc, ut in rows:
            mols.append((c, 0, gene, umi, rc, 0 if ut else 1))
        for gene in sorted({r[0] for r in rows}):
            trip.append((c, gene, sum(1 for r in rows if r[0] == gene)))
    return np.array(flags, dtype=np.uint8), mols, trip, col_of


def _run_oracle(barcodes, reads_per_barcode, umi_len, n_genes, with_select=True):
    r1, q1, feat, sel = _arrays(barcodes, reads_per_barcode, umi_len)
    o = cro.Oracle()
    wl = o.add_whitelist(np.array([np.frombuffer(b, dtype=np.uint8) for b in barcodes]))
    lib = o.add_library(wl, 0, 16, 16, umi_len)
    o.set_features(np.zeros(n_genes, dtype=np.int32))
    o.add_reads(lib, r1, q1, feat)
    if with_select:
        o.set_select_keys(sel)
    o.run(1)
    return o


def _triplets(matrix):
    ind = matrix["indptr"]
    return [(c, int(f), int(v)) for c in range(len(ind) - 1)
            for f, v in zip(matrix["indices"][ind[c]:ind[c + 1]], matrix["data"][ind[c]:ind[c + 1]])]


def test_oracle_reproduces_the_hand_case():
    bc = [b"AAAACCCCGGGGTTTT"]
    reads = [table_reads(b"G")]
    flags, mols, trip, _ = _expected(bc, reads)
    o = _run_oracle(bc, reads, 5, 2)
    assert np.array_equal(o.reads()["flags"], flags)
    assert [tuple(r) for r in o.molecules().tolist()] == mols == [(0, 0, G0, 0b1000000000, 5, 1), (0, 0, G0, 0b1001100101, 2, 1),
                                                                   (0, 0, G1, 0b1000001111, 1, 1)]
    assert _triplets(o.matrix()) == trip == [(0, G0, 2), (0, G1, 1)]
    o.close()


# ---------------------------------------------------------------------------
# random tiny barcodes: transliteration vs oracle (CPU), and vs the CUDA path (GPU)
# ---------------------------------------------------------------------------
def random_barcodes(rng, n_barcodes, umi_len, n_genes, max_reads=40, alphabet=2):
    """Tiny barcodes with few distinct UMIs (a 2- or 3-letter alphabet) so that corrections, chains, ties, zero
    counts and cross-gene low-support groups occur in almost every one."""
    barcodes, reads_per = [], []
    seen = set()
    qn = itertools.count()
    for _ in range(n_barcodes):
        while True:
            bc = bytes(rng.choice(list(NUCS), size=16).astype(np.uint8))
            if bc not in seen:
                seen.add(bc)
                break
        reads = []
        for _ in range(int(rng.integers(1, max_reads + 1))):
            umi = bytes(rng.choice(list(NUCS[:alphabet]), size=umi_len).astype(np.uint8))
            if len(set(umi)) == 1:  # homopolymer UMIs are invalid (umi/src/info.rs:20-37): keep the case simple
                umi = umi[:-1] + bytes([NUCS[(NUCS.index(umi[-1]) + 1) % 4]])
            reads.append((umi, int(rng.integers(0, n_genes)), (int(rng.integers(0, 2)), None)))
        # unique qname ranks in a random order
        ranks = rng.permutation(len(reads))
        reads = [(u, g, (ut, int(r) + 1_000_000 * len(barcodes))) for (u, g, (ut, _)), r in zip(reads, ranks)]
        barcodes.append(bc)
        reads_per.append(reads)
    del qn
    return barcodes, reads_per


@pytest.mark.parametrize("umi_len,alphabet,n_genes", [(3, 3, 2), (4, 2, 3), (4, 3, 2)])
def test_oracle_matches_transliteration_on_random_tiny_barcodes(umi_len, alphabet, n_genes):
    rng = np.random.default_rng(1000 * umi_len + 10 * alphabet + n_genes)
    barcodes, reads_per = random_barcodes(rng, 1500, umi_len, n_genes, alphabet=alphabet)
    flags, mols, trip, _ = _expected(barcodes, reads_per)
    o = _run_oracle(barcodes, reads_per, umi_len, n_genes)
    assert np.array_equal(o.reads()["flags"], flags)
    assert [tuple(r) for r in o.molecules().tolist()] == mols
    assert _triplets(o.matrix()) == trip
    assert (flags & 4).any() and (flags & 8).any()  # corrections and low-support UMIs did occur
    o.close()


def test_hypothesis_differential_oracle_vs_transliteration():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    umi = st.binary(min_size=3, max_size=3).map(lambda b: bytes(NUCS[x % 3] for x in b)).filter(lambda u: len(set(u)) > 1)
    read = st.tuples(umi, st.integers(0, 2), st.integers(0, 1))
    barcode_reads = st.lists(read, min_size=1, max_size=25)

    @hyp.settings(max_examples=400, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(st.lists(barcode_reads, min_size=1, max_size=6), st.randoms(use_true_random=False))
    def check(per_barcode, pyrandom):
        barcodes, reads_per = [], []
        for b, reads in enumerate(per_barcode):
            bc = bytes(NUCS[(b >> (2 * k)) & 3] for k in range(16))
            ranks = list(range(len(reads)))
            pyrandom.shuffle(ranks)
            barcodes.append(bc)
            reads_per.append([(u, g, (ut, 1000 * b + r)) for (u, g, ut), r in zip(reads, ranks)])
        flags, mols, trip, _ = _expected(barcodes, reads_per)
        o = _run_oracle(barcodes, reads_per, 3, 3)
        try:
            assert np.array_equal(o.reads()["flags"], flags)
            assert [tuple(r) for r in o.molecules().tolist()] == mols
            assert _triplets(o.matrix()) == trip
        finally:
            o.close()

    check()


@pytest.mark.gpu
@pytest.mark.parametrize("umi_correction", [True, False])
def test_gpu_matches_transliteration_on_hand_case_and_random_tiny_barcodes(umi_correction):
    """The CUDA path against the literal transliteration: the hand case plus 4000 random tiny barcodes (5-base UMIs
    over two letters, three genes, random select keys), all in one GEM well."""
    import cellranger_b200 as cb

    rng = np.random.default_rng(77)
    barcodes, reads_per = random_barcodes(rng, 4000, 5, 3, alphabet=2)
    barcodes.append(b"AAAACCCCGGGGTTTT")
    reads_per.append(table_reads(b"G"))
    flags, mols, trip, col_of = _expected(barcodes, reads_per, umi_correction=umi_correction)
    r1, q1, feat, sel = _arrays(barcodes, reads_per, 5)
    gw = cb.GemWell()
    wl = gw.add_whitelist(cb.Whitelist.plain(np.array([np.frombuffer(b, dtype=np.uint8) for b in barcodes])))
    lib = gw.add_library(wl, cb.ChemistryDef("tiny", 0, 16, 16, 5), umi_correction=umi_correction)
    gw.set_feature_reference(cb.FeatureReference(3))
    gw.add_reads(lib, r1, q1, feat, select_key=sel)
    gw.run(annotate_reads=True)
    assert np.array_equal(gw.reads(0)["flags"], flags)
    assert [tuple(r) for r in gw.molecules().tolist()] == mols
    m = gw.count_matrix()
    got = [(c, int(f), int(v)) for c in range(len(m.indptr) - 1)
           for f, v in zip(m.indices[m.indptr[c]:m.indptr[c + 1]], m.data[m.indptr[c]:m.indptr[c + 1]])]
    assert got == trip
    if umi_correction:
        hand_col = col_of[len(barcodes) - 1]
        assert [r[1:] for r in mols if r[0] == hand_col] == [(0, G0, 0b1000000000, 5, 1), (0, G0, 0b1001100101, 2, 1),
                                                            (0, G1, 0b1000001111, 1, 1)]
    gw.close()
