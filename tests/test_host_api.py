"""Host-side pieces of cellranger_b200.api that need no GPU."""
import gzip

import numpy as np
import pytest

from cellranger_b200 import api


def test_tethered_offset_follows_compile_pattern_rules():
    # FeatureExtractor::compile_pattern — lib/rust/cr_types/src/reference/feature_extraction.rs:306-342
    assert api.tethered_offset("5PNNNNNNNNNN(BC)") == 10
    assert api.tethered_offset("5P(BC)") == 0
    assert api.tethered_offset("^(BC)") == 0
    assert api.tethered_offset("5p-NN(BC)") == 2
    for bad in ("(BC)", "5PAGT(BC)", "5PNN(BC)TTT", "NN(BC)", "5PNN(BC)(BC)"):
        with pytest.raises(ValueError):
            api.tethered_offset(bad)


def test_whitelist_from_txt_plain_and_translation(tmp_path):
    p = tmp_path / "wl.txt"
    p.write_text("ACGT\nTTTT\n")
    wl = api.Whitelist.from_txt(str(p))
    assert wl.translated is None and wl.length == 4 and wl.seqs.shape == (2, 4)
    (tmp_path / "translation").mkdir()  # what makes a file a translation whitelist in the reference
    g = tmp_path / "translation" / "tr.txt.gz"
    with gzip.open(g, "wt") as f:
        f.write("ACGT\tAAAA\nTTTT\tCCCC\n")
    tr = api.Whitelist.from_txt(str(g))
    assert bytes(tr.translated[1]) == b"CCCC" and bytes(tr.seqs[0]) == b"ACGT"
    with pytest.raises(ValueError):
        api.Whitelist.trans(["ACGT"], ["AAAA", "CCCC"])


def test_ascii_matrix_and_unpack_roundtrip():
    m = api.ascii_matrix(["ACGT", "TTGA"])
    assert m.shape == (2, 4) and bytes(m[1]) == b"TTGA"
    with pytest.raises(ValueError):
        api.ascii_matrix(["ACGT", "AC"])
    packed = np.array([27, 0b11110010], dtype=np.uint32)  # ACGT, TTAG
    assert [bytes(r) for r in api.unpack_2bit(packed, 4)] == [b"ACGT", b"TTAG"]


def test_count_matrix_mtx_lines_and_lazy_barcodes():
    calls = []

    def resolve(ranks):
        calls.append(1)
        return np.frombuffer(b"AAAACCCC", dtype=np.uint8).reshape(2, 4)[: len(ranks)]

    m = api.CountMatrix(np.array([3, 9], dtype=np.uint32), np.array([0, 2, 3], dtype=np.int64),
                        np.array([0, 5, 2], dtype=np.uint32), np.array([1, 4, 7], dtype=np.int32), 10,
                        resolve_barcodes=resolve)
    assert m.shape == (10, 2) and not calls
    # MtxWriter::write_matrix_mtx: `feature+1 barcode+1 count`
    assert m.mtx_lines() == ["1 1 1", "6 1 4", "3 2 7"]
    assert m.barcode_strings() == ["AAAA-1", "CCCC-1"] and len(calls) == 1
    m.barcodes
    assert len(calls) == 1


def test_feature_reference_and_chemistry_presets():
    fr = api.FeatureReference(3)
    assert fr.add_feature_barcode("CD3", "ACGTACGTACGTACG", 1) == 3 and fr.n_features == 4
    with pytest.raises(ValueError):
        fr.add_feature_barcode("x", "ACGT", 0)
    # lib/python/cellranger/chemistry_defs.json: SC3Pv2 UMI 10 @16, SC3Pv3 UMI 12 @16
    assert (api.ChemistryDef.SC3Pv2().umi_length, api.ChemistryDef.SC3Pv3().umi_length) == (10, 12)
    assert api.Posterior().bc_confidence_threshold == 0.975


def test_bench_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU port of the reference algorithm) prints one JSON line with the keys
    the driver reads; here on a tiny sample so that it runs in seconds."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-sample", "30000",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("cfg2")


def test_csc_fingerprint_adds_up_over_column_blocks():
    """bench.py's result fingerprint: the per-rank values of a barcode-owner sharded run (contiguous column blocks)
    add up to the fingerprint of the whole matrix; a changed count, feature or barcode changes it."""
    import bench

    def cm(rank, indptr, indices, data):
        return api.CountMatrix(np.array(rank, dtype=np.uint32), np.array(indptr, dtype=np.int64),
                               np.array(indices, dtype=np.uint32), np.array(data, dtype=np.int32), 10)

    whole = bench.combine_fingerprints([bench.csc_fingerprint(cm([3, 9, 12], [0, 2, 3, 3], [0, 5, 2], [1, 4, 7]), 5)])
    parts = bench.combine_fingerprints([bench.csc_fingerprint(cm([3], [0, 2], [0, 5], [1, 4]), 2),
                                        bench.csc_fingerprint(cm([9, 12], [0, 1, 1], [2], [7]), 3)])
    assert whole == parts and whole["nnz"] == 3 and whole["n_barcodes"] == 3 and whole["umis"] == 12
    for other in (cm([3, 9, 12], [0, 2, 3, 3], [0, 5, 2], [1, 4, 8]), cm([3, 9, 12], [0, 2, 3, 3], [0, 6, 2], [1, 4, 7]),
                  cm([3, 9, 13], [0, 2, 3, 3], [0, 5, 2], [1, 4, 7]), cm([3, 9, 12], [0, 1, 3, 3], [0, 5, 2], [1, 4, 7])):
        assert bench.combine_fingerprints([bench.csc_fingerprint(other, 5)])["csc_hash"] != whole["csc_hash"]


def test_tethered_offset_against_the_reference_pattern_compiler(kats):
    """tethered_offset() / tethered_min_read_length() against vectors GENERATED by the reference's own Python
    compile_pattern (lib/python/cellranger/rna/feature_ref.py:426-465, the twin of feature_extraction.rs:306-342):
    the capture of its regex on a probe read starts where tethered_offset says, wildcards behind (BC) included
    (TotalSeq-B), and its regex stops matching exactly below tethered_min_read_length cycles."""
    import re

    import pytest

    from cellranger_b200.api import tethered_min_read_length, tethered_offset

    for c in kats["tethered_patterns"]["cases"]:
        off = tethered_offset(c["pattern"])
        assert off == c["capture_start"], c
        assert c["probe"][off:off + c["length"]] == c["capture"]
        need = tethered_min_read_length(c["pattern"], c["length"])
        rx = re.compile(c["regex"])
        assert rx.search(c["probe"][:need]) is not None and rx.search(c["probe"][:need - 1]) is None, c
    assert len(kats["tethered_patterns"]["cases"]) >= 16
    for bad in ("(BC)", "NN(BC)", "5PACGT(BC)", "5PNN(BC)ACGT", "5P(BC)(BC)", "5PNN(BC)3P", "^(BC)$"):
        with pytest.raises(ValueError):
            tethered_offset(bad)


def test_barcode_strings_against_the_reference_formatter(kats):
    """CountMatrix.barcode_strings (what barcodes.tsv holds) against vectors GENERATED by the reference's own
    format_barcode_seq (lib/python/cellranger/utils.py:44-47)."""
    import numpy as np

    from cellranger_b200.api import CountMatrix

    for c in kats["format_barcode_seq"]["cases"]:
        bcs = np.frombuffer(c["barcode"].encode(), dtype=np.uint8).reshape(1, -1)
        m = CountMatrix(np.zeros(1, dtype=np.uint32), np.zeros(2, dtype=np.int64), np.zeros(0, dtype=np.uint32),
                        np.zeros(0, dtype=np.int32), 1, barcodes=bcs)
        assert m.barcode_strings(c["gem_group"]) == [c["expect"]]


def test_molecule_info_columns_against_the_reference_module(kats):
    """molecule_info_columns(): names, order and dtypes of molecule_info.h5's datasets and the Txomic code, against
    MOLECULE_INFO_COLUMNS / UMI_TYPE_TXOMIC evaluated from the reference's own module
    (lib/python/cellranger/molecule_counter.py:74-103)."""
    import numpy as np

    from cellranger_b200 import api

    k = kats["molecule_info_columns"]
    assert [[n, np.dtype(d).name] for n, d in api.MOLECULE_INFO_COLUMNS] == k["columns"]
    assert api.UMI_TYPE_TXOMIC == k["umi_type_txomic"]
    rows = np.array([[0, 0, 7, 0x2C, 3, 1], [0, 1, 9, 0x1B, 1, 0], [4, 0, 7, 0x2C, 2, 1]], dtype=np.uint32)
    cols = api.molecule_info_columns(rows, gem_group=2)
    assert list(cols) == [n for n, _ in k["columns"]]
    for n, d in k["columns"]:
        assert cols[n].dtype == np.dtype(d) and cols[n].shape == (3,)
    assert cols["gem_group"].tolist() == [2, 2, 2] and cols["barcode_idx"].tolist() == [0, 0, 4]
    assert cols["library_idx"].tolist() == [0, 1, 0] and cols["feature_idx"].tolist() == [7, 9, 7]
    assert cols["umi"].tolist() == [0x2C, 0x1B, 0x2C] and cols["count"].tolist() == [3, 1, 2]
    assert cols["umi_type"].tolist() == [1, 0, 1]


def test_chemistry_presets_against_the_reference_chemistry_defs(kats):
    """ChemistryDef.SC3Pv2 / SC3Pv3 (and the generic constructor) against the entries of the reference's
    chemistry_defs.json: barcode R1[0:16], UMI R1[16:26] (v2) / R1[16:28] (v3)."""
    from cellranger_b200.api import ChemistryDef

    e = kats["chemistry_defs"]["entries"]
    assert ChemistryDef.from_chemistry_defs_entry("SC3Pv2", e["SC3Pv2"]) == ChemistryDef.SC3Pv2()
    assert ChemistryDef.from_chemistry_defs_entry("SC3Pv3", e["SC3Pv3"]) == ChemistryDef.SC3Pv3()
    lt = ChemistryDef.from_chemistry_defs_entry("SC3Pv3LT", e["SC3Pv3LT"])
    assert (lt.bc_offset, lt.bc_length, lt.umi_offset, lt.umi_length) == (0, 16, 16, 12)
    assert e["SC3Pv2"]["barcode"][0]["whitelist"]["name"] == "737K-august-2016"
    assert e["SC3Pv3"]["barcode"][0]["whitelist"]["name"] == "3M-february-2018"


def test_whitelist_from_txt_follows_the_reference_reader(tmp_path):
    """WhitelistSource::{iter, is_translation, as_set, as_translation} (barcode/src/whitelist.rs:242-337): first
    column = sequence; a translation whitelist is one that lives in a directory called `translation`, and then
    every line carries the translated sequence too."""
    import gzip

    import pytest

    from cellranger_b200.api import Whitelist

    plain_dir, trans_dir = tmp_path / "barcodes", tmp_path / "barcodes" / "translation"
    trans_dir.mkdir(parents=True)
    (plain_dir / "wl.txt").write_text("AAAACCCCGGGGTTTT\nACGTACGTACGTACGT\n")
    w = Whitelist.from_txt(str(plain_dir / "wl.txt"))
    assert w.translated is None and [bytes(s) for s in w.seqs] == [b"AAAACCCCGGGGTTTT", b"ACGTACGTACGTACGT"]
    # extra columns outside a translation directory are ignored (as_set takes the left column only)
    (plain_dir / "wl_ids.txt").write_text("AAAACCCCGGGGTTTT\tTTTTGGGGCCCCAAAA\tBC001\nACGTACGTACGTACGT\tTGCATGCATGCATGCA\tBC002\n")
    w = Whitelist.from_txt(str(plain_dir / "wl_ids.txt"))
    assert w.translated is None and w.seqs.shape == (2, 16)
    with gzip.open(trans_dir / "wl.txt.gz", "wt") as f:
        f.write("AAAACCCCGGGGTTTT TTTTGGGGCCCCAAAA\nACGTACGTACGTACGT TGCATGCATGCATGCA\n")
    w = Whitelist.from_txt(str(trans_dir / "wl.txt.gz"))
    assert [bytes(s) for s in w.translated] == [b"TTTTGGGGCCCCAAAA", b"TGCATGCATGCATGCA"]
    (trans_dir / "broken.txt").write_text("AAAACCCCGGGGTTTT TTTTGGGGCCCCAAAA\nACGTACGTACGTACGT\n")
    with pytest.raises(ValueError, match="not a translation whitelist"):
        Whitelist.from_txt(str(trans_dir / "broken.txt"))
    assert Whitelist.from_txt(str(plain_dir / "wl_ids.txt"), translation=True).translated is not None
