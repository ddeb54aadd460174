"""Host side of the B200 barcode / UMI path, mirroring the reference's interfaces for this path.

Names follow the reference (paths under lib/rust/ of the reference checkout):
  Whitelist            barcode/src/whitelist.rs:452-525        Plain / Trans
  Posterior            barcode/src/corrector.rs:93-109          the correction strategy's two parameters
  BarcodeCorrector     barcode/src/corrector.rs:15-71           batch form of correct_barcode
  ChemistryDef         cr_types/src/chemistry/mod.rs:718-751    only the read layout the path needs
  FeatureReference     cr_types/src/reference/feature_reference.rs  genes + tethered feature barcodes
  GemWell              the three hot stages over one GEM well:
     .make_shard()          cr_lib/src/stages/make_shard.rs          exact match + priors
     .barcode_correction()  cr_lib/src/stages/barcode_correction.rs  Hamming-1 posterior correction
     .align_and_count()     cr_lib/src/stages/align_and_count.rs     UMI correction / dedup / counts
     .count_matrix()        cr_h5/src/count_matrix.rs:382-448        CSC arrays, barcode index
Everything computes on the GPU through libcrgpu.so (include/crgpu.h); there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import NO_FEATURE, NO_RANK, LibraryDef, ReadBatch, check, ptr

F64_MAX = 1.7976931348623157e308

# BarcodeSegmentState, barcode/src/lib.rs:270-283
NOT_CHECKED, VALID_BEFORE_CORRECTION, VALID_AFTER_CORRECTION, INVALID = 0, 1, 2, 3
# per-read flags (DupInfo, tx_annotation/src/mark_dups.rs:61-72)
F_UMI_VALID, F_HAS_DUPINFO, F_UMI_CORRECTED, F_LOW_SUPPORT, F_UMI_COUNT = 1, 2, 4, 8, 16

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def ascii_matrix(seqs, L: Optional[int] = None) -> np.ndarray:
    """list of str/bytes of equal length, or an (n, L) uint8 array -> (n, L) uint8 ASCII matrix."""
    if isinstance(seqs, np.ndarray):
        a = np.ascontiguousarray(seqs, dtype=np.uint8)
        if a.ndim == 1 and L:
            a = a.reshape(-1, L)
        return a
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    if L is None:
        L = len(bs[0]) if bs else 0
    if any(len(b) != L for b in bs):
        raise ValueError("sequences must all have the same length")
    return np.frombuffer(b"".join(bs), dtype=np.uint8).reshape(len(bs), L).copy()


def unpack_2bit(packed: np.ndarray, L: int) -> np.ndarray:
    """2-bit packed sequences (first base most significant) -> (n, L) ASCII."""
    p = np.asarray(packed).astype(np.uint64)
    out = np.empty((p.shape[0], L), dtype=np.uint8)
    for i in range(L):
        out[:, i] = _BASES[((p >> np.uint64(2 * (L - 1 - i))) & np.uint64(3)).astype(np.int64)]
    return out


@dataclass
class Whitelist:
    """Whitelist::Plain (translated is None) or Whitelist::Trans (raw -> translated)."""
    seqs: np.ndarray
    translated: Optional[np.ndarray] = None

    @staticmethod
    def plain(seqs) -> "Whitelist":
        return Whitelist(ascii_matrix(seqs))

    @staticmethod
    def trans(raw, translated) -> "Whitelist":
        r, t = ascii_matrix(raw), ascii_matrix(translated)
        if r.shape != t.shape:
            raise ValueError("raw and translated whitelists differ in shape")
        return Whitelist(r, t)

    @staticmethod
    def from_txt(path: str, translation: Optional[bool] = None) -> "Whitelist":
        """A whitelist .txt[.gz] as WhitelistSource reads it (barcode/src/whitelist.rs:242-337): every line is split
        on whitespace, the first column is the sequence, the second (when present) its translation. The file is a
        translation whitelist exactly when its parent directory is called `translation` (is_translation, :242-253;
        `translation=` overrides that); then every line needs both columns (as_translation, :299-309), otherwise
        only the first column counts (as_set, :288-290)."""
        import gzip
        import os

        if translation is None:
            translation = os.path.basename(os.path.dirname(os.path.abspath(path))) == "translation"
        op = gzip.open if path.endswith(".gz") else open
        raw, tr = [], []
        with op(path, "rt") as f:
            for line in f:
                parts = line.split()
                if not parts:
                    raise ValueError(f"{path}: empty line in a whitelist")  # iter.next().unwrap() in the reference
                raw.append(parts[0])
                if translation:
                    if len(parts) < 2:
                        raise ValueError(f"not a translation whitelist: {path}")
                    tr.append(parts[1])
        return Whitelist.trans(raw, tr) if translation else Whitelist.plain(raw)

    @property
    def length(self) -> int:
        return int(self.seqs.shape[1])


@dataclass
class Posterior:
    """Posterior { max_expected_barcode_errors, bc_confidence_threshold } with the reference defaults."""
    max_expected_barcode_errors: float = F64_MAX
    bc_confidence_threshold: float = 0.975


@dataclass
class ChemistryDef:
    """Read layout of a chemistry (lib/python/cellranger/chemistry_defs.json)."""
    name: str
    bc_offset: int = 0
    bc_length: int = 16
    umi_offset: int = 16
    umi_length: int = 12

    @staticmethod
    def SC3Pv2() -> "ChemistryDef":
        return ChemistryDef("SC3Pv2", 0, 16, 16, 10)

    @staticmethod
    def SC3Pv3() -> "ChemistryDef":
        return ChemistryDef("SC3Pv3", 0, 16, 16, 12)

    @staticmethod
    def from_chemistry_defs_entry(name: str, entry: dict) -> "ChemistryDef":
        """From one entry of the reference's chemistry_defs.json (lib/python/cellranger/chemistry_defs.json): the
        layouts this path handles keep one gel-bead barcode segment and the UMI on R1."""
        bcs, umis = entry["barcode"], entry["umi"]
        if len(bcs) != 1 or len(umis) != 1 or bcs[0]["read_type"] != "R1" or umis[0]["read_type"] != "R1":
            raise ValueError(f"chemistry {name}: only a single R1 barcode segment with the UMI on R1 is supported here")
        return ChemistryDef(name, int(bcs[0]["offset"]), int(bcs[0]["length"]), int(umis[0]["offset"]),
                            int(umis[0]["length"]))


_PATTERN_RE = re.compile(r"^(?:5[Pp]?[-_]?|\^)?([ACGTN]*)\(BC\)([ACGTN]*)(?:[-_]?3[Pp]?|\$)?$")


def _tethered_parts(pattern: str):
    m = _PATTERN_RE.match(pattern)
    if not m or not (pattern.startswith("5") or pattern.startswith("^")):
        raise ValueError(f"unsupported feature pattern {pattern!r}: need a 5'-tethered pattern with one (BC)")
    pre, post = m.group(1), m.group(2)
    if set(pre) - {"N"} or set(post) - {"N"}:
        raise ValueError(f"unsupported feature pattern {pattern!r}: only N wildcards around (BC) are supported")
    if re.search(r"(?:[-_]?3[Pp]?|\$)$", pattern):
        raise ValueError(f"unsupported feature pattern {pattern!r}: anchored at both ends (the read length itself "
                         "would decide the match)")
    return len(pre), len(post)


def tethered_offset(pattern: str) -> int:
    """Offset of the (BC) capture of a 5'-tethered pattern such as `5PNNNNNNNNNN(BC)`, `^(BC)` or TotalSeq-B's
    `5PNNNNNNNNNN(BC)NNNNNNNNN` (FeatureExtractor::compile_pattern, cr_types/src/reference/feature_extraction.rs:
    306-342; Python twin: lib/python/cellranger/rna/feature_ref.py:426-465). Only wildcard (N) bases may surround
    the capture; anything else is outside this path."""
    return _tethered_parts(pattern)[0]


def tethered_min_read_length(pattern: str, fb_length: int) -> int:
    """Cycles a read needs before the reference's anchored regex can match it at all: the wildcards before the
    capture, the capture, and the wildcards behind it. Hand it to add_library(fb_min_read_length=...)."""
    pre, post = _tethered_parts(pattern)
    return pre + int(fb_length) + post


FEATURE_TYPE_NAMES = {1: "Antibody Capture", 2: "CRISPR Guide Capture", 3: "Multiplexing Capture", 4: "Custom"}


@dataclass
class FeatureReference:
    """Genes (indices 0..n_genes-1) followed by feature-barcode features."""
    n_genes: int
    fb_ids: list = field(default_factory=list)
    fb_seqs: list = field(default_factory=list)       # str, equal length per feature type
    fb_types: list = field(default_factory=list)      # feature-type id (>= 1) per feature barcode
    fb_patterns: list = field(default_factory=list)   # pattern string per feature barcode

    @property
    def n_features(self) -> int:
        return self.n_genes + len(self.fb_seqs)

    def add_feature_barcode(self, fid: str, seq: str, feature_type: int, pattern: str = "5PNNNNNNNNNN(BC)") -> int:
        if feature_type < 1:
            raise ValueError("feature_type 0 is reserved for genes")
        self.fb_ids.append(fid)
        self.fb_seqs.append(seq)
        self.fb_types.append(feature_type)
        self.fb_patterns.append(pattern)
        return self.n_features - 1


class CountMatrix:
    """Feature x barcode matrix in CSC, as write_matrix_h5_helper lays it out
    (cr_h5/src/count_matrix.rs:382-448): data i32, indices = feature index, indptr i64 per barcode.
    `barcodes` ((n_barcodes, L) ASCII, sorted) is materialised on first use from the content ranks."""

    def __init__(self, barcode_rank, indptr, indices, data, n_features, resolve_barcodes=None, barcodes=None):
        self.barcode_rank = barcode_rank   # uint32 content rank of each column
        self.indptr = indptr               # int64[n_barcodes + 1]
        self.indices = indices             # uint32[nnz]
        self.data = data                   # int32[nnz]
        self.n_features = n_features
        self._resolve = resolve_barcodes
        self._barcodes = barcodes

    @property
    def barcodes(self) -> np.ndarray:
        if self._barcodes is None:
            self._barcodes = self._resolve(self.barcode_rank)
        return self._barcodes

    @property
    def shape(self):
        return (self.n_features, self.barcode_rank.shape[0])

    def barcode_strings(self, gem_group: int = 1):
        return [f"{bytes(b).decode()}-{gem_group}" for b in self.barcodes]

    def mtx_lines(self):
        """`feature barcode count` triplets, 1-based, in file order
        (MtxWriter::write_matrix_mtx, cr_lib/src/stages/write_matrix_market.rs:81-120)."""
        cols = np.repeat(np.arange(self.barcode_rank.shape[0], dtype=np.int64), np.diff(self.indptr))
        return [f"{int(f) + 1} {int(c) + 1} {int(v)}" for f, c, v in zip(self.indices, cols, self.data)]


def comm_unique_id() -> bytes:
    """128 bytes that identify a new communicator (crgpu_comm_unique_id): rank 0 makes it, every rank gets it."""
    L = _lib.load()
    buf = (C.c_uint8 * 128)()
    check(L.crgpu_comm_unique_id(buf), "crgpu_comm_unique_id")
    return bytes(buf)


class GemWell:
    """One GEM well on one GPU: whitelist tables, library types, read batches and the three stages."""

    def __init__(self, device: int = 0, posterior: Optional[Posterior] = None, filter_umis: bool = True):
        self.L = _lib.load()
        self._ctx = C.c_void_p()
        check(self.L.crgpu_ctx_create(int(device), C.byref(self._ctx)), "crgpu_ctx_create")
        self.device = device
        self.posterior = posterior or Posterior()
        check(self.L.crgpu_set_params(self._ctx, C.c_double(self.posterior.bc_confidence_threshold),
                                      C.c_double(self.posterior.max_expected_barcode_errors), int(filter_umis)))
        self._keep = []
        self._pinned_bufs = {}
        self._whitelists = []
        self._libs = []
        self._batches = []
        self._dev_owned = []   # device arrays allocated by add_fastq, released with the reads
        self.bc_length = None
        self.umi_length = None
        self.feature_reference: Optional[FeatureReference] = None

    # ---- lifecycle ----
    def close(self):
        if getattr(self, "_ctx", None):
            for p, _ in self._pinned_bufs.values():
                self.L.crgpu_host_free_pinned(C.c_void_p(p))
            self._pinned_bufs = {}
            for p in getattr(self, "_dev_owned", []):
                self.L.crgpu_dev_free(self._ctx, C.c_void_p(p))
            self._dev_owned = []
            self.L.crgpu_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def ctx(self):
        return self._ctx

    # ---- setup ----
    def add_whitelist(self, wl: Whitelist) -> int:
        out = C.c_int(-1)
        check(self.L.crgpu_whitelist_add(self._ctx, ptr(wl.seqs), C.c_uint64(wl.seqs.shape[0]), wl.length,
                                         ptr(wl.translated), C.byref(out)), "crgpu_whitelist_add")
        self._whitelists.append(wl)
        self.bc_length = wl.length
        return out.value

    def add_library(self, whitelist: int, chemistry: ChemistryDef, umi_correction: bool = True,
                    feature_type: int = 0, fb_offset: int = 0, fb_length: int = 0, fb_min_read_length: int = 0) -> int:
        """A library type. feature_type 0 = Gene Expression (features come with the reads);
        otherwise a feature-barcode library whose features carry that feature_type. fb_min_read_length
        (tethered_min_read_length): a pattern with wildcards behind (BC) only matches reads that long - batches
        with a shorter R2 stride are refused, since the reference would match none of their reads."""
        d = LibraryDef(whitelist, chemistry.bc_offset, chemistry.bc_length, chemistry.umi_offset,
                       chemistry.umi_length, int(umi_correction), int(feature_type != 0), feature_type,
                       fb_offset, fb_length)
        out = C.c_int(-1)
        check(self.L.crgpu_library_add(self._ctx, C.byref(d), C.byref(out)), "crgpu_library_add")
        self._libs.append(d)
        self._fb_min_read_length = getattr(self, "_fb_min_read_length", {})
        self._fb_min_read_length[out.value] = max(int(fb_min_read_length), int(fb_offset + fb_length))
        self.umi_length = chemistry.umi_length
        return out.value

    def set_feature_reference(self, fr: FeatureReference):
        ft = np.zeros(fr.n_features, dtype=np.int32)
        stride = max([len(s) for s in fr.fb_seqs], default=1)
        seqs = np.full((fr.n_features, stride), ord("A"), dtype=np.uint8)
        for i, (s, t) in enumerate(zip(fr.fb_seqs, fr.fb_types)):
            ft[fr.n_genes + i] = t
            seqs[fr.n_genes + i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
        check(self.L.crgpu_features_set(self._ctx, fr.n_features, ptr(ft), ptr(seqs), stride), "crgpu_features_set")
        self.feature_reference = fr

    def total_barcode_counts(self, min_reads_to_report_bc: int = 1):
        """total_barcode_counts of BARCODE_CORRECTION (cr_lib/src/stages/barcode_correction.rs:327-362): the reads
        that were not valid before correction, by barcode after correction -> (seqs (n, L) ASCII, valid, counts),
        invalid sequences first. Needs barcode_correction()."""
        n = C.c_uint64()
        check(self.L.crgpu_total_barcode_counts(self._ctx, C.c_uint64(int(min_reads_to_report_bc)), C.byref(n)),
              "crgpu_total_barcode_counts")
        seqs = np.zeros((n.value, self.bc_length or 0), dtype=np.uint8)
        valid = np.zeros(n.value, dtype=np.uint8)
        counts = np.zeros(n.value, dtype=np.uint64)
        check(self.L.crgpu_total_barcode_counts_get(self._ctx, ptr(seqs), ptr(valid), ptr(counts)),
              "crgpu_total_barcode_counts_get")
        return seqs, valid, counts

    def set_target_filter(self, on_target, targeted_umi_min_read_count: Optional[int]):
        """DupBuilder::build(filter_umis, umi_correction, targeted_umi_min_read_count) with the feature
        reference's target set (tx_annotation/src/mark_dups.rs:156-170,311-320): on_target = bool per feature;
        None / 0 switches the filter off."""
        t = np.ascontiguousarray(on_target, dtype=np.uint8)
        check(self.L.crgpu_set_target_filter(self._ctx, ptr(t), C.c_int32(t.shape[0]),
                                             C.c_uint64(int(targeted_umi_min_read_count or 0))), "crgpu_set_target_filter")

    def add_reads(self, library: int, r1_seq, r1_qual, feature=None, r2_seq=None, r2_qual=None, select_key=None) -> int:
        """Host arrays: r1_seq/r1_qual (n, r1_len) uint8, feature uint32[n] (GEX) or r2_* (feature barcode);
        select_key: optional uint64[n], UmiSelectKey{utype, qname} per read as one order-preserving word (bit 63
        = NonTxomic, low bits = rank of the qname; tx_annotation/src/mark_dups.rs:110-152)."""
        r1_seq = np.ascontiguousarray(r1_seq, dtype=np.uint8)
        r1_qual = np.ascontiguousarray(r1_qual, dtype=np.uint8)
        if r1_seq.ndim != 2 or r1_seq.shape != r1_qual.shape:
            raise ValueError("r1_seq and r1_qual must be (n, r1_len) arrays of the same shape")
        n, r1_len = r1_seq.shape
        rb = ReadBatch()
        rb.n, rb.r1_len = n, r1_len
        rb.r1_seq, rb.r1_qual = r1_seq.ctypes.data, r1_qual.ctypes.data
        keep = [r1_seq, r1_qual]
        if feature is not None:
            feature = np.ascontiguousarray(feature, dtype=np.uint32)
            if feature.shape != (n,):
                raise ValueError("feature must have one entry per read")
            rb.feature = feature.ctypes.data
            keep.append(feature)
        if r2_seq is not None:
            r2_seq = np.ascontiguousarray(r2_seq, dtype=np.uint8)
            r2_qual = np.ascontiguousarray(r2_qual, dtype=np.uint8)
            rb.r2_len = r2_seq.shape[1]
            need = getattr(self, "_fb_min_read_length", {}).get(library, 0)
            if self._libs[library].is_feature_barcode and rb.r2_len < need:
                raise ValueError(f"R2 of {rb.r2_len} cycles is shorter than the {need} cycles the feature pattern "
                                 "needs: no read could match")
            rb.r2_seq, rb.r2_qual = r2_seq.ctypes.data, r2_qual.ctypes.data
            keep += [r2_seq, r2_qual]
        if select_key is not None:
            select_key = np.ascontiguousarray(select_key, dtype=np.uint64)
            if select_key.shape != (n,):
                raise ValueError("select_key must have one entry per read")
            rb.select_key = select_key.ctypes.data
            keep.append(select_key)
        rb.on_device = 0
        out = C.c_int(-1)
        check(self.L.crgpu_reads_add(self._ctx, library, C.byref(rb), C.byref(out)), "crgpu_reads_add")
        self._keep.append(keep)
        self._batches.append((library, n))
        return out.value

    def add_reads_device(self, library: int, n: int, r1_len: int, r1_seq: int, r1_qual: int, feature: int = 0,
                         r2_len: int = 0, r2_seq: int = 0, r2_qual: int = 0, select_key: int = 0) -> int:
        """Device pointers (ints), borrowed until clear_reads()."""
        rb = ReadBatch()
        rb.n, rb.r1_len, rb.r1_seq, rb.r1_qual = n, r1_len, r1_seq, r1_qual
        rb.feature = feature or None
        rb.r2_len, rb.r2_seq, rb.r2_qual = r2_len, r2_seq or None, r2_qual or None
        rb.select_key = select_key or None
        rb.on_device = 1
        out = C.c_int(-1)
        check(self.L.crgpu_reads_add(self._ctx, library, C.byref(rb), C.byref(out)), "crgpu_reads_add")
        self._batches.append((library, n))
        return out.value

    @staticmethod
    def _fastq_bytes(fastq) -> np.ndarray:
        """FASTQ text as a uint8 array; gzip members (.fastq.gz, magic 1f 8b) are inflated on the host first, as the
        reference's readers do (fastq_set over flate2) - decompression is outside the accelerated path."""
        if isinstance(fastq, (bytes, bytearray, memoryview)):
            raw = bytes(fastq)
            if raw[:2] == b"\x1f\x8b":
                import gzip

                raw = gzip.decompress(raw)
            return np.frombuffer(raw, dtype=np.uint8)
        a = np.ascontiguousarray(fastq, dtype=np.uint8)
        if a.shape[0] >= 2 and a[0] == 0x1F and a[1] == 0x8B:
            import gzip

            return np.frombuffer(gzip.decompress(a.tobytes()), dtype=np.uint8)
        return a

    def _fastq_to_device(self, text: np.ndarray, read_len: int):
        cap = int(np.count_nonzero(text == 10) // 4 + 1)
        dev = []
        for _ in range(2):
            p = C.c_void_p()
            check(self.L.crgpu_dev_alloc(self._ctx, C.c_uint64(cap * read_len + 16), C.byref(p)), "crgpu_dev_alloc")
            dev.append(p.value)
        n_rec, n_short, n_bad = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self.L.crgpu_fastq_extract(self._ctx, ptr(text), C.c_uint64(text.shape[0]), 0, read_len, C.c_void_p(dev[0]),
                                         C.c_void_p(dev[1]), C.c_uint64(cap), C.byref(n_rec), C.byref(n_short),
                                         C.byref(n_bad)), "crgpu_fastq_extract")
        return dev, int(n_rec.value), int(n_short.value), int(n_bad.value)

    def add_fastq(self, library: int, r1_fastq, feature=None, read_len: Optional[int] = None, r2_fastq=None,
                  select_key=None) -> dict:
        """One chunk of R1 FASTQ text (bytes or uint8 array, whole 4-line records; plain or gzip) as a read batch:
        crgpu_fastq_extract slices the first `read_len` bases / qualities of every record on the device
        (RnaProcessor::process_read, cr_types/src/rna_read.rs:363-467). A gene-expression library gets `feature`
        (uint32 per record, from the aligner) as in add_reads; a feature-barcode library gets `r2_fastq`, the R2
        chunk of the same read pairs, whose first fb_offset + fb_length cycles hold the capture sequence
        (FeatureExtractor::match_read, cr_types/src/reference/feature_extraction.rs:358-471).
        Returns {batch, n_records, n_short, n_malformed, ...}."""
        d = self._libs[library]
        text = self._fastq_bytes(r1_fastq)
        rl = int(read_len or max(d.bc_offset + d.bc_length, d.umi_offset + d.umi_length))
        dev, n, n_short, n_bad = self._fastq_to_device(text, rl)
        self._dev_owned.extend(dev)
        info = {"n_records": n, "n_short": n_short, "n_malformed": n_bad, "read_len": rl, "dev_seq": dev[0],
                "dev_qual": dev[1]}
        fdev, r2_len, r2s, r2q = 0, 0, 0, 0
        if d.is_feature_barcode:
            if r2_fastq is None:
                raise ValueError("a feature-barcode library needs the R2 FASTQ of the read pairs")
            r2_len = int(d.fb_offset + d.fb_length)
            dev2, n2, short2, bad2 = self._fastq_to_device(self._fastq_bytes(r2_fastq), r2_len)
            self._dev_owned.extend(dev2)
            if n2 != n:
                raise ValueError(f"{n} R1 records but {n2} R2 records: the two files are not the same read pairs")
            r2s, r2q = dev2
            info.update(r2_len=r2_len, n_short_r2=short2, n_malformed_r2=bad2, dev_r2_seq=r2s, dev_r2_qual=r2q)
        elif feature is not None:
            f = np.ascontiguousarray(feature, dtype=np.uint32)
            if f.shape[0] != n:
                raise ValueError(f"{n} FASTQ records but {f.shape[0]} feature assignments")
            p = C.c_void_p()
            check(self.L.crgpu_dev_alloc(self._ctx, C.c_uint64(max(n, 1) * 4), C.byref(p)), "crgpu_dev_alloc")
            check(self.L.crgpu_memcpy_h2d(self._ctx, p, ptr(f), C.c_uint64(f.nbytes)), "crgpu_memcpy_h2d")
            fdev = p.value
            self._dev_owned.append(fdev)
        sdev = 0
        if select_key is not None:
            k = np.ascontiguousarray(select_key, dtype=np.uint64)
            if k.shape[0] != n:
                raise ValueError("select_key must have one entry per record")
            p = C.c_void_p()
            check(self.L.crgpu_dev_alloc(self._ctx, C.c_uint64(max(n, 1) * 8), C.byref(p)), "crgpu_dev_alloc")
            check(self.L.crgpu_memcpy_h2d(self._ctx, p, ptr(k), C.c_uint64(k.nbytes)), "crgpu_memcpy_h2d")
            sdev = p.value
            self._dev_owned.append(sdev)
        info["batch"] = self.add_reads_device(library, n, rl, dev[0], dev[1], fdev, r2_len, r2s, r2q, select_key=sdev)
        return info

    def read_device(self, dev: int, shape, dtype=np.uint8) -> np.ndarray:
        """Copy a device array of this context to the host (inspection, tests)."""
        out = np.zeros(shape, dtype=dtype)
        if out.nbytes:
            check(self.L.crgpu_memcpy_d2h(self._ctx, ptr(out), C.c_void_p(dev), C.c_uint64(out.nbytes)), "crgpu_memcpy_d2h")
        return out

    def clear_reads(self):
        check(self.L.crgpu_reads_clear(self._ctx))
        for p in self._dev_owned:
            self.L.crgpu_dev_free(self._ctx, C.c_void_p(p))
        self._dev_owned = []
        self._keep.clear()
        self._batches.clear()

    # ---- stages ----
    def make_shard(self):
        check(self.L.crgpu_pass1(self._ctx), "crgpu_pass1")

    def barcode_correction(self):
        check(self.L.crgpu_pass2(self._ctx), "crgpu_pass2")

    def align_and_count(self, annotate_reads: bool = False):
        check(self.L.crgpu_count(self._ctx), "crgpu_count")
        if annotate_reads:
            check(self.L.crgpu_annotate_reads(self._ctx), "crgpu_annotate_reads")

    def run(self, annotate_reads: bool = False):
        self.make_shard()
        self.barcode_correction()
        self.align_and_count(annotate_reads)

    def sync(self):
        check(self.L.crgpu_sync(self._ctx), "crgpu_sync")

    # ---- cross-chunk state ----
    def n_content(self) -> int:
        n = C.c_uint64()
        check(self.L.crgpu_whitelist_size(self._ctx, C.byref(n), None))
        return int(n.value)

    def set_prior(self, library: int, counts):
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        check(self.L.crgpu_prior_set(self._ctx, library, ptr(c), C.c_uint64(c.shape[0])), "crgpu_prior_set")

    def prior(self, library: int) -> np.ndarray:
        return self._counts(library, 0)

    def corrected_counts(self, library: int) -> np.ndarray:
        return self._counts(library, 1)

    def _counts(self, library, which):
        out = np.zeros(self.n_content(), dtype=np.uint32)
        check(self.L.crgpu_bc_counts_get(self._ctx, library, which, ptr(out), C.c_uint64(out.shape[0])))
        return out

    def fb_exact_counts(self) -> np.ndarray:
        n = self.feature_reference.n_features if self.feature_reference else 0
        out = np.zeros(n, dtype=np.int64)
        check(self.L.crgpu_fb_counts_get(self._ctx, ptr(out), n))
        return out

    def prior_dev(self, library: int):
        p, n = C.c_void_p(), C.c_uint64()
        check(self.L.crgpu_prior_dev(self._ctx, library, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def valid_counts_dev(self, library: int):
        p, n = C.c_void_p(), C.c_uint64()
        check(self.L.crgpu_valid_counts_dev(self._ctx, library, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def corrected_dev(self, library: int):
        p, n = C.c_void_p(), C.c_uint64()
        check(self.L.crgpu_corrected_dev(self._ctx, library, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def valid_counts_refresh(self):
        check(self.L.crgpu_valid_counts_refresh(self._ctx), "crgpu_valid_counts_refresh")

    def fb_counts_dev(self):
        p, n = C.c_void_p(), C.c_int32()
        check(self.L.crgpu_fb_counts_dev(self._ctx, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def keys_dev(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(self.L.crgpu_keys_dev(self._ctx, C.byref(p), C.byref(n)), "crgpu_keys_dev")
        return p.value, int(n.value)

    def keys_partition(self, bounds) -> np.ndarray:
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        out = np.zeros(b.shape[0] - 1, dtype=np.uint64)
        check(self.L.crgpu_keys_partition(self._ctx, b.shape[0] - 1, ptr(b), ptr(out)), "crgpu_keys_partition")
        return out

    def keys_set(self, dev_ptr: int, n: int):
        check(self.L.crgpu_keys_set(self._ctx, C.c_void_p(dev_ptr), C.c_uint64(n)), "crgpu_keys_set")

    # ---- fused key exchange over peer memory (CUDA IPC + NVLink stores) ----
    def exchange_init(self, capacity_keys: int) -> bytes:
        h = (C.c_uint8 * 128)()
        check(self.L.crgpu_exchange_init(self._ctx, C.c_uint64(capacity_keys), h), "crgpu_exchange_init")
        return bytes(h)

    def exchange_connect(self, n_ranks: int, my_rank: int, handles: bytes):
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        check(self.L.crgpu_exchange_connect(self._ctx, n_ranks, my_rank, buf), "crgpu_exchange_connect")

    def exchange_reset(self):
        check(self.L.crgpu_exchange_reset(self._ctx), "crgpu_exchange_reset")

    def keys_scatter_peers_begin(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        check(self.L.crgpu_keys_scatter_peers_begin(self._ctx, b.shape[0] - 1, ptr(b)), "crgpu_keys_scatter_peers_begin")

    def keys_scatter_peers(self, bounds) -> np.ndarray:
        b = np.ascontiguousarray(bounds, dtype=np.uint32)
        out = np.zeros(b.shape[0] - 1, dtype=np.uint64)
        check(self.L.crgpu_keys_scatter_peers(self._ctx, b.shape[0] - 1, ptr(b), ptr(out)), "crgpu_keys_scatter_peers")
        return out

    def exchange_finish(self) -> int:
        n = C.c_uint64()
        check(self.L.crgpu_exchange_finish(self._ctx, C.byref(n)), "crgpu_exchange_finish")
        return int(n.value)

    # ---- the sharded run inside the library (NCCL communicator, device owner ranges, fused exchange) ----
    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int, exchange_capacity_keys: int):
        """Collective over the ranks (one process per GPU): joins the communicator created from `unique_id`
        (comm_unique_id() on rank 0, handed to every rank by the host) and maps every rank's receive buffer."""
        buf = (C.c_uint8 * len(unique_id)).from_buffer_copy(unique_id)
        check(self.L.crgpu_comm_init(self._ctx, buf, int(n_ranks), int(rank), C.c_uint64(int(exchange_capacity_keys))),
              "crgpu_comm_init")

    def sharded_run(self):
        """One step of the sharded path, collective: pass 1, all-reduce of the priors, pass 2, all-reduce of the
        corrected counts, owner ranges, key exchange over peer memory, count on the owned barcode range."""
        check(self.L.crgpu_sharded_run(self._ctx), "crgpu_sharded_run")

    def owner_bounds(self) -> np.ndarray:
        out = np.zeros(17, dtype=np.uint32)
        check(self.L.crgpu_owner_bounds_get(self._ctx, ptr(out), 17), "crgpu_owner_bounds_get")
        return out[: self.shard_stats()["n_ranks"] + 1]

    def shard_stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        check(self.L.crgpu_shard_stats(self._ctx, out))
        return {"sent_remote_keys": int(out[0]), "received_keys": int(out[1]), "n_ranks": int(out[2]), "rank": int(out[3])}

    def owner_bounds_compute(self, counts, n_parts: int) -> np.ndarray:
        """Owner ranges for a vector of per-rank read counts, with the device arithmetic of the sharded run."""
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        out = np.zeros(n_parts + 1, dtype=np.uint32)
        check(self.L.crgpu_owner_bounds_compute(self._ctx, ptr(c), C.c_uint64(c.shape[0]), int(n_parts), ptr(out)),
              "crgpu_owner_bounds_compute")
        return out

    def set_owned_range(self, lo: int, hi: int):
        check(self.L.crgpu_set_owned_range(self._ctx, C.c_uint32(lo), C.c_uint32(hi)))

    def key_layout(self) -> dict:
        v = [C.c_int32() for _ in range(4)]
        check(self.L.crgpu_key_layout(self._ctx, *[C.byref(x) for x in v]))
        return dict(rank_shift=v[0].value, feature_shift=v[1].value, lib_shift=v[2].value, umi_bits=v[3].value)

    def stream(self) -> int:
        p = C.c_void_p()
        check(self.L.crgpu_stream(self._ctx, C.byref(p)))
        return p.value or 0

    # ---- results ----
    def stats(self) -> dict:
        out = (C.c_uint64 * len(_lib.STAT_NAMES))()
        check(self.L.crgpu_stats(self._ctx, out))
        return {k: int(v) for k, v in zip(_lib.STAT_NAMES, out) if not k.startswith("_")}

    def phase_times(self) -> dict:
        cap = 64
        ms = (C.c_float * cap)()
        n = C.c_int32()
        names = C.c_void_p()
        check(self.L.crgpu_phase_times(self._ctx, ms, cap, C.byref(n), C.byref(names)))
        out, addr = {}, names.value or 0
        for i in range(min(n.value, cap)):  # the names are NUL-separated; read each up to its own terminator
            name = C.string_at(addr)
            addr += len(name) + 1
            out[name.decode()] = float(ms[i])
        return out

    def barcode_seqs(self, ranks) -> np.ndarray:
        r = np.ascontiguousarray(ranks, dtype=np.uint32)
        out = np.zeros((r.shape[0], self.bc_length), dtype=np.uint8)
        check(self.L.crgpu_barcode_seqs(self._ctx, ptr(r), C.c_uint64(r.shape[0]), ptr(out)))
        return out

    def reads(self, batch: int = 0) -> dict:
        """Per-read results of one batch: bc_rank, bc_state, umi (2-bit), flags, feature."""
        n = self._batches[batch][1]
        bc_rank = np.zeros(n, dtype=np.uint32)
        state = np.zeros(n, dtype=np.uint8)
        umi = np.zeros(n, dtype=np.uint32)
        flags = np.zeros(n, dtype=np.uint8)
        feature = np.zeros(n, dtype=np.uint32)
        check(self.L.crgpu_reads_get(self._ctx, batch, ptr(bc_rank), ptr(state), ptr(umi), ptr(flags), ptr(feature)),
              "crgpu_reads_get")
        return dict(bc_rank=bc_rank, state=state, umi=umi, flags=flags, feature=feature)

    def _pinned(self, name: str, shape, dtype) -> np.ndarray:
        """A numpy view of a grow-only page-locked host buffer owned by this GemWell."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        cur = self._pinned_bufs.get(name)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                self.L.crgpu_host_free_pinned(C.c_void_p(cur[0]))
            p = C.c_void_p()
            cap = max(nbytes + nbytes // 8, 4096)
            check(self.L.crgpu_host_alloc_pinned(C.c_uint64(cap), C.byref(p)), "crgpu_host_alloc_pinned")
            cur = (p.value, cap)
            self._pinned_bufs[name] = cur
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(cur[0])
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def count_matrix(self, pinned: bool = False) -> CountMatrix:
        """The CSC arrays. pinned=True reads them into page-locked buffers that this GemWell reuses on
        the next call (the arrays of an earlier call are then overwritten)."""
        nb, nnz, nf = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self.L.crgpu_matrix_dims(self._ctx, C.byref(nb), C.byref(nnz), C.byref(nf)), "crgpu_matrix_dims")
        mk = (lambda nm, sh, dt: self._pinned(nm, sh, dt)) if pinned else (lambda nm, sh, dt: np.zeros(sh, dtype=dt))
        rank = mk("rank", (nb.value,), np.uint32)
        indptr = mk("indptr", (nb.value + 1,), np.int64)
        indices = mk("indices", (nnz.value,), np.uint32)
        data = mk("data", (nnz.value,), np.int32)
        check(self.L.crgpu_matrix_get(self._ctx, ptr(rank), ptr(indptr), ptr(indices), ptr(data)), "crgpu_matrix_get")
        return CountMatrix(rank, indptr, indices, data, int(nf.value), resolve_barcodes=self.barcode_seqs)

    def barcode_summary(self, library: int = 0) -> np.ndarray:
        """BarcodeSummary rows of one library (cr_lib/src/aligner.rs:33-68): a structured array in matrix
        column order with fields barcode_rank, reads, umis, candidate_dup_reads, umi_corrected_reads. Like the
        reference (visit_read_annotation, cr_lib/src/align_metrics.rs:705-721) only barcodes with at least one
        read in this library get a row."""
        nb = C.c_uint64()
        check(self.L.crgpu_matrix_dims(self._ctx, C.byref(nb), None, None), "crgpu_matrix_dims")
        raw = np.zeros((nb.value, 4), dtype=np.uint32)
        check(self.L.crgpu_barcode_summary(self._ctx, library, ptr(raw)), "crgpu_barcode_summary")
        rank = np.zeros(nb.value, dtype=np.uint32)
        check(self.L.crgpu_matrix_get(self._ctx, ptr(rank), None, None, None), "crgpu_matrix_get")
        keep = raw[:, 0] > 0
        out = np.zeros(int(keep.sum()), dtype=[("barcode_rank", np.uint32), ("reads", np.uint64), ("umis", np.uint64),
                                               ("candidate_dup_reads", np.uint64), ("umi_corrected_reads", np.uint64)])
        out["barcode_rank"] = rank[keep]
        for i, f in enumerate(("reads", "umis", "candidate_dup_reads", "umi_corrected_reads")):
            out[f] = raw[keep, i]
        return out

    def barcode_correction_metrics(self, library: int = 0) -> dict:
        """InnerBarcodeCorrectionMetrics of one library (cr_lib/src/barcode_correction_metrics.rs:16-39,62-86):
        corrected_bc = corrected reads / all reads, good_bc = (valid before + corrected) / all reads."""
        total = sum(n for lib, n in self._batches if lib == library)
        valid_before = int(self.prior(library).sum(dtype=np.uint64))
        corrected = int(self.corrected_counts(library).sum(dtype=np.uint64))
        frac = (lambda a: a / total) if total else (lambda a: float("nan"))
        return {"total_reads": total, "valid_before": valid_before, "corrected": corrected,
                "corrected_bc": frac(corrected), "good_bc": frac(valid_before + corrected)}

    def barcode_diversity(self, library: int = 0) -> dict:
        """BarcodeDiversityMetrics (cr_lib/src/stages/barcode_correction.rs:428-441): barcodes_detected and
        effective_barcode_diversity = inverse Simpson index of the valid-barcode read counts
        (SimpleHistogram::effective_diversity, metric/src/histogram.rs:161-171)."""
        n, d = C.c_uint64(), C.c_double()
        check(self.L.crgpu_barcode_diversity(self._ctx, library, C.byref(n), C.byref(d)), "crgpu_barcode_diversity")
        return {"barcodes_detected": int(n.value), "effective_barcode_diversity": float(d.value)}

    def write_mex(self, folder: str, software_version: str = "Cell Ranger cellranger_b200", gem_group: int = 1):
        """raw_feature_bc_matrix/{matrix.mtx.gz, barcodes.tsv.gz, features.tsv.gz}
        (MtxWriter, cr_lib/src/stages/write_matrix_market.rs:41-120)."""
        fr = self.feature_reference
        rows = []
        if fr is not None:
            names = getattr(fr, "gene_names", None)
            for g in range(fr.n_genes):
                gid = names[g] if names else f"GENE{g:06d}"
                rows.append(f"{gid}\t{gid}\tGene Expression")
            for fid, ft in zip(fr.fb_ids, fr.fb_types):
                rows.append(f"{fid}\t{fid}\t{FEATURE_TYPE_NAMES.get(ft, 'Custom')}")
        tsv = ("\n".join(rows) + "\n").encode() if rows else None
        check(self.L.crgpu_matrix_write_mex(self._ctx, folder.encode(), software_version.encode(), int(gem_group), tsv),
              "crgpu_matrix_write_mex")

    def molecules(self) -> np.ndarray:
        """UmiCount rows in molecule_info order (by barcode, then library_idx, feature_idx, umi, read_count;
        cr_types/src/types.rs:152-160): (barcode column, library_idx, feature_idx, umi 2-bit, read_count,
        umi_type with 1 = Txomic, 0 = NonTxomic)."""
        n = C.c_uint64()
        check(self.L.crgpu_molecules_count(self._ctx, C.byref(n)))
        out = np.zeros((n.value, 6), dtype=np.uint32)
        check(self.L.crgpu_molecules_get(self._ctx, ptr(out)), "crgpu_molecules_get")
        return out


# molecule_info.h5 column set: names, order and dtypes of MOLECULE_INFO_COLUMNS
# (lib/python/cellranger/molecule_counter.py:90-103; written by cr_h5/src/molecule_info.rs:28-47,972-1033)
MOLECULE_INFO_COLUMNS = (("gem_group", np.uint16), ("barcode_idx", np.uint64), ("feature_idx", np.uint32),
                         ("library_idx", np.uint16), ("umi", np.uint32), ("count", np.uint32), ("umi_type", np.uint32))
UMI_TYPE_TXOMIC = 1  # molecule_counter.py:86; a non-transcriptomic UMI carries 0


def molecule_info_columns(rows: np.ndarray, gem_group: int = 1) -> dict:
    """The UmiCount rows of GemWell.molecules() as the datasets of molecule_info.h5: one array per column of
    MOLECULE_INFO_COLUMNS, in row order (barcode_idx = column of the barcode in the barcode index)."""
    rows = np.asarray(rows)
    src = {"gem_group": np.full(rows.shape[0], gem_group), "barcode_idx": rows[:, 0], "library_idx": rows[:, 1],
           "feature_idx": rows[:, 2], "umi": rows[:, 3], "count": rows[:, 4], "umi_type": rows[:, 5]}
    return {name: src[name].astype(dt) for name, dt in MOLECULE_INFO_COLUMNS}


MAX_READS_BARCODE_COMPATIBILITY = 1_000_000   # check_barcodes_compatibility.rs:79
MIN_BARCODE_SIMILARITY = 0.1                  # lib/bin/parameters.toml


class WhitelistHistogram:
    """A device histogram over the entries of one whitelist (sorted raw sequences), filled by
    sample_valid_barcodes (cr_lib/src/stages/check_barcodes_compatibility.rs:98-120)."""

    def __init__(self, gw: "GemWell", whitelist: int):
        self.gw, self.whitelist = gw, whitelist
        n = C.c_uint64()
        check(gw.L.crgpu_whitelist_entries(gw.ctx, whitelist, C.byref(n)), "crgpu_whitelist_entries")
        self.n = int(n.value)
        p = C.c_void_p()
        check(gw.L.crgpu_dev_alloc(gw.ctx, C.c_uint64(max(self.n, 1) * 4), C.byref(p)), "crgpu_dev_alloc")
        self.dev = p.value
        check(gw.L.crgpu_dev_memset(gw.ctx, C.c_void_p(self.dev), 0, C.c_uint64(self.n * 4)), "crgpu_dev_memset")
        self.reads = 0
        self.reads_in_whitelist = 0

    def observe(self, seqs, bc_offset: int = 0) -> int:
        """seqs: (n, stride) uint8 ASCII records holding the barcode at bc_offset. At most
        MAX_READS_BARCODE_COMPATIBILITY reads are looked at in total, as in the reference. Returns the reads matched."""
        a = ascii_matrix(seqs)
        room = MAX_READS_BARCODE_COMPATIBILITY - self.reads
        a = np.ascontiguousarray(a[:max(room, 0)])
        if a.shape[0] == 0:
            return 0
        m = C.c_uint64()
        check(self.gw.L.crgpu_sample_valid_barcodes(self.gw.ctx, self.whitelist, ptr(a), C.c_uint64(a.shape[0]),
                                                    int(a.shape[1]), int(bc_offset), 0, C.c_void_p(self.dev),
                                                    C.byref(m)), "crgpu_sample_valid_barcodes")
        self.reads += a.shape[0]
        self.reads_in_whitelist += int(m.value)
        return int(m.value)

    def fraction(self) -> float:
        """WhitelistMatchStats::fraction (detect_chemistry/whitelist_filter.rs:61-110): matched / reads with a barcode."""
        return self.reads_in_whitelist / self.reads if self.reads else 0.0

    def counts(self) -> np.ndarray:
        return self.gw.read_device(self.dev, (self.n,), np.uint32)

    def nx(self, fraction: float) -> int:
        out = C.c_uint32()
        check(self.gw.L.crgpu_hist_nx(self.gw.ctx, C.c_void_p(self.dev), C.c_uint64(self.n), C.c_double(fraction),
                                      C.byref(out)), "crgpu_hist_nx")
        return int(out.value)

    def robust_cosine_similarity(self, other: "WhitelistHistogram", translate_whitelist: int = -1) -> float:
        out = C.c_double()
        check(self.gw.L.crgpu_robust_cosine_similarity(self.gw.ctx, C.c_void_p(self.dev), C.c_void_p(other.dev),
                                                       C.c_uint64(self.n), int(translate_whitelist), C.byref(out)),
              "crgpu_robust_cosine_similarity")
        return float(out.value)

    def close(self):
        if self.dev:
            self.gw.L.crgpu_dev_free(self.gw.ctx, C.c_void_p(self.dev))
            self.dev = 0


def check_barcodes_compatibility(gw: "GemWell", plain_whitelist: int, translation_whitelist: Optional[int],
                                 gex_reads, other_reads: dict, bc_offset: int = 0, check_library_compatibility=True,
                                 min_barcode_similarity: float = MIN_BARCODE_SIMILARITY) -> dict:
    """CHECK_BARCODES_COMPATIBILITY's main for two or more library types (check_barcodes_compatibility.rs:161-262):
    gel-bead barcode histograms of the Gene Expression reads and of every other library type's reads against the
    plain whitelist, robust cosine similarity with and without translation, `libraries_to_translate`, and the
    insufficient-overlap error. gex_reads / other_reads[name]: (n, stride) ASCII read arrays (or lists of them)."""
    def hist_of(reads):
        h = WhitelistHistogram(gw, plain_whitelist)
        for part in (reads if isinstance(reads, (list, tuple)) else [reads]):
            h.observe(part, bc_offset)
        return h

    gex = hist_of(gex_reads)
    out = {"libraries_to_translate": [], "similarity": {}, "gex_fraction_in_whitelist": gex.fraction()}
    try:
        for name, reads in other_reads.items():
            h = hist_of(reads)
            try:
                sim = gex.robust_cosine_similarity(h)
                rec = {"without_translation": sim, "with_translation": None, "fraction_in_whitelist": h.fraction()}
                if translation_whitelist is not None:
                    tsim = gex.robust_cosine_similarity(h, translation_whitelist)
                    rec["with_translation"] = tsim
                    if tsim > sim:  # :243-246
                        out["libraries_to_translate"].append(name)
                        sim = tsim
                rec["similarity"] = sim
                out["similarity"][name] = rec
                if check_library_compatibility and not sim >= min_barcode_similarity:  # :248-253
                    raise ValueError(f"Barcodes from the [Gene Expression] library and the [{name}] library have "
                                     f"insufficient overlap (similarity {sim:.4f} < {min_barcode_similarity}).")
            finally:
                h.close()
    finally:
        gex.close()
    return out


class BarcodeCorrector:
    """BarcodeCorrector::new(whitelist, bc_counts, strategy) — barcode/src/corrector.rs:15-71 — in batch
    form: correct_barcodes() takes many invalid (or unchecked) segments at once."""

    def __init__(self, whitelist: Whitelist, bc_counts: Optional[dict] = None,
                 strategy: Optional[Posterior] = None, device: int = 0):
        self.gw = GemWell(device=device, posterior=strategy or Posterior())
        wl = self.gw.add_whitelist(whitelist)
        self.lib = self.gw.add_library(wl, ChemistryDef("segment", 0, whitelist.length, 0, 0))
        self.L = whitelist.length
        content = whitelist.translated if whitelist.translated is not None else whitelist.seqs
        self._content_sorted = np.unique(content, axis=0)
        if bc_counts:
            self.set_counts(bc_counts)

    def set_counts(self, bc_counts: dict):
        """bc_counts: {content sequence: count} (SimpleHistogram<BcSegSeq>)."""
        index = {bytes(s): i for i, s in enumerate(self._content_sorted)}
        prior = np.zeros(len(index), dtype=np.uint32)
        for k, v in bc_counts.items():
            kb = k.encode() if isinstance(k, str) else bytes(k)
            if kb in index:
                prior[index[kb]] = v
        self.gw.set_prior(self.lib, prior)

    def correct_barcodes(self, seqs, quals=None):
        """Returns (corrected ASCII (n, L), state uint8[n]); rows of reads left Invalid keep the input."""
        s = ascii_matrix(seqs, self.L)
        q = None if quals is None else ascii_matrix(quals, self.L)
        n = s.shape[0]
        rank = np.zeros(n, dtype=np.uint32)
        state = np.zeros(n, dtype=np.uint8)
        check(self.gw.L.crgpu_correct_barcodes(self.gw.ctx, self.lib, ptr(s), ptr(q), C.c_uint64(n), ptr(rank),
                                               ptr(state)), "crgpu_correct_barcodes")
        out = s.copy()
        ok = rank != NO_RANK
        if ok.any():
            out[ok] = self.gw.barcode_seqs(rank[ok])
        return out, state

    def close(self):
        self.gw.close()


class SegmentedBarcodeCorrector:
    """BarcodeConstruct<BarcodeCorrector> over a segmented barcode (GelBeadAndProbe, Segmented;
    barcode/src/lib.rs:510-514): every segment has its own whitelist, its own segment counts
    (valid_bc_segment_counts, cr_lib/src/make_shard_metrics.rs:171-188) and is checked and corrected on its own,
    exactly as correct_barcode_in_read walks the segments (cr_lib/src/stages/barcode_correction.rs:88-99); the
    barcode is valid when every segment is (SegmentedBarcode::is_valid, barcode/src/lib.rs:818-823). One
    BarcodeCorrector - one resident whitelist table - per segment."""

    def __init__(self, whitelists: Sequence[Whitelist], segment_counts: Optional[Sequence[Optional[dict]]] = None,
                 strategy: Optional[Posterior] = None, device: int = 0):
        counts = list(segment_counts) if segment_counts is not None else [None] * len(whitelists)
        if len(counts) != len(whitelists):
            raise ValueError("one count histogram (or None) per segment")
        self.segments = [BarcodeCorrector(w, c, strategy, device) for w, c in zip(whitelists, counts)]

    def correct_barcodes(self, segment_seqs: Sequence, segment_quals: Optional[Sequence] = None):
        """segment_seqs[k]: the n sequences of segment k (with segment_quals[k] or None). Returns
        (corrected sequences per segment, states per segment, barcode_valid bool[n])."""
        if len(segment_seqs) != len(self.segments):
            raise ValueError("one array of sequences per segment")
        outs, states = [], []
        for k, seg in enumerate(self.segments):
            q = None if segment_quals is None else segment_quals[k]
            o, st = seg.correct_barcodes(segment_seqs[k], q)
            outs.append(o)
            states.append(st)
        valid = np.ones(states[0].shape[0], dtype=bool)
        for st in states:
            valid &= (st == VALID_BEFORE_CORRECTION) | (st == VALID_AFTER_CORRECTION)
        return outs, states, valid

    def close(self):
        for seg in self.segments:
            seg.close()
