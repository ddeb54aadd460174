"""Deterministic synthetic reads of the shapes BASELINE.json names.

The generator is counter based and integer only: every random draw of read
`gi` is splitmix64(seed_mix + gi*64 + k), every categorical draw is a binary
search in a uint32 threshold table built once on the host. The same function is
implemented twice — here in numpy (tests, CPU baselines) and in
csrc/synth.cuh (bench at 200 M reads, generated straight into HBM) — and
tests/test_gpu_parity.py checks that the two produce identical bytes.

Read layout (what the hot path consumes; alignment happens upstream and is out
of scope, so the gene assignment is part of the input):
  r1_seq  u8[n, r1_len]   ASCII bases of R1: barcode at [0:16], UMI at [16:16+L]
  r1_qual u8[n, r1_len]   ASCII Phred+33
  feature u32[n]          gene index, 0xFFFFFFFF = not confidently mapped
  (feature-barcode libraries)  r2_seq/r2_qual u8[n, r2_len] with the feature
  barcode at [fb_offset : fb_offset+fb_len]
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

NO_FEATURE = 0xFFFFFFFF
MASK64 = (1 << 64) - 1
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _mix_seed(seed: int, stream: int) -> np.uint64:
    s = (seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & MASK64
    return splitmix64(np.array([s], dtype=np.uint64))[0]


def pack_2bit(seqs: np.ndarray) -> np.ndarray:
    """(n, L) ASCII ACGT -> uint64 packed, first base most significant (A0 C1 G2 T3)."""
    lut = np.zeros(256, dtype=np.uint64)
    lut[ord("C")] = 1
    lut[ord("G")] = 2
    lut[ord("T")] = 3
    out = np.zeros(seqs.shape[0], dtype=np.uint64)
    for p in range(seqs.shape[1]):
        out = (out << np.uint64(2)) | lut[seqs[:, p]]
    return out


def unpack_2bit(packed: np.ndarray, L: int) -> np.ndarray:
    """uint packed -> (n, L) ASCII."""
    packed = packed.astype(np.uint64)
    out = np.zeros((packed.shape[0], L), dtype=np.uint8)
    for p in range(L):
        out[:, p] = _BASES[((packed >> np.uint64(2 * (L - 1 - p))) & np.uint64(3)).astype(np.int64)]
    return out


def _thr32(p: float) -> int:
    return int(min(max(p, 0.0), 1.0) * 4294967296.0) if p < 1.0 else 0xFFFFFFFF


def _cdf_u32(weights: np.ndarray) -> np.ndarray:
    """weights -> uint32 upper thresholds; draw u32 `u` picks first index with thr > u."""
    c = np.cumsum(weights.astype(np.float64))
    c = c / c[-1]
    t = np.minimum(np.floor(c * 4294967296.0), 4294967295.0).astype(np.uint64)
    t[-1] = 0xFFFFFFFF
    return t.astype(np.uint32)


@dataclass
class SynthConfig:
    name: str = "cfg1"
    seed: int = 0xC3110001
    wl_seed: int = 737
    n_whitelist: int = 737_280
    bc_len: int = 16
    umi_len: int = 10
    n_cells: int = 3000
    n_genes: int = 30_000
    zipf_s: float = 1.1
    top_genes: int = 0  # >0: expression restricted to this many genes (cfg5)
    cell_sigma: float = 1.0
    dup_factor: float = 4.0
    reads_per_cell_hint: float = 0.0  # 0 → n_reads*(1-ambient)/n_cells
    ambient_frac: float = 0.10
    unmapped_frac: float = 0.15
    bc_err: float = 0.01
    umi_err: float = 0.01
    n_frac: float = 0.001
    # quality classes: (phred, probability); error bases draw from err_quals
    quals: tuple = ((37, 0.85), (25, 0.10), (11, 0.05))
    err_quals: tuple = ((25, 0.6), (11, 0.4))
    n_qual: int = 2
    umi_homopolymer_frac: float = 0.0
    # feature-barcode library (cfg4)
    fb_frac: float = 0.0
    fb_len: int = 15
    fb_offset: int = 10
    n_fb_features: int = 0
    fb_err: float = 0.01

    @property
    def r1_len(self) -> int:
        return self.bc_len + self.umi_len

    @property
    def r2_len(self) -> int:
        return self.fb_offset + self.fb_len


def preset(name: str, n_reads: int) -> SynthConfig:
    """The five BASELINE.json configs, scalable in read count (cells scale with reads
    so reads/cell stays at the full-size value)."""
    def scaled(cells_full, reads_full):
        return max(8, int(round(cells_full * n_reads / reads_full)))

    if name == "cfg1":  # 3' v2, 737K whitelist
        return SynthConfig(name=name, seed=0xC3110001, wl_seed=737, n_whitelist=737_280, umi_len=10,
                           n_cells=scaled(3000, 1_000_000))
    if name == "cfg2":  # 3' v3, 3M whitelist, 200 M reads
        return SynthConfig(name=name, seed=0xC3110002, wl_seed=3018, n_whitelist=6_794_880, umi_len=12,
                           n_cells=scaled(10_000, 200_000_000))
    if name == "cfg3":  # NovaSeq S4 lane scale, sharded
        return SynthConfig(name=name, seed=0xC3110003, wl_seed=3018, n_whitelist=6_794_880, umi_len=12,
                           n_cells=scaled(80_000, 1_600_000_000))
    if name == "cfg4":  # GEX + antibody capture
        return SynthConfig(name=name, seed=0xC3110004, wl_seed=3018, n_whitelist=6_794_880, umi_len=12,
                           n_cells=scaled(10_000, 200_000_000), fb_frac=0.2, n_fb_features=140)
    if name == "cfg5":  # high-error stress
        return SynthConfig(name=name, seed=0xC3110005, wl_seed=3018, n_whitelist=6_794_880, umi_len=12,
                           n_cells=scaled(2000, 200_000_000), bc_err=0.05, top_genes=200, dup_factor=1.0,
                           quals=((37, 0.45), (25, 0.15), (20, 0.10), (15, 0.15), (11, 0.10), (10, 0.05)),
                           err_quals=((20, 0.3), (15, 0.3), (11, 0.3), (8, 0.1)),
                           umi_homopolymer_frac=0.01)
    raise ValueError(name)


@dataclass
class SynthTables:
    cfg: SynthConfig
    whitelist: np.ndarray  # (W, L) ASCII, sorted
    wl_packed: np.ndarray  # uint32[W] sorted
    cell_rank: np.ndarray  # uint32[n_cells]
    cell_cdf: np.ndarray  # uint32[n_cells]
    n_mol: np.ndarray  # uint32[n_cells]
    gene_cdf: np.ndarray  # uint32[n_genes]
    qual_thr: np.ndarray  # uint16-range thresholds as uint32[k]
    qual_val: np.ndarray  # uint8[k] ASCII
    equal_thr: np.ndarray
    equal_val: np.ndarray
    fb_seqs: np.ndarray = field(default_factory=lambda: np.zeros((0, 15), dtype=np.uint8))
    fb_cdf: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.uint32))
    trans: np.ndarray | None = None  # (W, L) FB-library raw sequences (translation whitelist), or None
    trans_packed: np.ndarray | None = None
    n_reads: int = 0


def make_whitelist(n: int, L: int, seed: int) -> np.ndarray:
    """n distinct random L-mers as sorted packed integers."""
    assert 2 * L <= 32
    space = 1 << (2 * L)
    assert n <= space
    if n > space // 2:
        rng = np.random.Generator(np.random.PCG64(seed))
        return np.sort(rng.permutation(space)[:n]).astype(np.uint32)
    got = np.zeros(0, dtype=np.uint64)
    ctr = 0
    s = _mix_seed(seed, 1)
    while got.shape[0] < n:
        m = int((n - got.shape[0]) * 1.05) + 1024
        x = splitmix64(s + np.arange(ctr, ctr + m, dtype=np.uint64)) & np.uint64(space - 1)
        ctr += m
        got = np.unique(np.concatenate([got, x]))
    if got.shape[0] > n:
        # drop a deterministic subset so exactly n remain
        drop = splitmix64(_mix_seed(seed, 2) + got)
        keep = np.argsort(drop, kind="stable")[:n]
        got = np.sort(got[keep])
    return got.astype(np.uint32)


def _qual_table(classes):
    probs = np.array([p for _, p in classes], dtype=np.float64)
    c = np.cumsum(probs) / probs.sum()
    thr = np.minimum(np.floor(c * 65536.0), 65535).astype(np.uint32)
    thr[-1] = 0xFFFF
    val = np.array([q + 33 for q, _ in classes], dtype=np.uint8)
    return thr, val


def make_tables(cfg: SynthConfig, n_reads: int) -> SynthTables:
    wl_packed = make_whitelist(cfg.n_whitelist, cfg.bc_len, cfg.wl_seed)
    whitelist = unpack_2bit(wl_packed, cfg.bc_len)
    rng = np.random.Generator(np.random.PCG64(cfg.seed))
    cell_rank = np.sort(rng.choice(cfg.n_whitelist, size=min(cfg.n_cells, cfg.n_whitelist), replace=False)).astype(np.uint32)
    n_cells = cell_rank.shape[0]
    w = np.exp(rng.normal(0.0, cfg.cell_sigma, size=n_cells))
    cell_cdf = _cdf_u32(w)
    reads_in_cells = n_reads * (1.0 - cfg.ambient_frac)
    exp_reads = w / w.sum() * reads_in_cells
    n_mol = np.maximum(1, np.round(exp_reads / cfg.dup_factor)).astype(np.uint32)
    ng = cfg.n_genes
    ranks = np.arange(1, ng + 1, dtype=np.float64)
    gw = ranks ** (-cfg.zipf_s)
    if cfg.top_genes:
        gw[cfg.top_genes:] = 0.0
    perm = rng.permutation(ng)  # expression rank is not the gene index
    gene_w = np.zeros(ng)
    gene_w[perm] = gw
    gene_cdf = _cdf_u32(gene_w + 1e-300)
    qt, qv = _qual_table(cfg.quals)
    et, ev = _qual_table(cfg.err_quals)
    t = SynthTables(cfg, whitelist, wl_packed, cell_rank, cell_cdf, n_mol, gene_cdf, qt, qv, et, ev, n_reads=n_reads)
    if cfg.n_fb_features:
        # the 12 real CellPlex CMO sequences (lib/python/cellranger/feature/multiplexing/cmo_sets/
        # SC3P_CellPlex_SetA.csv:2-13) plus random 15-mers at pairwise Hamming distance >= 3
        cmo = ["ATGAGGAATTCCTGC", "CATGCCAATAGAGCG", "CCGTCGTCCAAGCAT", "AACGTTAATCACTCA", "CGCGATATGGTCGGA",
               "AAGATGAGGTCTGTG", "AAGCTCGTTGGAAGA", "CGGATTCCACATCAT", "GTTGATCTATAACAG", "GCAGGAGGTATCAAT",
               "GAATCGTGATTCTTC", "ACATGGTCAACGCTG"]
        seqs = [np.frombuffer(s.encode(), dtype=np.uint8) for s in cmo if len(s) == cfg.fb_len][: cfg.n_fb_features]
        while len(seqs) < cfg.n_fb_features:
            cand = _BASES[rng.integers(0, 4, size=cfg.fb_len)]
            if all(int((cand != s).sum()) >= 3 for s in seqs):
                seqs.append(cand)
        t.fb_seqs = np.stack(seqs).astype(np.uint8)
        fw = np.arange(1, cfg.n_fb_features + 1, dtype=np.float64) ** -0.7
        t.fb_cdf = _cdf_u32(fw)
        # translation whitelist for the FB library: raw = a permutation pairing of the entries
        pairing = rng.permutation(cfg.n_whitelist)
        t.trans_packed = wl_packed[pairing]  # raw sequence whose translation is whitelist[i]
        t.trans = unpack_2bit(t.trans_packed, cfg.bc_len)
    return t


def _draw(seed_mix: np.uint64, gi: np.ndarray, k: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(seed_mix + gi * np.uint64(64) + np.uint64(k))


def _pick(cdf: np.ndarray, u32: np.ndarray) -> np.ndarray:
    return np.minimum(np.searchsorted(cdf, u32.astype(np.uint32), side="right"), cdf.shape[0] - 1)


def generate_reads(t: SynthTables, start: int, n: int, library: str = "gex") -> dict:
    """Reads [start, start+n) of the given library ('gex' or 'fb'). Returns numpy arrays."""
    cfg = t.cfg
    Lb, Lu = cfg.bc_len, cfg.umi_len
    fb = library == "fb"
    sm = _mix_seed(cfg.seed, 11 if fb else 7)
    sm_mol = _mix_seed(cfg.seed, 13)  # molecules are shared by both libraries of a cell
    gi = np.arange(start, start + n, dtype=np.uint64)
    W = np.uint64(cfg.n_whitelist)
    n_cells = t.cell_rank.shape[0]

    w0 = _draw(sm, gi, 0)
    ambient = (w0 & np.uint64(0xFFFFFFFF)) < np.uint64(_thr32(cfg.ambient_frac))
    cell = _pick(t.cell_cdf, w0 >> np.uint64(32))
    w1 = _draw(sm, gi, 1)
    amb_rank = (w1 >> np.uint64(11)) % W
    rank = np.where(ambient, amb_rank, t.cell_rank[cell].astype(np.uint64))
    cid = np.where(ambient, np.uint64(n_cells) + amb_rank, cell.astype(np.uint64))
    w2 = _draw(sm, gi, 2)
    nm = np.where(ambient, np.uint64(1 << 20), t.n_mol[cell].astype(np.uint64))
    mol = (w2 >> np.uint64(11)) % nm
    with np.errstate(over="ignore"):
        mkey = (cid << np.uint64(27)) + mol + (np.uint64(1 << 26) if fb else np.uint64(0))
        m0 = splitmix64(sm_mol + mkey * np.uint64(4))
        m1 = splitmix64(sm_mol + mkey * np.uint64(4) + np.uint64(1))
    umi = m1 & np.uint64((1 << (2 * Lu)) - 1)
    if cfg.umi_homopolymer_frac > 0:
        homo = ((m1 >> np.uint64(32)) & np.uint64(0xFFFFFFFF)) < np.uint64(_thr32(cfg.umi_homopolymer_frac))
        hb = (m1 >> np.uint64(30)) & np.uint64(3)
        rep = np.uint64(int("01" * Lu, 2))
        umi = np.where(homo, hb * rep, umi)

    out = {}
    if fb:
        feat_i = _pick(t.fb_cdf, m0 >> np.uint64(32))
        out["fb_true"] = (feat_i + cfg.n_genes).astype(np.uint32)
    else:
        gene = _pick(t.gene_cdf, m0 >> np.uint64(32))
        w3 = _draw(sm, gi, 3)
        unmapped = (w3 & np.uint64(0xFFFFFFFF)) < np.uint64(_thr32(cfg.unmapped_frac))
        out["feature"] = np.where(unmapped, np.uint64(NO_FEATURE), gene.astype(np.uint64)).astype(np.uint32)

    # true barcode: the FB library carries the *raw* (pre-translation) sequence
    if fb and t.trans_packed is not None:
        bc = t.trans_packed[rank.astype(np.int64)].astype(np.uint64)
    else:
        bc = t.wl_packed[rank.astype(np.int64)].astype(np.uint64)

    L1 = cfg.r1_len
    seq = np.zeros((n, L1), dtype=np.uint8)
    qual = np.zeros((n, L1), dtype=np.uint8)
    err_thr_bc = np.uint64(_thr32(cfg.bc_err))
    err_thr_umi = np.uint64(_thr32(cfg.umi_err))
    for p in range(L1):
        if p < Lb:
            base = (bc >> np.uint64(2 * (Lb - 1 - p))) & np.uint64(3)
            thr = err_thr_bc
        else:
            q = p - Lb
            base = (umi >> np.uint64(2 * (Lu - 1 - q))) & np.uint64(3)
            thr = err_thr_umi
        wb = _draw(sm, gi, 8 + p)
        err = (wb & np.uint64(0xFFFFFFFF)) < thr
        sub = (base + np.uint64(1) + ((wb >> np.uint64(32)) & np.uint64(0xFF)) % np.uint64(3)) & np.uint64(3)
        base = np.where(err, sub, base)
        uq = ((wb >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.uint32)
        qn = t.qual_val[np.minimum(np.searchsorted(t.qual_thr, uq, side="left"), len(t.qual_val) - 1)]
        qe = t.equal_val[np.minimum(np.searchsorted(t.equal_thr, uq, side="left"), len(t.equal_val) - 1)]
        seq[:, p] = _BASES[base.astype(np.int64)]
        qual[:, p] = np.where(err, qe, qn)
    wn = _draw(sm, gi, 5)
    has_n = (wn & np.uint64(0xFFFFFFFF)) < np.uint64(_thr32(cfg.n_frac))
    npos = ((wn >> np.uint64(32)) % np.uint64(L1)).astype(np.int64)
    idx = np.nonzero(has_n)[0]
    seq[idx, npos[idx]] = ord("N")
    qual[idx, npos[idx]] = cfg.n_qual + 33
    out["r1_seq"] = seq
    out["r1_qual"] = qual
    out["true_rank"] = rank.astype(np.uint32)

    if fb:
        L2 = cfg.r2_len
        r2 = np.zeros((n, L2), dtype=np.uint8)
        q2 = np.zeros((n, L2), dtype=np.uint8)
        fbp = pack_2bit(t.fb_seqs)[feat_i]
        thr = np.uint64(_thr32(cfg.fb_err))
        for p in range(L2):
            wb = _draw(sm, gi, 40 + p)
            if p < cfg.fb_offset:
                base = (wb >> np.uint64(34)) & np.uint64(3)
                err = np.zeros(n, dtype=bool)
            else:
                q = p - cfg.fb_offset
                base = (fbp >> np.uint64(2 * (cfg.fb_len - 1 - q))) & np.uint64(3)
                err = (wb & np.uint64(0xFFFFFFFF)) < thr
                sub = (base + np.uint64(1) + ((wb >> np.uint64(32)) & np.uint64(0xFF)) % np.uint64(3)) & np.uint64(3)
                base = np.where(err, sub, base)
            uq = ((wb >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.uint32)
            qn = t.qual_val[np.minimum(np.searchsorted(t.qual_thr, uq, side="left"), len(t.qual_val) - 1)]
            qe = t.equal_val[np.minimum(np.searchsorted(t.equal_thr, uq, side="left"), len(t.equal_val) - 1)]
            r2[:, p] = _BASES[base.astype(np.int64)]
            q2[:, p] = np.where(err, qe, qn)
        out["r2_seq"] = r2
        out["r2_qual"] = q2
    return out
