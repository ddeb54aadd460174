"""pyref.py — a second, structurally different restatement of the path, in Python (TEST INFRASTRUCTURE).

The C++ oracle (cr_oracle.cpp) follows the reference's own shape: ASCII keys, hash maps, per-read
`process()`. This module restates the same semantics from the specification in SURVEY.md Appendix A in a
different shape — 2-bit packed integers, content ranks, one sorted key table per GEM well and closed-form
set expressions for the counts — so that an error of reading in one of them shows up as a disagreement
(tests/test_oracle_crosscheck.py). It also serves as the stand-in count engine of the world_size-2 gloo
sharding tests. Pure Python loops: small inputs only. Never imported by the product.

Reference lines restated: lib/rust/barcode/src/corrector.rs:111-171, lib/rust/umi/src/info.rs:20-74,
lib/rust/tx_annotation/src/mark_dups.rs:19-59,87-108,201-363, lib/rust/cr_types/src/types.rs:180-188,
lib/rust/cr_types/src/barcode_index.rs:39-53, lib/rust/cr_h5/src/count_matrix.rs:382-448,
lib/rust/cr_types/src/reference/feature_extraction.rs:34-117, .../feature_checker.rs:8-50.
"""
from __future__ import annotations

import math
from collections import defaultdict

import numpy as np

NO_FEATURE = 0xFFFFFFFF
_CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def pack(seq) -> tuple:
    """ASCII bytes -> (packed int with non-ACGT as 0, bitmask of non-ACGT positions)."""
    v, bad = 0, 0
    for i, c in enumerate(bytes(seq)):
        k = _CODE.get(c)
        if k is None:
            bad |= 1 << i
            k = 0
        v = (v << 2) | k
    return v, bad


def probability(q: int) -> float:
    return math.pow(10.0, -(float(q) - 33.0) / 10.0)


def correct_barcode(lookup: dict, prior, seq: bytes, qual, threshold=0.975, max_expected_errors=None):
    """lookup: packed raw sequence -> content rank. Returns the accepted content rank or None."""
    L = len(seq)
    q, bad = pack(seq)
    nbad = bin(bad).count("1")
    cands = []
    if nbad == 0:
        positions = range(L)
    elif nbad == 1:
        positions = [bad.bit_length() - 1]
    else:
        positions = []
    for pos in positions:
        sh = 2 * (L - 1 - pos)
        orig = None if bad else (q >> sh) & 3
        for b in range(4):
            if b == orig:
                continue
            t = (q & ~(3 << sh)) | (b << sh)
            r = lookup.get(t)
            if r is not None:
                qv = 66 if qual is None else min(qual[pos], 66)
                cands.append((probability(qv) * float(1 + int(prior[r])), r))
    if not cands:
        return None
    total = 0.0
    for lik, _ in cands:  # enumeration order = (position, A C G T)
        total += lik
    best = max(cands)
    if max_expected_errors is not None and qual is not None:
        ee = 0.0
        for x in qual:
            ee += probability(x)
        if not ee < max_expected_errors:
            return None
    return best[1] if best[0] / total >= threshold else None


def umi_valid(seq: bytes, qual: bytes) -> bool:
    if b"N" in seq:
        return False
    if len(set(seq)) <= 1:
        return False
    return all(((x - 33) & 0xFF) >= 10 for x in qual)


def feature_dist(raw, ftype):
    sums = defaultdict(int)
    for c, t in zip(raw, ftype):
        sums[t] += int(c)
    p = [float(int(c)) / float(sums[t]) if sums[t] > 0 else 0.0 for c, t in zip(raw, ftype)]
    if all(x == 0.0 for x in p):
        p = [1.0 / len(p)] * len(p)
    return p


def match_feature(fb_lookup: dict, dist, seq: bytes, qual: bytes, threshold=0.975):
    """fb_lookup: packed capture -> feature index. dist None = exact only."""
    L = len(seq)
    q, bad = pack(seq)
    if not bad and q in fb_lookup:
        return fb_lookup[q]
    if dist is None or bin(bad).count("1") > 1:
        return None
    positions = range(L) if not bad else [bad.bit_length() - 1]
    best, best_f, total = -1.0, None, 0.0
    for pos in positions:
        sh = 2 * (L - 1 - pos)
        orig = None if bad else (q >> sh) & 3
        for b in range(4):
            if b == orig:
                continue
            f = fb_lookup.get((q & ~(3 << sh)) | (b << sh))
            if f is None:
                continue
            qv = min((qual[pos] - 33) & 0xFF, 33)
            lik = dist[f] * math.pow(10.0, -float(qv) / 10.0)
            total += lik
            if lik > best:
                best, best_f = lik, f
    if best_f is None or total == 0.0:
        return None
    return best_f if best / total >= threshold else None


def dedup_count(keys, key_fields, umi_correction, filter_umis=True, on_target=None, target_min_reads=0):
    """keys: iterable of (rank, lib, feature, umi) tuples, one per read entering dedup (any order).
    umi_correction: {lib: bool}. Returns dict with
      table   {(rank, lib, feature, umi): dict(c0, dest, low, c2)}
      entries sorted [(rank, feature, count)]
      molecules sorted [(rank, lib, feature, umi, read_count)]"""
    del key_fields
    c0 = defaultdict(int)
    for k in keys:
        c0[k] += 1
    by_seg = defaultdict(dict)  # (rank, lib, feature) -> {umi: count}
    for (r, l, f, u), c in c0.items():
        by_seg[(r, l, f)][u] = c
    dest = {}
    for (r, l, f), umis in by_seg.items():
        for u, c in umis.items():
            best = (c, u)
            if umi_correction.get(l, True):
                for v, cv in umis.items():
                    x = u ^ v
                    y = (x | (x >> 1)) & 0x5555555555555555
                    if y and not (y & (y - 1)):
                        best = max(best, (cv, v))
            dest[(r, l, f, u)] = best[1]
    c1, c2 = dict(c0), {k: 0 for k in c0}
    for k, d in dest.items():
        r, l, f, u = k
        dk = (r, l, f, d)
        if d != u:
            c1[k] -= 1
            c1[dk] += 1
        c2[dk] += c0[k]
    low = {k: False for k in c0}
    if filter_umis:
        by_umi = defaultdict(list)  # (rank, lib, umi) -> [(feature, c1)]
        for (r, l, f, u), c in c1.items():
            by_umi[(r, l, u)].append((f, c))
        for (r, l, u), lst in by_umi.items():
            m = max(c for _, c in lst)
            tied = sum(1 for _, c in lst if c == m) >= 2
            for f, c in lst:
                low[(r, l, f, u)] = tied or c < m
    targets = {(k[0], k[1], k[2], d) for k, d in dest.items()}

    def filtered(r, l, f, u):  # is_filtered_target_umi, mark_dups.rs:311-320
        return bool(target_min_reads) and on_target is not None and f < len(on_target) and bool(on_target[f]) and \
            c2[(r, l, f, u)] < target_min_reads and not low[(r, l, f, u)]

    mols = sorted((r, l, f, u, c2[(r, l, f, u)]) for (r, l, f, u) in targets
                  if not low[(r, l, f, u)] and not filtered(r, l, f, u))
    ent = defaultdict(int)
    for r, l, f, u, _ in mols:
        ent[(r, f)] += 1
    table = {k: dict(c0=c0[k], dest=dest[k], low=low[(k[0], k[1], k[2], dest[k])], c2=c2[(k[0], k[1], k[2], dest[k])])
             for k in c0}
    return dict(table=table, entries=sorted((r, f, c) for (r, f), c in ent.items()), molecules=mols)


def run_pipeline(whitelists, libraries, feature_type, fb_seqs, batches, threshold=0.975, filter_umis=True):
    """whitelists: list of (raw (n,L) ASCII, translated (n,L) ASCII or None); the first defines the content space.
    libraries: list of dict(wl, bc_off, bc_len, umi_off, umi_len, umi_correction, is_fb, ftype, fb_offset, fb_len)
    batches: list of dict(lib, r1_seq, r1_qual, feature | r2_seq, r2_qual)."""
    def rows(a):
        return [bytes(x) for x in np.asarray(a)]

    content = sorted({pack(s)[0] for s in rows(whitelists[0][1] if whitelists[0][1] is not None else whitelists[0][0])})
    crank = {c: i for i, c in enumerate(content)}
    lookups = []
    for raw, tr in whitelists:
        src = rows(raw)
        dst = rows(tr) if tr is not None else src
        lookups.append({pack(a)[0]: crank[pack(b)[0]] for a, b in zip(src, dst)})
    n_content = len(content)
    n_libs = len(libraries)
    prior = [np.zeros(n_content, dtype=np.int64) for _ in range(n_libs)]
    corrected = [np.zeros(n_content, dtype=np.int64) for _ in range(n_libs)]
    fb_lookup = []
    for lib in libraries:
        d = {}
        if lib["is_fb"]:
            for f, t in enumerate(feature_type):
                if t == lib["ftype"]:
                    d[pack(bytes(fb_seqs[f][: lib["fb_len"]]))[0]] = f
        fb_lookup.append(d)
    fb_counts = [0] * len(feature_type)
    reads = []  # per read dict
    for b in batches:
        lib = libraries[b["lib"]]
        lk = lookups[lib["wl"]]
        s1, q1 = rows(b["r1_seq"]), rows(b["r1_qual"])
        for i in range(len(s1)):
            bc = s1[i][lib["bc_off"]: lib["bc_off"] + lib["bc_len"]]
            bq = q1[i][lib["bc_off"]: lib["bc_off"] + lib["bc_len"]]
            um = s1[i][lib["umi_off"]: lib["umi_off"] + lib["umi_len"]]
            uq = q1[i][lib["umi_off"]: lib["umi_off"] + lib["umi_len"]]
            q, bad = pack(bc)
            r = None if bad else lk.get(q)
            rd = dict(lib=b["lib"], bc=bc, bq=bq, rank=r, state=1 if r is not None else 3, umi=um,
                      umi_valid=umi_valid(um, uq), feature=NO_FEATURE)
            if r is not None:
                prior[b["lib"]][r] += 1
            if lib["is_fb"]:
                cap = bytes(b["r2_seq"][i])[lib["fb_offset"]: lib["fb_offset"] + lib["fb_len"]]
                capq = bytes(b["r2_qual"][i])[lib["fb_offset"]: lib["fb_offset"] + lib["fb_len"]]
                rd["cap"], rd["capq"] = cap, capq
                if len(cap) == lib["fb_len"]:
                    f = match_feature(fb_lookup[b["lib"]], None, cap, capq)
                    if f is not None:
                        fb_counts[f] += 1
            else:
                rd["feature"] = int(b["feature"][i])
            reads.append(rd)
    dist = feature_dist(fb_counts, feature_type) if len(feature_type) else []
    for rd in reads:
        lib = libraries[rd["lib"]]
        if rd["state"] == 3:
            r = correct_barcode(lookups[lib["wl"]], prior[rd["lib"]], rd["bc"], rd["bq"], threshold)
            if r is not None:
                rd["rank"], rd["state"] = r, 2
                corrected[rd["lib"]][r] += 1
        if lib["is_fb"] and len(rd["cap"]) == lib["fb_len"]:
            f = match_feature(fb_lookup[rd["lib"]], dist, rd["cap"], rd["capq"])
            rd["feature"] = NO_FEATURE if f is None else f
    keys = []
    for gi, rd in enumerate(reads):
        rd["key"] = None
        if rd["rank"] is not None and rd["umi_valid"] and rd["feature"] != NO_FEATURE:
            rd["key"] = (rd["rank"], rd["lib"], rd["feature"], pack(rd["umi"])[0])
            keys.append(rd["key"])
    dd = dedup_count(keys, None, {i: bool(l["umi_correction"]) for i, l in enumerate(libraries)}, filter_umis)
    table = dd["table"]
    # representative read: smallest read index of raw key rep(D); rep(D) = min(S_D) if S_D else D
    min_read = {}
    for gi, rd in enumerate(reads):
        if rd["key"] is not None and rd["key"] not in min_read:
            min_read[rd["key"]] = gi
    s_d = defaultdict(list)
    for k, row in table.items():
        d = row["dest"]
        if d != k[3]:
            dk = (k[0], k[1], k[2], d)
            if k[3] < d or table[dk]["dest"] != d:
                s_d[dk].append(k[3])
    for gi, rd in enumerate(reads):
        fl = 1 if rd["umi_valid"] else 0
        rd["proc_umi"] = None
        if rd["key"] is not None:
            row = table[rd["key"]]
            d = row["dest"]
            dk = (rd["key"][0], rd["key"][1], rd["key"][2], d)
            rep = min(s_d[dk]) if s_d.get(dk) else d
            is_rep = rep == rd["key"][3] and min_read[rd["key"]] == gi
            fl |= 2 | (4 if d != rd["key"][3] else 0) | (8 if row["low"] else 0) | (16 if (not row["low"] and is_rep) else 0)
            rd["proc_umi"] = d
        rd["flags"] = fl
    valid_ranks = sorted({rd["rank"] for rd in reads if rd["rank"] is not None})
    col = {r: i for i, r in enumerate(valid_ranks)}
    indptr = [0] * (len(valid_ranks) + 1)
    for r, f, c in dd["entries"]:
        indptr[col[r] + 1] += 1
    for i in range(len(valid_ranks)):
        indptr[i + 1] += indptr[i]
    return dict(reads=reads, prior=prior, corrected=corrected, fb_counts=fb_counts, barcode_ranks=valid_ranks,
                content=content, indptr=indptr, indices=[f for _, f, _ in dd["entries"]],
                data=[c for _, _, c in dd["entries"]],
                molecules=[(col[r], l, f, u, c) for r, l, f, u, c in dd["molecules"]])


def barcode_summary(bc_ascii, state, flags):
    """BarcodeSummary::observe (cr_lib/src/aligner.rs:33-68) over the reads of ONE library type, as driven by
    visit_read_annotation (cr_lib/src/align_metrics.rs:705-721): one row per valid barcode, in barcode order.

    bc_ascii: (n, L) uint8 processed barcode of every read; state: BarcodeSegmentState per read (1 = valid
    before, 2 = valid after correction); flags: per-read DupInfo bits as cro_get_reads returns them
    (bit1 dup_info is Some, bit2 is_corrected, bit3 is_low_support_umi, bit4 is_umi_count).
    Returns (barcodes (k, L) uint8 sorted, reads, umis, candidate_dup_reads, umi_corrected_reads)."""
    valid = (state == 1) | (state == 2)
    if not valid.any():
        z = np.zeros(0, dtype=np.int64)
        return bc_ascii[:0], z, z, z, z
    uniq, inv = np.unique(bc_ascii[valid], axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    f = flags[valid]
    has = (f & 2) != 0
    k = len(uniq)
    reads = np.bincount(inv, minlength=k)                                   # self.reads += 1
    cand = np.bincount(inv[has & ((f & 8) == 0)], minlength=k)              # !dup_info.is_low_support_umi
    corr = np.bincount(inv[has & ((f & 4) != 0)], minlength=k)              # dup_info.is_corrected
    umis = np.bincount(inv[has & ((f & 16) != 0)], minlength=k)             # dup_info.is_umi_count()
    return uniq, reads, umis, cand, corr
