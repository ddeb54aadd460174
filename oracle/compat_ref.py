"""CPU restatement of the barcode-compatibility check and the whitelist match rate.

TEST INFRASTRUCTURE, like everything under oracle/: only tests/ may import it. Plain Python over dicts and sets,
following the reference line by line:
  match_to_whitelist        lib/rust/barcode/src/whitelist.rs:526-545
  sample_valid_barcodes     lib/rust/cr_lib/src/stages/check_barcodes_compatibility.rs:98-120
  nx                        lib/rust/stats/src/nx.rs:6-38 (pinned by its tests, :113-131, and doc examples :60-101)
  robust_cosine_similarity  lib/rust/cr_lib/src/stages/check_barcodes_compatibility.rs:122-158
  main (>= 2 library types) lib/rust/cr_lib/src/stages/check_barcodes_compatibility.rs:225-256
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

MAX_READS_BARCODE_COMPATIBILITY = 1_000_000  # check_barcodes_compatibility.rs:79
ROBUST_FRACTION_THRESHOLD = 0.925            # :80


def match_to_whitelist(wl: set, seq: bytes) -> Optional[bytes]:
    if seq in wl:  # :533-535
        return seq
    pos_n = seq.find(b"N")  # :537-539 position of the first N
    if pos_n < 0:
        return None
    for base in b"ACGT":  # :541-544 the first replacement that is on the whitelist
        trial = seq[:pos_n] + bytes([base]) + seq[pos_n + 1:]
        if trial in wl:
            return trial
    return None


def sample_valid_barcodes(wl: set, seqs: Iterable[bytes], hist: Optional[Dict[bytes, int]] = None,
                          max_reads: int = MAX_READS_BARCODE_COMPATIBILITY):
    """Returns (histogram, reads looked at, reads matched)."""
    hist = {} if hist is None else hist
    num_reads = matched = 0
    for seq in seqs:
        bc = match_to_whitelist(wl, seq)  # :109
        if bc is not None:
            hist[bc] = hist.get(bc, 0) + 1  # :110
            matched += 1
        num_reads += 1  # :112
        if num_reads >= max_reads:  # :114
            break
    return hist, num_reads, matched


def nx(items: Iterable[int], fraction: float) -> Optional[int]:
    assert 0.0 < fraction < 1.0  # :11
    owned, s = [], 0.0
    for item in items:
        assert item > 0, "Found a number that is not positive while computing Nx"  # :17-21
        s += float(item)
        owned.append(item)
    owned.sort(reverse=True)  # :25
    cumulative, cutoff = 0.0, s * fraction  # :29-30
    for item in owned:
        cumulative += float(item)
        if cumulative >= cutoff:  # :33
            return item
    return None


def robust_cosine_similarity(c1: Dict[bytes, int], c2: Dict[bytes, int]) -> float:
    thresh1 = nx(c1.values(), ROBUST_FRACTION_THRESHOLD)  # :130-133
    if thresh1 is None:
        return 0.0
    thresh2 = nx(c2.values(), ROBUST_FRACTION_THRESHOLD)  # :134-137
    if thresh2 is None:
        return 0.0
    mag1 = math.sqrt(sum(float(min(c, thresh1) * min(c, thresh1)) for c in c1.values()))  # :139-143
    mag2 = math.sqrt(sum(float(min(c, thresh2) * min(c, thresh2)) for c in c2.values()))  # :145-149
    dot_prod = sum(float(min(c, thresh1) * min(c2.get(bc, 0), thresh2)) for bc, c in c1.items())  # :151-155
    return dot_prod / (mag1 * mag2)  # :157


def map_key(hist: Dict[bytes, int], translate: Dict[bytes, bytes]) -> Dict[bytes, int]:
    out: Dict[bytes, int] = {}
    for k, v in hist.items():  # SimpleHistogram::map_key: observe_by(f(k), count)
        t = translate[k]
        out[t] = out.get(t, 0) + v
    return out


def libraries_to_translate(gex_hist, other_hists: Dict[str, Dict[bytes, int]], translate: Optional[Dict[bytes, bytes]],
                           min_barcode_similarity: float = 0.1, check_library_compatibility: bool = True):
    """:225-256. Returns (set of library names to translate, {name: (similarity, translated similarity)})."""
    to_translate, sims = set(), {}
    for name, this_hist in other_hists.items():
        similarity = robust_cosine_similarity(gex_hist, this_hist)  # :238
        trans_similarity = None
        if translate is not None:  # :240-247
            trans_similarity = robust_cosine_similarity(gex_hist, map_key(this_hist, translate))
            sims[name] = (similarity, trans_similarity)
            if trans_similarity > similarity:
                to_translate.add(name)
                similarity = trans_similarity
        else:
            sims[name] = (similarity, None)
        if check_library_compatibility and not similarity >= min_barcode_similarity:  # :248-253
            raise ValueError(f"insufficient overlap: {name}")
    return to_translate, sims
