// cr_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain restatement, on the host, of the algorithms on Cell Ranger's
// barcode / UMI correction and counting path, written from a reading of the
// reference sources (cited per function as lib/rust/<crate>/src/<file>:<lines>,
// relative to /root/reference). It keeps the reference's structure: ASCII
// sequences as hash-map keys, hash-map priors, 3*L trial sequences per invalid
// barcode with one pow() per hit, per-barcode hash maps of (UMI, gene) counts.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library. The product (cellranger_b200,
// libcrgpu.so) never links, imports or calls it.
//
// Parity pinning: the Rust reference cannot be compiled in this image (no
// cargo/rustc, un-vendored git dependencies), so this restatement is pinned
// against every self-contained known-answer test the reference holds for the
// path (tests/test_oracle_kat.py; vectors in tests/golden/reference_kats.json,
// each citing its reference test) and cross-checked against a structurally
// different Python restatement (oracle/pyref.py). The parts of the path the
// reference has no test for (low-support filter, one-read pre-move,
// representative read, feature_counts, BarcodeIndex, CSC assembly) are pinned
// by code reading only: PARITY UNPINNED for those, as DESIGN.md states.
//
// Build: make -C oracle   (g++ -O2 -std=c++17 -shared -fPIC -pthread)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// BcSegSeq / UmiSeq / feature-barcode sequences as the reference holds them: SSeqGen<N> of fastq_set 0.5.3 - a
// fixed-capacity array of ASCII bytes plus a length byte, stored inline (no heap), Ord = byte-lexicographic,
// Hash over the bytes (lib/rust/barcode/src/lib.rs:33-52, lib/rust/umi/src/lib.rs:12-14). Unused bytes stay zero,
// so equality and hashing run over the whole 32-byte value.
struct Seq {
  static const size_t CAP = 31;
  char b[CAP] = {0};
  uint8_t n = 0;
  Seq() {}
  Seq(const char* p, size_t len) {
    if (len > CAP) {
      fprintf(stderr, "cr_oracle: sequence of %zu bytes exceeds the %zu-byte capacity\n", len, CAP);
      abort();
    }
    n = (uint8_t)len;
    memcpy(b, p, len);
  }
  size_t size() const { return n; }
  const char* data() const { return b; }
  char& operator[](size_t i) { return b[i]; }
  char operator[](size_t i) const { return b[i]; }
  bool operator==(const Seq& o) const { return memcmp(this, &o, sizeof(Seq)) == 0; }
  bool operator!=(const Seq& o) const { return !(*this == o); }
  bool operator<(const Seq& o) const {  // lexicographic over unsigned bytes, the shorter first on a tie
    const int c = memcmp(b, o.b, n < o.n ? n : o.n);
    return c < 0 || (c == 0 && n < o.n);
  }
  bool operator>(const Seq& o) const { return o < *this; }
  bool operator>=(const Seq& o) const { return !(*this < o); }
  bool operator<=(const Seq& o) const { return !(o < *this); }
  const char* begin() const { return b; }
  const char* end() const { return b + n; }
  static const size_t npos = (size_t)-1;
  size_t find(char c) const {
    for (size_t i = 0; i < n; i++)
      if (b[i] == c) return i;
    return npos;
  }
};
static_assert(sizeof(Seq) == 32, "Seq is hashed and compared as 32 bytes");
struct SeqHash {
  size_t operator()(const Seq& s) const {
    uint64_t w[4];
    memcpy(w, &s, 32);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 4; i++) {
      h ^= w[i];
      h *= 0xFF51AFD7ED558CCDull;
      h ^= h >> 32;
    }
    return (size_t)h;
  }
};

// The reference's HashMap / HashSet (TxHashMap = hashbrown behind std, lib/rust/metric/src/lib.rs:61-111): an
// open-addressing table with the entries stored inline - one probe is one cache line, not the two pointer hops of
// std::unordered_map's node lists. The subset of the std interface this file uses; iteration order is arbitrary,
// as it is for the reference's maps.
template <typename K, typename V, typename H>
class FlatMap {
 public:
  struct Entry {
    K first;
    V second;
  };

 private:
  std::vector<Entry> slots_;
  std::vector<uint8_t> used_;
  size_t size_ = 0, mask_ = 0;
  H hash_;

  size_t probe(const K& k) const {  // slot of k, or the free slot where it would go (capacity > 0)
    size_t i = hash_(k) & mask_;
    while (used_[i] && !(slots_[i].first == k)) i = (i + 1) & mask_;
    return i;
  }
  void grow() {
    const size_t cap = slots_.empty() ? 16 : slots_.size() * 2;
    std::vector<Entry> old;
    std::vector<uint8_t> old_used;
    old.swap(slots_);
    old_used.swap(used_);
    slots_.resize(cap);
    used_.assign(cap, 0);
    mask_ = cap - 1;
    for (size_t i = 0; i < old.size(); i++)
      if (old_used[i]) {
        const size_t j = probe(old[i].first);
        slots_[j] = old[i];
        used_[j] = 1;
      }
  }

 public:
  template <bool CONST>
  class Iter {
    typedef typename std::conditional<CONST, const FlatMap, FlatMap>::type Map;
    typedef typename std::conditional<CONST, const Entry, Entry>::type E;
    Map* m_;
    size_t i_;
    void skip() {
      while (i_ < m_->slots_.size() && !m_->used_[i_]) i_++;
    }

   public:
    Iter(Map* m, size_t i) : m_(m), i_(i) { skip(); }
    E& operator*() const { return m_->slots_[i_]; }
    E* operator->() const { return &m_->slots_[i_]; }
    Iter& operator++() {
      i_++;
      skip();
      return *this;
    }
    bool operator==(const Iter& o) const { return i_ == o.i_; }
    bool operator!=(const Iter& o) const { return i_ != o.i_; }
  };
  typedef Iter<false> iterator;
  typedef Iter<true> const_iterator;
  iterator begin() { return iterator(this, 0); }
  iterator end() { return iterator(this, slots_.size()); }
  const_iterator begin() const { return const_iterator(this, 0); }
  const_iterator end() const { return const_iterator(this, slots_.size()); }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  void clear() {
    slots_.clear();
    used_.clear();
    size_ = mask_ = 0;
  }
  void reserve(size_t n) {
    while (slots_.size() < 2 * n + 16) grow();
  }
  iterator find(const K& k) {
    if (slots_.empty()) return end();
    const size_t i = probe(k);
    return used_[i] ? iterator(this, i) : end();
  }
  const_iterator find(const K& k) const {
    if (slots_.empty()) return end();
    const size_t i = probe(k);
    return used_[i] ? const_iterator(this, i) : end();
  }
  size_t count(const K& k) const { return find(k) != end() ? 1 : 0; }
  std::pair<iterator, bool> emplace(const K& k, const V& v) {
    if (2 * (size_ + 1) > slots_.size()) grow();
    const size_t i = probe(k);
    if (used_[i]) return {iterator(this, i), false};
    slots_[i].first = k;
    slots_[i].second = v;
    used_[i] = 1;
    size_++;
    return {iterator(this, i), true};
  }
  void insert(const K& k) { emplace(k, V()); }
  V& operator[](const K& k) { return emplace(k, V()).first->second; }
  V& at(const K& k) {
    auto it = find(k);
    if (it == end()) {
      fprintf(stderr, "cr_oracle: FlatMap::at on a missing key\n");
      abort();
    }
    return it->second;
  }
  const V& at(const K& k) const { return const_cast<FlatMap*>(this)->at(k); }
};

// ---------------------------------------------------------------------------
// Barcode segment state — lib/rust/barcode/src/lib.rs:270-310
// ---------------------------------------------------------------------------
enum BcState : uint8_t {
  NOT_CHECKED = 0,
  VALID_BEFORE_CORRECTION = 1,
  VALID_AFTER_CORRECTION = 2,
  INVALID = 3,
};

static inline bool state_is_valid(uint8_t s) {
  return s == VALID_BEFORE_CORRECTION || s == VALID_AFTER_CORRECTION;
}

// BarcodeSegmentState::change — lib/rust/barcode/src/lib.rs:291-309
static inline uint8_t state_change(uint8_t s, bool in_wl) {
  if (s == NOT_CHECKED) return in_wl ? VALID_BEFORE_CORRECTION : INVALID;
  if (s == INVALID) return in_wl ? VALID_AFTER_CORRECTION : INVALID;
  // the reference panics here; the oracle never reaches it
  return s;
}

// ---------------------------------------------------------------------------
// Whitelist::{Plain,Trans} — lib/rust/barcode/src/whitelist.rs:452-525
// ---------------------------------------------------------------------------
struct Whitelist {
  int L = 0;
  bool is_trans = false;
  FlatMap<Seq, char, SeqHash> plain;  // HashSet
  FlatMap<Seq, Seq, SeqHash> trans;

  // check_and_update: membership; a translation whitelist replaces the content
  // by the translated sequence (whitelist.rs:494-516).
  bool check(const Seq& s, Seq* content) const {
    if (!is_trans) {
      if (plain.count(s)) {
        *content = s;
        return true;
      }
      return false;
    }
    auto it = trans.find(s);
    if (it == trans.end()) return false;
    *content = it->second;
    return true;
  }
  bool contains(const Seq& s) const {
    return is_trans ? trans.count(s) > 0 : plain.count(s) > 0;
  }
};

// SimpleHistogram<K>::get → 0 when absent — lib/rust/metric/src/histogram.rs:26-145
typedef FlatMap<Seq, int64_t, SeqHash> Hist;
static inline int64_t hist_get(const Hist& h, const Seq& k) {
  auto it = h.find(k);
  return it == h.end() ? 0 : it->second;
}

// probability(qual) — lib/rust/barcode/src/corrector.rs:167-171
static inline double probability(uint8_t qual) {
  double q = (double)qual;
  return pow(10.0, -(q - 33.0) / 10.0);
}

const uint8_t BC_MAX_QV = 66;  // corrector.rs:8
const char BASE_OPTS[4] = {'A', 'C', 'G', 'T'};  // corrector.rs:9

// Posterior::correct_barcode — lib/rust/barcode/src/corrector.rs:111-165.
// Returns true and fills *out_content when a correction is accepted.
static bool posterior_correct(const Whitelist& wl, const Hist& bc_counts,
                              const Seq& observed, const uint8_t* qual /*nullable*/,
                              double max_expected_barcode_errors, double threshold,
                              Seq* out_content) {
  Seq a = observed;
  bool have_best = false;
  double best_like = 0.0;
  Seq best_bc;
  double total = 0.0;
  for (size_t pos = 0; pos < a.size(); pos++) {
    uint8_t qv = qual ? std::min(qual[pos], BC_MAX_QV) : BC_MAX_QV;
    char existing = a[pos];
    for (char val : BASE_OPTS) {
      if (val == existing) continue;
      a[pos] = val;
      Seq content;
      if (wl.check(a, &content)) {
        int64_t raw = hist_get(bc_counts, content);
        int64_t c = 1 + raw;  // Laplace smoothing
        double like = probability(qv) * (double)c;
        if (!have_best) {
          have_best = true;
          best_like = like;
          best_bc = content;
        } else {
          // Option<(Of64, BarcodeSegment)>::max: tuple order, ties → larger sequence
          if (like > best_like || (like == best_like && content >= best_bc)) {
            best_like = like;
            best_bc = content;
          }
        }
        total += like;
      }
    }
    a[pos] = existing;
  }
  double expected_errors = 0.0;
  if (qual)
    for (size_t i = 0; i < observed.size(); i++) expected_errors += probability(qual[i]);
  if (have_best) {
    if (expected_errors < max_expected_barcode_errors && best_like / total >= threshold) {
      *out_content = best_bc;
      return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------
// UmiInfo::new — lib/rust/umi/src/info.rs:20-74
// ---------------------------------------------------------------------------
static bool umi_is_valid(const uint8_t* seq, const uint8_t* qual, int L) {
  bool has_n = false;
  for (int i = 0; i < L; i++) has_n |= (seq[i] == 'N');
  bool homopolymer = true;
  for (int i = 1; i < L; i++)
    if (seq[i - 1] != seq[i]) {
      homopolymer = false;
      break;
    }
  bool low_min_qual = false;
  for (int i = 0; i < L; i++)
    if ((uint8_t)(qual[i] - 33) < 10) low_min_qual = true;
  return !(has_n || homopolymer || low_min_qual);
}

// SSeqGen::encode_2bit_u32 (fastq_set 0.5.3, un-vendored): first base most
// significant, A0 C1 G2 T3 — pinned by lib/python/cellranger/utils.py:230-246
// and lib/python/tenkit/seq.py:10-11.
static uint32_t encode_2bit_u32(const Seq& s) {
  uint32_t r = 0;
  for (char c : s) {
    uint32_t v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
    r = (r << 2) | v;
  }
  return r;
}

// ---------------------------------------------------------------------------
// (UMI, gene) keyed maps — lib/rust/tx_annotation/src/mark_dups.rs
// ---------------------------------------------------------------------------
struct UG {
  Seq umi;
  uint32_t gene;
  bool operator==(const UG& o) const { return gene == o.gene && umi == o.umi; }
  bool operator<(const UG& o) const {
    if (umi != o.umi) return umi < o.umi;
    return gene < o.gene;
  }
};
struct UGHash {
  size_t operator()(const UG& k) const {
    return SeqHash()(k.umi) * 1000003u ^ (size_t)k.gene * 0x9E3779B97F4A7C15ull;
  }
};
typedef FlatMap<UG, uint64_t, UGHash> UGCounts;
typedef FlatMap<UG, Seq, UGHash> UGCorr;
typedef FlatMap<UG, char, UGHash> UGSet;

// correct_umis — mark_dups.rs:19-59
static UGCorr correct_umis(const UGCounts& counts) {
  static const char nucs[4] = {'A', 'C', 'G', 'T'};
  UGCorr corrections;
  for (const auto& kv : counts) {
    const Seq& umi = kv.first.umi;
    uint32_t gene = kv.first.gene;
    Seq test = umi;
    uint64_t best_count = kv.second;
    Seq best_umi = umi;
    for (size_t pos = 0; pos < umi.size(); pos++) {
      for (char c : nucs) {
        if (c == umi[pos]) continue;
        test[pos] = c;
        auto it = counts.find(UG{test, gene});
        uint64_t tc = it == counts.end() ? 0 : it->second;
        if (tc > best_count || (tc == best_count && test > best_umi)) {
          best_umi = test;
          best_count = tc;
        }
      }
      test[pos] = umi[pos];
    }
    if (umi != best_umi) corrections.emplace(UG{umi, gene}, best_umi);
  }
  return corrections;
}

// determine_low_support_umigenes — mark_dups.rs:87-108
static UGSet determine_low_support(const UGCounts& counts) {
  UGSet low;
  struct Row {
    Seq umi;
    uint32_t gene;
    uint64_t count;
  };
  std::vector<Row> v;
  v.reserve(counts.size());
  for (const auto& kv : counts) v.push_back(Row{kv.first.umi, kv.first.gene, kv.second});
  std::sort(v.begin(), v.end(), [](const Row& a, const Row& b) {
    if (a.umi != b.umi) return a.umi < b.umi;
    if (a.gene != b.gene) return a.gene < b.gene;
    return a.count < b.count;
  });
  size_t i = 0;
  while (i < v.size()) {
    size_t j = i;
    while (j < v.size() && v[j].umi == v[i].umi) j++;
    uint64_t mx = 0;
    for (size_t k = i; k < j; k++) mx = std::max(mx, v[k].count);
    size_t n_at_max = 0;
    for (size_t k = i; k < j; k++) n_at_max += (v[k].count == mx);
    bool tied = n_at_max >= 2;
    for (size_t k = i; k < j; k++)
      if (tied || v[k].count < mx) low.insert(UG{v[k].umi, v[k].gene});
    i = j;
  }
  return low;
}

// UmiSelectKey{utype, qname} (mark_dups.rs:110-152) as the order-preserving word Ctx::select_key holds per read
// (without caller-supplied keys: every read Txomic, the qname ordered like the global read index).
typedef FlatMap<UG, uint64_t, UGHash> UGMinKey;

// BarcodeDupMarker — mark_dups.rs:183-364
struct DupMarker {
  UGCounts counts;
  UGSet low_support;
  UGCorr corrections;
  UGMinKey min_key;

  // BarcodeDupMarker::new — mark_dups.rs:201-277
  void build(bool filter_umis, bool umi_correction) {
    if (umi_correction) corrections = correct_umis(counts);
    struct Move {
      UG raw, corr;
      uint64_t raw_count;
    };
    std::vector<Move> moves;
    for (const auto& kv : corrections)
      moves.push_back(Move{kv.first, UG{kv.second, kv.first.gene}, counts.at(kv.first)});
    for (const auto& m : moves) {  // one read first (:226-232)
      counts.at(m.raw) -= 1;
      counts.at(m.corr) += 1;
    }
    if (filter_umis) low_support = determine_low_support(counts);
    for (const auto& m : moves) {  // the rest (:241-246)
      counts.at(m.raw) -= m.raw_count - 1;
      counts.at(m.corr) += m.raw_count - 1;
    }
    // lowest raw UMI that would be corrected onto (corr, gene) and could be
    // the UMI count (:248-268)
    UGCorr min_raw;
    for (const auto& kv : corrections) {
      const Seq& raw_seq = kv.first.umi;
      uint32_t gene = kv.first.gene;
      const Seq& corr_seq = kv.second;
      if (raw_seq < corr_seq || corrections.count(UG{corr_seq, gene})) {
        UG ck{corr_seq, gene};
        auto it = min_raw.find(ck);
        if (it == min_raw.end())
          min_raw.emplace(ck, raw_seq);
        else if (raw_seq < it->second)
          it->second = raw_seq;
      }
    }
    std::vector<std::pair<UG, uint64_t>> upd;
    for (const auto& kv : min_raw)
      upd.emplace_back(kv.first, min_key.at(UG{kv.second, kv.first.gene}));
    for (const auto& u : upd) min_key[u.first] = u.second;
  }
};

// ---------------------------------------------------------------------------
// Feature-barcode correction — lib/rust/cr_types/src/reference/feature_extraction.rs:34-117
// for a single tethered capture (one (seq, qual) candidate).
// ---------------------------------------------------------------------------
const uint8_t FEATURE_MAX_QV = 33;
const double FEATURE_CONF_THRESHOLD = 0.975;

struct FeaturePattern {
  int offset = 0;  // capture starts at R2[offset]
  int len = 0;
  FlatMap<Seq, int, SeqHash> features;  // sequence → feature index
};

static bool correct_feature_barcode(const FeaturePattern& pat, const std::vector<double>& feat_dist,
                                    const Seq& seq, const uint8_t* qual, Seq* hit) {
  auto check = [&](const Seq& s, int i, double* like) -> bool {
    auto it = pat.features.find(s);
    if (it == pat.features.end()) return false;
    double p_wl = feat_dist[it->second];
    if (i >= 0) {
      double qv = (double)std::min((uint8_t)(qual[i] - 33), FEATURE_MAX_QV);
      double p_edit = pow(10.0, -qv / 10.0);
      *like = p_wl * p_edit;
    } else {
      *like = p_wl;
    }
    return true;
  };
  // whitelist_likelihoods for a single candidate: every test sequence is
  // distinct, so each hit is a vacant insert and the sum accumulates in
  // enumeration order (:62-98).
  std::vector<std::pair<Seq, double>> hits;
  double sum = 0.0;
  double like;
  if (check(seq, -1, &like)) {
    hits.emplace_back(seq, like);
    sum += like;
  } else {
    Seq t = seq;
    for (size_t i = 0; i < t.size(); i++) {
      char orig = t[i];
      for (char b : BASE_OPTS) {
        if (b != orig) {
          t[i] = b;
          if (check(t, (int)i, &like)) {
            hits.emplace_back(t, like);
            sum += like;
          }
        }
      }
      t[i] = orig;
    }
  }
  double max_like = -1.0;
  Seq best;
  for (const auto& h : hits)  // HashMap iteration order is arbitrary; ties cannot pass 0.975
    if (h.second > max_like) {
      max_like = h.second;
      best = h.first;
    }
  if (max_like / sum >= FEATURE_CONF_THRESHOLD) {
    *hit = best;
    return true;
  }
  return false;
}

// FeatureExtractor::find_closest for one tethered capture — feature_extraction.rs:447-471
static int find_closest(const FeaturePattern& pat, const std::vector<double>* feat_dist,
                        const Seq& bc, const uint8_t* qual) {
  auto it = pat.features.find(bc);
  if (it != pat.features.end()) return it->second;
  if (feat_dist) {
    Seq hit;
    if (correct_feature_barcode(pat, *feat_dist, bc, qual, &hit)) return pat.features.at(hit);
  }
  return -1;
}

// compute_feature_dist — lib/rust/cr_types/src/reference/feature_checker.rs:8-50
static std::vector<double> compute_feature_dist(const std::vector<int64_t>& raw,
                                                const std::vector<int>& ftype) {
  std::map<int, int64_t> sums;
  for (size_t i = 0; i < raw.size(); i++) sums[ftype[i]] += raw[i];
  std::vector<double> p(raw.size(), 0.0);
  for (size_t i = 0; i < raw.size(); i++) {
    int64_t s = sums[ftype[i]];
    if (s > 0) p[i] = (double)raw[i] / (double)s;
  }
  bool all_zero = true;
  for (double x : p) all_zero &= (x == 0.0);
  if (all_zero)
    for (double& x : p) x = 1.0 / (double)p.size();
  return p;
}

// ---------------------------------------------------------------------------
// Pipeline state
// ---------------------------------------------------------------------------
struct Library {
  int wl = 0;
  int bc_off = 0, bc_len = 16, umi_off = 16, umi_len = 12;
  bool umi_correction = true;  // off for Multiplexing Capture, lib/rust/cr_lib/src/aligner.rs:313-318
  bool is_fb = false;
  int ftype = 0;  // feature type id owning this library's features
  FeaturePattern pat;
  Hist prior;             // valid_bc_segment_counts (make_shard_metrics.rs:171-187)
  Hist corrected_counts;  // bc_counts_corrected (barcode_correction.rs:335-340)
};

struct Batch {
  int lib = 0;
  uint64_t n = 0;
  int r1_len = 0, r2_len = 0;
  const uint8_t *r1_seq = nullptr, *r1_qual = nullptr, *r2_seq = nullptr, *r2_qual = nullptr;
  const uint32_t* feature = nullptr;
  uint64_t base = 0;  // global read index of read 0
};

const uint32_t NO_FEATURE = 0xFFFFFFFFu;

struct ReadOut {
  uint8_t bc_state = NOT_CHECKED;
  uint8_t umi_valid = 0;
  uint8_t has_dup = 0, is_corrected = 0, is_low_support = 0, is_umi_count = 0, is_filtered_target = 0;
  uint32_t feature = NO_FEATURE;
  uint32_t read_count = 0;
};

struct Molecule {
  uint32_t bc_idx, lib, feature, umi, read_count, utype;  // utype as UmiType::to_u32: 1 Txomic, 0 NonTxomic
};

struct Ctx {
  std::vector<Whitelist> wls;
  std::vector<Library> libs;
  int n_features = 0;
  std::vector<int> feature_type;  // per feature: owning feature-type id (0 = gene)
  std::vector<Batch> batches;
  uint64_t n_reads = 0;
  // UmiSelectKey per read (mark_dups.rs:110-114) as one order-preserving word: bit 63 = UmiType (0 Txomic <
  // 1 NonTxomic, the derive(Ord) order of umi/src/lib.rs:101-107), low bits = the rank of the qname. Empty: every
  // read is Txomic and its qname orders like its global read index.
  std::vector<uint64_t> select_key;
  double threshold = 0.975;
  double max_expected_errors = 1.7976931348623157e308;  // f64::MAX, corrector.rs:104-106
  bool filter_umis = true;                              // lib/rust/cr_lib/src/aligner.rs:270
  // targeted_umi_min_read_count (None = 0) and the panel's target set (mark_dups.rs:189-191,311-320)
  uint64_t target_min_reads = 0;
  std::vector<uint8_t> on_target;

  // outputs
  std::vector<ReadOut> out;
  std::vector<Seq> bc_content;  // per read content (translated) sequence
  std::vector<Seq> umi_out;     // per read processed UMI (raw if no DupInfo)
  std::vector<int64_t> fb_exact_counts;
  std::vector<double> feat_dist;
  std::vector<Seq> barcodes;  // BarcodeIndex: sorted unique valid barcodes
  std::vector<int64_t> indptr;
  std::vector<uint32_t> indices;
  std::vector<int32_t> data;
  std::vector<Molecule> molecules;
  uint64_t n_valid_before = 0, n_corrected = 0, n_invalid = 0;
  uint64_t n_umi_corrected_reads = 0, n_low_support_reads = 0, n_umis = 0, n_dup_reads = 0;
};

static inline Seq bytes(const uint8_t* p, int n) { return Seq((const char*)p, (size_t)n); }

template <typename F>
static void parallel_for(uint64_t n, int threads, F f) {
  if (threads <= 1 || n < 4096) {
    f(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  uint64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    uint64_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
    th.emplace_back([=] { f(t, lo, hi); });
  }
  for (auto& x : th) x.join();
}

// Pass 1 = MAKE_SHARD: exact whitelist check + priors + exact feature-barcode
// counts — lib/rust/cr_types/src/rna_read.rs:356-365, lib/rust/cr_lib/src/make_shard_metrics.rs:171-187,337-345
static void pass1(Ctx& c, int threads) {
  c.out.assign(c.n_reads, ReadOut());
  c.bc_content.assign(c.n_reads, Seq());
  c.umi_out.assign(c.n_reads, Seq());
  c.fb_exact_counts.assign(c.n_features, 0);
  for (auto& b : c.batches) {
    Library& lib = c.libs[b.lib];
    const Whitelist& wl = c.wls[lib.wl];
    std::vector<Hist> hp(std::max(threads, 1));
    std::vector<std::vector<int64_t>> fp(std::max(threads, 1), std::vector<int64_t>(c.n_features, 0));
    parallel_for(b.n, threads, [&](int t, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        uint64_t g = b.base + i;
        const uint8_t* s = b.r1_seq + i * b.r1_len;
        const uint8_t* q = b.r1_qual + i * b.r1_len;
        Seq raw = bytes(s + lib.bc_off, lib.bc_len);
        Seq content;
        bool hit = wl.check(raw, &content);
        c.out[g].bc_state = state_change(NOT_CHECKED, hit);
        c.bc_content[g] = hit ? content : raw;
        if (hit) hp[t][content] += 1;
        c.out[g].umi_valid = umi_is_valid(s + lib.umi_off, q + lib.umi_off, lib.umi_len);
        c.umi_out[g] = bytes(s + lib.umi_off, lib.umi_len);
        if (lib.is_fb) {
          // MAKE_SHARD extractor has feature_dist = None: exact captures only
          int f = -1;
          if (b.r2_len >= lib.pat.offset + lib.pat.len)
            f = find_closest(lib.pat, nullptr, bytes(b.r2_seq + i * b.r2_len + lib.pat.offset, lib.pat.len),
                             b.r2_qual + i * b.r2_len + lib.pat.offset);
          if (f >= 0) fp[t][f] += 1;
        } else {
          c.out[g].feature = b.feature ? b.feature[i] : NO_FEATURE;
        }
      }
    });
    for (auto& h : hp)
      for (auto& kv : h) lib.prior[kv.first] += kv.second;
    for (auto& v : fp)
      for (int f = 0; f < c.n_features; f++) c.fb_exact_counts[f] += v[f];
  }
  c.feat_dist = compute_feature_dist(c.fb_exact_counts, c.feature_type);
}

// Pass 2 = BARCODE_CORRECTION over invalid reads — lib/rust/cr_lib/src/stages/barcode_correction.rs:76-99,327-345
// plus the ALIGN_AND_COUNT feature-barcode extraction with feat_dist
// (lib/rust/cr_lib/src/aligner.rs:477-515,624-630).
static void pass2(Ctx& c, int threads) {
  for (auto& b : c.batches) {
    Library& lib = c.libs[b.lib];
    const Whitelist& wl = c.wls[lib.wl];
    std::vector<Hist> hc(std::max(threads, 1));
    parallel_for(b.n, threads, [&](int t, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; i++) {
        uint64_t g = b.base + i;
        const uint8_t* s = b.r1_seq + i * b.r1_len;
        const uint8_t* q = b.r1_qual + i * b.r1_len;
        if (c.out[g].bc_state == INVALID) {
          Seq content;
          if (posterior_correct(wl, lib.prior, bytes(s + lib.bc_off, lib.bc_len), q + lib.bc_off,
                                c.max_expected_errors, c.threshold, &content)) {
            c.out[g].bc_state = VALID_AFTER_CORRECTION;
            c.bc_content[g] = content;
            hc[t][content] += 1;
          }
        }
        if (lib.is_fb) {
          int f = -1;
          if (b.r2_len >= lib.pat.offset + lib.pat.len)
            f = find_closest(lib.pat, &c.feat_dist,
                             bytes(b.r2_seq + i * b.r2_len + lib.pat.offset, lib.pat.len),
                             b.r2_qual + i * b.r2_len + lib.pat.offset);
          c.out[g].feature = f >= 0 ? (uint32_t)f : NO_FEATURE;
        }
      }
    });
    for (auto& h : hc)
      for (auto& kv : h) lib.corrected_counts[kv.first] += kv.second;
  }
  c.n_valid_before = c.n_corrected = c.n_invalid = 0;
  for (auto& o : c.out) {
    c.n_valid_before += o.bc_state == VALID_BEFORE_CORRECTION;
    c.n_corrected += o.bc_state == VALID_AFTER_CORRECTION;
    c.n_invalid += o.bc_state == INVALID;
  }
}

// ALIGN_AND_COUNT per barcode: DupBuilder::observe (mark_dups.rs:128-155) per
// library type (aligner.rs:292-304), BarcodeDupMarker::new / ::process
// (mark_dups.rs:201-363), umi_counts.sort() + BcUmiInfo::feature_counts
// (stages/align_and_count.rs:298-333, cr_types/src/types.rs:180-188).
struct BcResult {
  std::vector<std::pair<uint32_t, uint32_t>> feature_counts;  // sorted by feature
  std::vector<Molecule> molecules;                            // bc_idx filled later
};

static void process_barcode(Ctx& c, const std::vector<uint64_t>& reads, const std::vector<int>& read_lib,
                            BcResult* res) {
  std::map<int, DupMarker> marker;  // keyed by library type
  auto select_key = [&](uint64_t g) { return c.select_key.empty() ? g : c.select_key[g]; };
  for (size_t k = 0; k < reads.size(); k++) {
    uint64_t g = reads[k];
    const ReadOut& o = c.out[g];
    if (o.umi_valid && o.feature != NO_FEATURE) {
      DupMarker& m = marker[read_lib[k]];
      UG key{c.umi_out[g], o.feature};
      m.counts[key] += 1;
      const uint64_t sk = select_key(g);  // old_min.min(ann_key), mark_dups.rs:147-151
      auto it = m.min_key.find(key);
      if (it == m.min_key.end())
        m.min_key.emplace(key, sk);
      else
        it->second = std::min(it->second, sk);
    } else {
      marker[read_lib[k]];  // dup_builder.entry(lib).or_default()
    }
  }
  for (auto& kv : marker) kv.second.build(c.filter_umis, c.libs[kv.first].umi_correction);
  struct UC {  // UmiCount, ordered as its derive(Ord) orders it (cr_types/src/types.rs:152-160)
    uint32_t lib, feature, umi, read_count, utype_ord;  // utype_ord: 0 Txomic < 1 NonTxomic
    bool operator<(const UC& o) const {
      if (lib != o.lib) return lib < o.lib;
      if (feature != o.feature) return feature < o.feature;
      if (umi != o.umi) return umi < o.umi;
      if (read_count != o.read_count) return read_count < o.read_count;
      return utype_ord < o.utype_ord;
    }
  };
  std::vector<UC> umi_counts;
  for (size_t k = 0; k < reads.size(); k++) {
    uint64_t g = reads[k];
    ReadOut& o = c.out[g];
    if (!o.umi_valid || o.feature == NO_FEATURE) continue;  // process() → None
    DupMarker& m = marker[read_lib[k]];
    UG raw{c.umi_out[g], o.feature};
    auto ci = m.corrections.find(raw);
    Seq corrected = ci == m.corrections.end() ? raw.umi : ci->second;
    bool is_corrected = ci != m.corrections.end();
    UG ck{corrected, raw.gene};
    bool low = m.low_support.count(ck) > 0;
    // is_min_qname compares the header with the qname of the key's UmiSelectKey (mark_dups.rs:300-303)
    const uint64_t sk = select_key(g);
    bool is_min = (m.min_key.at(ck) & 0x7FFFFFFFFFFFFFFFull) == (sk & 0x7FFFFFFFFFFFFFFFull);
    uint64_t rc = m.counts.at(ck);
    // is_filtered_target_umi (mark_dups.rs:311-320): on-target feature, fewer reads than the threshold, not low support
    bool filtered = c.target_min_reads != 0 && raw.gene < c.on_target.size() && c.on_target[raw.gene] &&
                    rc < c.target_min_reads && !low;
    bool is_umi_count = !low && is_min && !filtered;  // sampling_factor is always true (stages/stubs.rs:6-8)
    o.is_filtered_target = filtered;
    o.has_dup = 1;
    o.is_corrected = is_corrected;
    o.is_low_support = low;
    o.is_umi_count = is_umi_count;
    o.read_count = (uint32_t)rc;
    c.umi_out[g] = corrected;
    // umi_type is the one of the read that carries the count (mark_dups.rs:322-326)
    if (is_umi_count)
      umi_counts.push_back(UC{(uint32_t)read_lib[k], raw.gene, encode_2bit_u32(corrected), (uint32_t)rc, (uint32_t)(sk >> 63)});
  }
  std::sort(umi_counts.begin(), umi_counts.end());
  std::map<uint32_t, uint32_t> fc;
  for (auto& u : umi_counts) fc[u.feature] += 1;
  res->feature_counts.assign(fc.begin(), fc.end());
  for (auto& u : umi_counts)
    res->molecules.push_back(Molecule{0, u.lib, u.feature, u.umi, u.read_count, u.utype_ord ? 0u : 1u});
}

static void count_stage(Ctx& c, int threads) {
  // Barcode index: sorted unique union over library types of raw-valid and
  // corrected barcodes — barcode_correction.rs:401-407, cr_types/src/barcode_index.rs:39-53
  // The reads of a barcode are brought together (shardio's sort by Barcode in the reference, which hands
  // ALIGN_AND_COUNT one barcode at a time in ascending order): grouped through a hash map in read order, then the
  // distinct barcodes sorted.
  std::vector<int> lib_of(c.n_reads);
  for (auto& b : c.batches)
    for (uint64_t i = 0; i < b.n; i++) lib_of[b.base + i] = b.lib;
  FlatMap<Seq, uint32_t, SeqHash> group_of;
  std::vector<Seq> group_bc;
  std::vector<std::vector<uint64_t>> members;
  for (uint64_t g = 0; g < c.n_reads; g++)
    if (state_is_valid(c.out[g].bc_state)) {
      auto r = group_of.emplace(c.bc_content[g], (uint32_t)group_bc.size());
      if (r.second) {
        group_bc.push_back(c.bc_content[g]);
        members.emplace_back();
      }
      members[r.first->second].push_back(g);
    }
  std::vector<uint32_t> order(group_bc.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return group_bc[a] < group_bc[b]; });
  c.barcodes.clear();
  std::vector<const std::vector<uint64_t>*> groups;
  for (uint32_t i : order) {
    c.barcodes.push_back(group_bc[i]);
    groups.push_back(&members[i]);
  }
  size_t nb = groups.size();
  std::vector<BcResult> results(nb);
  parallel_for(nb, threads, [&](int, uint64_t lo, uint64_t hi) {
    for (uint64_t k = lo; k < hi; k++) {
      std::vector<int> rl;
      rl.reserve(groups[k]->size());
      for (uint64_t g : *groups[k]) rl.push_back(lib_of[g]);
      process_barcode(c, *groups[k], rl, &results[k]);
    }
  });
  // write_matrix_h5_helper — lib/rust/cr_h5/src/count_matrix.rs:382-448
  c.indptr.assign(1, 0);
  c.indices.clear();
  c.data.clear();
  c.molecules.clear();
  for (size_t k = 0; k < nb; k++) {
    for (auto& fcount : results[k].feature_counts) {
      c.indices.push_back(fcount.first);
      c.data.push_back((int32_t)fcount.second);
    }
    c.indptr.push_back((int64_t)c.indices.size());
    for (auto m : results[k].molecules) {
      m.bc_idx = (uint32_t)k;
      c.molecules.push_back(m);
    }
  }
  c.n_umi_corrected_reads = c.n_low_support_reads = c.n_umis = c.n_dup_reads = 0;
  for (auto& o : c.out) {
    if (!o.has_dup) continue;
    c.n_dup_reads++;
    c.n_umi_corrected_reads += o.is_corrected;
    c.n_low_support_reads += o.is_low_support;
    c.n_umis += o.is_umi_count;
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// C API (ctypes) — used by tests/, smoke() and bench.py's CPU legs only.
// ---------------------------------------------------------------------------
extern "C" {

double cro_probability(uint8_t q) { return probability(q); }

void* cro_ctx_new() { return new Ctx(); }
void cro_ctx_free(void* p) { delete (Ctx*)p; }

// DupBuilder::build(.., targeted_umi_min_read_count) + FeatureReference::target_set; min_reads = 0: None
void cro_set_target_filter(void* p, const uint8_t* on_target, int32_t n, uint64_t min_reads) {
  Ctx& c = *(Ctx*)p;
  c.on_target.assign(on_target, on_target + (on_target ? n : 0));
  c.target_min_reads = on_target ? min_reads : 0;
}

void cro_set_params(void* p, double threshold, double max_expected_errors, int filter_umis) {
  Ctx& c = *(Ctx*)p;
  c.threshold = threshold;
  c.max_expected_errors = max_expected_errors;
  c.filter_umis = filter_umis != 0;
}

// seqs: n*L ASCII bytes; trans: n*L ASCII bytes or NULL
int cro_add_whitelist(void* p, const uint8_t* seqs, uint64_t n, int L, const uint8_t* trans) {
  Ctx& c = *(Ctx*)p;
  Whitelist w;
  w.L = L;
  w.is_trans = trans != nullptr;
  for (uint64_t i = 0; i < n; i++) {
    if (trans)
      w.trans[bytes(seqs + i * L, L)] = bytes(trans + i * L, L);
    else
      w.plain.insert(bytes(seqs + i * L, L));
  }
  c.wls.push_back(std::move(w));
  return (int)c.wls.size() - 1;
}

int cro_add_library(void* p, int wl, int bc_off, int bc_len, int umi_off, int umi_len, int umi_correction,
                    int is_fb, int ftype, int fb_offset, int fb_len) {
  Ctx& c = *(Ctx*)p;
  Library l;
  l.wl = wl;
  l.bc_off = bc_off;
  l.bc_len = bc_len;
  l.umi_off = umi_off;
  l.umi_len = umi_len;
  l.umi_correction = umi_correction != 0;
  l.is_fb = is_fb != 0;
  l.ftype = ftype;
  l.pat.offset = fb_offset;
  l.pat.len = fb_len;
  c.libs.push_back(std::move(l));
  return (int)c.libs.size() - 1;
}

// feature_type[f]: 0 = gene, otherwise the ftype id of the owning FB library.
// fb_seqs: for features with type != 0, fb_len ASCII bytes at fb_seqs[f*fb_stride].
void cro_set_features(void* p, int n_features, const int32_t* feature_type, const uint8_t* fb_seqs, int fb_stride) {
  Ctx& c = *(Ctx*)p;
  c.n_features = n_features;
  c.feature_type.assign(feature_type, feature_type + n_features);
  for (auto& l : c.libs) {
    if (!l.is_fb) continue;
    l.pat.features.clear();
    for (int f = 0; f < n_features; f++)
      if (feature_type[f] == l.ftype) l.pat.features[bytes(fb_seqs + (size_t)f * fb_stride, l.pat.len)] = f;
  }
}

// Buffers are borrowed: the caller keeps them alive until cro_ctx_free.
void cro_add_reads(void* p, int lib, uint64_t n, int r1_len, const uint8_t* r1_seq, const uint8_t* r1_qual,
                   const uint32_t* feature, int r2_len, const uint8_t* r2_seq, const uint8_t* r2_qual) {
  Ctx& c = *(Ctx*)p;
  Batch b;
  b.lib = lib;
  b.n = n;
  b.r1_len = r1_len;
  b.r1_seq = r1_seq;
  b.r1_qual = r1_qual;
  b.feature = feature;
  b.r2_len = r2_len;
  b.r2_seq = r2_seq;
  b.r2_qual = r2_qual;
  b.base = c.n_reads;
  c.n_reads += n;
  c.batches.push_back(b);
}

// drop the read batches and every per-run result; whitelists, libraries and features stay
void cro_reset_reads(void* p) {
  Ctx& c = *(Ctx*)p;
  c.batches.clear();
  c.n_reads = 0;
  c.out.clear();
  c.bc_content.clear();
  c.umi_out.clear();
  c.select_key.clear();
  for (auto& l : c.libs) {
    l.prior.clear();
    l.corrected_counts.clear();
  }
}

void cro_pass1(void* p, int threads) { pass1(*(Ctx*)p, threads); }
void cro_pass2(void* p, int threads) { pass2(*(Ctx*)p, threads); }
void cro_count(void* p, int threads) { count_stage(*(Ctx*)p, threads); }
void cro_run(void* p, int threads) {
  pass1(*(Ctx*)p, threads);
  pass2(*(Ctx*)p, threads);
  count_stage(*(Ctx*)p, threads);
}

// multi-chunk priors: add counts to a library's prior (the MAKE_SHARD join sums
// the histograms of every chunk, lib/rust/cr_lib/src/stages/make_shard.rs:303-358)
void cro_prior_add(void* p, int lib, const uint8_t* seqs, uint64_t n, int L, const int64_t* counts) {
  Ctx& c = *(Ctx*)p;
  for (uint64_t i = 0; i < n; i++)
    if (counts[i]) c.libs[lib].prior[bytes(seqs + i * L, L)] += counts[i];
}
void cro_prior_clear(void* p, int lib) { ((Ctx*)p)->libs[lib].prior.clear(); }
void cro_fb_counts_set(void* p, const int64_t* counts) {
  Ctx& c = *(Ctx*)p;
  c.fb_exact_counts.assign(counts, counts + c.n_features);
  c.feat_dist = compute_feature_dist(c.fb_exact_counts, c.feature_type);
}

// which: 0 = prior (raw valid), 1 = corrected
void cro_get_counts(void* p, int lib, int which, const uint8_t* seqs, uint64_t n, int L, int64_t* out) {
  Ctx& c = *(Ctx*)p;
  const Hist& h = which == 0 ? c.libs[lib].prior : c.libs[lib].corrected_counts;
  for (uint64_t i = 0; i < n; i++) out[i] = hist_get(h, bytes(seqs + i * L, L));
}
void cro_get_fb_counts(void* p, int64_t* out) {
  Ctx& c = *(Ctx*)p;
  for (int f = 0; f < c.n_features; f++) out[f] = c.fb_exact_counts[f];
}
void cro_get_feat_dist(void* p, double* out) {
  Ctx& c = *(Ctx*)p;
  for (int f = 0; f < c.n_features; f++) out[f] = c.feat_dist[f];
}

uint64_t cro_n_reads(void* p) { return ((Ctx*)p)->n_reads; }

// Per-read outputs. bc: n*bc_len ASCII content; umi: n*umi_len ASCII processed UMI.
// flags bit0 umi_valid, bit1 has_dup, bit2 is_corrected, bit3 is_low_support, bit4 is_umi_count,
// bit5 is_filtered_target_umi
void cro_get_reads(void* p, int bc_len, int umi_len, uint8_t* bc, uint8_t* state, uint8_t* umi, uint8_t* flags,
                   uint32_t* feature, uint32_t* read_count) {
  Ctx& c = *(Ctx*)p;
  for (uint64_t g = 0; g < c.n_reads; g++) {
    const ReadOut& o = c.out[g];
    if (bc) memcpy(bc + g * bc_len, c.bc_content[g].data(), std::min((size_t)bc_len, c.bc_content[g].size()));
    if (state) state[g] = o.bc_state;
    if (umi) memcpy(umi + g * umi_len, c.umi_out[g].data(), std::min((size_t)umi_len, c.umi_out[g].size()));
    if (flags)
      flags[g] = (uint8_t)(o.umi_valid | (o.has_dup << 1) | (o.is_corrected << 2) | (o.is_low_support << 3) |
                           (o.is_umi_count << 4) | (o.is_filtered_target << 5));
    if (feature) feature[g] = o.feature;
    if (read_count) read_count[g] = o.read_count;
  }
}

void cro_get_stats(void* p, uint64_t* out8) {
  Ctx& c = *(Ctx*)p;
  out8[0] = c.n_valid_before;
  out8[1] = c.n_corrected;
  out8[2] = c.n_invalid;
  out8[3] = c.n_dup_reads;
  out8[4] = c.n_umi_corrected_reads;
  out8[5] = c.n_low_support_reads;
  out8[6] = c.n_umis;
  out8[7] = c.molecules.size();
}

uint64_t cro_matrix_n_barcodes(void* p) { return ((Ctx*)p)->barcodes.size(); }
uint64_t cro_matrix_nnz(void* p) { return ((Ctx*)p)->indices.size(); }
void cro_matrix_get(void* p, int bc_len, uint8_t* barcodes, int64_t* indptr, uint32_t* indices, int32_t* data) {
  Ctx& c = *(Ctx*)p;
  for (size_t k = 0; k < c.barcodes.size(); k++) memcpy(barcodes + k * bc_len, c.barcodes[k].data(), bc_len);
  memcpy(indptr, c.indptr.data(), c.indptr.size() * sizeof(int64_t));
  if (!c.indices.empty()) {
    memcpy(indices, c.indices.data(), c.indices.size() * sizeof(uint32_t));
    memcpy(data, c.data.data(), c.data.size() * sizeof(int32_t));
  }
}
// bc_counts_total of BarcodeCorrection::main (cr_lib/src/stages/barcode_correction.rs:327-362): every read of the
// invalid shard observed under its barcode after correction, entries below min_reads_to_report_bc dropped; the
// whole read set as one chunk. Entries in Barcode order (derive(Ord): valid = false first, then the sequence).
// Returns the number of entries; a second call with buffers copies them out.
uint64_t cro_total_barcode_counts(void* p, uint64_t min_reads, int bc_len, uint8_t* seqs, uint8_t* valid, uint64_t* counts) {
  Ctx& c = *(Ctx*)p;
  std::map<std::pair<bool, Seq>, uint64_t> hist;
  for (uint64_t g = 0; g < c.n_reads; g++) {
    const uint8_t st = c.out[g].bc_state;
    if (st == VALID_BEFORE_CORRECTION) continue;  // never reaches BARCODE_CORRECTION's reader
    hist[{st == VALID_AFTER_CORRECTION, c.bc_content[g]}] += 1;  // bc_counts_total.observe_owned(bc), :344
  }
  uint64_t n = 0;
  for (const auto& kv : hist) {
    if (kv.second < min_reads) continue;  // retain(|_, v| min_reads_to_report_bc <= v.count()), :357
    if (seqs) memcpy(seqs + n * bc_len, kv.first.second.data(), std::min((size_t)bc_len, kv.first.second.size()));
    if (valid) valid[n] = kv.first.first ? 1 : 0;
    if (counts) counts[n] = kv.second;
    n++;
  }
  return n;
}

uint64_t cro_n_molecules(void* p) { return ((Ctx*)p)->molecules.size(); }
// rows in the order ALIGN_AND_COUNT emits them: by barcode, then umi_counts.sort() (align_and_count.rs:314)
void cro_molecules_get(void* p, uint32_t* out6) {
  Ctx& c = *(Ctx*)p;
  for (size_t k = 0; k < c.molecules.size(); k++) {
    out6[6 * k + 0] = c.molecules[k].bc_idx;
    out6[6 * k + 1] = c.molecules[k].lib;
    out6[6 * k + 2] = c.molecules[k].feature;
    out6[6 * k + 3] = c.molecules[k].umi;
    out6[6 * k + 4] = c.molecules[k].read_count;
    out6[6 * k + 5] = c.molecules[k].utype;
  }
}
// one UmiSelectKey word per read, in the order the reads were added (n must equal the read count); n = 0 clears
void cro_set_select_keys(void* p, const uint64_t* keys, uint64_t n) {
  Ctx& c = *(Ctx*)p;
  c.select_key.assign(keys, keys + n);
}

// ---- single-function entry points for the reference's known-answer tests ----

// Posterior::correct_barcode on one segment. wl: n*L ASCII (+trans or NULL);
// count_seqs/counts: the prior histogram. Returns 1 and writes L bytes on accept.
int cro_kat_correct_barcode(const uint8_t* wl_seqs, uint64_t n_wl, int L, const uint8_t* trans,
                            const uint8_t* count_seqs, const int64_t* counts, uint64_t n_counts,
                            const uint8_t* observed, const uint8_t* qual, double max_expected_errors,
                            double threshold, uint8_t* out) {
  Whitelist w;
  w.L = L;
  w.is_trans = trans != nullptr;
  for (uint64_t i = 0; i < n_wl; i++) {
    if (trans)
      w.trans[bytes(wl_seqs + i * L, L)] = bytes(trans + i * L, L);
    else
      w.plain.insert(bytes(wl_seqs + i * L, L));
  }
  Hist h;
  for (uint64_t i = 0; i < n_counts; i++) h[bytes(count_seqs + i * L, L)] += counts[i];
  Seq content;
  if (posterior_correct(w, h, bytes(observed, L), qual, max_expected_errors, threshold, &content)) {
    memcpy(out, content.data(), L);
    return 1;
  }
  return 0;
}

// Whitelist::match_to_whitelist — lib/rust/barcode/src/whitelist.rs:532-545
int cro_kat_match_to_whitelist(const uint8_t* wl_seqs, uint64_t n_wl, int L, const uint8_t* seq, uint8_t* out) {
  Whitelist w;
  w.L = L;
  for (uint64_t i = 0; i < n_wl; i++) w.plain.insert(bytes(wl_seqs + i * L, L));
  Seq s = bytes(seq, L);
  if (w.contains(s)) {
    memcpy(out, s.data(), L);
    return 1;
  }
  size_t pos_n = s.find('N');
  if (pos_n == Seq::npos) return 0;
  for (char b : BASE_OPTS) {
    s[pos_n] = b;
    if (w.contains(s)) {
      memcpy(out, s.data(), L);
      return 1;
    }
  }
  return 0;
}

// correct_umis on an explicit (umi, gene, count) table. out_corr: n*L bytes,
// equal to the input UMI where no correction is recorded. Returns #corrections.
int cro_kat_correct_umis(const uint8_t* umis, const uint32_t* genes, const uint64_t* counts, uint64_t n, int L,
                         uint8_t* out_corr) {
  UGCounts c;
  for (uint64_t i = 0; i < n; i++) c[UG{bytes(umis + i * L, L), genes[i]}] = counts[i];
  UGCorr corr = correct_umis(c);
  for (uint64_t i = 0; i < n; i++) {
    auto it = corr.find(UG{bytes(umis + i * L, L), genes[i]});
    memcpy(out_corr + i * L, it == corr.end() ? (const char*)(umis + i * L) : it->second.data(), L);
  }
  return (int)corr.size();
}

// determine_low_support_umigenes on an explicit table; out_low[i] = 1 if low support
void cro_kat_low_support(const uint8_t* umis, const uint32_t* genes, const uint64_t* counts, uint64_t n, int L,
                         uint8_t* out_low) {
  UGCounts c;
  for (uint64_t i = 0; i < n; i++) c[UG{bytes(umis + i * L, L), genes[i]}] = counts[i];
  auto low = determine_low_support(c);
  for (uint64_t i = 0; i < n; i++) out_low[i] = low.count(UG{bytes(umis + i * L, L), genes[i]}) ? 1 : 0;
}

int cro_kat_umi_is_valid(const uint8_t* seq, const uint8_t* qual, int L) { return umi_is_valid(seq, qual, L); }
uint32_t cro_kat_encode_2bit(const uint8_t* seq, int L) { return encode_2bit_u32(bytes(seq, L)); }

void cro_kat_feature_dist(const int64_t* raw, const int32_t* ftype, int n, double* out) {
  std::vector<double> p = compute_feature_dist(std::vector<int64_t>(raw, raw + n), std::vector<int>(ftype, ftype + n));
  for (int i = 0; i < n; i++) out[i] = p[i];
}

// FeatureExtractor::find_closest for one tethered capture of length L.
// feat_seqs: n*L ASCII, feat_idx[n]; feat_dist over all features (or NULL for exact-only).
int cro_kat_feature_match(const uint8_t* feat_seqs, const int32_t* feat_idx, int n, int L, const double* feat_dist,
                          int n_dist, const uint8_t* seq, const uint8_t* qual) {
  FeaturePattern pat;
  pat.len = L;
  for (int i = 0; i < n; i++) pat.features[bytes(feat_seqs + (size_t)i * L, L)] = feat_idx[i];
  std::vector<double> d;
  if (feat_dist) d.assign(feat_dist, feat_dist + n_dist);
  return find_closest(pat, feat_dist ? &d : nullptr, bytes(seq, L), qual);
}

}  // extern "C"
