// crgpu.hpp — C++17 host layer over the C ABI of crgpu.h (header only).
//
// The reference's host side is compiled code (Rust); this image has no Rust toolchain, so the compiled host
// layer that mirrors the reference's interfaces for the path is C++ (the Python mirror in cellranger_b200/api.py
// is what the tests and bench.py drive). Names, argument meaning and error behaviour follow the reference:
//
//   crgpu::Whitelist            barcode::Whitelist::{Plain, Trans}            lib/rust/barcode/src/whitelist.rs:452-525
//   crgpu::Posterior            barcode::corrector::Posterior                 lib/rust/barcode/src/corrector.rs:83-109
//   crgpu::BarcodeCorrector     barcode::corrector::BarcodeCorrector + trait CorrectBarcode (batch form)   :15-81
//   crgpu::ChemistryDef         barcode / UMI read components                 lib/rust/cr_types/src/chemistry/mod.rs:718-751
//   crgpu::GemWell::make_shard / barcode_correction / align_and_count
//                               the three MartianStage mains                  lib/rust/cr_lib/src/stages/{make_shard,barcode_correction,align_and_count}.rs
//   crgpu::CountMatrix          CSC layout of write_matrix_h5_helper          lib/rust/cr_h5/src/count_matrix.rs:382-448
//   crgpu::UmiCount             cr_types::types::UmiCount                     lib/rust/cr_types/src/types.rs:148-160
//   crgpu::BarcodeSummary       cr_lib::aligner::BarcodeSummary               lib/rust/cr_lib/src/aligner.rs:33-68
//
// Errors: every failing C call becomes a crgpu::Error carrying crgpu_last_error() (the reference returns
// anyhow::Error from the stage). There is no CPU fallback: without a CUDA device GemWell's constructor throws.
#ifndef CRGPU_HPP
#define CRGPU_HPP

#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "crgpu.h"

namespace crgpu {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int rc, const char* what) {
  if (rc != CRGPU_OK) throw Error(rc, std::string(what) + " failed (" + std::to_string(rc) + "): " + crgpu_last_error());
}

// barcode::BarcodeSegmentState (lib/rust/barcode/src/lib.rs:270-283)
enum class BarcodeSegmentState : uint8_t {
  NotChecked = CRGPU_BC_NOT_CHECKED,
  ValidBeforeCorrection = CRGPU_BC_VALID_BEFORE_CORRECTION,
  ValidAfterCorrection = CRGPU_BC_VALID_AFTER_CORRECTION,
  Invalid = CRGPU_BC_INVALID,
};

// Whitelist::Plain(set) or Whitelist::Trans(raw -> translated)
struct Whitelist {
  int length = 0;
  std::vector<uint8_t> seqs;        // n * length ASCII
  std::vector<uint8_t> translated;  // empty (Plain) or n * length ASCII (Trans)

  static Whitelist plain(const std::vector<std::string>& s) {
    Whitelist w;
    w.length = s.empty() ? 0 : (int)s[0].size();
    for (const auto& x : s) {
      if ((int)x.size() != w.length) throw Error(CRGPU_E_INVALID, "whitelist sequences of different lengths");
      w.seqs.insert(w.seqs.end(), x.begin(), x.end());
    }
    return w;
  }
  static Whitelist trans(const std::vector<std::pair<std::string, std::string>>& raw_to_translated) {
    Whitelist w;
    w.length = raw_to_translated.empty() ? 0 : (int)raw_to_translated[0].first.size();
    for (const auto& p : raw_to_translated) {
      if ((int)p.first.size() != w.length || (int)p.second.size() != w.length)
        throw Error(CRGPU_E_INVALID, "whitelist sequences of different lengths");
      w.seqs.insert(w.seqs.end(), p.first.begin(), p.first.end());
      w.translated.insert(w.translated.end(), p.second.begin(), p.second.end());
    }
    return w;
  }
  uint64_t size() const { return length ? seqs.size() / (size_t)length : 0; }
};

// Posterior { max_expected_barcode_errors, bc_confidence_threshold } with the reference's defaults
struct Posterior {
  double max_expected_barcode_errors = std::numeric_limits<double>::max();
  double bc_confidence_threshold = 0.975;
};

// where barcode and UMI sit in R1
struct ChemistryDef {
  std::string name;
  int bc_offset = 0, bc_length = 16, umi_offset = 16, umi_length = 12;
  static ChemistryDef SC3Pv2() { return {"SC3Pv2", 0, 16, 16, 10}; }
  static ChemistryDef SC3Pv3() { return {"SC3Pv3", 0, 16, 16, 12}; }
  int r1_length() const {
    const int a = bc_offset + bc_length, b = umi_offset + umi_length;
    return a > b ? a : b;
  }
};

struct CountMatrix {  // feature x barcode, CSC: one column per valid barcode, sorted
  uint64_t n_features = 0;
  std::vector<uint32_t> barcode_rank;  // content rank of each column
  std::vector<std::string> barcodes;   // the sequences (no "-1" suffix)
  std::vector<int64_t> indptr;         // n_barcodes + 1
  std::vector<uint32_t> indices;       // feature index
  std::vector<int32_t> data;           // UMI count
};

struct UmiCount {
  uint32_t barcode_column, library_idx, feature_idx, umi /* 2 bit / base */, read_count;
  uint32_t umi_type;  // as molecule_info stores it: 1 = Txomic, 0 = NonTxomic
};

struct BarcodeSummary {
  uint32_t barcode_rank;
  uint64_t reads, umis, candidate_dup_reads, umi_corrected_reads;
};

struct ReadResults {  // per read of one batch
  std::vector<uint32_t> bc_rank;  // CRGPU_NO_RANK when the barcode is invalid
  std::vector<BarcodeSegmentState> bc_state;
  std::vector<uint32_t> umi;  // processed UMI, 2 bit / base
  std::vector<uint8_t> flags;  // CRGPU_F_*
  std::vector<uint32_t> feature;
};

// One GEM well on one GPU.
class GemWell {
 public:
  explicit GemWell(int device = 0, Posterior posterior = Posterior(), bool filter_umis = true) {
    check(crgpu_ctx_create(device, &ctx_), "crgpu_ctx_create");
    check(crgpu_set_params(ctx_, posterior.bc_confidence_threshold, posterior.max_expected_barcode_errors,
                           filter_umis ? 1 : 0),
          "crgpu_set_params");
  }
  ~GemWell() {
    if (ctx_) crgpu_ctx_destroy(ctx_);
  }
  GemWell(const GemWell&) = delete;
  GemWell& operator=(const GemWell&) = delete;

  crgpu_ctx* ctx() { return ctx_; }

  // DupBuilder::build(filter_umis, umi_correction, targeted_umi_min_read_count) with the panel's target set
  // (tx_annotation/src/mark_dups.rs:156-170,311-320); min_read_count 0 = None
  void set_target_filter(const std::vector<uint8_t>& on_target, uint64_t targeted_umi_min_read_count) {
    check(crgpu_set_target_filter(ctx_, on_target.data(), (int32_t)on_target.size(), targeted_umi_min_read_count),
          "crgpu_set_target_filter");
  }

  int add_whitelist(const Whitelist& w) {
    int id = -1;
    check(crgpu_whitelist_add(ctx_, w.seqs.data(), w.size(), w.length,
                              w.translated.empty() ? nullptr : w.translated.data(), &id),
          "crgpu_whitelist_add");
    return id;
  }
  // umi_correction = false for Multiplexing Capture libraries (cr_lib/src/aligner.rs:313-318)
  int add_library(int whitelist, const ChemistryDef& chem, bool umi_correction = true) {
    crgpu_library_def d;
    std::memset(&d, 0, sizeof(d));
    d.whitelist = whitelist;
    d.bc_offset = chem.bc_offset;
    d.bc_length = chem.bc_length;
    d.umi_offset = chem.umi_offset;
    d.umi_length = chem.umi_length;
    d.umi_correction = umi_correction ? 1 : 0;
    int id = -1;
    check(crgpu_library_add(ctx_, &d, &id), "crgpu_library_add");
    libs_.push_back(chem);
    return id;
  }
  // a feature-barcode library: capture sequence at R2[fb_offset : +fb_length] (tethered patterns)
  int add_feature_barcode_library(int whitelist, const ChemistryDef& chem, int feature_type, int fb_offset,
                                  int fb_length, bool umi_correction = true) {
    crgpu_library_def d;
    std::memset(&d, 0, sizeof(d));
    d.whitelist = whitelist;
    d.bc_offset = chem.bc_offset;
    d.bc_length = chem.bc_length;
    d.umi_offset = chem.umi_offset;
    d.umi_length = chem.umi_length;
    d.umi_correction = umi_correction ? 1 : 0;
    d.is_feature_barcode = 1;
    d.feature_type = feature_type;
    d.fb_offset = fb_offset;
    d.fb_length = fb_length;
    int id = -1;
    check(crgpu_library_add(ctx_, &d, &id), "crgpu_library_add");
    libs_.push_back(chem);
    return id;
  }
  // FeatureReference: feature_type[f] = 0 for genes; fb_seqs = n_features * stride capture sequences (or empty)
  void set_features(const std::vector<int32_t>& feature_type, const std::vector<uint8_t>& fb_seqs = {},
                    int fb_stride = 0) {
    check(crgpu_features_set(ctx_, (int32_t)feature_type.size(), feature_type.data(),
                             fb_seqs.empty() ? nullptr : fb_seqs.data(), fb_stride),
          "crgpu_features_set");
  }
  // host arrays: r1_seq / r1_qual = n * r1_len ASCII, feature = gene index per read (CRGPU_NO_FEATURE = unmapped)
  int add_reads(int library, uint64_t n, int r1_len, const uint8_t* r1_seq, const uint8_t* r1_qual,
                const uint32_t* feature, int r2_len = 0, const uint8_t* r2_seq = nullptr,
                const uint8_t* r2_qual = nullptr, const uint64_t* select_key = nullptr) {
    crgpu_read_batch b;
    std::memset(&b, 0, sizeof(b));
    b.n = n;
    b.r1_len = r1_len;
    b.r1_seq = r1_seq;
    b.r1_qual = r1_qual;
    b.feature = feature;
    b.r2_len = r2_len;
    b.r2_seq = r2_seq;
    b.r2_qual = r2_qual;
    b.select_key = select_key;  // UmiSelectKey{utype, qname} per read as one word, or nullptr
    int id = -1;
    check(crgpu_reads_add(ctx_, library, &b, &id), "crgpu_reads_add");
    batch_sizes_.push_back(n);
    return id;
  }
  void clear_reads() {
    check(crgpu_reads_clear(ctx_), "crgpu_reads_clear");
    batch_sizes_.clear();
  }

  // the three stages
  void make_shard() { check(crgpu_pass1(ctx_), "crgpu_pass1"); }
  void barcode_correction() { check(crgpu_pass2(ctx_), "crgpu_pass2"); }
  void align_and_count(bool annotate_reads = false) {
    check(crgpu_count(ctx_), "crgpu_count");
    if (annotate_reads) check(crgpu_annotate_reads(ctx_), "crgpu_annotate_reads");
  }
  void run(bool annotate_reads = false) {
    make_shard();
    barcode_correction();
    align_and_count(annotate_reads);
  }

  CountMatrix count_matrix() {
    CountMatrix m;
    uint64_t nb = 0, nnz = 0;
    check(crgpu_matrix_dims(ctx_, &nb, &nnz, &m.n_features), "crgpu_matrix_dims");
    m.barcode_rank.resize(nb);
    m.indptr.resize(nb + 1);
    m.indices.resize(nnz);
    m.data.resize(nnz);
    check(crgpu_matrix_get(ctx_, m.barcode_rank.data(), m.indptr.data(), m.indices.data(), m.data.data()),
          "crgpu_matrix_get");
    m.barcodes = barcode_seqs(m.barcode_rank);
    return m;
  }
  std::vector<std::string> barcode_seqs(const std::vector<uint32_t>& ranks) {
    uint64_t n_content = 0;
    int L = 0;
    check(crgpu_whitelist_size(ctx_, &n_content, &L), "crgpu_whitelist_size");
    std::vector<uint8_t> ascii(ranks.size() * (size_t)L);
    if (!ranks.empty()) check(crgpu_barcode_seqs(ctx_, ranks.data(), ranks.size(), ascii.data()), "crgpu_barcode_seqs");
    std::vector<std::string> out(ranks.size());
    for (size_t i = 0; i < ranks.size(); i++) out[i].assign(reinterpret_cast<const char*>(ascii.data()) + i * L, (size_t)L);
    return out;
  }
  std::vector<UmiCount> molecules() {
    uint64_t n = 0;
    check(crgpu_molecules_count(ctx_, &n), "crgpu_molecules_count");
    std::vector<uint32_t> raw(n * 6);
    if (n) check(crgpu_molecules_get(ctx_, raw.data()), "crgpu_molecules_get");
    std::vector<UmiCount> out(n);
    for (uint64_t i = 0; i < n; i++)
      out[i] = {raw[6 * i], raw[6 * i + 1], raw[6 * i + 2], raw[6 * i + 3], raw[6 * i + 4], raw[6 * i + 5]};
    return out;
  }
  // rows for the barcodes with at least one read in `library`, in barcode order
  std::vector<BarcodeSummary> barcode_summary(int library) {
    uint64_t nb = 0;
    check(crgpu_matrix_dims(ctx_, &nb, nullptr, nullptr), "crgpu_matrix_dims");
    std::vector<uint32_t> raw(nb * 4), rank(nb);
    if (nb) {
      check(crgpu_barcode_summary(ctx_, library, raw.data()), "crgpu_barcode_summary");
      check(crgpu_matrix_get(ctx_, rank.data(), nullptr, nullptr, nullptr), "crgpu_matrix_get");
    }
    std::vector<BarcodeSummary> out;
    for (uint64_t c = 0; c < nb; c++)
      if (raw[4 * c]) out.push_back({rank[c], raw[4 * c], raw[4 * c + 1], raw[4 * c + 2], raw[4 * c + 3]});
    return out;
  }
  ReadResults reads(int batch) {
    const uint64_t n = batch_sizes_.at((size_t)batch);
    ReadResults r;
    r.bc_rank.resize(n);
    r.umi.resize(n);
    r.flags.resize(n);
    r.feature.resize(n);
    std::vector<uint8_t> st(n);
    check(crgpu_reads_get(ctx_, batch, r.bc_rank.data(), st.data(), r.umi.data(), r.flags.data(), r.feature.data()),
          "crgpu_reads_get");
    r.bc_state.resize(n);
    for (uint64_t i = 0; i < n; i++) r.bc_state[i] = static_cast<BarcodeSegmentState>(st[i]);
    return r;
  }
  void write_mex(const std::string& folder, const std::string& software_version, int gem_group = 1,
                 const std::string& features_tsv = std::string()) {
    check(crgpu_matrix_write_mex(ctx_, folder.c_str(), software_version.c_str(), gem_group,
                                 features_tsv.empty() ? nullptr : features_tsv.c_str()),
          "crgpu_matrix_write_mex");
  }
  void set_prior(int library, const std::vector<uint32_t>& counts_by_rank) {
    check(crgpu_prior_set(ctx_, library, counts_by_rank.data(), counts_by_rank.size()), "crgpu_prior_set");
  }
  uint64_t n_content() {
    uint64_t n = 0;
    check(crgpu_whitelist_size(ctx_, &n, nullptr), "crgpu_whitelist_size");
    return n;
  }

 private:
  crgpu_ctx* ctx_ = nullptr;
  std::vector<ChemistryDef> libs_;
  std::vector<uint64_t> batch_sizes_;
};

// BarcodeCorrector::new(whitelist, bc_counts, strategy) with the trait's correct_barcode in batch form:
// observed segments + optional qualities in, Some((corrected segment, distance 1)) / None out.
class BarcodeCorrector {
 public:
  struct Corrected {
    std::string segment;  // the whitelist sequence (translated for Whitelist::Trans)
    BarcodeSegmentState state;
  };
  // bc_counts: (whitelist sequence, count) pairs = SimpleHistogram<BcSegSeq> of the valid-before reads
  BarcodeCorrector(const Whitelist& whitelist, const std::vector<std::pair<std::string, uint32_t>>& bc_counts,
                   Posterior strategy = Posterior(), int device = 0)
      : gw_(device, strategy), length_(whitelist.length) {
    const int wl = gw_.add_whitelist(whitelist);
    lib_ = gw_.add_library(wl, ChemistryDef{"segment", 0, whitelist.length, 0, 0});
    // content ranks: ask the library for the sorted content sequences once
    const uint64_t n = gw_.n_content();
    std::vector<uint32_t> all(n);
    for (uint64_t i = 0; i < n; i++) all[i] = (uint32_t)i;
    content_ = gw_.barcode_seqs(all);
    if (!bc_counts.empty()) {
      std::vector<uint32_t> prior(n, 0u);
      for (const auto& kv : bc_counts) {
        // binary search in the sorted content
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
          const uint64_t mid = (lo + hi) / 2;
          if (content_[mid] < kv.first) lo = mid + 1; else hi = mid;
        }
        if (lo < n && content_[lo] == kv.first) prior[lo] += kv.second;
      }
      gw_.set_prior(lib_, prior);
    }
  }
  // one entry per observed segment; qualities may be empty (no quality information)
  std::vector<std::optional<Corrected>> correct_barcodes(const std::vector<std::string>& observed,
                                                         const std::vector<std::string>& quals = {}) {
    const uint64_t n = observed.size();
    std::vector<uint8_t> seq(n * (size_t)length_), q;
    for (uint64_t i = 0; i < n; i++) {
      if ((int)observed[i].size() != length_) throw Error(CRGPU_E_INVALID, "segment of the wrong length");
      std::memcpy(seq.data() + i * length_, observed[i].data(), (size_t)length_);
    }
    if (!quals.empty()) {
      if (quals.size() != n) throw Error(CRGPU_E_INVALID, "one quality string per segment");
      q.resize(n * (size_t)length_);
      for (uint64_t i = 0; i < n; i++) {
        if ((int)quals[i].size() != length_) throw Error(CRGPU_E_INVALID, "quality string of the wrong length");
        std::memcpy(q.data() + i * length_, quals[i].data(), (size_t)length_);
      }
    }
    std::vector<uint32_t> rank(n);
    std::vector<uint8_t> state(n);
    check(crgpu_correct_barcodes(gw_.ctx(), lib_, seq.data(), q.empty() ? nullptr : q.data(), n, rank.data(),
                                 state.data()),
          "crgpu_correct_barcodes");
    std::vector<std::optional<Corrected>> out(n);
    for (uint64_t i = 0; i < n; i++)
      if (state[i] == CRGPU_BC_VALID_BEFORE_CORRECTION || state[i] == CRGPU_BC_VALID_AFTER_CORRECTION)
        out[i] = Corrected{content_[rank[i]], static_cast<BarcodeSegmentState>(state[i])};
    return out;
  }

 private:
  GemWell gw_;
  int lib_ = -1;
  int length_ = 0;
  std::vector<std::string> content_;
};

}  // namespace crgpu

#endif  // CRGPU_HPP
