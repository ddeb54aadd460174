// kernels.h — host-callable launchers of the libcrgpu kernels (internal).
#pragma once
#include "common.cuh"

struct Pass1Args {
  uint64_t n;         // reads of this launch: a whole batch, or one chunk of it (seq / qual / feature / bc_out /
  uint64_t idx_base;  // umi_out then point at the chunk, and idx_base is its first read's index in the batch)
  int r1_len;
  const uint8_t* seq;
  const uint8_t* qual;
  const uint32_t* feature;  // nullptr: features are resolved later (feature-barcode library)
  int bc_off, bc_len, umi_off, umi_len;
  DevWhitelist wl;
  uint32_t* prior;  // [n_content]; nullptr = do not count
  uint32_t* bc_out;
  uint32_t* umi_out;
  unsigned long long* keys;
  unsigned long long* counters;  // [0] keys written so far, [1] invalid entries of this batch
  uint32_t* inv_idx;
  uint32_t* inv_bc;
  uint32_t* inv_nmask;
  uint4* inv_qual;
  KeyLayout kl;
  uint32_t lib;
  int emit_keys;
  int have_qual;  // 0: qualities absent (crgpu_correct_barcodes with qual == NULL)
  int debug_flags;  // profiling ablations (CRGPU_P1_DBG); 0 in production
  uint32_t n_features;             // rows of the matrix: a feature index >= n_features (and != NO_FEATURE) is an error
  unsigned long long* bad_feature;  // device counter of such reads (they are treated as unmapped)
};

struct Pass2Args {
  uint64_t n_invalid;  // entries of the side list (an upper bound when n_invalid_packed_dev is set)
  // optional: a device word holding the entry count (>> n_invalid_dev_shift: 32 for the packed counter word
  // pass 1 works on, 0 for a plain count). When set, the kernel reads the count from there, so that no host
  // round trip separates the two passes.
  const unsigned long long* n_invalid_dev;
  int n_invalid_dev_shift;
  const uint32_t* inv_idx;
  const uint32_t* inv_bc;
  const uint32_t* inv_nmask;
  const uint4* inv_qual;
  DevWhitelist wl;
  const uint32_t* prior;
  uint32_t* corrected;  // [n_content] counts of corrected reads; may be nullptr
  uint32_t* bc_out;
  const uint32_t* umi_out;
  const uint32_t* feature;  // nullptr when keys are emitted later
  unsigned long long* keys;
  unsigned long long* counters;  // [0] keys
  KeyLayout kl;
  uint32_t lib;
  int emit_keys;
  int have_qual;
  double threshold;
  double max_expected_errors;
  int check_expected_errors;
  uint32_t n_features;  // as in Pass1Args (pass 1 has counted the offenders; here they are only masked)
};

struct FbArgs {
  uint64_t n;
  int r2_len, fb_off, fb_len;
  const uint8_t* r2_seq;
  const uint8_t* r2_qual;
  const uint32_t* fb_keys;   // sorted packed feature sequences of this library [n_fb]
  const uint32_t* fb_index;  // feature index of each [n_fb]
  int n_fb;
  unsigned long long* exact_counts;  // [n_features] (pass 1) or nullptr
  const double* feat_dist;           // [n_features] (pass 2) or nullptr
  uint32_t* feature_out;             // [n] resolved feature (pass 2)
  double threshold;
};

struct EmitArgs {
  uint64_t n;
  const uint32_t* bc_out;
  const uint32_t* umi_out;
  const uint32_t* feature;
  unsigned long long* keys;
  unsigned long long* counters;
  KeyLayout kl;
  uint32_t lib;
};

void upload_prob_luts(const double* bc_lut256, const double* fb_lut64, cudaStream_t st);
int launch_pass1(const Pass1Args& a, int n_sms, cudaStream_t st);  // returns #kernel launches
int launch_pass2(const Pass2Args& a, cudaStream_t st);
int launch_fb(const FbArgs& a, cudaStream_t st);
int launch_emit_keys(const EmitArgs& a, cudaStream_t st);
int launch_collect_invalid(const uint32_t* inv_idx, const uint32_t* inv_bc, const uint32_t* inv_nmask,
                           const unsigned long long* n_invalid_dev, uint64_t n_max, const uint32_t* bc_out,
                           unsigned long long* out, unsigned long long* counter, cudaStream_t st);
int launch_valid_counts(const uint32_t* prior, const uint32_t* corrected, uint32_t* out, uint64_t n, cudaStream_t st);

// ---- FASTQ text -> fixed-stride read arrays (fastq.cu) ----
size_t fastq_temp_bytes(uint64_t n_bytes);
int launch_fastq_extract(const uint8_t* text, uint64_t n_bytes, int read_len, uint8_t* out_seq, uint8_t* out_qual,
                         uint64_t capacity, void* temp, unsigned long long* counters, cudaStream_t st);

// look-back watchdog flags (one per translation unit with a chained scan, see common.cuh): fetch = asynchronous
// copy of the flag to *host_out on the stream; clear = reset after a reported timeout
void sort_lb_flag_fetch(unsigned int* host_out, cudaStream_t st);
void sort_lb_flag_clear(cudaStream_t st);
void dedup_lb_flag_fetch(unsigned int* host_out, cudaStream_t st);
void dedup_lb_flag_clear(cudaStream_t st);
void fastq_lb_flag_fetch(unsigned int* host_out, cudaStream_t st);
void fastq_lb_flag_clear(cudaStream_t st);

// ---- sort (sort.cu) ----
// sorts n 64-bit keys on bits [begin_bit, end_bit) (stable: lower bits keep their input order); result in
// *out (one of the two buffers)
int sort_keys(unsigned long long* keys, unsigned long long* alt, uint64_t n, int end_bit, void* temp,
              size_t temp_bytes, unsigned long long** out, cudaStream_t st, int begin_bit = 0);
size_t sort_temp_bytes(uint64_t n);
int sort_num_passes(int end_bit);
int sort_histograms(const unsigned long long* keys, uint64_t n, int end_bit, void* temp, cudaStream_t st,
                    int begin_bit = 0);
int sort_passes(unsigned long long* keys, unsigned long long* alt, uint64_t n, int end_bit, void* temp,
                unsigned long long** out, cudaStream_t st, int begin_bit = 0);

// ---- per-segment UMI sort + run-length encoding (finish_kernels.cu) ----
// keys sorted on the bits above the UMI -> the distinct-key table (dkeys, c0) in full key order
bool finish_supported(int umi_bits);
int run_finish(unsigned long long* keys, unsigned long long* alt, uint64_t n, int umi_bits, unsigned long long* dkeys,
               uint32_t* c0, unsigned long long* desc, unsigned long long* total_out, cudaStream_t st,
               void (*mark)(void*, const char*) = nullptr, void* mark_user = nullptr);
size_t finish_desc_bytes(uint64_t n);
void finish_lb_flag_fetch(unsigned int* host_out, cudaStream_t st);
void finish_lb_flag_clear(cudaStream_t st);

// ---- dedup / count (dedup_kernels.cu) ----
struct DedupBuffers {
  // inputs
  unsigned long long* sorted;      // [n_keys] sorted keys: on every bit, or (finish_umi) on the bits above the UMI
  unsigned long long* sorted_alt;  // the spare buffer of the sort (scratch of the finishing sort)
  int finish_umi;                  // 1: the UMI bits are still to be sorted, per segment (finish_kernels.cu)
  uint64_t n_keys;
  KeyLayout kl;
  uint32_t umi_correction_mask;  // bit lib set: UMI correction enabled for that library
  int filter_umis;
  // targeted-panel filter (mark_dups.rs:311-320): on_target[feature] != 0 and fewer than target_min_reads reads
  // (0 = no filter) and not low support: not a UMI count. Sets bit 1 of low[].
  const uint8_t* on_target;
  uint32_t n_on_target;
  unsigned long long target_min_reads;
  // work / outputs (device)
  unsigned long long* dkeys;  // [cap] distinct keys
  uint32_t* c0;               // [cap] raw read counts
  uint32_t* best;             // [cap] index of the correction target (self if uncorrected)
  unsigned long long* inc;    // [cap] incoming: count << 40 | reads
  uint8_t* low;               // [cap] bit 0: low support, bit 1: filtered target UMI
  unsigned long long* key2;   // [cap] (rank, lib, umi, feature) order for the low-support grouping
  unsigned long long* key2_alt;
  unsigned long long* lb_desc;  // look-back descriptors
  uint32_t* tickets;            // atomic tickets (several)
  unsigned long long* scalars;  // device scalars: [0] n_distinct [1] nnz [2] n_molecules [3] corrected keys [4] low keys
                                // [12] filtered target UMIs
  void* sort_temp;
  size_t sort_temp_bytes;
  void (*mark)(void* user, const char* phase);  // optional: called at the start of each sub-phase
  void* mark_user;
  int verify;          // CRGPU_VERIFY: count order violations after the sort and the run-length encoding
  uint32_t* slots;     // 2-bit hash slots of the low-support pre-filter
  size_t slots_bytes;
  // matrix entries
  uint32_t* ent_rank;     // [cap]
  uint32_t* ent_feature;  // [cap]
  uint32_t* ent_count;    // [cap]
  // molecules: read count (c2) of each; their keys end up in key2
  uint32_t* mol;      // [cap]
  uint32_t* mol_idx;  // [cap] index of each molecule in the distinct-key table (nullptr: not kept)
  uint64_t cap;
};
int run_dedup(DedupBuffers& b, uint64_t* n_distinct_host, cudaStream_t st);
struct MatrixArgs {
  const uint32_t* valid_counts[CRGPU_MAX_LIBS];
  int n_libs;
  uint32_t n_content;
  uint32_t own_lo, own_hi;
  uint32_t* col_of_rank;   // [n_content+1] exclusive scan of seen flags
  uint32_t* barcode_rank;  // [n_barcodes]
  long long* indptr;       // [n_barcodes+1]
};
int run_matrix(DedupBuffers& b, MatrixArgs& m, uint64_t nnz, uint64_t n_mol, uint64_t* n_barcodes_host,
               cudaStream_t st);

int run_owner_partition(const unsigned long long* keys, uint64_t n, int rank_shift, const uint32_t* bounds, int n_parts,
                        unsigned long long* out, unsigned long long* scratch, uint64_t* counts_host, cudaStream_t st);
int run_owner_scatter_peers(const unsigned long long* keys, uint64_t n, int rank_shift, const uint32_t* bounds,
                            int n_parts, unsigned long long* const* peer_buf, unsigned long long* const* peer_cursor,
                            unsigned long long capacity, unsigned long long* d_sent, cudaStream_t st);
int run_owner_scatter_peers_dev(const unsigned long long* keys, const unsigned long long* begin_dev,
                                const unsigned long long* end_dev, uint64_t n_max, int rank_shift,
                                const uint32_t* bounds_dev, int n_parts, unsigned long long* const* peer_buf,
                                unsigned long long* const* peer_cursor, unsigned long long capacity,
                                unsigned long long* d_sent, cudaStream_t st);
// owner ranges balanced by read count, computed on the device into bounds_dev[n_parts + 1]
int run_owner_bounds(const uint32_t* const* vectors, int n_vectors, uint32_t n, int n_parts, unsigned long long* scratch,
                     uint32_t* bounds_dev, cudaStream_t st);
size_t owner_bounds_scratch_bytes(uint32_t n);
int launch_diversity(const uint32_t* counts, uint64_t n, unsigned long long* out4, cudaStream_t st);
int launch_state_counts(const uint32_t* bc_out, uint64_t n, unsigned long long* out4, cudaStream_t st);

struct AnnotateArgs {
  uint64_t n;
  const unsigned long long* select;  // UmiSelectKey word per read, or nullptr (= the global read index, Txomic)
  const uint32_t* bc_out;
  const uint32_t* umi_out;  // raw packed UMI words
  uint32_t* umi_proc;       // out: processed (corrected) UMI words
  const uint32_t* feature;
  uint8_t* flags_out;
  uint32_t lib;
  uint64_t read_base;  // global index of read 0 of this batch
};
// UmiCount rows {column, library, feature, umi, read_count, umi_type}. min_key / rep_raw: the per-distinct-key
// tables of the annotation (nullptr: every molecule Txomic). reorder != 0: the rows are re-sorted into
// (barcode, library, feature, umi) order through sort_a / sort_b (n_mol keys each) - needed only when the key
// order (feature above library) is not already that order.
int run_molecule_rows(DedupBuffers& b, const uint32_t* col_of_rank, uint64_t n_mol, const unsigned long long* min_key,
                      const uint32_t* rep_raw, int reorder, unsigned long long* sort_a, unsigned long long* sort_b,
                      void* sort_temp, size_t sort_temp_bytes, uint32_t* out6, cudaStream_t st);
int run_barcode_summary(DedupBuffers& b, uint64_t n_distinct, uint32_t lib, const uint32_t* barcode_rank,
                        const uint32_t* valid, const uint32_t* col_of_rank, uint64_t n_bc, uint32_t* out4, cudaStream_t st);
int run_annotate_prepare(DedupBuffers& b, uint64_t n_distinct, unsigned long long* min_key, uint32_t* rep_raw,
                         cudaStream_t st);
int run_annotate_min(DedupBuffers& b, uint64_t n_distinct, const AnnotateArgs& a, unsigned long long* min_key,
                     cudaStream_t st);
int run_annotate_final(DedupBuffers& b, uint64_t n_distinct, const AnnotateArgs& a, const unsigned long long* min_key,
                       const uint32_t* rep_raw, unsigned long long* read_stats, cudaStream_t st);

// ---- synth (synth.cu) ----
struct crgpu_synth_params;
int launch_synth(const crgpu_synth_params* p, const uint32_t* d_wl, const uint32_t* d_cell_rank,
                 const uint32_t* d_cell_cdf, const uint32_t* d_n_mol, const uint32_t* d_gene_cdf,
                 const uint32_t* d_fb_cdf, const uint32_t* d_fb_packed, uint64_t start, uint64_t n, uint8_t* r1_seq,
                 uint8_t* r1_qual, uint32_t* feature, uint8_t* r2_seq, uint8_t* r2_qual, cudaStream_t st);
