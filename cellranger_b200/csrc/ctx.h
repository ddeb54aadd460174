// ctx.h — the context behind the opaque crgpu_ctx handle and the small host helpers shared by the translation
// units that implement the C ABI (crgpu.cu: single-device stages; shard.cu: the sharded run). Internal.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/crgpu.h"
#include "kernels.h"


// message of the calling thread's last failure (crgpu_last_error); defined in crgpu.cu
__attribute__((visibility("hidden"))) int fail(int code, const std::string& msg);

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess)                                                                               \
      return fail(CRGPU_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                      \
  } while (0)

#define CHECK_KERNEL()                      \
  do {                                      \
    cudaError_t e__ = cudaPeekAtLastError(); \
    if (e__ != cudaSuccess) return fail(CRGPU_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return CRGPU_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(CRGPU_E_NOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    cap = want;
    return CRGPU_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct HostWhitelist {
  int L = 0;
  uint32_t W = 0;
  bool is_trans = false;
  DevWhitelist dev;
  std::vector<DevBuf> bufs;
};

struct Library {
  crgpu_library_def def;
  DevBuf prior, corrected, valid;  // u32[n_content]
  // feature-barcode table
  std::vector<uint32_t> fb_keys, fb_index;
  DevBuf d_fb_keys, d_fb_index;
};

struct Batch {
  int lib = 0;
  uint64_t n = 0;
  int r1_len = 0, r2_len = 0;
  bool on_device = false;
  // a host batch is copied by the first crgpu_pass1 that sees it, chunk by chunk under the kernels of the
  // chunks before: until then these are the caller's host pointers
  bool host_pending = false;
  const uint8_t *h_r1_seq = nullptr, *h_r1_qual = nullptr, *h_r2_seq = nullptr, *h_r2_qual = nullptr;
  const uint32_t* h_feature = nullptr;
  const unsigned long long* h_select = nullptr;
  const unsigned long long* select = nullptr;  // device: UmiSelectKey word per read, or nullptr
  DevBuf own_select;
  const uint8_t *r1_seq = nullptr, *r1_qual = nullptr, *r2_seq = nullptr, *r2_qual = nullptr;
  const uint32_t* feature = nullptr;  // device pointers (borrowed or owned)
  DevBuf own_seq, own_qual, own_feat, own_r2s, own_r2q;
  DevBuf feature_res;  // resolved features of a feature-barcode batch
  DevBuf inv_idx, inv_bc, inv_nmask, inv_qual;
  uint64_t n_invalid = 0;
  uint64_t base = 0;
};

// fixed slots at the end of the 1024-word counter block
constexpr int CTR_KEYS_PASS1 = 1022;   // keys_total as pass 1 left it (pass 2 restarts from here when re-run)
constexpr int CTR_BAD_FEATURE = 1023;  // reads whose feature index has no row in the matrix

inline int bits_for(uint64_t n_values) {  // bits to hold values 0..n_values-1
  int b = 0;
  while (b < 63 && (1ull << b) < n_values) b++;
  return b;
}


struct crgpu_ctx {
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host -> device copies of read batches, overlapped with pass 1
  cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  double threshold = 0.975;
  double max_expected_errors = 1.7976931348623157e308;
  int filter_umis = 1;
  DevBuf on_target;  // u8[n_on_target]: the panel's target set (crgpu_set_target_filter)
  uint32_t n_on_target = 0;
  uint64_t target_min_reads = 0;
  // crgpu_total_barcode_counts: the last result, host side
  std::vector<uint8_t> tbc_seqs, tbc_valid;
  std::vector<uint64_t> tbc_counts;

  // content space
  int L = 0;
  std::vector<uint32_t> content;  // sorted packed content sequences
  std::vector<HostWhitelist*> wls;
  std::vector<Library*> libs;
  int n_features = 0;
  std::vector<int32_t> feature_type;
  std::vector<uint8_t> fb_seqs;
  int fb_stride = 0;
  bool have_fb = false;
  DevBuf d_fb_counts, d_feat_dist;

  std::vector<Batch*> batches;
  std::vector<Batch*> batch_pool;  // retired batches whose device buffers are reused
  uint64_t n_reads = 0;

  KeyLayout kl{};
  bool layout_ready = false;
  int stage = 0;  // 0 nothing, 1 pass1 done, 2 pass2 done, 3 count done

  DevBuf bc_out, umi_out, umi_proc, flags;
  DevBuf keys, keys_alt, sort_temp;
  DevBuf counters;  // [0] packed scratch, [1] keys_total, [2..] per-batch invalid totals, [CTR_*] below
  unsigned long long* sorted = nullptr;
  uint64_t n_keys = 0;
  bool keys_external = false;
  unsigned long long* key_src = nullptr;  // where crgpu_count finds the keys (nullptr = the keys buffer)

  // dedup
  DevBuf dkeys, c0, best, inc, low, key2, key2_alt, lb_desc, tickets, scalars, ent_rank, ent_feature, ent_count, mol;
  DevBuf col_of_rank, barcode_rank, indptr, mol_rows, min_read, rep_raw, ls_slots, summary, fastq_text, fastq_tmp;
  DevBuf mol_idx, mol_sort, mol_sort_alt;  // distinct-key index of each molecule; scratch of the row reordering
  bool have_select = false;                // some batch carries UmiSelectKey words
  uint64_t n_distinct = 0, n_mol = 0, nnz = 0, n_barcodes = 0;
  uint32_t own_lo = 0, own_hi = 0xFFFFFFFFu;
  bool annotated = false;

  // fused exchange over peer memory
  void* xchg_buf = nullptr;     // this rank's receive buffer (cudaMalloc, exported through CUDA IPC)
  void* xchg_cursor = nullptr;  // u64[2]: keys received, overflow flag
  uint64_t xchg_capacity = 0;
  int xchg_ranks = 0, xchg_rank = -1;
  unsigned long long* peer_buf[CRGPU_MAX_PARTS] = {nullptr};
  unsigned long long* peer_cursor[CRGPU_MAX_PARTS] = {nullptr};
  bool peer_opened[CRGPU_MAX_PARTS] = {false};
  // early part of the exchange (keys of pass 1, sent on a second stream while pass 2 runs)
  cudaStream_t xchg_stream = nullptr;
  cudaEvent_t xchg_ready = nullptr, xchg_done = nullptr;
  bool xchg_early = false;
  uint64_t xchg_early_keys = 0;
  uint32_t xchg_early_bounds[CRGPU_MAX_PARTS + 1] = {0};

  // sharded run inside the library (shard.cu): NCCL communicator, owner bounds, exchange statistics
  void* nccl_comm = nullptr;
  int comm_n = 0, comm_rank = -1;
  DevBuf shard_buf;                      // device scratch: bounds, barrier word, gathered peer records, partial sums
  uint32_t h_bounds[CRGPU_MAX_PARTS + 1] = {0};
  uint64_t shard_sent_remote_keys = 0, shard_received_keys = 0;

  uint64_t stats[CRGPU_STAT_COUNT] = {0};
  uint64_t launches = 0;

  // phase timing (events come from a pool: none is created or destroyed in the steady state)
  std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> phases;
  std::vector<cudaEvent_t> event_pool;
  std::string phase_names;

  // scratch of crgpu_correct_barcodes (grow-only, so that the plugin seam allocates nothing per call)
  DevBuf cb_seq, cb_qual, cb_bc, cb_umi, cb_keys, cb_ctr, cb_idx, cb_ibc, cb_inm, cb_iq;
  std::vector<uint32_t> cb_host;
};

// phase timing on the context stream (crgpu.cu)
__attribute__((visibility("hidden"))) int phase_begin(crgpu_ctx* c, const char* name);
__attribute__((visibility("hidden"))) int phase_end(crgpu_ctx* c);
__attribute__((visibility("hidden"))) void phases_clear(crgpu_ctx* c, const char* prefix);
// frees the communicator and the scratch of the sharded run (shard.cu)
__attribute__((visibility("hidden"))) void shard_release(crgpu_ctx* c);
