// dedup_kernels.cu — UMI correction, low-support filter, dedup and counting on sorted packed keys. sm_100a.
//
// Input: the 64-bit keys (rank | feature | library | umi) of every read that enters dedup, sorted. The
// reference does this per barcode with hash maps (lib/rust/tx_annotation/src/mark_dups.rs); here every
// step is a data-parallel pass over the run-length-encoded key table:
//   c0      raw read count per distinct key                         DupBuilder::observe       :128-155
//   best    argmax (count, umi) over the closed Hamming-1 ball      correct_umis              :19-59
//   c1      counts after moving ONE read per corrected key          BarcodeDupMarker::new     :226-232
//   low     per (barcode, library, umi): ties / sub-maximal genes   determine_low_support_... :87-108
//   c2      counts after moving every read                          BarcodeDupMarker::new     :241-246
//   UMIs    distinct correction targets that are not low support    BarcodeDupMarker::process :280-363
//           → UmiCount rows → per (barcode, feature) counts (cr_types/src/types.rs:180-188)
#include <algorithm>

#include "kernels.h"

#define INC_READS_MASK 0xFFFFFFFFFFull

// ---------------------------------------------------------------------------
// generic ordered stream compaction: single pass, decoupled look-back
// ---------------------------------------------------------------------------
template <int THREADS, int ITEMS, typename Op>
__global__ void __launch_bounds__(THREADS) compact_kernel(Op op, uint64_t n, unsigned long long* desc,
                                                          uint32_t* ticket, unsigned long long* total_out) {
  __shared__ uint32_t scan_s[THREADS / 32 + 1];
  __shared__ unsigned long long bcast;
  __shared__ uint32_t tile_s;
  constexpr int TILE = THREADS * ITEMS;
  if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = tile_s;
  const uint64_t first = (uint64_t)tile * TILE + (uint64_t)threadIdx.x * ITEMS;
  bool f[ITEMS];
  uint32_t cnt = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    uint64_t i = first + k;
    f[k] = (i < n) && op.flag(i);
    cnt += f[k];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan<THREADS>(cnt, &total, scan_s);
  unsigned long long excl = lookback_exclusive(desc, tile, (unsigned long long)total, &bcast);
  uint64_t pos = excl + off;
#pragma unroll
  for (int k = 0; k < ITEMS; k++)
    if (f[k]) op.emit(first + k, pos++);
  const uint64_t n_tiles = (n + TILE - 1) / TILE;
  if (tile == n_tiles - 1 && threadIdx.x == 0) *total_out = excl + total;
}

struct ScanScratch {
  unsigned long long* desc;
  uint32_t* ticket;
  uint64_t desc_cap;
};

template <typename Op>
static int run_compact(const Op& op, uint64_t n, ScanScratch s, unsigned long long* total_out, cudaStream_t st) {
  constexpr int THREADS = 256, ITEMS = 8, TILE = THREADS * ITEMS;
  if (n == 0) {
    cudaMemsetAsync(total_out, 0, 8, st);
    return 0;
  }
  uint64_t tiles = (n + TILE - 1) / TILE;
  cudaMemsetAsync(s.desc, 0, tiles * 8, st);
  cudaMemsetAsync(s.ticket, 0, 4, st);
  compact_kernel<THREADS, ITEMS, Op><<<(unsigned)tiles, THREADS, 0, st>>>(op, n, s.desc, s.ticket, total_out);
  return 1;
}

// ---- run-length encoding of sorted 64-bit keys (shifted right by `shift`) ----
struct RleOp {
  const unsigned long long* keys;
  int shift;
  unsigned long long* out_keys;  // un-shifted key of each run head (may be nullptr)
  uint32_t* out_pos;             // index of each run head
  __device__ bool flag(uint64_t i) const {
    return i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift);
  }
  __device__ void emit(uint64_t i, uint64_t pos) const {
    if (out_keys) out_keys[pos] = keys[i];
    out_pos[pos] = (uint32_t)i;
  }
};

__global__ void run_lengths_kernel(const uint32_t* pos, uint64_t n_runs, uint64_t n_items, uint32_t* len) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n_runs; j += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t nxt = (j + 1 < n_runs) ? pos[j + 1] : n_items;
    len[j] = (uint32_t)(nxt - pos[j]);
  }
}

static inline int grid_for(uint64_t n, int threads = 256, int cap = 148 * 16) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, (uint64_t)cap));
}

// ---------------------------------------------------------------------------
// correct_umis: for each distinct key find argmax (c0, umi) over itself and its Hamming-1 neighbours
// inside the same (rank, feature, library) segment — raw counts, single hop.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool hamming1_2bit(unsigned long long a, unsigned long long b) {
  unsigned long long x = a ^ b;
  unsigned long long y = (x | (x >> 1)) & 0x5555555555555555ull;
  return y != 0ull && (y & (y - 1ull)) == 0ull;
}

__global__ void __launch_bounds__(256) correct_umis_kernel(const unsigned long long* __restrict__ dkeys,
                                                           const uint32_t* __restrict__ c0, uint64_t m, KeyLayout kl,
                                                           uint32_t corr_mask, uint32_t* __restrict__ best,
                                                           unsigned long long* __restrict__ inc,
                                                           unsigned long long* __restrict__ scalars) {
  const int ub = kl.umi_bits;
  const unsigned long long umask = (1ull << ub) - 1ull;
  const uint32_t lmask = (1u << (kl.feature_shift - kl.lib_shift)) - 1u;
  unsigned long long n_corr = 0, n_corr_reads = 0;
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = dkeys[j];
    const unsigned long long seg = key >> ub;
    const uint32_t lib = (uint32_t)(key >> kl.lib_shift) & lmask;
    uint32_t bj = (uint32_t)j;
    if ((corr_mask >> lib) & 1u) {
      uint32_t bcount = c0[j];
      unsigned long long bumi = key & umask;
      constexpr int WINDOW = 48;
      bool big = (j >= WINDOW && (dkeys[j - WINDOW] >> ub) == seg) || (j + WINDOW < m && (dkeys[j + WINDOW] >> ub) == seg);
      if (!big) {
        for (int64_t k = (int64_t)j - 1; k >= 0; k--) {
          unsigned long long o = dkeys[k];
          if ((o >> ub) != seg) break;
          if (hamming1_2bit(o, key)) {
            uint32_t tc = c0[k];
            unsigned long long tu = o & umask;
            if (tc > bcount || (tc == bcount && tu > bumi)) {
              bcount = tc;
              bumi = tu;
              bj = (uint32_t)k;
            }
          }
        }
        for (uint64_t k = j + 1; k < m; k++) {
          unsigned long long o = dkeys[k];
          if ((o >> ub) != seg) break;
          if (hamming1_2bit(o, key)) {
            uint32_t tc = c0[k];
            unsigned long long tu = o & umask;
            if (tc > bcount || (tc == bcount && tu > bumi)) {
              bcount = tc;
              bumi = tu;
              bj = (uint32_t)k;
            }
          }
        }
      } else {
        // large segment: locate it, then binary-search each of the 3L mutants
        uint64_t lo = 0, hi = j;
        const unsigned long long seg_first = seg << ub;
        while (lo < hi) {
          uint64_t mid = (lo + hi) >> 1;
          if (dkeys[mid] < seg_first)
            lo = mid + 1;
          else
            hi = mid;
        }
        const uint64_t s_lo = lo;
        lo = j;
        hi = m;
        while (lo < hi) {
          uint64_t mid = (lo + hi) >> 1;
          if ((dkeys[mid] >> ub) <= seg)
            lo = mid + 1;
          else
            hi = mid;
        }
        const uint64_t s_hi = lo;
        for (int sh = 0; sh < ub; sh += 2) {
          for (unsigned long long d = 1; d < 4; d++) {
            unsigned long long t = key ^ (d << sh);
            uint64_t a = s_lo, b = s_hi;
            while (a < b) {
              uint64_t mid = (a + b) >> 1;
              if (dkeys[mid] < t)
                a = mid + 1;
              else
                b = mid;
            }
            if (a < s_hi && dkeys[a] == t) {
              uint32_t tc = c0[a];
              unsigned long long tu = t & umask;
              if (tc > bcount || (tc == bcount && tu > bumi)) {
                bcount = tc;
                bumi = tu;
                bj = (uint32_t)a;
              }
            }
          }
        }
      }
    }
    best[j] = bj;
    if (bj != (uint32_t)j) {
      uint32_t c = c0[j];
      atomicAdd(inc + bj, (1ull << 40) | (unsigned long long)c);
      n_corr++;
      n_corr_reads += c;
    }
  }
  // block reduction of the statistics
  __shared__ unsigned long long s_a, s_b;
  if (threadIdx.x == 0) {
    s_a = 0;
    s_b = 0;
  }
  __syncthreads();
  if (n_corr) {
    atomicAdd(&s_a, n_corr);
    atomicAdd(&s_b, n_corr_reads);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_a) {
    atomicAdd(scalars + 3, s_a);
    atomicAdd(scalars + 5, s_b);
  }
}

// ---------------------------------------------------------------------------
// low support: regroup by (rank, library, umi) across features
// ---------------------------------------------------------------------------
struct FieldMasks {
  int rbits, fbits, lbits, ubits;
};
__host__ __device__ inline FieldMasks field_masks(const KeyLayout& kl) {
  FieldMasks f;
  f.ubits = kl.umi_bits;
  f.lbits = kl.feature_shift - kl.lib_shift;
  f.fbits = kl.rank_shift - kl.feature_shift;
  f.rbits = kl.total_bits - kl.rank_shift;
  return f;
}

__global__ void make_key2_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m, KeyLayout kl,
                                 unsigned long long* __restrict__ key2) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long k = dkeys[j];
    unsigned long long umi = k & ((1ull << fm.ubits) - 1ull);
    unsigned long long lib = (k >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
    unsigned long long feat = (k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull);
    unsigned long long rank = k >> kl.rank_shift;
    key2[j] = (((rank << fm.lbits | lib) << fm.ubits | umi) << fm.fbits) | feat;
  }
}

__device__ __forceinline__ uint64_t lower_bound_u64(const unsigned long long* a, uint64_t n, unsigned long long v) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) low_support_kernel(const unsigned long long* __restrict__ key2s, uint64_t m,
                                                          KeyLayout kl, const unsigned long long* __restrict__ dkeys,
                                                          const uint32_t* __restrict__ c0,
                                                          const uint32_t* __restrict__ best,
                                                          const unsigned long long* __restrict__ inc,
                                                          uint8_t* __restrict__ low,
                                                          unsigned long long* __restrict__ scalars) {
  const FieldMasks fm = field_masks(kl);
  const unsigned long long fmask = (1ull << fm.fbits) - 1ull;
  unsigned long long n_low = 0;
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long g = key2s[k] >> fm.fbits;
    if (k > 0 && (key2s[k - 1] >> fm.fbits) == g) continue;          // not the group head
    if (k + 1 >= m || (key2s[k + 1] >> fm.fbits) != g) continue;     // a single gene: never low support
    // group head with >= 2 features
    const unsigned long long umi = g & ((1ull << fm.ubits) - 1ull);
    const unsigned long long lib = (g >> fm.ubits) & ((1ull << fm.lbits) - 1ull);
    const unsigned long long rank = g >> (fm.ubits + fm.lbits);
    uint64_t mx = 0, n_at_max = 0;
    for (int pass = 0; pass < 2; pass++) {
      for (uint64_t t = k; t < m && (key2s[t] >> fm.fbits) == g; t++) {
        unsigned long long feat = key2s[t] & fmask;
        unsigned long long pk = (rank << kl.rank_shift) | (feat << kl.feature_shift) | (lib << kl.lib_shift) | umi;
        uint64_t j = lower_bound_u64(dkeys, m, pk);
        uint64_t c1 = (uint64_t)c0[j] - (best[j] != (uint32_t)j ? 1u : 0u) + (inc[j] >> 40);
        if (pass == 0) {
          if (c1 > mx) {
            mx = c1;
            n_at_max = 1;
          } else if (c1 == mx) {
            n_at_max++;
          }
        } else {
          if (n_at_max >= 2 || c1 < mx) {
            low[j] = 1;
            n_low++;
          }
        }
      }
    }
  }
  if (n_low) atomicAdd(scalars + 4, n_low);
}

// ---------------------------------------------------------------------------
// molecules (UmiCount rows) and matrix entries
// ---------------------------------------------------------------------------
struct MolOp {
  const unsigned long long* dkeys;
  const uint32_t* c0;
  const uint32_t* best;
  const unsigned long long* inc;
  const uint8_t* low;
  unsigned long long* mol_key;  // packed key of the molecule (corrected UMI)
  uint32_t* mol_reads;          // c2
  __device__ bool flag(uint64_t j) const {
    bool is_target = best[j] == (uint32_t)j || (inc[j] >> 40) != 0ull;
    return is_target && !low[j];
  }
  __device__ void emit(uint64_t j, uint64_t pos) const {
    mol_key[pos] = dkeys[j];
    mol_reads[pos] = (uint32_t)((best[j] == (uint32_t)j ? c0[j] : 0u) + (inc[j] & INC_READS_MASK));
  }
};

__global__ void entries_kernel(const unsigned long long* __restrict__ mol_key, const uint32_t* __restrict__ pos,
                               uint64_t n_ent, uint64_t n_mol, KeyLayout kl, uint32_t* __restrict__ ent_rank,
                               uint32_t* __restrict__ ent_feature, uint32_t* __restrict__ ent_count) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_ent; e += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t p = pos[e];
    uint64_t nxt = (e + 1 < n_ent) ? pos[e + 1] : n_mol;
    unsigned long long k = mol_key[p];
    ent_rank[e] = (uint32_t)(k >> kl.rank_shift);
    ent_feature[e] = (uint32_t)((k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull));
    ent_count[e] = (uint32_t)(nxt - p);
  }
}

__global__ void low_reads_kernel(const uint32_t* __restrict__ c0, const uint32_t* __restrict__ best,
                                 const uint8_t* __restrict__ low, uint64_t m, unsigned long long* scalars) {
  unsigned long long s = 0;
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x)
    if (low[best[j]]) s += c0[j];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(scalars + 6, s);
}

// Work buffers reused across phases:
//   key2 / key2_alt : sort scratch for the (rank, lib, umi, feature) grouping, later molecule keys
int run_dedup(DedupBuffers& b, uint64_t* n_distinct_host, cudaStream_t st) {
  int launches = 0;
  ScanScratch ss{b.lb_desc, b.tickets, 0};
  cudaMemsetAsync(b.scalars, 0, 16 * 8, st);
  // 1. run-length encode: distinct keys + raw counts (head positions parked in `best`)
  RleOp rle{b.sorted, 0, b.dkeys, b.best};
  launches += run_compact(rle, b.n_keys, ss, b.scalars + 0, st);
  unsigned long long m = 0;
  cudaMemcpyAsync(&m, b.scalars + 0, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  *n_distinct_host = m;
  if (m == 0) return launches;
  run_lengths_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, b.n_keys, b.c0);
  launches++;
  // 2. UMI correction targets + incoming counts
  cudaMemsetAsync(b.inc, 0, m * 8, st);
  cudaMemsetAsync(b.low, 0, m, st);
  correct_umis_kernel<<<grid_for(m, 256, 148 * 32), 256, 0, st>>>(b.dkeys, b.c0, m, b.kl, b.umi_correction_mask,
                                                                  b.best, b.inc, b.scalars);
  launches++;
  // 3. low-support filter: regroup by (rank, library, umi)
  if (b.filter_umis) {
    make_key2_kernel<<<grid_for(m), 256, 0, st>>>(b.dkeys, m, b.kl, b.key2);
    launches++;
    unsigned long long* sorted2 = nullptr;
    launches += sort_keys(b.key2, b.key2_alt, m, b.kl.total_bits, b.sort_temp, b.sort_temp_bytes, &sorted2, st);
    low_support_kernel<<<grid_for(m, 256, 148 * 32), 256, 0, st>>>(sorted2, m, b.kl, b.dkeys, b.c0, b.best, b.inc,
                                                                   b.low, b.scalars);
    launches++;
  }
  low_reads_kernel<<<grid_for(m), 256, 0, st>>>(b.c0, b.best, b.low, m, b.scalars);
  launches++;
  // 4. molecules = correction targets that are not low support (key2 now holds their keys)
  MolOp mol{b.dkeys, b.c0, b.best, b.inc, b.low, b.key2, b.mol};
  launches += run_compact(mol, m, ss, b.scalars + 2, st);
  return launches;
}

// ---------------------------------------------------------------------------
// matrix: barcode index (seen whitelist ranks), CSC arrays, UmiCount rows
// ---------------------------------------------------------------------------
struct SeenOp {
  const uint32_t* vc[CRGPU_MAX_LIBS];
  int n_libs;
  uint32_t lo, hi;
  uint32_t* col_of_rank;
  uint32_t* barcode_rank;
  __device__ bool flag(uint64_t r) const {
    if (r < lo || r >= hi) return false;
    uint32_t s = 0;
    for (int l = 0; l < n_libs; l++) s |= vc[l][r];
    return s != 0u;
  }
  __device__ void emit(uint64_t r, uint64_t pos) const {
    col_of_rank[r] = (uint32_t)pos;
    barcode_rank[pos] = (uint32_t)r;
  }
};

__global__ void column_counts_kernel(const uint32_t* __restrict__ ent_rank, const uint32_t* __restrict__ pos,
                                     uint64_t n_runs, uint64_t n_ent, const uint32_t* __restrict__ col_of_rank,
                                     long long* __restrict__ indptr) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_runs; r += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t p = pos[r];
    uint64_t nxt = (r + 1 < n_runs) ? pos[r + 1] : n_ent;
    indptr[col_of_rank[ent_rank[p]] + 1] = (long long)(nxt - p);
  }
}

struct RleU32Op {
  const uint32_t* keys;
  uint32_t* out_pos;
  __device__ bool flag(uint64_t i) const { return i == 0 || keys[i] != keys[i - 1]; }
  __device__ void emit(uint64_t i, uint64_t pos) const { out_pos[pos] = (uint32_t)i; }
};

// in-place inclusive prefix sum of per-column entry counts (each < 2^32) into int64 offsets:
// single pass, decoupled look-back
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) inclusive_scan_i64_kernel(long long* a, uint64_t n, unsigned long long* desc,
                                                                     uint32_t* ticket) {
  __shared__ uint32_t scan_s[THREADS / 32 + 1];
  __shared__ unsigned long long bcast;
  __shared__ uint32_t tile_s;
  constexpr int TILE = THREADS * ITEMS;
  if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = tile_s;
  const uint64_t first = (uint64_t)tile * TILE + (uint64_t)threadIdx.x * ITEMS;
  uint32_t v[ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    uint64_t i = first + k;
    v[k] = i < n ? (uint32_t)a[i] : 0u;
    sum += v[k];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan<THREADS>(sum, &total, scan_s);
  unsigned long long run = lookback_exclusive(desc, tile, (unsigned long long)total, &bcast) + off;
#pragma unroll
  for (int k = 0; k < ITEMS; k++) {
    uint64_t i = first + k;
    run += v[k];
    if (i < n) a[i] = (long long)run;
  }
}

__global__ void molecules_kernel(const unsigned long long* __restrict__ mol_key, const uint32_t* __restrict__ mol_reads,
                                 uint64_t n_mol, KeyLayout kl, const uint32_t* __restrict__ col_of_rank,
                                 uint32_t* __restrict__ out5) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_mol; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long k = mol_key[i];
    uint32_t rank = (uint32_t)(k >> kl.rank_shift);
    out5[5 * i + 0] = col_of_rank[rank];
    out5[5 * i + 1] = (uint32_t)((k >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull));
    out5[5 * i + 2] = (uint32_t)((k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull));
    out5[5 * i + 3] = (uint32_t)(k & ((1ull << fm.ubits) - 1ull));
    out5[5 * i + 4] = mol_reads[i];
  }
}

// After run_dedup: b.key2 = molecule keys, b.mol = molecule read counts (c2), scalars[2] = n_mol.
// Produces entries (ent_*), the barcode index and indptr. `b.best` is reused for run positions.
int run_matrix(DedupBuffers& b, MatrixArgs& ma, uint64_t /*nnz_unused*/, uint64_t n_mol, uint64_t* n_barcodes_host,
               cudaStream_t st) {
  int launches = 0;
  ScanScratch ss{b.lb_desc, b.tickets, 0};
  // barcode index
  SeenOp seen;
  for (int l = 0; l < CRGPU_MAX_LIBS; l++) seen.vc[l] = ma.valid_counts[l];
  seen.n_libs = ma.n_libs;
  seen.lo = ma.own_lo;
  seen.hi = ma.own_hi;
  seen.col_of_rank = ma.col_of_rank;
  seen.barcode_rank = ma.barcode_rank;
  launches += run_compact(seen, ma.n_content, ss, b.scalars + 7, st);
  // entries = runs of (rank, feature) among the molecules
  uint32_t* run_pos = reinterpret_cast<uint32_t*>(b.key2_alt);
  RleOp rle{b.key2, b.kl.feature_shift, nullptr, run_pos};
  launches += run_compact(rle, n_mol, ss, b.scalars + 1, st);
  unsigned long long h[2] = {0, 0};
  cudaMemcpyAsync(&h[0], b.scalars + 1, 8, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&h[1], b.scalars + 7, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  const uint64_t n_ent = h[0], n_bc = h[1];
  *n_barcodes_host = n_bc;
  cudaMemsetAsync(ma.indptr, 0, (n_bc + 1) * sizeof(long long), st);
  if (n_ent) {
    entries_kernel<<<grid_for(n_ent), 256, 0, st>>>(b.key2, run_pos, n_ent, n_mol, b.kl, b.ent_rank, b.ent_feature,
                                                    b.ent_count);
    launches++;
    // nnz per barcode column = runs of equal rank among the entries
    uint32_t* rank_run_pos = run_pos + n_ent;  // second half of the scratch
    RleU32Op r2{b.ent_rank, rank_run_pos};
    launches += run_compact(r2, n_ent, ss, b.scalars + 8, st);
    unsigned long long n_runs = 0;
    cudaMemcpyAsync(&n_runs, b.scalars + 8, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    column_counts_kernel<<<grid_for(n_runs), 256, 0, st>>>(b.ent_rank, rank_run_pos, n_runs, n_ent, ma.col_of_rank,
                                                          ma.indptr);
    launches++;
  }
  {
    constexpr int THREADS = 256, ITEMS = 8, TILE = THREADS * ITEMS;
    uint64_t tiles = (n_bc + 1 + TILE - 1) / TILE;
    cudaMemsetAsync(ss.desc, 0, tiles * 8, st);
    cudaMemsetAsync(ss.ticket, 0, 4, st);
    inclusive_scan_i64_kernel<THREADS, ITEMS><<<(unsigned)tiles, THREADS, 0, st>>>(ma.indptr, n_bc + 1, ss.desc, ss.ticket);
    launches++;
  }
  return launches;
}

int run_molecule_rows(DedupBuffers& b, const uint32_t* col_of_rank, uint64_t n_mol, uint32_t* out5, cudaStream_t st) {
  if (!n_mol) return 0;
  molecules_kernel<<<grid_for(n_mol), 256, 0, st>>>(b.key2, b.mol, n_mol, b.kl, col_of_rank, out5);
  return 1;
}

// ---------------------------------------------------------------------------
// per-read DupInfo (optional): BarcodeDupMarker::process (mark_dups.rs:280-363)
// ---------------------------------------------------------------------------
__global__ void annotate_prepare_kernel(const uint32_t* __restrict__ best, uint64_t m, uint32_t* __restrict__ min_read,
                                        uint32_t* __restrict__ rep_raw) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    min_read[j] = 0xFFFFFFFFu;
    rep_raw[j] = 0xFFFFFFFFu;
  }
}
// lowest raw UMI corrected onto d that is smaller than d, or any if d is itself corrected (:248-259)
__global__ void annotate_rep_kernel(const uint32_t* __restrict__ best, uint64_t m, uint32_t* __restrict__ rep_raw) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t d = best[j];
    if (d != (uint32_t)j && ((uint32_t)j < d || best[d] != d)) atomicMin(rep_raw + d, (uint32_t)j);
  }
}

__device__ __forceinline__ bool read_key(const AnnotateArgs& a, const KeyLayout& kl, uint64_t i,
                                         unsigned long long* key) {
  uint32_t bw = a.bc_out[i], uw = a.umi_out[i];
  uint32_t f = a.feature ? a.feature[i] : NO_FEATURE;
  uint32_t st = bw >> BC_STATE_SHIFT;
  if (!((st == ST_VALID_BEFORE || st == ST_VALID_AFTER) && (uw & UMI_VALID_BIT) && f != NO_FEATURE)) return false;
  *key = ((unsigned long long)(bw & BC_RANK_MASK) << kl.rank_shift) | ((unsigned long long)f << kl.feature_shift) |
         ((unsigned long long)a.lib << kl.lib_shift) | (unsigned long long)(uw & UMI_SEQ_MASK);
  return true;
}

__global__ void annotate_min_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m, KeyLayout kl,
                                    AnnotateArgs a, uint32_t* __restrict__ min_read) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long key;
    if (!read_key(a, kl, i, &key)) continue;
    uint64_t j = lower_bound_u64(dkeys, m, key);
    atomicMin(min_read + j, (uint32_t)(a.read_base + i));
  }
}

__global__ void annotate_final_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m, KeyLayout kl,
                                      AnnotateArgs a, const uint32_t* __restrict__ best,
                                      const uint8_t* __restrict__ low, const uint32_t* __restrict__ min_read,
                                      const uint32_t* __restrict__ rep_raw) {
  const unsigned long long umask = (1ull << kl.umi_bits) - 1ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long key;
    uint32_t uw = a.umi_out[i];
    uint8_t fl = (uw & UMI_VALID_BIT) ? 1u : 0u;
    if (read_key(a, kl, i, &key)) {
      uint64_t j = lower_bound_u64(dkeys, m, key);
      uint32_t d = best[j];
      bool corrected = d != (uint32_t)j;
      bool is_low = low[d] != 0;
      uint32_t rep = rep_raw[d] == 0xFFFFFFFFu ? d : rep_raw[d];
      bool is_rep = rep == (uint32_t)j && min_read[j] == (uint32_t)(a.read_base + i);
      fl |= 2u | (corrected ? 4u : 0u) | (is_low ? 8u : 0u) | ((!is_low && is_rep) ? 16u : 0u);
      uw = (uw & ~UMI_SEQ_MASK) | (uint32_t)(dkeys[d] & umask);
    }
    a.umi_proc[i] = uw;
    a.flags_out[i] = fl;
  }
}

int run_annotate_prepare(DedupBuffers& b, uint64_t m, uint32_t* min_read, uint32_t* rep_raw, cudaStream_t st) {
  if (!m) return 0;
  annotate_prepare_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, min_read, rep_raw);
  annotate_rep_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, rep_raw);
  return 2;
}
int run_annotate_min(DedupBuffers& b, uint64_t m, const AnnotateArgs& a, uint32_t* min_read, cudaStream_t st) {
  if (!a.n || !m) return 0;
  annotate_min_kernel<<<grid_for(a.n, 256, 148 * 32), 256, 0, st>>>(b.dkeys, m, b.kl, a, min_read);
  return 1;
}
int run_annotate_final(DedupBuffers& b, uint64_t m, const AnnotateArgs& a, const uint32_t* min_read,
                       const uint32_t* rep_raw, unsigned long long*, cudaStream_t st) {
  if (!a.n) return 0;
  annotate_final_kernel<<<grid_for(a.n, 256, 148 * 32), 256, 0, st>>>(b.dkeys, m, b.kl, a, b.best, b.low, min_read,
                                                                     rep_raw);
  return 1;
}
