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
#include <cstdio>

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
  const uint32_t tile = acquire_tile(ticket, &tile_s);
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
  compact_kernel<THREADS, ITEMS, Op><<<(unsigned)tiles, THREADS, 0, st>>>(op, n, s.desc, tile_ticket(s.ticket), total_out);
  return 1;
}

// ---- run-length encoding of sorted 64-bit keys (shifted right by `shift`) ----

// ---- specialised single-pass compactions (vector loads, flags in registers) ----
constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 8;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

// ticket + block scan + look-back shared by the specialised kernels: the thread's offset inside the tile's
// output, the tile's exclusive prefix and its total; writes the grand total once
struct CpPlace {
  uint32_t off, total;
  uint64_t excl;
};
__device__ __forceinline__ CpPlace cp_place(uint32_t cnt, uint32_t tile, uint64_t n_tiles, unsigned long long* desc,
                                            unsigned long long* total_out, uint32_t* scan_s,
                                            unsigned long long* bcast) {
  CpPlace r;
  r.off = block_exclusive_scan<CP_THREADS>(cnt, &r.total, scan_s);
  r.excl = lookback_exclusive(desc, tile, (unsigned long long)r.total, bcast);
  if (tile == n_tiles - 1 && threadIdx.x == 0) *total_out = r.excl + r.total;
  return r;
}

// run-length encoding of sorted keys compared after `>> shift`; heads go to out_pos (and out_keys)
// A chained scan retires at most ~32 tiles per L2 round trip (the look-back window), 60-85 tiles per
// microsecond measured, so the byte rate of a cheap compaction is set by the bytes per tile: RLE_ITEMS = 16.
constexpr int RLE_ITEMS = 16;
constexpr int RLE_TILE = CP_THREADS * RLE_ITEMS;
template <bool WRITE_KEYS>
__global__ void __launch_bounds__(CP_THREADS) rle_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                                         int shift, unsigned long long* __restrict__ out_keys,
                                                         uint32_t* __restrict__ out_pos, unsigned long long* desc,
                                                         uint32_t* ticket, unsigned long long* total_out) {
  __shared__ uint32_t scan_s[CP_THREADS / 32 + 1];
  __shared__ unsigned long long bcast;
  __shared__ uint32_t tile_s;
  const uint32_t tile = acquire_tile(ticket, &tile_s);
  const uint64_t first = (uint64_t)tile * RLE_TILE + (uint64_t)threadIdx.x * RLE_ITEMS;
  unsigned long long k[RLE_ITEMS];
  if (first + RLE_ITEMS <= n) {
    const ulonglong2* v = reinterpret_cast<const ulonglong2*>(keys + first);  // 64-byte aligned
#pragma unroll
    for (int i = 0; i < RLE_ITEMS / 2; i++) {
      ulonglong2 t = __ldcs(v + i);
      k[2 * i] = t.x;
      k[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < RLE_ITEMS; i++) k[i] = first + i < n ? keys[first + i] : 0ull;
  }
  // the key before this thread's first: the neighbour lane's last key, or one global load for lane 0
  unsigned long long prev = __shfl_up_sync(0xFFFFFFFFu, k[RLE_ITEMS - 1], 1);
  if ((threadIdx.x & 31) == 0 && first > 0 && first < n) prev = keys[first - 1];
  bool f[RLE_ITEMS];
  uint32_t cnt = 0;
#pragma unroll
  for (int i = 0; i < RLE_ITEMS; i++) {
    const unsigned long long before = i == 0 ? prev : k[i - 1];
    f[i] = first + i < n && ((first + i == 0) || (k[i] >> shift) != (before >> shift));
    cnt += f[i];
  }
  const uint64_t n_tiles = (n + RLE_TILE - 1) / RLE_TILE;
  const CpPlace pl = cp_place(cnt, tile, n_tiles, desc, total_out, scan_s, &bcast);
  // stage the tile's heads in shared memory so the global stores are contiguous
  __shared__ unsigned long long st_k[WRITE_KEYS ? RLE_TILE : 1];
  __shared__ uint16_t st_p[RLE_TILE];
  uint32_t o = pl.off;
#pragma unroll
  for (int i = 0; i < RLE_ITEMS; i++)
    if (f[i]) {
      if (WRITE_KEYS) st_k[o] = k[i];
      st_p[o] = (uint16_t)(threadIdx.x * RLE_ITEMS + i);
      o++;
    }
  __syncthreads();
  const uint64_t tile_base = (uint64_t)tile * RLE_TILE;
  for (uint32_t i = threadIdx.x; i < pl.total; i += CP_THREADS) {
    if (WRITE_KEYS) out_keys[pl.excl + i] = st_k[i];
    out_pos[pl.excl + i] = (uint32_t)(tile_base + st_p[i]);
  }
}

template <bool WRITE_KEYS>
static int run_rle(const unsigned long long* keys, uint64_t n, int shift, unsigned long long* out_keys,
                   uint32_t* out_pos, unsigned long long* desc, uint32_t* ticket, unsigned long long* total_out,
                   cudaStream_t st) {
  if (n == 0) {
    cudaMemsetAsync(total_out, 0, 8, st);
    return 0;
  }
  uint64_t tiles = (n + RLE_TILE - 1) / RLE_TILE;
  cudaMemsetAsync(desc, 0, tiles * 8, st);
  cudaMemsetAsync(ticket, 0, 4, st);
  rle_kernel<WRITE_KEYS><<<(unsigned)tiles, CP_THREADS, 0, st>>>(keys, n, shift, out_keys, out_pos, desc,
                                                                tile_ticket(ticket), total_out);
  return 1;
}

__global__ void run_lengths_kernel(const uint32_t* pos, uint64_t n_runs, uint64_t n_items, uint32_t* len) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n_runs; j += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t nxt = (j + 1 < n_runs) ? pos[j + 1] : n_items;
    len[j] = (uint32_t)(nxt - pos[j]);
  }
}

// grid for a grid-stride kernel: at most per_sm blocks per SM of the current device
static inline int grid_for(uint64_t n, int threads = 256, int per_sm = 16) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + threads - 1) / threads, (uint64_t)sm_count() * per_sm));
}

void dedup_lb_flag_fetch(unsigned int* host_out, cudaStream_t st) {
  cudaMemcpyFromSymbolAsync(host_out, lb_timeout_flag, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st);
}
void dedup_lb_flag_clear(cudaStream_t st) {
  static const unsigned int zero = 0;
  cudaMemcpyToSymbolAsync(lb_timeout_flag, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st);
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

// candidate update shared by every search strategy: argmax of the tuple (count, umi)
struct BestPick {
  uint32_t count;
  unsigned long long umi;
  uint32_t idx;
  __device__ __forceinline__ void consider(uint32_t tc, unsigned long long tu, uint32_t ti) {
    if (tc > count || (tc == count && tu > umi)) {
      count = tc;
      umi = tu;
      idx = ti;
    }
  }
};

// Tiled kernel: a block answers CU_TILE consecutive keys out of a shared-memory window that also holds
// CU_HALO keys on each side, so every segment of up to CU_HALO keys that touches the tile is complete in
// shared memory. All window keys go into a shared-memory hash set (full 64-bit keys, so a mutant can only
// match a key of its own segment) and into a 64 Kbit presence bitmap. Keys whose segment is a singleton
// need nothing; the others are compacted into a work list and answered densely: each of the 3L mutants
// is first tested against the bitmap (one shared load; almost always absent) and only then probed in the
// hash set. A segment longer than the halo is cut by the window edge: its keys are probed against
// further windows until the whole segment has been seen.
constexpr int CU_THREADS = 512;
constexpr int CU_TILE = 3072;
constexpr int CU_HALO = 512;
constexpr int CU_WIN = CU_TILE + 2 * CU_HALO;  // 4096
constexpr int CU_SLOTS = 8192;                 // 32-bit slots: native shared-memory CAS
constexpr int CU_BITS = 1 << 18;               // presence bitmap (32 KB)
constexpr uint32_t CU_EMPTY = 0xFFFFFFFFu;

// 32-bit hash of a 64-bit key: only the low word changes between a key and its UMI mutants, so the
// contribution of the high word is computed once per key (hi_mix) and each mutant costs one multiply.
__device__ __forceinline__ uint32_t cu_hi_mix(unsigned long long k) { return (uint32_t)(k >> 32) * 0x85EBCA77u; }
__device__ __forceinline__ uint32_t cu_mix(uint32_t lo, uint32_t hi_mix) {
  uint32_t h = (lo * 0x9E3779B1u) ^ hi_mix;
  return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t cu_slot(uint32_t h) { return (h * 0x2C1B3C6Du) >> 19; }  // 13 bits
static_assert(CU_SLOTS == 8192 && CU_BITS == (1 << 18), "hash field widths");
// Presence-bitmap index: XOR-linear in the low key word (an 18-bit fold of it) plus a per-segment term, so
// the index of a mutant is the index of the key XOR a compile-time constant - one instruction per mutant.
// Bit b lives in word b >> 5 at position 31 - (b & 31), which lets a funnel shift bring it to the sign bit.
__device__ __forceinline__ uint32_t cu_fold(uint32_t lo) { return (lo ^ (lo >> 18)) & (CU_BITS - 1); }
__device__ __forceinline__ uint32_t cu_bit_of(uint32_t lo, uint32_t hi_mix) { return cu_fold(lo) ^ (hi_mix >> 14); }

// insert the keys w_key[0..n) that pass `take` into the hash set and the bitmap (tables already cleared)
template <typename Take>
__device__ __forceinline__ void cu_build(const unsigned long long* w_key, int n, uint32_t* table, uint32_t* bitmap,
                                         Take take) {
  for (int i = threadIdx.x; i < n; i += CU_THREADS) {
    const unsigned long long k = w_key[i];
    if (!take(i, k)) continue;
    const uint32_t hm = cu_hi_mix(k);
    const uint32_t b = cu_bit_of((uint32_t)k, hm);
    atomicOr(&bitmap[b >> 5], 0x80000000u >> (b & 31));
    uint32_t sl = cu_slot(cu_mix((uint32_t)k, hm));
    while (atomicCAS(&table[sl], CU_EMPTY, (uint32_t)i) != CU_EMPTY) sl = (sl + 1) & (CU_SLOTS - 1);
  }
}

// probe the 3L mutants of `key`; `hit(idx)` is called with the window index of every neighbour found.
// Two phases so that a warp does not diverge on every mutant: first all mutants are tested against the
// bitmap (straight-line code, one shared load each) into a candidate mask, then only the few candidates
// (bitmap false positives and true neighbours) are looked up in the hash set.
// UB > 0: the UMI width in bits is a compile-time constant and the mutant loop is fully unrolled
// (constant shifts, constant mask bits); UB == 0: generic width `ub`.
template <int UB, typename Hit>
__device__ __forceinline__ void cu_probe(const unsigned long long* w_key, const uint32_t* table,
                                         const uint32_t* bitmap, unsigned long long key, int ub, Hit hit) {
  const uint32_t lo = (uint32_t)key, hm = cu_hi_mix(key);
  // candidate bits are shifted in from the right: after the loop mutant m sits at bit (n_mut - 1 - m)
  uint32_t cand_lo = 0u, cand_hi = 0u;
  const uint32_t bk = cu_bit_of(lo, hm);
  int n_mut;
  if constexpr (UB > 0) {
    n_mut = 3 * (UB / 2);
#pragma unroll
    for (int m = 0; m < 3 * (UB / 2); m++) {
      const uint32_t delta = (uint32_t)(m % 3 + 1) << (2 * (m / 3));
      const uint32_t b = bk ^ ((delta ^ (delta >> 18)) & (CU_BITS - 1));  // constant folded
      const uint32_t t = __funnelshift_l(0u, bitmap[b >> 5], b);           // presence bit -> sign bit
      cand_hi = __funnelshift_l(cand_lo, cand_hi, 1);
      cand_lo = __funnelshift_l(t, cand_lo, 1);
    }
  } else {
    n_mut = 0;
    for (int sh = 0; sh < ub; sh += 2)
      for (uint32_t d = 1; d < 4; d++, n_mut++) {
        const uint32_t b = bk ^ cu_fold(d << sh);
        const uint32_t t = __funnelshift_l(0u, bitmap[b >> 5], b);
        cand_hi = __funnelshift_l(cand_lo, cand_hi, 1);
        cand_lo = __funnelshift_l(t, cand_lo, 1);
      }
  }
  unsigned long long cand = ((unsigned long long)cand_hi << 32) | cand_lo;
  while (cand) {
    const int mi = n_mut - 1 - (__ffsll((long long)cand) - 1);
    cand &= cand - 1ull;
    const uint32_t tlo = lo ^ ((uint32_t)(mi % 3 + 1) << (2 * (mi / 3)));
    const unsigned long long t = (key & 0xFFFFFFFF00000000ull) | tlo;
    uint32_t sl = cu_slot(cu_mix(tlo, hm));
    while (true) {
      const uint32_t idx = table[sl];
      if (idx == CU_EMPTY) break;
      if (w_key[idx] == t) {
        hit(idx);
        break;
      }
      sl = (sl + 1) & (CU_SLOTS - 1);
    }
  }
}

// Segments of at most CU_PAIR keys that lie completely inside the window are answered by comparing the key
// with every other key of its segment (a few XOR / popcount-style tests each); only longer segments, and the
// ones cut by the window edge, use the hash set. Segment bounds come from a bitmap of segment heads.
constexpr int CU_PAIR = 32;  // default; CRGPU_CU_PAIR overrides it for profiling
constexpr int CU_HEAD_WORDS = CU_WIN / 32;

// segment [a, b) of window position i from the head bitmap; false if it is longer than CU_PAIR (or its bounds
// are further than the scan limit)
__device__ __forceinline__ bool cu_small_segment(const uint32_t* heads, int i, int wn, int pair_max, int* a_out,
                                                 int* b_out) {
  const int LIMIT = pair_max / 32 + 1;  // words looked at on each side
  int word = i >> 5;
  const int bit = i & 31;
  uint32_t mk = heads[word] & (0xFFFFFFFFu >> (31 - bit));
  int steps = 0;
  while (mk == 0u) {
    if (++steps > LIMIT || word == 0) return false;
    mk = heads[--word];
  }
  const int a = word * 32 + 31 - __clz(mk);
  word = i >> 5;
  mk = bit == 31 ? 0u : (heads[word] & (0xFFFFFFFEu << bit));
  steps = 0;
  int b = wn;
  while (true) {
    if (mk != 0u) {
      b = word * 32 + __ffs(mk) - 1;
      break;
    }
    if (++steps > LIMIT) return false;
    if (++word >= CU_HEAD_WORDS) break;  // no head up to the end of the window: the segment ends at wn
    mk = heads[word];
  }
  if (b > wn) b = wn;
  *a_out = a;
  *b_out = b;
  return b - a <= pair_max;
}

template <int UB>
__global__ void __launch_bounds__(CU_THREADS, 2) correct_umis_kernel(const unsigned long long* __restrict__ dkeys,
                                                                     const uint32_t* __restrict__ c0, uint64_t m,
                                                                     KeyLayout kl, uint32_t corr_mask,
                                                                     uint32_t* __restrict__ best,
                                                                     unsigned long long* __restrict__ inc,
                                                                     unsigned long long* __restrict__ scalars,
                                                                     int pair_max) {
  extern __shared__ __align__(16) unsigned char cu_smem[];
  unsigned long long* w_key = reinterpret_cast<unsigned long long*>(cu_smem);  // CU_WIN
  uint32_t* table = reinterpret_cast<uint32_t*>(w_key + CU_WIN);                // CU_SLOTS
  uint32_t* bitmap = table + CU_SLOTS;                                          // CU_BITS / 32
  unsigned short* work = reinterpret_cast<unsigned short*>(bitmap + CU_BITS / 32);  // CU_TILE
  unsigned short* w_cnt = work + CU_TILE;  // CU_WIN raw counts of the home window, saturated at 0xFFFF
  __shared__ uint32_t heads[CU_HEAD_WORDS];  // bit i: window key i starts a segment
  __shared__ uint32_t bigs[CU_HEAD_WORDS];   // bit i: window key i belongs to a long or cut segment
  __shared__ uint32_t s_nwork, s_any_big, s_ncut;

  const int ub = kl.umi_bits;
  const unsigned long long umask = (1ull << ub) - 1ull;
  const uint32_t lmask = (1u << (kl.feature_shift - kl.lib_shift)) - 1u;
  const int tid = threadIdx.x;
  const uint64_t q_lo = (uint64_t)blockIdx.x * CU_TILE;
  const uint64_t q_hi = q_lo + CU_TILE < m ? q_lo + CU_TILE : m;
  const uint64_t w_lo = q_lo >= CU_HALO ? q_lo - CU_HALO : 0;
  const uint64_t w_hi = q_hi + CU_HALO < m ? q_hi + CU_HALO : m;
  const int wn = (int)(w_hi - w_lo);

  if (tid == 0) {
    s_nwork = 0;
    s_any_big = 0;
    s_ncut = 0;
  }
  for (int i = tid; i < wn; i += CU_THREADS) {
    w_key[i] = dkeys[w_lo + i];
    const uint32_t c = c0[w_lo + i];
    w_cnt[i] = (unsigned short)(c < 0xFFFFu ? c : 0xFFFFu);
  }
  // raw count of home-window key i: from shared memory unless it saturated the 16-bit cache
  auto count_of = [&](int i) -> uint32_t {
    const uint32_t c = w_cnt[i];
    return c != 0xFFFFu ? c : c0[w_lo + i];
  };
  __syncthreads();
  // is the first / last segment of the window cut by the window edge?
  const unsigned long long seg_l = w_key[0] >> ub, seg_r = w_key[wn - 1] >> ub;
  const bool cut_l = w_lo > 0 && (dkeys[w_lo - 1] >> ub) == seg_l;
  const bool cut_r = w_hi < m && (dkeys[w_hi] >> ub) == seg_r;
  // segment heads (position wn counts as a head so that the last segment is closed)
  for (int i = tid; i < CU_WIN; i += CU_THREADS) {  // warp-uniform trip count: ballots inside
    const bool head = i < wn ? (i == 0 || (w_key[i] >> ub) != (w_key[i - 1] >> ub)) : (i == wn);
    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, head);
    if ((tid & 31) == 0) heads[i >> 5] = mk;
  }
  __syncthreads();

  unsigned long long n_corr = 0, n_corr_reads = 0;
  // classify every window key; tile keys of short segments are answered here, pairwise
  const int q_off = (int)(q_lo - w_lo);
  const int q_end = q_off + (int)(q_hi - q_lo);
  bool any_big = false;
  for (int i = tid; i < CU_WIN; i += CU_THREADS) {
    bool big = false;
    if (i < wn) {
      const unsigned long long key = w_key[i];
      const unsigned long long seg = key >> ub;
      const bool is_cut = (cut_l && seg == seg_l) || (cut_r && seg == seg_r);
      int a = i, b = i + 1;
      const bool small = !is_cut && cu_small_segment(heads, i, wn, pair_max, &a, &b);
      const bool in_tile = i >= q_off && i < q_end;
      const uint32_t lib = (uint32_t)(key >> kl.lib_shift) & lmask;
      const bool correct = ((corr_mask >> lib) & 1u) != 0u;
      big = !small && correct;
      if (in_tile) {
        const uint64_t j = w_lo + i;
        if (!correct || (small && b - a == 1)) {
          best[j] = (uint32_t)j;
        } else if (small) {
          const uint32_t lo = (uint32_t)key;
          BestPick bp{0u, key & umask, (uint32_t)j};
          bool have_own = false;
          for (int t = a; t < b; t++) {
            const uint32_t x = (uint32_t)w_key[t] ^ lo;  // same segment: only UMI bits can differ
            const uint32_t y = (x | (x >> 1)) & 0x55555555u;
            if (y != 0u && (y & (y - 1u)) == 0u) {
              if (!have_own) {
                bp.count = count_of(i);
                have_own = true;
              }
              bp.consider(count_of(t), w_key[t] & umask, (uint32_t)(w_lo + t));
            }
          }
          best[j] = bp.idx;
          if (bp.idx != (uint32_t)j) {
            const uint32_t own = count_of(i);
            atomicAdd(inc + bp.idx, (1ull << 40) | (unsigned long long)own);
            n_corr++;
            n_corr_reads += own;
          }
        } else {
          work[atomicAdd(&s_nwork, 1u)] = (unsigned short)(i | (is_cut ? 0x8000 : 0));
          if (is_cut) s_ncut = 1u;
        }
      }
    }
    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, big);
    if ((tid & 31) == 0) bigs[i >> 5] = mk;
    any_big |= mk != 0u;
  }
  if (any_big && (tid & 31) == 0) s_any_big = 1u;
  __syncthreads();
  const int nwork = (int)s_nwork;
  if (s_any_big) {  // block-uniform
    for (int i = tid; i < CU_SLOTS; i += CU_THREADS) table[i] = CU_EMPTY;
    for (int i = tid; i < CU_BITS / 32; i += CU_THREADS) bitmap[i] = 0u;
    __syncthreads();
    cu_build(w_key, wn, table, bitmap, [&](int i, unsigned long long) { return ((bigs[i >> 5] >> (i & 31)) & 1u) != 0u; });
    __syncthreads();
  }

  for (int w = tid; w < nwork; w += CU_THREADS) {
    const int i = work[w] & 0x7FFF;
    const bool is_cut = (work[w] & 0x8000) != 0;
    const uint64_t j = w_lo + i;
    const unsigned long long key = w_key[i];
    BestPick bp{0u, key & umask, (uint32_t)j};
    bool have_own = false;
    cu_probe<UB>(w_key, table, bitmap, key, ub, [&](uint32_t idx) {
      if (!have_own) {
        bp.count = count_of(i);
        have_own = true;
      }
      bp.consider(count_of((int)idx), w_key[idx] & umask, (uint32_t)(w_lo + idx));
    });
    best[j] = bp.idx;  // provisional for a cut segment: the rest of it is still to come
    if (!is_cut && bp.idx != (uint32_t)j) {
      const uint32_t own = count_of(i);
      atomicAdd(inc + bp.idx, (1ull << 40) | (unsigned long long)own);
      n_corr++;
      n_corr_reads += own;
    }
  }
  // further windows for the cut segments (block-uniform conditions); a window edge that cuts a segment of the
  // halo only - none of the tile's keys - needs nothing
  if ((cut_l || cut_r) && s_ncut) {
    const uint64_t home_lo = w_lo;
    for (int side = 0; side < 2; side++) {
      if (side == 0 ? !cut_l : !cut_r) continue;
      const unsigned long long seg_s = side == 0 ? seg_l : seg_r;
      uint64_t edge = side == 0 ? w_lo : w_hi;  // the window moves away from the home window
      bool more = true;
      while (more) {
        uint64_t a, b;
        if (side == 0) {
          b = edge;
          a = b >= (uint64_t)CU_WIN ? b - CU_WIN : 0;
          more = a > 0 && (dkeys[a - 1] >> ub) == seg_s;
          edge = a;
        } else {
          a = edge;
          b = a + CU_WIN < m ? a + CU_WIN : m;
          more = b < m && (dkeys[b] >> ub) == seg_s;
          edge = b;
        }
        const int xn = (int)(b - a);
        __syncthreads();  // everyone is done with the previous window
        for (int i = tid; i < xn; i += CU_THREADS) w_key[i] = dkeys[a + i];
        for (int i = tid; i < CU_SLOTS; i += CU_THREADS) table[i] = CU_EMPTY;
        for (int i = tid; i < CU_BITS / 32; i += CU_THREADS) bitmap[i] = 0u;
        __syncthreads();
        cu_build(w_key, xn, table, bitmap, [=](int, unsigned long long k) { return (k >> ub) == seg_s; });
        __syncthreads();
        for (int w = tid; w < nwork; w += CU_THREADS) {
          if (!(work[w] & 0x8000)) continue;
          const uint64_t j = home_lo + (work[w] & 0x7FFF);
          const unsigned long long key = dkeys[j];
          if ((key >> ub) != seg_s) continue;
          const uint32_t cur = best[j];
          BestPick bp{c0[cur], dkeys[cur] & umask, cur};
          cu_probe<UB>(w_key, table, bitmap, key, ub, [&](uint32_t idx) {
            bp.consider(c0[a + idx], w_key[idx] & umask, (uint32_t)(a + idx));
          });
          if (bp.idx != cur) best[j] = bp.idx;
        }
      }
    }
    __syncthreads();
    // the cut keys have seen their whole segment: account for the corrected ones
    for (int w = tid; w < nwork; w += CU_THREADS) {
      if (!(work[w] & 0x8000)) continue;
      const uint64_t j = home_lo + (work[w] & 0x7FFF);
      const uint32_t bj = best[j];
      if (bj != (uint32_t)j) {
        const uint32_t own = c0[j];
        atomicAdd(inc + bj, (1ull << 40) | (unsigned long long)own);
        n_corr++;
        n_corr_reads += own;
      }
    }
  }
  // statistics: warp reduction, then one global atomic per warp (64-bit shared-memory atomics are a
  // compare-and-swap loop on this hardware: never use them on a hot address)
  for (int d = 16; d > 0; d >>= 1) {
    n_corr += __shfl_xor_sync(0xFFFFFFFFu, n_corr, d);
    n_corr_reads += __shfl_xor_sync(0xFFFFFFFFu, n_corr_reads, d);
  }
  if ((tid & 31) == 0 && n_corr) {
    atomicAdd(scalars + 3, n_corr);
    atomicAdd(scalars + 5, n_corr_reads);
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

// ---- pre-filter for the (rank, library, umi) regrouping ----
// A UMI is only ever low support when the same (barcode, library, UMI) occurs with two or more
// features, which is rare. Each distinct key hashes its (rank, library, umi) into a table of 2-bit
// slots: the first visitor sets bit 0, any later visitor sets bit 1. Keys whose slot has bit 1 are the
// candidates (all true groups plus hash collisions); only they are sorted and regrouped exactly.
// Slot of a key's (rank, library, umi) group. The table is split into regions of 2^17 slots (32 KB; larger
// regions give fewer false candidates but measured slower - locality wins): the
// region is chosen by the barcode rank alone, so the keys of one barcode - which are neighbours in the
// sorted table and therefore processed together - stay inside one L2-resident region.
static int ls_region_bits_host() { return getenv("CRGPU_LS_REGION") ? atoi(getenv("CRGPU_LS_REGION")) : 17; }
__device__ __forceinline__ unsigned long long group_slot(unsigned long long key, const KeyLayout& kl,
                                                         const FieldMasks& fm, int slot_bits, int LS_REGION_BITS) {
  unsigned long long umi = key & ((1ull << fm.ubits) - 1ull);
  unsigned long long lib = (key >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
  unsigned long long rank = key >> kl.rank_shift;
  unsigned long long g = (lib << fm.ubits) | umi;
  g ^= g >> 15;
  g *= 0x9E3779B97F4A7C15ull;
  g ^= g >> 29;
  unsigned long long r = (rank + 0x632BE59BD9B4E019ull) * 0xBF58476D1CE4E5B9ull;
  r ^= r >> 31;
  if (slot_bits <= LS_REGION_BITS) return (g ^ r) >> (64 - slot_bits);
  const int rbits = slot_bits - LS_REGION_BITS;
  return ((r >> (64 - rbits)) << LS_REGION_BITS) | ((g ^ (r << 7)) >> (64 - LS_REGION_BITS));
}

__global__ void __launch_bounds__(256) ls_mark_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m,
                                                      KeyLayout kl, uint32_t* __restrict__ slots, int slot_bits,
                                                      int region_bits) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long h = group_slot(dkeys[j], kl, fm, slot_bits, region_bits);
    uint32_t bit = 1u << (2 * (h & 15));
    uint32_t old = atomicOr(slots + (h >> 4), bit);
    if (old & bit) atomicOr(slots + (h >> 4), bit << 1);
  }
}

__global__ void __launch_bounds__(256) ls_collect_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m,
                                                         KeyLayout kl, const uint32_t* __restrict__ slots,
                                                         int slot_bits, int region_bits,
                                                         unsigned long long* __restrict__ cand,
                                                         unsigned long long* __restrict__ n_cand) {
  __shared__ uint32_t scan_s[9];
  __shared__ unsigned long long base_s;
  const FieldMasks fm = field_masks(kl);
  const uint64_t n_blocks_work = (m + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < n_blocks_work; blk += gridDim.x) {
    uint64_t j = blk * 256 + threadIdx.x;
    bool hit = false;
    unsigned long long k2 = 0;
    if (j < m) {
      unsigned long long k = dkeys[j];
      unsigned long long h = group_slot(k, kl, fm, slot_bits, region_bits);
      hit = (slots[h >> 4] >> (2 * (h & 15) + 1)) & 1u;
      if (hit) {
        unsigned long long umi = k & ((1ull << fm.ubits) - 1ull);
        unsigned long long lib = (k >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
        unsigned long long feat = (k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull);
        unsigned long long rank = k >> kl.rank_shift;
        k2 = (((rank << fm.lbits | lib) << fm.ubits | umi) << fm.fbits) | feat;
      }
    }
    uint32_t tot;
    uint32_t off = block_exclusive_scan<256>(hit ? 1u : 0u, &tot, scan_s);
    if (threadIdx.x == 0) base_s = tot ? atomicAdd(n_cand, (unsigned long long)tot) : 0ull;
    __syncthreads();
    if (hit) cand[base_s + off] = k2;
    __syncthreads();
  }
}

// ---- the same pre-filter in shared memory ----
// The (rank, library, umi) groups never cross a barcode, and the distinct-key table is sorted by barcode: a block
// takes the barcodes that START in its tile of the table (the last one may run far past the tile; a tile in the
// middle of a long barcode has nothing to do), marks their keys in a 2-bit filter in shared memory, and reads the
// keys a second time (out of L2) to collect the ones whose slot was visited twice. No global atomics but one per
// warp of candidates, no 150 MB slot table to clear and to miss in L2. 32 slots per key of the range (at most
// LF_MAX_SLOTS): about 3 % of the keys are false candidates, the exact regrouping downstream drops them.
constexpr int LF_THREADS = 512;
constexpr int LF_TILE = 4096;
constexpr int LF_MAX_SLOTS = 1 << 18;  // 2 bits each: 64 KB
constexpr int LF_MIN_SLOTS = 1 << 12;
constexpr int LF_SUPER = 32768;  // keys whose candidates are collected behind one global atomic

__device__ __forceinline__ uint32_t lf_hash(unsigned long long key, const KeyLayout& kl, const FieldMasks& fm) {
  const unsigned long long umi = key & ((1ull << fm.ubits) - 1ull);
  const unsigned long long lib = (key >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
  const unsigned long long rank = key >> kl.rank_shift;
  unsigned long long g = ((rank << fm.lbits | lib) << fm.ubits) | umi;
  g ^= g >> 31;
  g *= 0x9E3779B97F4A7C15ull;
  g ^= g >> 29;
  g *= 0xBF58476D1CE4E5B9ull;
  return (uint32_t)(g >> 32);
}

__global__ void __launch_bounds__(LF_THREADS) ls_filter_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m,
                                                               KeyLayout kl, unsigned long long* __restrict__ cand,
                                                               unsigned long long* __restrict__ n_cand) {
  extern __shared__ uint32_t lf_slots[];  // LF_MAX_SLOTS / 16 words
  __shared__ unsigned long long s_lo, s_hi, s_base;
  __shared__ uint32_t s_bits[LF_SUPER / 32], s_scan[LF_THREADS / 32 + 1], s_total;
  const FieldMasks fm = field_masks(kl);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t t_lo = (uint64_t)blockIdx.x * LF_TILE;
  const uint64_t t_hi = t_lo + LF_TILE < m ? t_lo + LF_TILE : m;
  if (tid == 0) {
    s_lo = ~0ull;
    s_hi = ~0ull;
  }
  __syncthreads();
  // first barcode start inside the tile
  {
    unsigned long long first = ~0ull;
    for (uint64_t j = t_lo + tid; j < t_hi; j += LF_THREADS) {
      const bool start = j == 0 || (dkeys[j] >> kl.rank_shift) != (dkeys[j - 1] >> kl.rank_shift);
      if (start && j < first) first = j;
    }
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, first, d);
      first = o < first ? o : first;
    }
    if (lane == 0 && first != ~0ull) atomicMin(&s_lo, first);
  }
  __syncthreads();
  const uint64_t lo = s_lo;
  if (lo == ~0ull) return;  // the tile lies inside a barcode that an earlier tile owns (block-uniform)
  // end of the barcode that holds the tile's last key
  uint64_t hi = m;
  if (t_hi < m) {
    const unsigned long long last_rank = dkeys[t_hi - 1] >> kl.rank_shift;
    for (uint64_t base = t_hi;; base += LF_THREADS) {
      const uint64_t j = base + tid;
      const bool differs = j >= m || (dkeys[j] >> kl.rank_shift) != last_rank;
      const uint32_t mk = __ballot_sync(0xFFFFFFFFu, differs);
      if (mk && lane == 0) atomicMin(&s_hi, (unsigned long long)(base + (tid & ~31) + (__ffs(mk) - 1)));
      __syncthreads();
      const unsigned long long found = s_hi;
      if (found != ~0ull) {
        hi = found < m ? found : m;
        break;
      }
      __syncthreads();
    }
  }
  const uint64_t range = hi - lo;
  uint32_t n_slots = LF_MIN_SLOTS;
  while (n_slots < (uint32_t)LF_MAX_SLOTS && (uint64_t)n_slots < 32ull * range) n_slots <<= 1;
  for (uint32_t i = tid; i < n_slots / 16; i += LF_THREADS) lf_slots[i] = 0u;
  __syncthreads();
  for (uint64_t j = lo + tid; j < hi; j += LF_THREADS) {
    const uint32_t h = lf_hash(dkeys[j], kl, fm) & (n_slots - 1u);
    const uint32_t bit = 1u << (2u * (h & 15u));
    const uint32_t old = atomicOr(&lf_slots[h >> 4], bit);
    if (old & bit) atomicOr(&lf_slots[h >> 4], bit << 1);
  }
  __syncthreads();
  // Collect, LF_SUPER keys at a time: hit flags into a bitmap, ONE global atomic for the block's candidates of the
  // super-chunk (a global atomic per warp of candidates - 2.4 M on one address - was 45 % of this kernel's stalls),
  // then every thread writes the candidates of its two bitmap words behind the block's prefix.
  for (uint64_t sc_lo = lo; sc_lo < hi; sc_lo += LF_SUPER) {
    const uint32_t sc_n = (uint32_t)((hi - sc_lo) < (uint64_t)LF_SUPER ? (hi - sc_lo) : (uint64_t)LF_SUPER);
    if (tid == 0) s_total = 0u;
    __syncthreads();
    for (uint32_t r = 0; r < LF_SUPER / LF_THREADS; r++) {  // warp-uniform trip count: ballots inside
      const uint32_t o = r * LF_THREADS + tid;
      bool hit = false;
      if (o < sc_n) {
        const uint32_t h = lf_hash(dkeys[sc_lo + o], kl, fm) & (n_slots - 1u);
        hit = (lf_slots[h >> 4] >> (2u * (h & 15u) + 1u)) & 1u;
      }
      const uint32_t mk = __ballot_sync(0xFFFFFFFFu, hit);
      if (lane == 0) {
        s_bits[o >> 5] = mk;
        if (mk) atomicAdd(&s_total, (uint32_t)__popc(mk));
      }
      if ((r + 1u) * LF_THREADS >= sc_n) break;  // block-uniform: the rest of the super-chunk is past the range
    }
    __syncthreads();
    if (tid == 0) s_base = s_total ? atomicAdd(n_cand, (unsigned long long)s_total) : 0ull;
    // exclusive prefix of the candidate counts per pair of bitmap words
    constexpr int WPT = LF_SUPER / 32 / LF_THREADS;  // 2 words per thread
    uint32_t wbits[WPT], cnt = 0;
#pragma unroll
    for (int k = 0; k < WPT; k++) {
      const uint32_t w = tid * WPT + k;
      wbits[k] = w * 32u < sc_n ? s_bits[w] : 0u;
      cnt += (uint32_t)__popc(wbits[k]);
    }
    uint32_t tot;
    uint32_t off = block_exclusive_scan<LF_THREADS>(cnt, &tot, s_scan);  // syncs: s_base is visible behind it
    unsigned long long out = s_base + off;
#pragma unroll
    for (int k = 0; k < WPT; k++) {
      uint32_t bits = wbits[k];
      while (bits) {
        const uint32_t b = (uint32_t)__ffs(bits) - 1u;
        bits &= bits - 1u;
        const unsigned long long key = dkeys[sc_lo + (uint64_t)(tid * WPT + k) * 32u + b];
        const unsigned long long umi = key & ((1ull << fm.ubits) - 1ull);
        const unsigned long long lib = (key >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
        const unsigned long long feat = (key >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull);
        const unsigned long long rank = key >> kl.rank_shift;
        cand[out++] = (((rank << fm.lbits | lib) << fm.ubits | umi) << fm.fbits) | feat;
      }
    }
    __syncthreads();
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

__global__ void __launch_bounds__(256) low_support_kernel(const unsigned long long* __restrict__ key2s, uint64_t n2,
                                                          uint64_t m, KeyLayout kl,
                                                          const unsigned long long* __restrict__ dkeys,
                                                          const uint32_t* __restrict__ c0,
                                                          const uint32_t* __restrict__ best,
                                                          const unsigned long long* __restrict__ inc,
                                                          uint8_t* __restrict__ low,
                                                          unsigned long long* __restrict__ scalars) {
  const FieldMasks fm = field_masks(kl);
  const unsigned long long fmask = (1ull << fm.fbits) - 1ull;
  unsigned long long n_low = 0;
  for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n2; k += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long g = key2s[k] >> fm.fbits;
    if (k > 0 && (key2s[k - 1] >> fm.fbits) == g) continue;          // not the group head
    if (k + 1 >= n2 || (key2s[k + 1] >> fm.fbits) != g) continue;     // a single gene: never low support
    // group head with >= 2 features
    const unsigned long long umi = g & ((1ull << fm.ubits) - 1ull);
    const unsigned long long lib = (g >> fm.ubits) & ((1ull << fm.lbits) - 1ull);
    const unsigned long long rank = g >> (fm.ubits + fm.lbits);
    uint64_t mx = 0, n_at_max = 0;
    for (int pass = 0; pass < 2; pass++) {
      for (uint64_t t = k; t < n2 && (key2s[t] >> fm.fbits) == g; t++) {
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
// Targeted-panel filter (BarcodeDupMarker::process, mark_dups.rs:311-320): a correction target on a feature of the
// target set with read_count (c2) below the threshold, and not low support, is no UMI count. Bit 1 of low[].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) target_filter_kernel(const unsigned long long* __restrict__ dkeys,
                                                            const uint32_t* __restrict__ c0,
                                                            const uint32_t* __restrict__ best,
                                                            const unsigned long long* __restrict__ inc, uint64_t m,
                                                            KeyLayout kl, const uint8_t* __restrict__ on_target,
                                                            uint32_t n_on_target, unsigned long long min_reads,
                                                            uint8_t* __restrict__ low, unsigned long long* n_filtered) {
  const FieldMasks fm = field_masks(kl);
  unsigned long long mine = 0;
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    const bool self = best[j] == (uint32_t)j;
    const unsigned long long in = inc[j];
    if (!(self || (in >> 40) != 0ull) || (low[j] & 1)) continue;
    const uint32_t feature = (uint32_t)((dkeys[j] >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull));
    const unsigned long long reads = (self ? c0[j] : 0u) + (in & INC_READS_MASK);
    if (feature < n_on_target && on_target[feature] && reads < min_reads) {
      low[j] |= 2;
      mine++;
    }
  }
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_filtered, mine);
}

// ---------------------------------------------------------------------------
// molecules (UmiCount rows) and matrix entries
// ---------------------------------------------------------------------------
// molecules = correction targets that are not low support: key + read count (c2), compacted in order
__global__ void __launch_bounds__(CP_THREADS) molecules_compact_kernel(
    const unsigned long long* __restrict__ dkeys, const uint32_t* __restrict__ c0, const uint32_t* __restrict__ best,
    const unsigned long long* __restrict__ inc, const uint8_t* __restrict__ low, uint64_t m,
    unsigned long long* __restrict__ mol_key, uint32_t* __restrict__ mol_reads, uint32_t* __restrict__ mol_idx,
    unsigned long long* desc, uint32_t* ticket, unsigned long long* total_out, unsigned long long* low_reads_out) {
  __shared__ uint32_t scan_s[CP_THREADS / 32 + 1];
  __shared__ unsigned long long bcast;
  __shared__ uint32_t tile_s;
  const uint32_t tile = acquire_tile(ticket, &tile_s);
  const uint64_t first = (uint64_t)tile * CP_TILE + (uint64_t)threadIdx.x * CP_ITEMS;
  uint32_t b[CP_ITEMS];
  unsigned long long in[CP_ITEMS];
  uint8_t lw[CP_ITEMS];
  if (first + CP_ITEMS <= m) {
    const uint4* vb = reinterpret_cast<const uint4*>(best + first);
    const ulonglong2* vi = reinterpret_cast<const ulonglong2*>(inc + first);
    const uint2 vl = *reinterpret_cast<const uint2*>(low + first);
#pragma unroll
    for (int i = 0; i < CP_ITEMS / 4; i++) {
      uint4 t = vb[i];
      b[4 * i] = t.x, b[4 * i + 1] = t.y, b[4 * i + 2] = t.z, b[4 * i + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < CP_ITEMS / 2; i++) {
      ulonglong2 t = vi[i];
      in[2 * i] = t.x, in[2 * i + 1] = t.y;
    }
#pragma unroll
    for (int i = 0; i < CP_ITEMS; i++) lw[i] = (uint8_t)((i < 4 ? vl.x : vl.y) >> (8 * (i & 3)));
  } else {
#pragma unroll
    for (int i = 0; i < CP_ITEMS; i++) {
      const bool ok = first + i < m;
      b[i] = ok ? best[first + i] : 0u;
      in[i] = ok ? inc[first + i] : 0ull;
      lw[i] = ok ? low[first + i] : (uint8_t)1;
    }
  }
  bool f[CP_ITEMS];
  uint32_t cnt = 0;
  unsigned long long low_reads = 0;  // reads whose corrected key is low support (is_low_support_umi)
#pragma unroll
  for (int i = 0; i < CP_ITEMS; i++) {
    const uint64_t j = first + i;
    const bool self = b[i] == (uint32_t)j;
    const bool is_target = self || (in[i] >> 40) != 0ull;
    f[i] = j < m && is_target && !lw[i];
    cnt += f[i];
    if (j < m && (self ? (lw[i] & 1) != 0 : (low[b[i]] & 1) != 0)) low_reads += c0[j];
  }
  for (int d = 16; d > 0; d >>= 1) low_reads += __shfl_xor_sync(0xFFFFFFFFu, low_reads, d);
  if ((threadIdx.x & 31) == 0 && low_reads) atomicAdd(low_reads_out, low_reads);
  const uint64_t n_tiles = (m + CP_TILE - 1) / CP_TILE;
  const CpPlace pl = cp_place(cnt, tile, n_tiles, desc, total_out, scan_s, &bcast);
  __shared__ unsigned long long st_k[CP_TILE];
  __shared__ uint32_t st_r[CP_TILE];
  __shared__ uint16_t st_i[CP_TILE];  // position of the molecule's distinct key inside the tile
  uint32_t o = pl.off;
#pragma unroll
  for (int i = 0; i < CP_ITEMS; i++)
    if (f[i]) {
      const uint64_t j = first + i;
      st_k[o] = dkeys[j];
      st_r[o] = (uint32_t)((b[i] == (uint32_t)j ? c0[j] : 0u) + (in[i] & INC_READS_MASK));
      st_i[o] = (uint16_t)(threadIdx.x * CP_ITEMS + i);
      o++;
    }
  __syncthreads();
  const uint64_t tile_first = (uint64_t)tile * CP_TILE;
  for (uint32_t i = threadIdx.x; i < pl.total; i += CP_THREADS) {
    mol_key[pl.excl + i] = st_k[i];
    mol_reads[pl.excl + i] = st_r[i];
    if (mol_idx) mol_idx[pl.excl + i] = (uint32_t)(tile_first + st_i[i]);
  }
}

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

// ---------------------------------------------------------------------------
// barcode-owner partition of the local keys (multi-GPU): keys of owner p = content ranks
// [bounds[p], bounds[p+1]) end up contiguous, in owner order; order inside an owner is irrelevant
// (the receiver sorts). One counting pass, one scatter pass.
// ---------------------------------------------------------------------------
struct OwnerBounds {
  uint32_t b[CRGPU_MAX_PARTS + 1];
  int n;
};
__device__ __forceinline__ int owner_of(const OwnerBounds& ob, uint32_t rank) {
  int p = 0;
#pragma unroll
  for (int i = 1; i < CRGPU_MAX_PARTS; i++)
    if (i < ob.n && rank >= ob.b[i]) p = i;
  return p;
}
__global__ void __launch_bounds__(256) owner_count_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                                          int rank_shift, OwnerBounds ob,
                                                          unsigned long long* __restrict__ counts) {
  __shared__ uint32_t s_cnt[CRGPU_MAX_PARTS];
  if (threadIdx.x < CRGPU_MAX_PARTS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    int p = owner_of(ob, (uint32_t)(keys[i] >> rank_shift));
    uint32_t peers = __match_any_sync(__activemask(), p);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[p], (uint32_t)__popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < ob.n && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}
// cursors[p] starts at the exclusive prefix of counts; blocks claim ranges per owner
__global__ void __launch_bounds__(256) owner_scatter_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                                            int rank_shift, OwnerBounds ob,
                                                            unsigned long long* __restrict__ cursors,
                                                            unsigned long long* __restrict__ out) {
  __shared__ uint32_t s_cnt[CRGPU_MAX_PARTS];
  __shared__ unsigned long long s_base[CRGPU_MAX_PARTS];
  const uint64_t n_blocks_work = (n + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < n_blocks_work; blk += gridDim.x) {
    if (threadIdx.x < CRGPU_MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = blk * 256 + threadIdx.x;
    int p = -1;
    uint32_t off = 0;
    unsigned long long k = 0;
    if (i < n) {
      k = keys[i];
      p = owner_of(ob, (uint32_t)(k >> rank_shift));
    }
    {
      uint32_t peers = __match_any_sync(0xFFFFFFFFu, p);
      int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (p >= 0 && (int)(threadIdx.x & 31) == leader) base = atomicAdd(&s_cnt[p], (uint32_t)__popc(peers));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      off = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < ob.n)
      s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(cursors + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
    __syncthreads();
    if (p >= 0) out[s_base[p] + off] = k;
    __syncthreads();
  }
}

// counts[0..n_parts) <- keys per owner; out <- keys grouped by owner. scratch: 2 * CRGPU_MAX_PARTS u64.
int run_owner_partition(const unsigned long long* keys, uint64_t n, int rank_shift, const uint32_t* bounds, int n_parts,
                        unsigned long long* out, unsigned long long* scratch, uint64_t* counts_host, cudaStream_t st) {
  OwnerBounds ob;
  ob.n = n_parts;
  for (int i = 0; i <= CRGPU_MAX_PARTS; i++) ob.b[i] = i <= n_parts ? bounds[i] : 0xFFFFFFFFu;
  unsigned long long* d_counts = scratch;
  unsigned long long* d_cursors = scratch + CRGPU_MAX_PARTS;
  cudaMemsetAsync(scratch, 0, 2 * CRGPU_MAX_PARTS * 8, st);
  int launches = 0;
  if (n) {
    owner_count_kernel<<<grid_for(n), 256, 0, st>>>(keys, n, rank_shift, ob, d_counts);
    launches++;
  }
  unsigned long long h[CRGPU_MAX_PARTS] = {0};
  cudaMemcpyAsync(h, d_counts, n_parts * 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  unsigned long long cur[CRGPU_MAX_PARTS] = {0}, run = 0;
  for (int p = 0; p < n_parts; p++) {
    counts_host[p] = h[p];
    cur[p] = run;
    run += h[p];
  }
  cudaMemcpyAsync(d_cursors, cur, n_parts * 8, cudaMemcpyHostToDevice, st);
  if (n) {
    owner_scatter_kernel<<<grid_for(n), 256, 0, st>>>(keys, n, rank_shift, ob, d_cursors, out);
    launches++;
  }
  cudaStreamSynchronize(st);  // `cur` is a local
  return launches;
}

// ---------------------------------------------------------------------------
// Fused exchange: the same grouping by owner, but every key is stored straight into the owner GPU's
// receive buffer over NVLink (peer pointers mapped through CUDA IPC). A block takes a chunk of PX_CHUNK
// keys, orders it by owner in shared memory, claims one range per owner with a single remote atomicAdd
// on that owner's cursor and writes each owner's run with coalesced peer stores.
// ---------------------------------------------------------------------------
constexpr int PX_THREADS = 256;
constexpr int PX_ITEMS = 16;
constexpr int PX_CHUNK = PX_THREADS * PX_ITEMS;  // 4096 keys = 32 KB

struct PeerTargets {
  unsigned long long* buf[CRGPU_MAX_PARTS];
  unsigned long long* cursor[CRGPU_MAX_PARTS];  // [0] keys received so far, [1] overflow flag
  unsigned long long capacity;
};

__device__ __forceinline__ void owner_scatter_peers_body(const unsigned long long* __restrict__ keys, uint64_t n,
                                                         int rank_shift, const OwnerBounds& ob, const PeerTargets& pt,
                                                         unsigned long long* __restrict__ sent) {
  __shared__ unsigned long long s_keys[PX_CHUNK];
  __shared__ uint32_t s_cnt[CRGPU_MAX_PARTS], s_off[CRGPU_MAX_PARTS + 1], s_fill[CRGPU_MAX_PARTS];
  __shared__ unsigned long long s_base[CRGPU_MAX_PARTS];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t n_chunks = (n + PX_CHUNK - 1) / PX_CHUNK;
  for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    if (tid < CRGPU_MAX_PARTS) {
      s_cnt[tid] = 0;
      s_fill[tid] = 0;
    }
    __syncthreads();
    const uint64_t first = chunk * PX_CHUNK;
    unsigned long long k[PX_ITEMS];
    int own[PX_ITEMS];
#pragma unroll
    for (int i = 0; i < PX_ITEMS; i++) {
      const uint64_t g = first + (uint64_t)i * PX_THREADS + tid;
      own[i] = -1;
      if (g < n) {
        k[i] = __ldcs(keys + g);
        own[i] = owner_of(ob, (uint32_t)(k[i] >> rank_shift));
      }
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, own[i]);
      if (own[i] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[own[i]], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t run = 0;
      for (int p = 0; p < ob.n; p++) {
        s_off[p] = run;
        run += s_cnt[p];
      }
      s_off[ob.n] = run;
    }
    // one remote claim per owner and chunk
    if (tid < ob.n && s_cnt[tid]) {
      unsigned long long b = atomicAdd(pt.cursor[tid], (unsigned long long)s_cnt[tid]);
      if (b + s_cnt[tid] > pt.capacity) {
        atomicExch(pt.cursor[tid] + 1, 1ull);  // the receiver reports the overflow
        b = ~0ull;
      }
      s_base[tid] = b;
      atomicAdd(sent + tid, (unsigned long long)s_cnt[tid]);
    }
    __syncthreads();
    // order the chunk by owner in shared memory
#pragma unroll
    for (int i = 0; i < PX_ITEMS; i++) {
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, own[i]);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (own[i] >= 0 && lane == leader) base = atomicAdd(&s_fill[own[i]], (uint32_t)__popc(peers));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (own[i] >= 0) s_keys[s_off[own[i]] + base + __popc(peers & ((1u << lane) - 1u))] = k[i];
    }
    __syncthreads();
    // coalesced peer stores, one owner run after the other
    const uint32_t total = s_off[ob.n];
    for (uint32_t p = tid; p < total; p += PX_THREADS) {
      int o = 0;
      while (p >= s_off[o + 1]) o++;
      const unsigned long long b = s_base[o];
      if (b != ~0ull) pt.buf[o][b + (p - s_off[o])] = s_keys[p];
    }
    __syncthreads();
  }
  // the keys were stored into other GPUs' memory: make them visible system-wide before the kernel retires (the
  // owners read them after a cross-rank barrier that is ordered behind this kernel)
  __threadfence_system();
}

__global__ void __launch_bounds__(PX_THREADS) owner_scatter_peers_kernel(const unsigned long long* __restrict__ keys,
                                                                        uint64_t n, int rank_shift, OwnerBounds ob,
                                                                        PeerTargets pt,
                                                                        unsigned long long* __restrict__ sent) {
  owner_scatter_peers_body(keys, n, rank_shift, ob, pt, sent);
}

// The same with everything that used to need the host read on the device: the key range [*begin_dev, *end_dev)
// (begin_dev may be null = 0) and the owner bounds, so that the sharded step enqueues it without a round trip.
__global__ void __launch_bounds__(PX_THREADS) owner_scatter_peers_dev_kernel(
    const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ begin_dev,
    const unsigned long long* __restrict__ end_dev, int rank_shift, const uint32_t* __restrict__ bounds_dev, int n_parts,
    PeerTargets pt, unsigned long long* __restrict__ sent) {
  __shared__ OwnerBounds ob;
  if (threadIdx.x <= CRGPU_MAX_PARTS) ob.b[threadIdx.x] = threadIdx.x <= n_parts ? bounds_dev[threadIdx.x] : 0xFFFFFFFFu;
  if (threadIdx.x == 0) ob.n = n_parts;
  __syncthreads();
  const unsigned long long first = begin_dev ? *begin_dev : 0ull;
  const unsigned long long last = *end_dev;
  owner_scatter_peers_body(keys + first, last > first ? last - first : 0ull, rank_shift, ob, pt, sent);
}

int run_owner_scatter_peers(const unsigned long long* keys, uint64_t n, int rank_shift, const uint32_t* bounds,
                            int n_parts, unsigned long long* const* peer_buf, unsigned long long* const* peer_cursor,
                            unsigned long long capacity, unsigned long long* d_sent, cudaStream_t st) {
  OwnerBounds ob;
  ob.n = n_parts;
  for (int i = 0; i <= CRGPU_MAX_PARTS; i++) ob.b[i] = i <= n_parts ? bounds[i] : 0xFFFFFFFFu;
  PeerTargets pt;
  for (int i = 0; i < CRGPU_MAX_PARTS; i++) {
    pt.buf[i] = i < n_parts ? peer_buf[i] : nullptr;
    pt.cursor[i] = i < n_parts ? peer_cursor[i] : nullptr;
  }
  pt.capacity = capacity;
  cudaMemsetAsync(d_sent, 0, CRGPU_MAX_PARTS * 8, st);
  if (!n) return 0;
  uint64_t chunks = (n + PX_CHUNK - 1) / PX_CHUNK;
  int grid = (int)std::min<uint64_t>(chunks, (uint64_t)sm_count() * 4);
  owner_scatter_peers_kernel<<<grid, PX_THREADS, 0, st>>>(keys, n, rank_shift, ob, pt, d_sent);
  return 1;
}


// ---------------------------------------------------------------------------
// The exchange kernel of the sharded run (key range and owner bounds read on the device, so that the step
// enqueues it without a host round trip).
// ---------------------------------------------------------------------------
constexpr int PL_THREADS = 256;
constexpr int PL_ITEMS = 16;
constexpr int PL_CHUNK = PL_THREADS * PL_ITEMS;

// lanes of the warp whose owner (0..16, 16 = no key) equals this lane's: five ballots
__device__ __forceinline__ uint32_t match_owner(uint32_t own) {
  uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
  for (int b = 0; b < 5; b++) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, (own >> b) & 1u);
    peers &= ((own >> b) & 1u) ? m : ~m;
  }
  return peers;
}

// A 16-way partition per chunk, built like a radix pass (sort.cu): every key is ranked among the keys of its
// owner inside its warp (ballots, warp-private counters), the counters are scanned over the warps and the owners,
// the chunk is ordered by owner in shared memory and every owner's run leaves as coalesced peer stores behind one
// remote claim per owner and chunk, whose round trip runs under the ordering of the chunk. Measured on B200, 168 M
// keys per GPU (profiles/r02_bench_{2,8}gpu*.json): 1.61 ms at 2 GPUs, where the former kernel (two match.any
// rounds per key, owner_scatter_peers_body) took 2.62 ms; 3.13 ms at 8 GPUs (1.18 GB out and in per GPU). At 2
// GPUs the kernel is bound by its own instructions (the same time with every store kept local). Measured and
// dropped: a counting phase and one claim per owner and BLOCK (2.35 ms at 2 GPUs, 3.46 ms at 8: the claims are not
// what limits it); addresses planned ahead by a counting kernel, a scan and an all-gather of the G x G counts
// (2.59 ms at 2 GPUs); chunks of 2048 keys at five blocks per SM (1.78 ms).
// owner of a rank = number of inner bounds at or below it (the bounds ascend): n_parts - 1 compares, not 15
__device__ __forceinline__ uint32_t owner_of_n(const OwnerBounds& ob, int n_parts, uint32_t rank) {
  uint32_t p = 0;
  for (int i = 1; i < n_parts; i++) p += rank >= ob.b[i] ? 1u : 0u;
  return p;
}

template <int FS_ITEMS, int FS_MINB>
__global__ void __launch_bounds__(PL_THREADS, FS_MINB) owner_scatter_fast_kernel(
    const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ begin_dev,
    const unsigned long long* __restrict__ end_dev, int rank_shift, const uint32_t* __restrict__ bounds_dev, int n_parts,
    PeerTargets pt, unsigned long long* __restrict__ sent) {
  constexpr int P = CRGPU_MAX_PARTS, WARPS = PL_THREADS / 32, FS_CHUNK = PL_THREADS * FS_ITEMS;
  __shared__ OwnerBounds ob;
  __shared__ unsigned long long s_keys[FS_CHUNK];
  __shared__ uint8_t s_own[FS_CHUNK];
  __shared__ uint32_t s_warp_hist[WARPS * (P + 1)];
  __shared__ uint32_t s_cnt[P], s_off[P + 1];
  __shared__ unsigned long long s_run[P];  // where this chunk's run of every owner goes (~0: nowhere)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid <= P) ob.b[tid] = tid <= n_parts ? bounds_dev[tid] : 0xFFFFFFFFu;
  if (tid == 0) ob.n = n_parts;
  __syncthreads();
  const unsigned long long first_key = begin_dev ? *begin_dev : 0ull;
  const unsigned long long last_key = *end_dev;
  const uint64_t n = last_key > first_key ? last_key - first_key : 0ull;
  keys += first_key;
  // the chunks of this block: [c_lo, c_hi)
  const uint64_t n_chunks = (n + FS_CHUNK - 1) / FS_CHUNK;
  const uint64_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const uint64_t c_lo = (uint64_t)blockIdx.x * per < n_chunks ? (uint64_t)blockIdx.x * per : n_chunks;
  const uint64_t c_hi = c_lo + per < n_chunks ? c_lo + per : n_chunks;
  if (c_lo == c_hi) return;  // block-uniform
  uint32_t* my_hist = s_warp_hist + warp * (P + 1);
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (uint64_t chunk = c_lo; chunk < c_hi; chunk++) {
    for (int i = tid; i < WARPS * (P + 1); i += PL_THREADS) s_warp_hist[i] = 0u;
    __syncthreads();
    const uint64_t first = chunk * FS_CHUNK;
    unsigned long long k[FS_ITEMS];
    uint32_t dr[FS_ITEMS];  // owner | rank inside the warp << 8
    const int warp_first = warp * 32 * FS_ITEMS;
#pragma unroll
    for (int i = 0; i < FS_ITEMS; i++) {
      const uint64_t g = first + warp_first + i * 32 + lane;
      k[i] = g < n ? __ldcs(keys + g) : 0ull;
    }
#pragma unroll
    for (int i = 0; i < FS_ITEMS; i++) {
      const uint64_t g = first + warp_first + i * 32 + lane;
      const uint32_t own = g < n ? owner_of_n(ob, n_parts, (uint32_t)(k[i] >> rank_shift)) : (uint32_t)P;
      const uint32_t peers = match_owner(own);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (lane == leader) {
        base = my_hist[own];
        my_hist[own] = base + (uint32_t)__popc(peers);
      }
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      __syncwarp();
      dr[i] = own | ((base + (uint32_t)__popc(peers & lt_mask)) << 8);
    }
    __syncthreads();
    if (warp == 0) {  // per owner: exclusive scan of its counts over the warps, then over the owners (shuffles)
      uint32_t run = 0;
      if (lane < P) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
          const uint32_t c = s_warp_hist[w * (P + 1) + lane];
          s_warp_hist[w * (P + 1) + lane] = run;
          run += c;
        }
        s_cnt[lane] = run;
      }
      uint32_t inc = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane <= P) s_off[lane] = inc - run;  // lane P: the chunk's total
    }
    __syncthreads();
    // per-chunk claims: the round trip runs under the ordering of the chunk in shared memory
    unsigned long long claim = ~0ull;
    uint32_t claim_cnt = 0u;
    if (tid < n_parts) {
      claim_cnt = s_cnt[tid];
      if (claim_cnt) {
        claim = atomicAdd(pt.cursor[tid], (unsigned long long)claim_cnt);
        atomicAdd(sent + tid, (unsigned long long)claim_cnt);
      }
    }
#pragma unroll
    for (int i = 0; i < FS_ITEMS; i++) {
      const uint32_t own = dr[i] & 0xFFu;
      if (own < (uint32_t)P) {
        const uint32_t pos = s_off[own] + my_hist[own] + (dr[i] >> 8);
        s_keys[pos] = k[i];
        s_own[pos] = (uint8_t)own;
      }
    }
    if (tid < P) {
      if (claim != ~0ull && claim + claim_cnt > pt.capacity) {
        atomicExch(pt.cursor[tid] + 1, 1ull);  // the receiver reports the overflow
        claim = ~0ull;
      }
      s_run[tid] = claim;
    }
    __syncthreads();
    // coalesced peer stores, one owner run after the other
    const uint32_t total = s_off[P];
#pragma unroll 4
    for (uint32_t p = tid; p < total; p += PL_THREADS) {
      const int o = s_own[p];
      const unsigned long long b = s_run[o];
      if (b != ~0ull) pt.buf[o][b + (p - s_off[o])] = s_keys[p];
    }
    __syncthreads();
  }
  __threadfence_system();  // as in owner_scatter_peers_body
}

int run_owner_scatter_peers_dev(const unsigned long long* keys, const unsigned long long* begin_dev,
                                const unsigned long long* end_dev, uint64_t n_max, int rank_shift,
                                const uint32_t* bounds_dev, int n_parts, unsigned long long* const* peer_buf,
                                unsigned long long* const* peer_cursor, unsigned long long capacity,
                                unsigned long long* d_sent, cudaStream_t st) {
  PeerTargets pt;
  for (int i = 0; i < CRGPU_MAX_PARTS; i++) {
    pt.buf[i] = i < n_parts ? peer_buf[i] : nullptr;
    pt.cursor[i] = i < n_parts ? peer_cursor[i] : nullptr;
  }
  pt.capacity = capacity;
  cudaMemsetAsync(d_sent, 0, CRGPU_MAX_PARTS * 8, st);
  if (!n_max) return 0;
  if (getenv("CRGPU_SCATTER_CFG") && atoi(getenv("CRGPU_SCATTER_CFG")) == 1) {  // the former kernel, for profiling
    const uint64_t chunks = (n_max + PX_CHUNK - 1) / PX_CHUNK;
    const int grid = (int)std::min<uint64_t>(chunks, (uint64_t)sm_count() * 4);
    owner_scatter_peers_dev_kernel<<<grid, PX_THREADS, 0, st>>>(keys, begin_dev, end_dev, rank_shift, bounds_dev, n_parts,
                                                               pt, d_sent);
    return 1;
  }
  const uint64_t chunks = (n_max + PL_CHUNK - 1) / PL_CHUNK;
  const int grid = (int)std::min<uint64_t>(chunks, (uint64_t)sm_count() * 3);
  owner_scatter_fast_kernel<PL_ITEMS, 3><<<grid, PL_THREADS, 0, st>>>(keys, begin_dev, end_dev, rank_shift, bounds_dev,
                                                                     n_parts, pt, d_sent);
  return 1;
}

// ---------------------------------------------------------------------------
// Owner ranges on the device: contiguous content-rank ranges [bounds[r], bounds[r+1]) holding about equal numbers
// of valid reads - what ShardReader::make_chunks does for the reference's barcode-range chunks
// (cr_lib/src/stages/align_and_count.rs:519-524). With csum the inclusive prefix sum of the per-rank totals
// (summed over the library types) and target_r = (total * r) / G:  bounds[r] = min(first i with csum[i] >=
// target_r, ...) + 1, made non-decreasing; bounds[0] = 0, bounds[G] = n. Same arithmetic as owner_bounds() in
// cellranger_b200/dist.py, which the tests compare it with.
// ---------------------------------------------------------------------------
constexpr int OB_THREADS = 1024;
constexpr int OB_PER_THREAD = 32;
constexpr int OB_CHUNK = OB_THREADS * OB_PER_THREAD;  // 32768 ranks per block

struct CountVectors {
  const uint32_t* v[CRGPU_MAX_LIBS];
  int n;
};

__device__ __forceinline__ unsigned long long ob_block_reduce(unsigned long long x, unsigned long long* s_warp) {
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  unsigned long long t = 0;
  for (int w = 0; w < OB_THREADS / 32; w++) t += s_warp[w];
  return t;
}

__global__ void __launch_bounds__(OB_THREADS) ob_partial_kernel(CountVectors cv, uint32_t n,
                                                                unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long s_warp[OB_THREADS / 32];
  const uint64_t first = (uint64_t)blockIdx.x * OB_CHUNK + (uint64_t)threadIdx.x * OB_PER_THREAD;
  unsigned long long sum = 0;
  for (int k = 0; k < OB_PER_THREAD; k++) {
    const uint64_t i = first + k;
    if (i < n)
      for (int l = 0; l < cv.n; l++) sum += cv.v[l][i];
  }
  const unsigned long long tot = ob_block_reduce(sum, s_warp);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(OB_THREADS) ob_cuts_kernel(CountVectors cv, uint32_t n, uint32_t n_chunks,
                                                             const unsigned long long* __restrict__ partial, int n_parts,
                                                             uint32_t* __restrict__ bounds) {
  __shared__ unsigned long long s_warp[OB_THREADS / 32];
  __shared__ unsigned long long s_total, s_before;
  __shared__ uint32_t s_chunk, s_cut;
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (uint32_t c = 0; c < n_chunks; c++) t += partial[c];
    s_total = t;
  }
  __syncthreads();
  uint32_t prev = 0;
  for (int r = 1; r < n_parts; r++) {
    const unsigned long long target = (s_total * (unsigned long long)r) / (unsigned long long)n_parts;
    if (threadIdx.x == 0) {  // the chunk in which the inclusive prefix sum first reaches the target
      unsigned long long run = 0;
      uint32_t c = 0;
      while (c < n_chunks && run + partial[c] < target) run += partial[c++];
      s_chunk = c;
      s_before = run;
      s_cut = 0xFFFFFFFFu;
    }
    __syncthreads();
    uint32_t cut = n;  // no index reaches the target (cannot happen for target <= total): searchsorted returns n
    if (s_chunk < n_chunks) {
      const uint64_t first = (uint64_t)s_chunk * OB_CHUNK + (uint64_t)threadIdx.x * OB_PER_THREAD;
      unsigned long long v[OB_PER_THREAD], sum = 0;
      for (int k = 0; k < OB_PER_THREAD; k++) {
        const uint64_t i = first + k;
        unsigned long long x = 0;
        if (i < n)
          for (int l = 0; l < cv.n; l++) x += cv.v[l][i];
        v[k] = x;
        sum += x;
      }
      // exclusive prefix of the thread sums over the block
      unsigned long long inc = sum;
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
      }
      __syncthreads();
      if (lane == 31) s_warp[warp] = inc;
      __syncthreads();
      unsigned long long wsum = 0;
      for (int w = 0; w < warp; w++) wsum += s_warp[w];
      unsigned long long run = s_before + wsum + inc - sum;
      for (int k = 0; k < OB_PER_THREAD; k++) {
        run += v[k];
        if (first + k < n && run >= target) {
          atomicMin(&s_cut, (uint32_t)(first + k));
          break;
        }
      }
      __syncthreads();
      if (s_cut != 0xFFFFFFFFu) cut = s_cut;
    }
    uint32_t b = cut + 1u < n ? cut + 1u : n;  // searchsorted(...) + 1, clamped
    if (b < prev) b = prev;                    // non-decreasing
    prev = b;
    if (threadIdx.x == 0) bounds[r] = b;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    bounds[0] = 0;
    bounds[n_parts] = n;
  }
}

// scratch: ceil(n / OB_CHUNK) u64
int run_owner_bounds(const uint32_t* const* vectors, int n_vectors, uint32_t n, int n_parts, unsigned long long* scratch,
                     uint32_t* bounds_dev, cudaStream_t st) {
  CountVectors cv;
  cv.n = n_vectors;
  for (int l = 0; l < CRGPU_MAX_LIBS; l++) cv.v[l] = l < n_vectors ? vectors[l] : nullptr;
  const uint32_t n_chunks = (uint32_t)(((uint64_t)n + OB_CHUNK - 1) / OB_CHUNK);
  if (n_chunks) ob_partial_kernel<<<n_chunks, OB_THREADS, 0, st>>>(cv, n, scratch);
  ob_cuts_kernel<<<1, OB_THREADS, 0, st>>>(cv, n, n_chunks, scratch, n_parts, bounds_dev);
  return n_chunks ? 2 : 1;
}
size_t owner_bounds_scratch_bytes(uint32_t n) { return ((size_t)n / OB_CHUNK + 2) * 8; }

// per-read barcode states of a batch (local statistics; the histograms may hold global counts)
__global__ void state_counts_kernel(const uint32_t* __restrict__ bc_out, uint64_t n, unsigned long long* out4) {
  unsigned long long c1 = 0, c2 = 0, c3 = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t st = bc_out[i] >> BC_STATE_SHIFT;
    c1 += st == ST_VALID_BEFORE;
    c2 += st == ST_VALID_AFTER;
    c3 += st == ST_INVALID;
  }
  for (int d = 16; d > 0; d >>= 1) {
    c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, d);
    c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, d);
    c3 += __shfl_xor_sync(0xFFFFFFFFu, c3, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c1) atomicAdd(out4 + 1, c1);
    if (c2) atomicAdd(out4 + 2, c2);
    if (c3) atomicAdd(out4 + 3, c3);
  }
}
int launch_state_counts(const uint32_t* bc_out, uint64_t n, unsigned long long* out4, cudaStream_t st) {
  if (!n) return 0;
  state_counts_kernel<<<grid_for(n), 256, 0, st>>>(bc_out, n, out4);
  return 1;
}

// debug verification (CRGPU_VERIFY=1): order violations in a key array; strict = equal neighbours count too
__global__ void order_violations_kernel(const unsigned long long* __restrict__ a, uint64_t n, int strict, int shift,
                                        unsigned long long* out) {
  unsigned long long bad = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    bad += strict ? ((a[i] >> shift) <= (a[i - 1] >> shift)) : ((a[i] >> shift) < (a[i - 1] >> shift));
  for (int d = 16; d > 0; d >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, d);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, bad);
}
int launch_order_violations(const unsigned long long* a, uint64_t n, int strict, int shift, unsigned long long* out,
                            cudaStream_t st) {
  if (n < 2) return 0;
  order_violations_kernel<<<grid_for(n), 256, 0, st>>>(a, n, strict, shift, out);
  return 1;
}

// Work buffers reused across phases:
//   key2 / key2_alt : sort scratch for the (rank, lib, umi, feature) grouping, later molecule keys
int run_dedup(DedupBuffers& b, uint64_t* n_distinct_host, cudaStream_t st) {
  int launches = 0;
  ScanScratch ss{b.lb_desc, b.tickets, 0};
  auto mark = [&](const char* name) {
    if (b.mark) b.mark(b.mark_user, name);
  };
  cudaMemsetAsync(b.scalars, 0, 16 * 8, st);
  mark("count.dedup.rle");
  // 1. run-length encode: distinct keys + raw counts (head positions parked in `best`)
  if (b.verify && b.finish_umi)  // before the finishing sort rearranges the long segments
    launches += launch_order_violations(b.sorted, b.n_keys, 0, b.kl.umi_bits, b.scalars + 10, st);
  if (b.finish_umi)
    launches += run_finish(b.sorted, b.sorted_alt, b.n_keys, b.kl.umi_bits, b.dkeys, b.c0, ss.desc, b.scalars + 0, st,
                           b.mark, b.mark_user);
  else
    launches += run_rle<true>(b.sorted, b.n_keys, 0, b.dkeys, b.best, ss.desc, ss.ticket, b.scalars + 0, st);
  unsigned long long m = 0;
  cudaMemcpyAsync(&m, b.scalars + 0, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  *n_distinct_host = m;
  if (m == 0) return launches;
  if (!b.finish_umi) {
    run_lengths_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, b.n_keys, b.c0);
    launches++;
  }
  if (b.verify) {
    if (!b.finish_umi) launches += launch_order_violations(b.sorted, b.n_keys, 0, 0, b.scalars + 10, st);
    launches += launch_order_violations(b.dkeys, m, 1, 0, b.scalars + 11, st);
  }
  // 2. UMI correction targets + incoming counts
  mark("count.dedup.correct_umis");
  cudaMemsetAsync(b.inc, 0, m * 8, st);
  cudaMemsetAsync(b.low, 0, m, st);
  {
    const size_t smem = (size_t)CU_WIN * 8 + (size_t)CU_SLOTS * 4 + (size_t)CU_BITS / 8 + (size_t)CU_TILE * 2 +
                        (size_t)CU_WIN * 2;
    auto kern = b.kl.umi_bits == 24 ? correct_umis_kernel<24>
                : b.kl.umi_bits == 20 ? correct_umis_kernel<20> : correct_umis_kernel<0>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned blocks = (unsigned)((m + CU_TILE - 1) / CU_TILE);
    const int pair_max = getenv("CRGPU_CU_PAIR") ? atoi(getenv("CRGPU_CU_PAIR")) : CU_PAIR;
    kern<<<blocks, CU_THREADS, smem, st>>>(b.dkeys, b.c0, m, b.kl, b.umi_correction_mask, b.best, b.inc, b.scalars,
                                           pair_max < 1 ? 1 : (pair_max > CU_HALO ? CU_HALO : pair_max));
    launches++;
  }
  mark("count.dedup.low_support");
  // 3. low-support filter: candidates by hashing (rank, library, umi), exact regrouping of those only
  if (b.filter_umis) {
    if (getenv("CRGPU_LS_GLOBAL") && atoi(getenv("CRGPU_LS_GLOBAL"))) {
      // the round-1 pre-filter (global 2-bit slot table), kept for A/B measurements
      int slot_bits = 16;
      const int slot_cap = getenv("CRGPU_LS_SLOTCAP") ? atoi(getenv("CRGPU_LS_SLOTCAP")) : 29;
      const int region_bits = ls_region_bits_host();
      while (slot_bits < slot_cap && (1ull << slot_bits) < 8 * m) slot_bits++;
      const size_t slot_bytes = ((size_t)1 << slot_bits) / 4;  // 2 bits per slot
      if (b.slots_bytes < slot_bytes) return -1;
      cudaMemsetAsync(b.slots, 0, slot_bytes, st);
      ls_mark_kernel<<<grid_for(m, 256, 32), 256, 0, st>>>(b.dkeys, m, b.kl, b.slots, slot_bits, region_bits);
      ls_collect_kernel<<<grid_for(m, 256, 16), 256, 0, st>>>(b.dkeys, m, b.kl, b.slots, slot_bits, region_bits, b.key2,
                                                                   b.scalars + 9);
      launches += 2;
    } else {
      const size_t smem = (size_t)LF_MAX_SLOTS / 4;
      cudaFuncSetAttribute(ls_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      ls_filter_kernel<<<(unsigned)((m + LF_TILE - 1) / LF_TILE), LF_THREADS, smem, st>>>(b.dkeys, m, b.kl, b.key2,
                                                                                        b.scalars + 9);
      launches += 1;
    }
    unsigned long long n_cand = 0;
    cudaMemcpyAsync(&n_cand, b.scalars + 9, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (getenv("CRGPU_DEBUG")) fprintf(stderr, "[crgpu] low-support candidates: %llu of %llu distinct keys\n", n_cand, m);
    mark("count.dedup.low_support.regroup");
    if (n_cand) {
      unsigned long long* sorted2 = nullptr;
      // grouping by (rank, library, umi) only needs the bits above the feature field
      launches += sort_keys(b.key2, b.key2_alt, n_cand, b.kl.total_bits, b.sort_temp, b.sort_temp_bytes, &sorted2, st,
                            field_masks(b.kl).fbits);
      low_support_kernel<<<grid_for(n_cand, 256, 32), 256, 0, st>>>(sorted2, n_cand, m, b.kl, b.dkeys, b.c0,
                                                                          b.best, b.inc, b.low, b.scalars);
      launches++;
    }
  }
  mark("count.dedup.molecules");
  if (b.on_target && b.target_min_reads) {
    target_filter_kernel<<<grid_for(m, 256, 16), 256, 0, st>>>(b.dkeys, b.c0, b.best, b.inc, m, b.kl, b.on_target,
                                                                   b.n_on_target, b.target_min_reads, b.low, b.scalars + 12);
    launches++;
  }
  // 4. molecules = correction targets that are not low support (key2 now holds their keys)
  {
    uint64_t tiles = (m + CP_TILE - 1) / CP_TILE;
    cudaMemsetAsync(ss.desc, 0, tiles * 8, st);
    cudaMemsetAsync(ss.ticket, 0, 4, st);
    molecules_compact_kernel<<<(unsigned)tiles, CP_THREADS, 0, st>>>(b.dkeys, b.c0, b.best, b.inc, b.low, m, b.key2,
                                                                    b.mol, b.mol_idx, ss.desc, tile_ticket(ss.ticket),
                                                                    b.scalars + 2, b.scalars + 6);
    launches++;
  }
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
  const uint32_t tile = acquire_tile(ticket, &tile_s);
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

// UmiType of a molecule as molecule_info stores it (1 Txomic, 0 NonTxomic): the type of its representative read,
// i.e. bit 63 of the smallest UmiSelectKey word of the representative raw key (rep_raw, or the key itself)
__device__ __forceinline__ uint32_t molecule_utype(uint32_t d, const unsigned long long* __restrict__ min_key,
                                                   const uint32_t* __restrict__ rep_raw) {
  if (!min_key) return 1u;
  const uint32_t r = rep_raw[d] == 0xFFFFFFFFu ? d : rep_raw[d];
  return (min_key[r] >> 63) ? 0u : 1u;
}

__global__ void molecules_kernel(const unsigned long long* __restrict__ mol_key, const uint32_t* __restrict__ mol_reads,
                                 const uint32_t* __restrict__ mol_idx, uint64_t n_mol, KeyLayout kl,
                                 const uint32_t* __restrict__ col_of_rank, const unsigned long long* __restrict__ min_key,
                                 const uint32_t* __restrict__ rep_raw, uint32_t* __restrict__ out6) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_mol; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long k = mol_key[i];
    uint32_t rank = (uint32_t)(k >> kl.rank_shift);
    out6[6 * i + 0] = col_of_rank[rank];
    out6[6 * i + 1] = (uint32_t)((k >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull));
    out6[6 * i + 2] = (uint32_t)((k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull));
    out6[6 * i + 3] = (uint32_t)(k & ((1ull << fm.ubits) - 1ull));
    out6[6 * i + 4] = mol_reads[i];
    out6[6 * i + 5] = mol_idx ? molecule_utype(mol_idx[i], min_key, rep_raw) : 1u;
  }
}

// (rank | feature | library | umi) <-> (rank | library | feature | umi): the order UmiCount sorts in
__device__ __forceinline__ unsigned long long key_lib_major(unsigned long long k, const KeyLayout& kl, const FieldMasks& fm) {
  const unsigned long long umi = k & ((1ull << fm.ubits) - 1ull);
  const unsigned long long lib = (k >> kl.lib_shift) & ((1ull << fm.lbits) - 1ull);
  const unsigned long long feat = (k >> kl.feature_shift) & ((1ull << fm.fbits) - 1ull);
  const unsigned long long rank = k >> kl.rank_shift;
  return (((rank << fm.lbits | lib) << fm.fbits | feat) << fm.ubits) | umi;
}
__global__ void molecules_remap_kernel(const unsigned long long* __restrict__ mol_key, uint64_t n_mol, KeyLayout kl,
                                       unsigned long long* __restrict__ out) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_mol; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = key_lib_major(mol_key[i], kl, fm);
}
// rows from the re-sorted (library-major) keys: each finds its molecule by binary search in the key-ordered table
__global__ void molecules_reordered_kernel(const unsigned long long* __restrict__ sorted_lm,
                                           const unsigned long long* __restrict__ mol_key,
                                           const uint32_t* __restrict__ mol_reads, const uint32_t* __restrict__ mol_idx,
                                           uint64_t n_mol, KeyLayout kl, const uint32_t* __restrict__ col_of_rank,
                                           const unsigned long long* __restrict__ min_key,
                                           const uint32_t* __restrict__ rep_raw, uint32_t* __restrict__ out6) {
  const FieldMasks fm = field_masks(kl);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_mol; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long lm = sorted_lm[i];
    const unsigned long long umi = lm & ((1ull << fm.ubits) - 1ull);
    const unsigned long long feat = (lm >> fm.ubits) & ((1ull << fm.fbits) - 1ull);
    const unsigned long long lib = (lm >> (fm.ubits + fm.fbits)) & ((1ull << fm.lbits) - 1ull);
    const unsigned long long rank = lm >> (fm.ubits + fm.fbits + fm.lbits);
    const unsigned long long k = (rank << kl.rank_shift) | (feat << kl.feature_shift) | (lib << kl.lib_shift) | umi;
    const uint64_t j = lower_bound_u64(mol_key, n_mol, k);
    out6[6 * i + 0] = col_of_rank[(uint32_t)rank];
    out6[6 * i + 1] = (uint32_t)lib;
    out6[6 * i + 2] = (uint32_t)feat;
    out6[6 * i + 3] = (uint32_t)umi;
    out6[6 * i + 4] = mol_reads[j];
    out6[6 * i + 5] = mol_idx ? molecule_utype(mol_idx[j], min_key, rep_raw) : 1u;
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
  launches += run_rle<false>(b.key2, n_mol, b.kl.feature_shift, nullptr, run_pos, ss.desc, ss.ticket, b.scalars + 1, st);
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
    inclusive_scan_i64_kernel<THREADS, ITEMS><<<(unsigned)tiles, THREADS, 0, st>>>(ma.indptr, n_bc + 1, ss.desc, tile_ticket(ss.ticket));
    launches++;
  }
  return launches;
}

// ---------------------------------------------------------------------------
// BarcodeSummary (cr_lib/src/aligner.rs:33-68, filled by visit_read_annotation, cr_lib/src/align_metrics.rs:705-721):
// per valid barcode and library {reads, umis, candidate_dup_reads, umi_corrected_reads}. Every read of a raw key
// shares its DupInfo flags, so the sums run over the distinct-key table:
//   candidate_dup_reads += c0[j] unless the corrected key best[j] is low support
//   umi_corrected_reads += c0[j] if best[j] != j
//   umis               += 1 for every correction target that is not low support (its representative read)
// `reads` is the valid-barcode read count of the library (prior + corrected).
// ---------------------------------------------------------------------------
__global__ void summary_reads_kernel(const uint32_t* __restrict__ barcode_rank, const uint32_t* __restrict__ valid,
                                     uint64_t n_bc, uint32_t* __restrict__ out) {
  for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n_bc; c += (uint64_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(out)[c] = make_uint4(valid[barcode_rank[c]], 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256) summary_keys_kernel(const unsigned long long* __restrict__ dkeys,
                                                           const uint32_t* __restrict__ c0,
                                                           const uint32_t* __restrict__ best,
                                                           const unsigned long long* __restrict__ inc,
                                                           const uint8_t* __restrict__ low, uint64_t m, KeyLayout kl,
                                                           uint32_t lib, const uint32_t* __restrict__ col_of_rank,
                                                           uint32_t* __restrict__ out) {
  const uint32_t lmask = (1u << (kl.feature_shift - kl.lib_shift)) - 1u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // warp-uniform trip count: the warp-level reductions below need every lane
  const uint64_t rounds = (m + stride - 1) / stride;
  for (uint64_t r = 0; r < rounds; r++) {
    const uint64_t j = r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t rank = 0xFFFFFFFFu, umis = 0, cand = 0, corr = 0;
    if (j < m) {
      const unsigned long long k = dkeys[j];
      if (((uint32_t)(k >> kl.lib_shift) & lmask) == lib) {
        rank = (uint32_t)(k >> kl.rank_shift);
        const uint32_t t = best[j];
        const uint32_t n = c0[j];
        if (!(low[t] & 1)) cand = n;  // candidate_dup_reads: not low support (aligner.rs:56-58)
        if (t != (uint32_t)j) corr = n;
        const bool is_target = t == (uint32_t)j || (inc[j] >> 40) != 0ull;
        umis = is_target && !low[j];
      }
    }
    // the table is sorted by rank: a warp sees a handful of barcodes; one atomic per barcode and counter
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, rank);
    const uint32_t s_umis = __reduce_add_sync(peers, umis);
    const uint32_t s_cand = __reduce_add_sync(peers, cand);
    const uint32_t s_corr = __reduce_add_sync(peers, corr);
    if (rank != 0xFFFFFFFFu && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) {
      uint32_t* o = out + 4 * (size_t)col_of_rank[rank];
      if (s_umis) atomicAdd(o + 1, s_umis);
      if (s_cand) atomicAdd(o + 2, s_cand);
      if (s_corr) atomicAdd(o + 3, s_corr);
    }
  }
}

int run_barcode_summary(DedupBuffers& b, uint64_t m, uint32_t lib, const uint32_t* barcode_rank, const uint32_t* valid,
                        const uint32_t* col_of_rank, uint64_t n_bc, uint32_t* out4, cudaStream_t st) {
  if (n_bc == 0) return 0;
  summary_reads_kernel<<<grid_for(n_bc), 256, 0, st>>>(barcode_rank, valid, n_bc, out4);
  if (m == 0) return 1;
  summary_keys_kernel<<<grid_for(m, 256, 8), 256, 0, st>>>(b.dkeys, b.c0, b.best, b.inc, b.low, m, b.kl, lib,
                                                                col_of_rank, out4);
  return 2;
}

// ---------------------------------------------------------------------------
// BarcodeDiversityMetrics of BARCODE_CORRECTION's join (cr_lib/src/stages/barcode_correction.rs:428-441):
// barcodes_detected = entries of the corrected barcode histogram, effective_barcode_diversity = its inverse
// Simpson index (sum c)^2 / sum c^2 (SimpleHistogram::effective_diversity, metric/src/histogram.rs:161-171).
// The sums are taken exactly in integers (sum c^2 in 128 bits) and converted once.
// out4: [0] barcodes with a count, [1] sum c, [2] low and [3] high word of sum c^2
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) diversity_kernel(const uint32_t* __restrict__ counts, uint64_t n,
                                                        unsigned long long* __restrict__ out4) {
  unsigned long long nz = 0, s = 0;
  unsigned __int128 s2 = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long c = counts[i];
    nz += c != 0ull;
    s += c;
    s2 += (unsigned __int128)(c * c);
  }
  unsigned long long lo = (unsigned long long)s2, hi = (unsigned long long)(s2 >> 64);
  for (int d = 16; d > 0; d >>= 1) {
    nz += __shfl_xor_sync(0xFFFFFFFFu, nz, d);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    const unsigned long long olo = __shfl_xor_sync(0xFFFFFFFFu, lo, d), ohi = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
    const unsigned long long nlo = lo + olo;
    hi += ohi + (nlo < lo ? 1ull : 0ull);
    lo = nlo;
  }
  if ((threadIdx.x & 31) == 0 && nz) {
    atomicAdd(out4 + 0, nz);
    atomicAdd(out4 + 1, s);
    const unsigned long long old = atomicAdd(out4 + 2, lo);
    atomicAdd(out4 + 3, hi + (old + lo < old ? 1ull : 0ull));
  }
}
int launch_diversity(const uint32_t* counts, uint64_t n, unsigned long long* out4, cudaStream_t st) {
  cudaMemsetAsync(out4, 0, 32, st);
  if (!n) return 0;
  diversity_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(counts, n, out4);
  return 1;
}

int run_molecule_rows(DedupBuffers& b, const uint32_t* col_of_rank, uint64_t n_mol, const unsigned long long* min_key,
                      const uint32_t* rep_raw, int reorder, unsigned long long* sort_a, unsigned long long* sort_b,
                      void* sort_temp, size_t sort_temp_bytes, uint32_t* out6, cudaStream_t st) {
  if (!n_mol) return 0;
  if (!reorder) {
    molecules_kernel<<<grid_for(n_mol), 256, 0, st>>>(b.key2, b.mol, b.mol_idx, n_mol, b.kl, col_of_rank, min_key, rep_raw,
                                                      out6);
    return 1;
  }
  molecules_remap_kernel<<<grid_for(n_mol), 256, 0, st>>>(b.key2, n_mol, b.kl, sort_a);
  unsigned long long* sorted = nullptr;
  int launches = 1 + sort_keys(sort_a, sort_b, n_mol, b.kl.total_bits, sort_temp, sort_temp_bytes, &sorted, st);
  molecules_reordered_kernel<<<grid_for(n_mol), 256, 0, st>>>(sorted, b.key2, b.mol, b.mol_idx, n_mol, b.kl, col_of_rank,
                                                              min_key, rep_raw, out6);
  return launches + 1;
}

// ---------------------------------------------------------------------------
// per-read DupInfo (optional): BarcodeDupMarker::process (mark_dups.rs:280-363)
// ---------------------------------------------------------------------------
__global__ void annotate_prepare_kernel(const uint32_t* __restrict__ best, uint64_t m,
                                        unsigned long long* __restrict__ min_key, uint32_t* __restrict__ rep_raw) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
    min_key[j] = ~0ull;
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

// UmiSelectKey word of read i: the caller's, or (Txomic, qname ordered like the global read index)
__device__ __forceinline__ unsigned long long select_word(const AnnotateArgs& a, uint64_t i) {
  return a.select ? a.select[i] : (unsigned long long)(a.read_base + i);
}

// the smallest UmiSelectKey of every raw key: old_min.min(ann_key), mark_dups.rs:147-151
__global__ void annotate_min_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m, KeyLayout kl,
                                    AnnotateArgs a, unsigned long long* __restrict__ min_key) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long key;
    if (!read_key(a, kl, i, &key)) continue;
    uint64_t j = lower_bound_u64(dkeys, m, key);
    atomicMin(min_key + j, select_word(a, i));
  }
}

__global__ void annotate_final_kernel(const unsigned long long* __restrict__ dkeys, uint64_t m, KeyLayout kl,
                                      AnnotateArgs a, const uint32_t* __restrict__ best,
                                      const uint8_t* __restrict__ low, const unsigned long long* __restrict__ min_key,
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
      bool is_low = (low[d] & 1) != 0;
      bool is_filtered = (low[d] & 2) != 0;  // is_filtered_target_umi, mark_dups.rs:311-320
      uint32_t rep = rep_raw[d] == 0xFFFFFFFFu ? d : rep_raw[d];
      // is_min_qname: the read's header equals the qname of the corrected key's UmiSelectKey (mark_dups.rs:300-303);
      // the key's select key is the one of raw key `rep` (:248-268), and only its own reads can carry that qname
      const unsigned long long q63 = 0x7FFFFFFFFFFFFFFFull;
      bool is_rep = rep == (uint32_t)j && (min_key[j] & q63) == (select_word(a, i) & q63);
      fl |= 2u | (corrected ? 4u : 0u) | (is_low ? 8u : 0u) | ((!is_low && !is_filtered && is_rep) ? 16u : 0u) |
            (is_filtered ? 32u : 0u);
      uw = (uw & ~UMI_SEQ_MASK) | (uint32_t)(dkeys[d] & umask);
    }
    a.umi_proc[i] = uw;
    a.flags_out[i] = fl;
  }
}

int run_annotate_prepare(DedupBuffers& b, uint64_t m, unsigned long long* min_key, uint32_t* rep_raw, cudaStream_t st) {
  if (!m) return 0;
  annotate_prepare_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, min_key, rep_raw);
  annotate_rep_kernel<<<grid_for(m), 256, 0, st>>>(b.best, m, rep_raw);
  return 2;
}
int run_annotate_min(DedupBuffers& b, uint64_t m, const AnnotateArgs& a, unsigned long long* min_key, cudaStream_t st) {
  if (!a.n || !m) return 0;
  annotate_min_kernel<<<grid_for(a.n, 256, 32), 256, 0, st>>>(b.dkeys, m, b.kl, a, min_key);
  return 1;
}
int run_annotate_final(DedupBuffers& b, uint64_t m, const AnnotateArgs& a, const unsigned long long* min_key,
                       const uint32_t* rep_raw, unsigned long long*, cudaStream_t st) {
  if (!a.n) return 0;
  annotate_final_kernel<<<grid_for(a.n, 256, 32), 256, 0, st>>>(b.dkeys, m, b.kl, a, b.best, b.low, min_key,
                                                                     rep_raw);
  return 1;
}
