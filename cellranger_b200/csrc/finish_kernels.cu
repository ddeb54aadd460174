// finish_kernels.cu — the last 2·L_umi bits of the key sort, done per (barcode, feature, library) segment in shared
// memory, fused with the run-length encoding. sm_100a.
//
// The device radix sort orders the keys on the bits ABOVE the UMI only (rank | feature | library: 38 of the 62
// bits of a 3' v3 key, 5 passes instead of 8). What DupBuilder::observe needs next (lib/rust/tx_annotation/src/
// mark_dups.rs:128-155: counts[(umi, gene)] += 1 per barcode and library) is, per segment of equal upper bits, the
// sorted distinct UMIs with their read counts. Segments are tiny on average (5.7 reads) and heavy-tailed (half the
// keys sit in segments of more than 32 reads; the top gene of a big cell holds ten thousand), so three regimes:
//
//   n <= 32      every key ranks itself against the other keys of its segment (pairwise, in the window)
//   n <= 2048    the medium segments of a window are sorted together, one bitonic network over
//                (segment, umi) words in shared memory
//   n  > 2048    sorted in place beforehand by finish_long_kernel (MSD partitions by UMI byte through the spare
//                sort buffer, then the same shared-memory network per bucket); the main kernel streams them
//
// Output: the distinct-key table (dkeys, c0) in key order, placed by a chained scan over the tiles - the table the
// former rle_kernel + run_lengths_kernel produced from fully sorted keys.
#include <algorithm>

#include "kernels.h"

namespace {

constexpr int FS_THREADS = 512;
constexpr int FS_TILE = 2048;         // keys a tile owns (segments that START in it)
constexpr int FS_WIN = 2 * FS_TILE;   // window in shared memory: the tile and its look-ahead
constexpr int FS_PAIR = 32;           // longest segment ranked pairwise
constexpr int FS_MED = FS_TILE;       // longest segment sorted in the window; longer ones are pre-sorted in place
constexpr int FS_WORDS = FS_WIN / 32;
constexpr uint32_t FS_NONE = 0xFFFFFFFFu;

// ascending bitonic sort of s[0..P), P a power of two, by the whole block
__device__ __forceinline__ void block_bitonic(uint32_t* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += FS_THREADS) {
        const int l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const int r = l | j;
        const uint32_t a = s[l], b = s[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) {
          s[l] = b;
          s[r] = a;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int pow2_at_least(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

// [first key of the segment that holds the tile's last key, or FS_NONE if no segment starts in the tile]:
// scans the tile backwards in chunks of FS_THREADS keys and stops at the first chunk with a head
__device__ __forceinline__ uint64_t last_head_in_tile(const unsigned long long* __restrict__ keys, uint64_t t_lo,
                                                      uint64_t t_hi, int ub, unsigned long long* s_tmp) {
  if (threadIdx.x == 0) *s_tmp = ~0ull;
  __syncthreads();
  for (uint64_t top = t_hi; top > t_lo;) {
    const uint64_t bot = top - t_lo > (uint64_t)FS_THREADS ? top - FS_THREADS : t_lo;
    const uint64_t j = bot + threadIdx.x;
    bool head = false;
    if (j < top) head = j == 0 || (keys[j] >> ub) != (keys[j - 1] >> ub);
    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, head);
    if (mk && (threadIdx.x & 31) == 0) {
      const unsigned long long cand = bot + (threadIdx.x & ~31) + (31 - __clz(mk));
      // the largest head wins: atomicMax on the complement-free value (~0 = none is handled by the flag below)
      unsigned long long old = *s_tmp;
      while (old == ~0ull || cand > old) {
        const unsigned long long seen = atomicCAS(s_tmp, old, cand);
        if (seen == old) break;
        old = seen;
      }
    }
    __syncthreads();
    if (*s_tmp != ~0ull) break;
    top = bot;
    __syncthreads();
  }
  const unsigned long long r = *s_tmp;
  __syncthreads();
  return r;
}

// first index >= from whose prefix differs from `pfx` (or n): forward scan by the whole block
__device__ __forceinline__ uint64_t segment_end_from(const unsigned long long* __restrict__ keys, uint64_t from, uint64_t n,
                                                     int ub, unsigned long long pfx, unsigned long long* s_tmp) {
  if (threadIdx.x == 0) *s_tmp = ~0ull;
  __syncthreads();
  uint64_t end = n;
  for (uint64_t base = from;; base += FS_THREADS) {
    const uint64_t j = base + threadIdx.x;
    const bool differs = j >= n || (keys[j] >> ub) != pfx;
    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, differs);
    if (mk && (threadIdx.x & 31) == 0) atomicMin(s_tmp, (unsigned long long)(base + (threadIdx.x & ~31) + (__ffs(mk) - 1)));
    __syncthreads();
    const unsigned long long found = *s_tmp;
    if (found != ~0ull) {
      end = found < n ? found : n;
      break;
    }
    __syncthreads();
  }
  __syncthreads();
  return end;
}

// ---------------------------------------------------------------------------
// long segments: sorted in place on their UMI bits
// ---------------------------------------------------------------------------
// sorts the keys src[lo, hi) (equal above bit `bits`) on their low `bits` bits into dst[lo, hi); src and dst may
// be the same array. hi - lo <= FS_WIN. Bitonic network in shared memory: only for large buckets.
__device__ __forceinline__ void smem_sort_range(const unsigned long long* src, unsigned long long* dst, uint64_t lo,
                                                uint64_t hi, int bits, uint32_t* scratch) {
  const int n = (int)(hi - lo);
  if (n <= 0) return;
  const int P = pow2_at_least(n);
  const unsigned long long upper = src[lo] & ~((1ull << bits) - 1ull);
  __syncthreads();  // everyone has read `upper` before dst (possibly = src) changes; scratch is free
  for (int i = threadIdx.x; i < P; i += FS_THREADS) scratch[i] = i < n ? (uint32_t)(src[lo + i] & ((1ull << bits) - 1ull)) : 0xFFFFFFFFu;
  __syncthreads();
  block_bitonic(scratch, P);
  for (int i = threadIdx.x; i < n; i += FS_THREADS) dst[lo + i] = upper | (unsigned long long)scratch[i];
  __syncthreads();
}

constexpr int FL_MAX_BUCKETS = 2048;  // buckets of one MSD step of a long segment
constexpr int FL_PAIR = 128;          // buckets up to this size are ranked pairwise (straight from L1 / L2)

// one MSD step: src[lo, hi) -> dst[lo, hi) grouped by the `kbits`-bit digit at `shift` (kbits <= 11); bucket starts
// (relative to lo) end up in s_start[0 .. 2^kbits]; s_cnt is scratch of the same size
__device__ __forceinline__ void msd_partition(const unsigned long long* src, unsigned long long* dst, uint64_t lo,
                                              uint64_t hi, int shift, int kbits, uint32_t* s_cnt, uint32_t* s_start,
                                              uint32_t* s_scan) {
  const int nb = 1 << kbits;
  const uint32_t dmask = (uint32_t)nb - 1u;
  for (int i = threadIdx.x; i < nb; i += FS_THREADS) s_cnt[i] = 0u;
  __syncthreads();
  for (uint64_t j = lo + threadIdx.x; j < hi; j += FS_THREADS) atomicAdd(&s_cnt[(uint32_t)(src[j] >> shift) & dmask], 1u);
  __syncthreads();
  // exclusive scan of the counters: every thread takes nb / FS_THREADS consecutive ones (at least one)
  {
    const int per = nb > FS_THREADS ? nb / FS_THREADS : 1;
    const int first = threadIdx.x * per;
    uint32_t sum = 0;
    for (int k = 0; k < per; k++)
      if (first + k < nb) sum += s_cnt[first + k];
    uint32_t total;
    uint32_t run = block_exclusive_scan<FS_THREADS>(sum, &total, s_scan);
    for (int k = 0; k < per; k++)
      if (first + k < nb) {
        s_start[first + k] = run;
        run += s_cnt[first + k];
      }
    if (threadIdx.x == 0) s_start[nb] = total;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += FS_THREADS) s_cnt[i] = s_start[i];  // cursors
  __syncthreads();
  for (uint64_t j = lo + threadIdx.x; j < hi; j += FS_THREADS) {
    const unsigned long long k = src[j];
    const uint32_t pos = atomicAdd(&s_cnt[(uint32_t)(k >> shift) & dmask], 1u);
    dst[lo + pos] = k;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FS_THREADS) finish_long_kernel(unsigned long long* keys, unsigned long long* alt,
                                                                 uint64_t n, int ub) {
  __shared__ uint32_t scratch[FS_WIN];
  __shared__ uint32_t s_cnt[FL_MAX_BUCKETS], s_start[FL_MAX_BUCKETS + 1], s_cnt2[256], s_start2[257];
  __shared__ uint32_t s_scan[FS_THREADS / 32 + 1];
  __shared__ unsigned long long s_tmp;
  const uint64_t t_lo = (uint64_t)blockIdx.x * FS_TILE;
  const uint64_t t_hi = t_lo + FS_TILE < n ? t_lo + FS_TILE : n;
  const unsigned long long a = last_head_in_tile(keys, t_lo, t_hi, ub, &s_tmp);
  if (a == ~0ull) return;  // no segment starts in this tile
  const unsigned long long pfx = keys[t_hi - 1] >> ub;
  const uint64_t b = t_hi < n ? segment_end_from(keys, t_hi, n, ub, pfx, &s_tmp) : n;
  const uint64_t len = b - a;
  if (len <= (uint64_t)FS_MED) return;  // the main kernel sorts it in its window
  // level 1: by the top kbits of the UMI (about 16 keys per bucket for uniform UMIs), keys -> alt
  int kbits = 8;
  while (kbits < 11 && kbits < ub && (len >> (kbits + 4)) > 0) kbits++;
  const int sh1 = ub - kbits;
  msd_partition(keys, alt, a, b, sh1, kbits, s_cnt, s_start, s_scan);
  const int nb = 1 << kbits;
  // small buckets: every key ranks itself among the keys of its bucket and goes straight to its sorted place
  const uint32_t low1 = sh1 >= 32 ? 0xFFFFFFFFu : ((1u << sh1) - 1u);
  for (uint64_t j = a + threadIdx.x; j < b; j += FS_THREADS) {
    const unsigned long long k = alt[j];
    const uint32_t d = (uint32_t)(k >> sh1) & (uint32_t)(nb - 1);
    const uint32_t bs = s_start[d], be = s_start[d + 1];
    if (be - bs > (uint32_t)FL_PAIR) continue;
    const uint32_t mine = (uint32_t)k & low1;
    uint32_t lt = 0, eq_before = 0;
    for (uint32_t q = bs; q < be; q++) {
      const uint32_t x = (uint32_t)alt[a + q] & low1;
      lt += x < mine;
      eq_before += (x == mine) & (a + q < j);
    }
    keys[a + bs + lt + eq_before] = k;
  }
  __syncthreads();
  // large buckets (skewed UMIs, or a segment of more than 64 k keys)
  for (int d = 0; d < nb; d++) {
    const uint64_t lo = a + s_start[d], hi = a + s_start[d + 1];
    if (hi - lo <= (uint64_t)FL_PAIR) continue;
    if (hi - lo <= (uint64_t)FS_WIN) {
      smem_sort_range(alt, keys, lo, hi, sh1, scratch);  // sorted bucket back into place
      continue;
    }
    // level 2: by the next byte, alt -> keys
    const int k2 = sh1 >= 8 ? 8 : sh1;
    const int sh2 = sh1 - k2;
    msd_partition(alt, keys, lo, hi, sh2, k2, s_cnt2, s_start2, s_scan);
    for (int e = 0; e < (1 << k2); e++) {
      const uint64_t lo2 = lo + s_start2[e], hi2 = lo + s_start2[e + 1];
      if (hi2 == lo2) continue;
      if (hi2 - lo2 <= (uint64_t)FS_WIN) {
        smem_sort_range(keys, keys, lo2, hi2, sh2, scratch);
        continue;
      }
      // level 3: at most 8 bits are left (ub <= 24 and kbits >= 8 leave sh1 <= 16, sh2 <= 8): count the values
      // and write them back in order
      const int nv = 1 << sh2;
      const unsigned long long upper = keys[lo2] & ~((1ull << sh2) - 1ull);
      __syncthreads();
      for (int i = threadIdx.x; i <= nv; i += FS_THREADS) scratch[i] = 0u;
      __syncthreads();
      for (uint64_t j = lo2 + threadIdx.x; j < hi2; j += FS_THREADS) atomicAdd(&scratch[(uint32_t)keys[j] & (uint32_t)(nv - 1)], 1u);
      __syncthreads();
      if (threadIdx.x == 0) {  // exclusive scan in place, nv <= 256
        uint32_t run = 0;
        for (int v = 0; v < nv; v++) {
          const uint32_t c = scratch[v];
          scratch[v] = run;
          run += c;
        }
        scratch[nv] = run;
      }
      __syncthreads();
      for (int v = 0; v < nv; v++)
        for (uint64_t j = lo2 + scratch[v] + threadIdx.x; j < lo2 + scratch[v + 1]; j += FS_THREADS)
          keys[j] = upper | (unsigned long long)v;
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
// main kernel: per-segment sort + run-length encoding + ordered placement
// ---------------------------------------------------------------------------
constexpr int FS_MSEG = FS_WIN / 32;     // medium segments are told apart by (start >> 5): they start > 32 apart
constexpr int FS_MBUCKETS = FS_WIN / 2;  // buckets of the medium segments of one window (len / 4 each, at least 1)

struct FinishSmem {
  unsigned long long w_key[FS_WIN];
  uint32_t st_c0[FS_WIN];     // read count of the distinct key staged at a window position
  uint32_t sc_umi[FS_WIN];    // the UMIs of the medium segments, grouped by bucket
  uint16_t st_src[FS_WIN];    // window position of a key that holds the staged distinct key
  uint16_t sc_pos[FS_WIN];    // window position of each sc_umi entry
  uint16_t seg_a[FS_WIN];     // window position -> start of its segment (0xFFFF: it started in an earlier tile)
  uint16_t seg_b[FS_WIN];     // ... -> end of its segment (0xFFFF: it runs past the window)
  uint32_t bucket[FS_MBUCKETS + 2];  // bucket counters, then bucket ends
  uint32_t heads[FS_WORDS];
  uint32_t outflag[FS_WORDS];
  uint32_t m_min[FS_MSEG], m_max[FS_MSEG], m_len[FS_MSEG], m_base[FS_MSEG], m_nb[FS_MSEG];
  float m_scale[FS_MSEG];
  uint32_t scan[FS_THREADS / 32 + 1];
  unsigned long long bcast, tmp;
  uint32_t any_med, n_buckets, own_lo, last_start;
};

// bucket of a UMI inside its medium segment: an order-preserving linear map of [min, max] onto the segment's buckets
__device__ __forceinline__ uint32_t medium_bucket(const FinishSmem& S, uint32_t s, uint32_t umi) {
  uint32_t r = (uint32_t)((float)(umi - S.m_min[s]) * S.m_scale[s]);
  const uint32_t nb = S.m_nb[s];
  if (r >= nb) r = nb - 1u;
  return S.m_base[s] + r;
}

// length of the run of keys equal to window key p (which starts it): in the window, then on in global memory
__device__ __forceinline__ uint32_t run_length_from(const FinishSmem& S, const unsigned long long* __restrict__ keys,
                                                    uint64_t t_lo, int wn, uint64_t n, int p) {
  const unsigned long long k = S.w_key[p];
  int q = p + 1;
  while (q < wn && S.w_key[q] == k) q++;
  uint64_t len = (uint64_t)(q - p);
  if (q == wn) {
    uint64_t g = t_lo + (uint64_t)wn;
    while (g < n && keys[g] == k) g++;
    len = g - (t_lo + (uint64_t)p);
  }
  return (uint32_t)len;
}

__global__ void __launch_bounds__(FS_THREADS) finish_rle_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                                                int ub, unsigned long long* __restrict__ dkeys,
                                                                uint32_t* __restrict__ c0, unsigned long long* desc,
                                                                unsigned long long* total_out) {
  extern __shared__ __align__(16) unsigned char fs_raw[];
  FinishSmem& S = *reinterpret_cast<FinishSmem*>(fs_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint64_t t_lo = (uint64_t)tile * FS_TILE;
  const int tile_n = (int)((n - t_lo) < (uint64_t)FS_TILE ? (n - t_lo) : (uint64_t)FS_TILE);
  const int wn = (int)((n - t_lo) < (uint64_t)FS_WIN ? (n - t_lo) : (uint64_t)FS_WIN);
  const bool closed = t_lo + (uint64_t)wn == n;  // the window reaches the end of the keys
  const uint32_t umask = (uint32_t)((1ull << ub) - 1ull);
  const uint64_t n_tiles = (n + FS_TILE - 1) / FS_TILE;

  for (int i = tid; i < wn; i += FS_THREADS) S.w_key[i] = __ldcs(keys + t_lo + i);
  for (int i = tid; i < FS_WORDS; i += FS_THREADS) S.outflag[i] = 0u;
  for (int i = tid; i < FS_MSEG; i += FS_THREADS) {
    S.m_min[i] = 0xFFFFFFFFu;
    S.m_max[i] = 0u;
    S.m_len[i] = 0u;
  }
  if (tid == 0) {
    S.any_med = 0u;
    S.own_lo = FS_NONE;
    S.last_start = FS_NONE;
  }
  const unsigned long long kprev = t_lo > 0 ? keys[t_lo - 1] : 0ull;
  __syncthreads();
  // segment heads: bit i set if window key i starts a segment
  for (int i = tid; i < FS_WIN; i += FS_THREADS) {  // warp-uniform trip count
    bool head = false;
    if (i < wn) head = i == 0 ? (t_lo == 0 || (S.w_key[0] >> ub) != (kprev >> ub)) : (S.w_key[i] >> ub) != (S.w_key[i - 1] >> ub);
    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, head);
    if (lane == 0) S.heads[i >> 5] = mk;
  }
  __syncthreads();
  // segment start of every position: running maximum of the head positions (per thread 8 consecutive positions,
  // block scan of the per-thread maxima); segment end: the next head, by the mirrored scan
  constexpr int PER = FS_WIN / FS_THREADS;  // 8
  static_assert(PER == 8, "a thread's positions share one flag byte");
  {
    const int first = tid * PER;
    const uint32_t hb = (S.heads[first >> 5] >> (first & 31)) & 0xFFu;  // the head bits of this thread's span
    int last_head = -1;
    int loc[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      if ((hb >> k) & 1u) last_head = first + k;
      loc[k] = last_head;
    }
    int run = last_head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, run, d);
      if (lane >= d && o > run) run = o;
    }
    if (lane == 31) S.scan[warp] = (uint32_t)run;
    __syncthreads();
    int carry = -1;
    for (int w = 0; w < warp; w++) carry = max(carry, (int)S.scan[w]);
    int before = __shfl_up_sync(0xFFFFFFFFu, run, 1);
    if (lane == 0) before = -1;
    before = max(before, carry);
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int a = loc[k] >= 0 ? loc[k] : before;  // -1: the position continues a segment of an earlier tile
      S.seg_a[first + k] = (uint16_t)(a < 0 ? 0xFFFF : a);
    }
    __syncthreads();
    int next_head = FS_WIN + 1;  // next head strictly after each position (mirrored)
    int locb[PER];
#pragma unroll
    for (int k = PER - 1; k >= 0; k--) {
      locb[k] = next_head;
      if ((hb >> k) & 1u) next_head = first + k;
    }
    int runb = next_head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_down_sync(0xFFFFFFFFu, runb, d);
      if (lane + d < 32 && o < runb) runb = o;
    }
    if (lane == 0) S.scan[warp] = (uint32_t)runb;
    __syncthreads();
    int carryb = FS_WIN + 1;
    for (int w = warp + 1; w < FS_THREADS / 32; w++) carryb = min(carryb, (int)S.scan[w]);
    int after = __shfl_down_sync(0xFFFFFFFFu, runb, 1);
    if (lane == 31) after = FS_WIN + 1;
    after = min(after, carryb);
#pragma unroll
    for (int k = 0; k < PER; k++) {
      int b = locb[k] <= FS_WIN ? locb[k] : after;
      if (b > wn) b = closed ? wn : FS_WIN + 1;  // no head up to the window end: closed only at the end of the keys
      S.seg_b[first + k] = (uint16_t)(b > FS_WIN ? 0xFFFF : b);
    }
  }
  // first and last segment start inside the tile
  for (int w = tid; w < FS_TILE / 32; w += FS_THREADS) {
    uint32_t mk = S.heads[w];
    const int left = tile_n - w * 32;
    if (left < 32) mk &= left > 0 ? (0xFFFFFFFFu >> (32 - left)) : 0u;
    if (mk) {
      atomicMin(&S.own_lo, (uint32_t)(w * 32 + __ffs(mk) - 1));
      atomicMin(&S.last_start, 0xFFFFFFFEu - (uint32_t)(w * 32 + 31 - __clz(mk)));  // min of the complement = max
    }
  }
  __syncthreads();
  const bool owns = S.own_lo != FS_NONE;
  const int own_lo = owns ? (int)S.own_lo : tile_n;  // the keys before it continue a segment of an earlier tile
  const int last_start = owns ? (int)(0xFFFFFFFEu - S.last_start) : 0;
  // The segments this tile ranks ("stages") are the ones that start in it and are at most FS_MED long. Longer ones
  // have been sorted in place by finish_long_kernel: every tile run-length encodes the part of them that lies
  // inside its own FS_TILE keys - its leading keys if they continue a long segment, its trailing keys if its last
  // segment is long - so that a segment of a million keys is spread over five hundred blocks.
  int stage_hi = own_lo;
  bool tail_long = false;
  if (owns) {
    const uint32_t eb = S.seg_b[last_start];
    if (eb != 0xFFFFu && (int)eb - last_start <= FS_MED) {
      stage_hi = (int)eb;
    } else {
      stage_hi = last_start;
      tail_long = true;
    }
  }
  bool lead_long = false;
  if (own_lo > 0 && t_lo > 0) {  // block-uniform: the leading keys [0, own_lo) continue the segment of key t_lo - 1
    // its end: the first head of the window (own_lo if the tile has one); its start: somewhere before the tile
    const uint32_t eb = owns ? (uint32_t)own_lo : (uint32_t)S.seg_b[0];
    if (eb == 0xFFFFu || (int)eb + 1 > FS_MED) {
      lead_long = true;  // at least one key before the tile and eb keys in it
    } else {
      // look back for its start, FS_MED - eb keys at most: more than that and it is long
      const unsigned long long pfx = S.w_key[0] >> ub;
      const uint64_t max_back = (uint64_t)(FS_MED - (int)eb);  // keys before t_lo that still make a medium segment
      if (tid == 0) S.tmp = ~0ull;
      __syncthreads();
      bool found = false;
      for (uint64_t off = 0; off <= max_back && !found; off += FS_THREADS) {
        const uint64_t back = off + (uint64_t)tid;  // the key t_lo - 1 - back
        bool differs = false;
        if (back <= max_back) differs = back >= t_lo || (keys[t_lo - 1 - back] >> ub) != pfx;
        const uint32_t mk = __ballot_sync(0xFFFFFFFFu, differs);
        if (mk && lane == 0) atomicMin(&S.tmp, (unsigned long long)(off + (uint64_t)(tid & ~31) + (uint64_t)(__ffs(mk) - 1)));
        __syncthreads();
        found = S.tmp != ~0ull;
        __syncthreads();
      }
      // S.tmp = number of keys of the segment before the tile (the first `back` at which the prefix differs)
      lead_long = !found || S.tmp + (unsigned long long)eb > (unsigned long long)FS_MED;
      __syncthreads();
    }
  }

  // ---- the keys of pre-sorted long segments inside this tile: run heads, counted where they stand ----
  {
    const int lead_hi = lead_long ? own_lo : 0;
    const int tail_lo = tail_long ? last_start : tile_n;
    for (int p = tid; p < tile_n; p += FS_THREADS) {
      if (p >= lead_hi && p < tail_lo) continue;
      const bool head = p == 0 ? (t_lo == 0 || S.w_key[0] != kprev) : S.w_key[p] != S.w_key[p - 1];
      if (!head) continue;
      S.st_src[p] = (uint16_t)p;
      S.st_c0[p] = run_length_from(S, keys, t_lo, wn, n, p);
      atomicOr(&S.outflag[p >> 5], 1u << (p & 31));
    }
  }

  // ---- rank the staged keys: pairwise in short segments; medium segments collect their UMI range first ----
  for (int i = own_lo + tid; i < stage_hi; i += FS_THREADS) {
    const int a = S.seg_a[i], b = S.seg_b[i];
    const uint32_t mine = (uint32_t)S.w_key[i] & umask;
    if (b - a <= FS_PAIR) {
      uint32_t lt = 0, eq = 0, eq_before = 0;
      for (int j = a; j < b; j++) {
        const uint32_t x = (uint32_t)S.w_key[j] & umask;
        lt += x < mine;
        eq += x == mine;
        eq_before += (x == mine) & (j < i);
      }
      if (eq_before == 0) {  // the first of its equals stands for the distinct key
        const int p = a + (int)lt;
        S.st_src[p] = (uint16_t)i;
        S.st_c0[p] = eq;
        atomicOr(&S.outflag[p >> 5], 1u << (p & 31));
      }
    } else {
      const uint32_t s = (uint32_t)a >> 5;
      atomicMin(&S.m_min[s], mine);
      atomicMax(&S.m_max[s], mine);
      if (i == a) {
        S.m_len[s] = (uint32_t)(b - a);
        S.any_med = 1u;
      }
    }
  }
  __syncthreads();
  if (S.any_med) {  // block-uniform
    // buckets: len / 4 per medium segment (about four keys each for uniform UMIs), numbered through the window
    if (warp == 0) {
      uint32_t nbk[FS_MSEG / 32], sum = 0;
#pragma unroll
      for (int k = 0; k < FS_MSEG / 32; k++) {
        const uint32_t len = S.m_len[lane * (FS_MSEG / 32) + k];
        nbk[k] = len ? max(1u, len >> 2) : 0u;
        sum += nbk[k];
      }
      uint32_t inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
      }
      uint32_t run = inc - sum;
#pragma unroll
      for (int k = 0; k < FS_MSEG / 32; k++) {
        const int s = lane * (FS_MSEG / 32) + k;
        S.m_base[s] = run;
        S.m_nb[s] = nbk[k];
        if (nbk[k]) S.m_scale[s] = (float)nbk[k] / ((float)(S.m_max[s] - S.m_min[s]) + 1.0f);
        run += nbk[k];
      }
      if (lane == 31) S.n_buckets = run;
    }
    for (int i = tid; i < FS_MBUCKETS + 2; i += FS_THREADS) S.bucket[i] = 0u;
    __syncthreads();
    // count
    for (int i = own_lo + tid; i < stage_hi; i += FS_THREADS) {
      const int a = S.seg_a[i];
      if ((int)S.seg_b[i] - a <= FS_PAIR) continue;
      atomicAdd(&S.bucket[medium_bucket(S, (uint32_t)a >> 5, (uint32_t)S.w_key[i] & umask)], 1u);
    }
    __syncthreads();
    // exclusive scan of the counters (FS_MBUCKETS / FS_THREADS = 4 per thread)
    {
      constexpr int BPT = FS_MBUCKETS / FS_THREADS;
      const int first = tid * BPT;
      uint32_t v[BPT], sum = 0;
#pragma unroll
      for (int k = 0; k < BPT; k++) {
        v[k] = S.bucket[first + k];
        sum += v[k];
      }
      uint32_t total;
      uint32_t run = block_exclusive_scan<FS_THREADS>(sum, &total, S.scan);
#pragma unroll
      for (int k = 0; k < BPT; k++) {
        S.bucket[first + k] = run;
        run += v[k];
      }
    }
    __syncthreads();
    // scatter: afterwards bucket[b] is the END of bucket b (and bucket[b - 1] its start)
    for (int i = own_lo + tid; i < stage_hi; i += FS_THREADS) {
      const int a = S.seg_a[i];
      if ((int)S.seg_b[i] - a <= FS_PAIR) continue;
      const uint32_t mine = (uint32_t)S.w_key[i] & umask;
      const uint32_t p = atomicAdd(&S.bucket[medium_bucket(S, (uint32_t)a >> 5, mine)], 1u);
      S.sc_umi[p] = mine;
      S.sc_pos[p] = (uint16_t)i;
    }
    __syncthreads();
    // rank inside the bucket
    for (int i = own_lo + tid; i < stage_hi; i += FS_THREADS) {
      const int a = S.seg_a[i];
      if ((int)S.seg_b[i] - a <= FS_PAIR) continue;
      const uint32_t s = (uint32_t)a >> 5;
      const uint32_t mine = (uint32_t)S.w_key[i] & umask;
      const uint32_t bk = medium_bucket(S, s, mine);
      const uint32_t bs = bk ? S.bucket[bk - 1] : 0u, be = S.bucket[bk];
      const uint32_t seg_first = S.m_base[s] ? S.bucket[S.m_base[s] - 1] : 0u;  // scratch index of the segment's first key
      uint32_t lt = 0, eq = 0, eq_before = 0;
      for (uint32_t q = bs; q < be; q++) {
        const uint32_t x = S.sc_umi[q];
        lt += x < mine;
        eq += x == mine;
        eq_before += (x == mine) & ((int)S.sc_pos[q] < i);
      }
      if (eq_before == 0) {
        const int p = a + (int)(bs - seg_first + lt);
        S.st_src[p] = (uint16_t)i;
        S.st_c0[p] = eq;
        atomicOr(&S.outflag[p >> 5], 1u << (p & 31));
      }
    }
  }
  __syncthreads();

  // ---- ordered placement ----
  const int first = tid * PER;
  const uint32_t bits = (S.outflag[first >> 5] >> (first & 31)) & 0xFFu;
  uint32_t total;
  const uint32_t off = block_exclusive_scan<FS_THREADS>((uint32_t)__popc(bits), &total, S.scan);
  const unsigned long long excl = lookback_exclusive(desc, tile, (unsigned long long)total, &S.bcast);
  if (tile == n_tiles - 1 && tid == 0) *total_out = excl + total;
  uint32_t todo = bits;
  uint64_t o = excl + off;
  while (todo) {
    const int k = __ffs(todo) - 1;
    todo &= todo - 1u;
    const int p = first + k;
    dkeys[o] = S.w_key[S.st_src[p]];
    c0[o] = S.st_c0[p];
    o++;
  }
}

}  // namespace

bool finish_supported(int umi_bits) { return umi_bits >= 16 && umi_bits <= 24; }

// keys: sorted on the bits above the UMI; alt: the spare buffer of the sort (same size); desc: (n / FS_TILE + 2) u64
int run_finish(unsigned long long* keys, unsigned long long* alt, uint64_t n, int umi_bits, unsigned long long* dkeys,
               uint32_t* c0, unsigned long long* desc, unsigned long long* total_out, cudaStream_t st,
               void (*mark)(void*, const char*), void* mark_user) {
  if (n == 0) {
    cudaMemsetAsync(total_out, 0, 8, st);
    return 0;
  }
  const uint64_t tiles = (n + FS_TILE - 1) / FS_TILE;
  cudaMemsetAsync(desc, 0, tiles * 8, st);
  finish_long_kernel<<<(unsigned)tiles, FS_THREADS, 0, st>>>(keys, alt, n, umi_bits);
  if (mark) mark(mark_user, "count.dedup.finish_rle");
  const size_t smem = sizeof(FinishSmem);
  cudaFuncSetAttribute(finish_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  finish_rle_kernel<<<(unsigned)tiles, FS_THREADS, smem, st>>>(keys, n, umi_bits, dkeys, c0, desc, total_out);
  return 2;
}
size_t finish_desc_bytes(uint64_t n) { return ((n + FS_TILE - 1) / FS_TILE + 2) * 8; }

void finish_lb_flag_fetch(unsigned int* host_out, cudaStream_t st) {
  cudaMemcpyFromSymbolAsync(host_out, lb_timeout_flag, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st);
}
void finish_lb_flag_clear(cudaStream_t st) {
  static const unsigned int zero = 0;
  cudaMemcpyToSymbolAsync(lb_timeout_flag, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st);
}
