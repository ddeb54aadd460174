// common.cuh — shared device helpers for libcrgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define CRGPU_MAX_ORD 8
#define CRGPU_MAX_LIBS 4
#define CRGPU_MAX_PARTS 16

// SM count of the current device (launch grids are sized in multiples of it); cached per device
inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// bc_out word: state in the top 2 bits, content rank below (CRGPU_NO_RANK when invalid)
#define BC_STATE_SHIFT 30
#define BC_RANK_MASK 0x3FFFFFFFu
#define ST_NOT_CHECKED 0u
#define ST_VALID_BEFORE 1u
#define ST_VALID_AFTER 2u
#define ST_INVALID 3u

// umi_out word: bit 31 = UmiInfo::is_valid, bit 30 = holds a non-ACGT base, low bits = 2-bit packed UMI
#define UMI_VALID_BIT 0x80000000u
#define UMI_HASN_BIT 0x40000000u
#define UMI_SEQ_MASK 0x3FFFFFFFu

#define NO_FEATURE 0xFFFFFFFFu

// Whitelist in device memory: the same W keys held in up to CRGPU_MAX_ORD "orderings". Ordering o keeps
// the keys rotated right by rot[o] bits (inside 2L bits) and sorted, with a prefix-offset table over the
// top p = 2L - s bits. Every Hamming-1 neighbour of a query whose differing base lies in the low s bits
// of the rotated key shares the query's bucket, so one bucket scan per ordering answers the membership
// of all 3L neighbours (Whitelist::check_and_update, lib/rust/barcode/src/whitelist.rs:494-516, called
// 3L times per invalid barcode by Posterior::correct_barcode, lib/rust/barcode/src/corrector.rs:125-149).
struct DevWhitelist {
  int L;       // bases per barcode (<= 16)
  int s;       // suffix bits per bucket (even)
  int n_ord;   // orderings
  uint32_t W;  // entries
  const uint32_t* keys[CRGPU_MAX_ORD];
  const uint32_t* vals[CRGPU_MAX_ORD];  // content rank of each entry; nullptr = the entry's own index
  const uint32_t* offs[CRGPU_MAX_ORD];  // (1 << p) + 1 bucket starts
  int rot[CRGPU_MAX_ORD];
  // sfx[o][i] = low s bits of keys[o][i] as 16 bits (s <= 16), what the neighbour scans read: half the bytes of
  // the full keys, so that all orderings together stay L2-resident. nullptr when s > 16; keys[o > 0] is then kept.
  const uint16_t* sfx[CRGPU_MAX_ORD];
  uint32_t resp[CRGPU_MAX_ORD];  // bit `pos` set: this ordering answers for mutations at base `pos`
  // exact membership: one 32-byte slot (one L2 sector) per bucket of the top (2L - slot_shift) key bits,
  // fetched with two 128-bit loads: word 0 = index of the bucket's first entry in keys[0] (low 27 bits) |
  // entry count (top 5 bits, 31 = more than WL_SLOT_CAP: the rest follows in keys[0]); words 1..7 = up to
  // fourteen 16-bit key suffixes (the low slot_shift bits), 0xFFFF-padded. About 6.5 entries per bucket, so
  // that fewer than 1 % of the lookups leave the slot.
  const uint4* slots;
  int slot_shift;
};

// how the 64-bit dedup key is laid out: rank | feature | library | umi
struct KeyLayout {
  int rank_shift, feature_shift, lib_shift, umi_bits;
  int total_bits;
};

__device__ __forceinline__ uint32_t mask_bits(int n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }

__device__ __forceinline__ uint32_t rotr_bits(uint32_t q, int r, int nbits) {
  if (r == 0) return q;
  return ((q >> r) | (q << (nbits - r))) & mask_bits(nbits);
}

// exact membership. Returns the index of the entry in keys[0] (sorted, rot 0) or -1.
// Split in three so that callers can issue the loads of several independent lookups before resolving any.
constexpr int WL_SLOT_CAP = 14;
struct WlProbe {
  uint4 s0, s1;
};
__device__ __forceinline__ uint32_t wl_find_begin(const DevWhitelist& wl, uint32_t q) {
  return wl.slot_shift >= 32 ? 0u : (q >> wl.slot_shift);  // bucket
}
__device__ __forceinline__ WlProbe wl_find_probe(const DevWhitelist& wl, uint32_t bucket) {
  WlProbe p;
  p.s0 = __ldg(wl.slots + 2 * (size_t)bucket);
  p.s1 = __ldg(wl.slots + 2 * (size_t)bucket + 1);
  return p;
}
__device__ __forceinline__ int wl_find_end(const DevWhitelist& wl, const WlProbe& p, uint32_t q) {
  const uint32_t cnt = p.s0.x >> 27, base = p.s0.x & 0x07FFFFFFu;
  if (cnt == 0u) return -1;
  const uint32_t qs = q & mask_bits(wl.slot_shift);
  const uint32_t qq = qs | (qs << 16);
  const uint32_t w[7] = {p.s0.y, p.s0.z, p.s0.w, p.s1.x, p.s1.y, p.s1.z, p.s1.w};
  int j = -1;
#pragma unroll
  for (int i = 6; i >= 0; i--) {
    const uint32_t x = w[i] ^ qq;
    if (!(x >> 16)) j = 2 * i + 1;
    if (!(x & 0xFFFFu)) j = 2 * i;
  }
  if (cnt <= (uint32_t)WL_SLOT_CAP) return (j >= 0 && (uint32_t)j < cnt) ? (int)(base + (uint32_t)j) : -1;
  // crowded bucket: the first fourteen entries are inline, the rest follow in the sorted keys (which end
  // with 0xFFFFFFFF sentinels)
  if (j >= 0) return (int)(base + (uint32_t)j);
  const uint32_t* __restrict__ keys = wl.keys[0];
  uint32_t idx = base + WL_SLOT_CAP;
  while (true) {
    uint32_t e = __ldg(keys + idx);
    if (e >= q) return (e == q && idx < wl.W) ? (int)idx : -1;
    idx++;
  }
}
__device__ __forceinline__ int wl_find(const DevWhitelist& wl, uint32_t q) {
  return wl_find_end(wl, wl_find_probe(wl, wl_find_begin(wl, q)), q);
}

__device__ __forceinline__ uint32_t wl_rank_of(const DevWhitelist& wl, int idx) {
  return wl.vals[0] ? __ldg(wl.vals[0] + idx) : (uint32_t)idx;
}

// Bit (pos*4 + base) set iff replacing base `pos` of q by `base` (!= the original) gives a whitelist entry.
__device__ __forceinline__ unsigned long long wl_neighbor_mask(const DevWhitelist& wl, uint32_t q) {
  unsigned long long m = 0ull;
  const int nbits = 2 * wl.L;
  for (int o = 0; o < wl.n_ord; o++) {
    const uint32_t* __restrict__ offs = wl.offs[o];
    const uint32_t* __restrict__ keys = wl.keys[o];
    const int r = wl.rot[o];
    uint32_t rq = rotr_bits(q, r, nbits);
    uint32_t b = wl.s >= 32 ? 0u : (rq >> wl.s);
    uint32_t lo = __ldg(offs + b), hi = __ldg(offs + b + 1);
    const uint32_t resp = wl.resp[o];
    const uint16_t* __restrict__ sfx = wl.sfx[o];
    const uint32_t smask = mask_bits(wl.s);
    for (uint32_t i = lo; i < hi; i++) {
      // the bucket fixes the prefix: only the suffix can differ
      const uint32_t e = sfx ? (uint32_t)__ldg(sfx + i) : (__ldg(keys + i) & smask);
      uint32_t x = e ^ (rq & smask);
      uint32_t y = (x | (x >> 1)) & 0x55555555u;
      if (y != 0u && (y & (y - 1u)) == 0u) {  // exactly one base differs
        int k = (31 - __clz(y)) >> 1;         // base index from the right in the rotated key
        int kk = k + (r >> 1);                // ... in the original key
        if (kk >= wl.L) kk -= wl.L;
        int pos = wl.L - 1 - kk;
        if ((resp >> pos) & 1u) {
          uint32_t base = (e >> (2 * k)) & 3u;
          m |= 1ull << (pos * 4 + base);
        }
      }
    }
  }
  return m;
}

// 4 ASCII bases in a little-endian word -> 8 bits, first base most significant; *bad gets 0x80 per byte
// that is not one of A,C,G,T.
__device__ __forceinline__ uint32_t pack4(uint32_t w, uint32_t* bad) {
  uint32_t t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
  uint32_t lo = t & 0x01010101u, hi = (t >> 1) & 0x01010101u;
  uint32_t e = 0x41414141u + 2u * lo + 6u * hi + 11u * (lo & hi);  // A 41, C 43, G 47, T 54
  uint32_t d = w ^ e;
  *bad = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
  return (t * 0x40100401u) >> 24;
}

// per byte 0x80 where 33 <= q < 43 (UMI_MIN_QV = 10, lib/rust/umi/src/info.rs:6,66-73; u8 wrap-around
// below 33 as in a release build). Bytes are assumed < 128.
__device__ __forceinline__ uint32_t lowqual4(uint32_t q) {
  uint32_t b = q + 0x5F5F5F5Fu;  // q + 95 per byte, no carry for q < 161
  return b & ~((b & 0x7F7F7F7Fu) + 0x76767676u) & 0x80808080u;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// block-wide exclusive scan of one small count per thread; returns the thread's offset, *total the sum.
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*THREADS/32+1*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += n;
  }
  __syncthreads();  // protect smem reuse across calls
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  constexpr int NW = THREADS / 32;
  uint32_t wsum = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NW; w++) {
    uint32_t c = smem[w];
    if (w < warp) wsum += c;
    tot += c;
  }
  *total = tot;
  return wsum + inc - v;
}

// ---------------------------------------------------------------------------
// single-pass chained scan (decoupled look-back) descriptors: a 64-bit word
// {status:2 | value:62}; tiles must be processed in an order in which every
// predecessor of a running tile is running or finished (tile ids are handed out
// by an atomic ticket).
// ---------------------------------------------------------------------------
#define LB_INVALID 0ull
#define LB_AGGREGATE 1ull
#define LB_PREFIX 2ull

__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long status, unsigned long long v) {
  unsigned long long w = (status << 62) | v;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// Watchdog of the chained scans. A block spins on its predecessors' descriptors; that terminates only if every
// predecessor of a resident block is itself resident or finished. With blockIdx-ordered dispatch (what the
// hardware does today) this holds; should a scheduler ever break it (MPS time slicing, preemption), the spin
// is bounded: after LB_SPIN_LIMIT polls the block raises this translation unit's flag and carries on with a
// zero prefix, and the host entry point that launched the scan fails with CRGPU_E_CUDA instead of hanging.
#define LB_SPIN_LIMIT (1u << 22)
static __device__ unsigned int lb_timeout_flag;
__device__ __forceinline__ void lb_raise_timeout() { atomicExch(&lb_timeout_flag, 1u); }

// Tile index of a block in a chained scan. Default: blockIdx.x - blocks are dispatched in index order, so every
// predecessor of a resident block is itself resident or finished and the look-back cannot starve (the same
// assumption CUB's single-pass scan makes). With ticket != nullptr (CRGPU_TICKETS=1) the index comes from an
// atomic counter instead, which costs one global round trip and a barrier before the first load of the tile.
__device__ __forceinline__ uint32_t acquire_tile(uint32_t* ticket, uint32_t* tile_s) {
  if (!ticket) return blockIdx.x;
  if (threadIdx.x == 0) *tile_s = atomicAdd(ticket, 1u);
  __syncthreads();
  return *tile_s;
}
inline uint32_t* tile_ticket(uint32_t* ticket) {  // host: which ticket pointer to hand to a kernel
  static const bool use = getenv("CRGPU_TICKETS") && atoi(getenv("CRGPU_TICKETS")) != 0;
  return use ? ticket : nullptr;
}

// Called by every thread of the block with the block aggregate; returns the exclusive prefix of the tile.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* desc, uint32_t tile,
                                                                 unsigned long long aggregate,
                                                                 unsigned long long* smem_bcast) {
  if (threadIdx.x < 32) {
    unsigned long long excl = 0ull;
    if (tile == 0) {
      if (threadIdx.x == 0) lb_store(desc, LB_PREFIX, aggregate);
    } else {
      if (threadIdx.x == 0) lb_store(desc + tile, LB_AGGREGATE, aggregate);
      int j = (int)tile - 1 - lane_id();
      while (true) {
        unsigned long long w = LB_PREFIX << 62;  // lanes before tile 0 act as a zero prefix
        if (j >= 0) {
          uint32_t spins = 0;
          while (true) {
            w = lb_load(desc + j);
            if ((w >> 62) != LB_INVALID) break;
            if (++spins > LB_SPIN_LIMIT) {  // watchdog: fail loudly instead of hanging
              lb_raise_timeout();
              w = LB_PREFIX << 62;
              break;
            }
            if (spins > 64u) __nanosleep(64);
          }
        }
        unsigned status = (unsigned)(w >> 62);
        unsigned long long val = w & 0x3FFFFFFFFFFFFFFFull;
        unsigned pmask = __ballot_sync(0xFFFFFFFFu, status == LB_PREFIX);
        // lanes up to and including the first PREFIX lane contribute
        int first = pmask ? (__ffs(pmask) - 1) : 32;
        if (lane_id() > first) val = 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
        excl += val;
        if (pmask) break;
        j -= 32;
      }
      if (threadIdx.x == 0) lb_store(desc + tile, LB_PREFIX, excl + aggregate);
    }
    if (threadIdx.x == 0) *smem_bcast = excl;
  }
  __syncthreads();
  return *smem_bcast;
}
