// sort.cu — device radix sort of the packed 64-bit dedup keys.
//
// Least-significant-digit radix sort restricted to the bits the key layout really uses
// (KeyLayout::total_bits). The digit width RB is a template parameter: 8 bits per pass; 9 (CRGPU_SORT_BITS=9:
// 7 passes instead of 8 for a 62-bit 3' v3 key) is kept as a measured experiment - it is slower on B200.
// One upfront kernel builds the digit histograms of every pass; each pass is then a single "onesweep"
// kernel: a tile of keys is ranked inside the block (warp match-any, warp-private digit counters), the
// tile's digit counts are published through a chained (decoupled look-back) scan, and the keys are
// scattered through shared memory so every digit bin is written as one contiguous run.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace {

constexpr int MAX_RADIX_BITS = 9;
constexpr int MAX_RADIX = 1 << MAX_RADIX_BITS;
constexpr int MAX_PASSES = 8;
constexpr int MIN_TILE = 2048;  // smallest tile of any configuration (sizes the descriptor scratch)

// ---- upfront histograms: hist[pass][digit] over all keys ----
template <int RADIX_BITS>
__global__ void __launch_bounds__(512) radix_hist_kernel(const unsigned long long* __restrict__ keys, uint64_t n,
                                                         int n_passes, int begin_bit, unsigned long long* __restrict__ hist) {
  constexpr int RADIX = 1 << RADIX_BITS;
  __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
  for (int i = threadIdx.x; i < n_passes * RADIX; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long k = keys[i];
    for (int p = 0; p < n_passes; p++) atomicAdd(&s_hist[p * RADIX + (int)((k >> (begin_bit + p * RADIX_BITS)) & (RADIX - 1))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_passes * RADIX; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}

// exclusive scan of each pass's bins (one block of RADIX threads per pass)
template <int RADIX>
__global__ void radix_scan_hist_kernel(unsigned long long* hist) {
  __shared__ unsigned long long s[RADIX];
  unsigned long long* h = hist + (size_t)blockIdx.x * RADIX;
  s[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int d = 0; d < RADIX; d++) {
      unsigned long long c = s[d];
      s[d] = run;
      run += c;
    }
  }
  __syncthreads();
  h[threadIdx.x] = s[threadIdx.x];
}

// ---- one onesweep pass ----
// desc[tile * RADIX + digit]: {status:2 | count:62} chained-scan descriptors of this pass.
// One tile of one pass. FULL = the tile holds SORT_TILE keys (no bounds checks in the hot loops).
// lanes of the warp holding the same 9-bit value: eight ballots instead of match.any, whose cost grows with
// the number of distinct values in the warp (about 30 for a uniformly distributed digit)
template <int NBITS>
__device__ __forceinline__ uint32_t match_digit(uint32_t digit) {
  uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
  for (int b = 0; b < NBITS; b++) {
    // one predicate, one vote, one conditional complement, one AND per bit (the C++ form compiled to seven)
    uint32_t m;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "@!p not.b32 %0, %0;\n\t}"
        : "=r"(m)
        : "r"(digit), "r"(1u << b));
    peers &= m;
  }
  return peers;
}

// RANK = 0: the lowest peer lane bumps the counter and shuffles the old value to its peers;
// RANK = 1: the lowest peer lane bumps the counter and the peers read the new value back from shared memory
// (one POPC per key instead of POPC + POPC + BREV + FLO + SHFL);
// RANK = 2: no ballots at all - every lane ORs its lane bit into a warp-private mask word of its digit
// (ATOMS.OR), reads the word back as its peer set, and the lowest peer bumps the counter and clears the word.
// LBK = descriptors fetched per look-back step.
template <int RADIX_BITS, int SORT_THREADS, int SORT_ITEMS, bool FULL, bool BALLOT, int RANK, int LBK>
__device__ __forceinline__ void onesweep_tile(const unsigned long long* __restrict__ in,
                                              unsigned long long* __restrict__ out, int cnt, uint64_t tile_first,
                                              uint32_t tile, int shift,
                                              const unsigned long long* __restrict__ bin_base,
                                              unsigned long long* __restrict__ desc, unsigned char* smem_raw) {
  constexpr int RADIX = 1 << RADIX_BITS;
  constexpr int WARPS = SORT_THREADS / 32;
  constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);               // SORT_TILE keys
  uint32_t* s_warp_hist = reinterpret_cast<uint32_t*>(smem_raw + SORT_TILE * 8);               // WARPS * RADIX
  uint32_t* s_bin_off = s_warp_hist + WARPS * RADIX;                                           // RADIX
  unsigned long long* s_delta = reinterpret_cast<unsigned long long*>(s_bin_off + RADIX);      // RADIX
  uint32_t* s_bin_cnt = reinterpret_cast<uint32_t*>(s_delta + RADIX);                          // RADIX
  uint32_t* s_warp_mask = s_bin_cnt + RADIX;                                                   // WARPS * RADIX (RANK 2)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // warp-striped load: warp w owns items [w*32*ITEMS, (w+1)*32*ITEMS), lane-strided
  unsigned long long key[SORT_ITEMS];
  uint32_t dr[SORT_ITEMS];  // digit (RADIX_BITS bits, RADIX = the invalid marker) | rank in warp << 16
  const int warp_first = warp * 32 * SORT_ITEMS;
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) {
    const int idx = warp_first + k * 32 + lane;
    key[k] = (FULL || idx < cnt) ? __ldcs(in + tile_first + idx) : ~0ull;
  }
  // rank each key among the keys of the same digit that precede it in this warp. The lowest peer lane
  // alone reads and bumps the warp-private counter and hands the old value to its peers by shuffle.
  uint32_t* my_hist = s_warp_hist + warp * RADIX;
  uint32_t* my_mask = s_warp_mask + warp * RADIX;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t gt_mask = lane == 31 ? 0u : (0xFFFFFFFEu << lane);
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) {
    const int idx = warp_first + k * 32 + lane;
    const bool valid = FULL || idx < cnt;
    const uint32_t digit = valid ? (uint32_t)((key[k] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;
    if (RANK == 2) {
      if (valid) atomicOr(my_mask + digit, 1u << lane);
      __syncwarp();
      const uint32_t peers = valid ? my_mask[digit] : 0u;
      const uint32_t base = valid ? my_hist[digit] : 0u;
      __syncwarp();
      if (valid && (peers & lt_mask) == 0u) {
        my_hist[digit] = base + (uint32_t)__popc(peers);
        my_mask[digit] = 0u;
      }
      __syncwarp();
      dr[k] = digit | ((base + (uint32_t)__popc(peers & lt_mask)) << 16);
      continue;
    }
    // a full tile has no invalid marker, so RADIX_BITS ballots tell the digits apart
    const uint32_t peers = BALLOT ? match_digit<FULL ? RADIX_BITS : RADIX_BITS + 1>(digit)
                                  : __match_any_sync(0xFFFFFFFFu, digit);
    if (RANK == 0) {
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (lane == leader && valid) {
        base = my_hist[digit];
        my_hist[digit] = base + (uint32_t)__popc(peers);
      }
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      __syncwarp();  // the next item may have another leader for the same digit
      dr[k] = digit | ((base + (uint32_t)__popc(peers & lt_mask)) << 16);
    } else {
      const uint32_t above = (uint32_t)__popc(peers & gt_mask);  // peers in higher lanes
      if ((peers & lt_mask) == 0u && valid) my_hist[digit] += above + 1u;
      __syncwarp();
      // shared-memory accesses of one warp are performed in program order: this read sees the bump of this
      // item and none of the later ones
      const uint32_t after = valid ? my_hist[digit] : 0u;
      dr[k] = digit | ((after - 1u - above) << 16);
    }
  }
  __syncthreads();
  // per digit: exclusive scan over the warps, tile total
  for (int d = tid; d < RADIX; d += SORT_THREADS) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      uint32_t c = s_warp_hist[w * RADIX + d];
      s_warp_hist[w * RADIX + d] = run;
      run += c;
    }
    s_bin_off[d] = run;
    s_bin_cnt[d] = run;
    // publish this tile's digit counts right away, so that successors can add them up while we scatter
    lb_store(desc + (size_t)tile * RADIX + d, tile == 0 ? LB_PREFIX : LB_AGGREGATE, (unsigned long long)run);
  }
  __syncthreads();
  // tile-local exclusive offsets of the digits (one warp, shuffles)
  if (warp == 0) {
    uint32_t carry = 0;
#pragma unroll
    for (int c = 0; c < RADIX / 32; c++) {
      uint32_t v = s_bin_off[c * 32 + lane];
      uint32_t inc = v;
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, dd);
        if (lane >= dd) inc += o;
      }
      s_bin_off[c * 32 + lane] = carry + inc - v;
      carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
  }
  __syncthreads();
  // fold the digit offsets into the per-warp offsets: one shared load per key in the scatter
  {
    uint4* h4 = reinterpret_cast<uint4*>(s_warp_hist);
    const uint4* o4 = reinterpret_cast<const uint4*>(s_bin_off);
    for (int i = tid; i < WARPS * RADIX / 4; i += SORT_THREADS) {
      uint4 h = h4[i];
      const uint4 o = o4[i & (RADIX / 4 - 1)];
      h.x += o.x, h.y += o.y, h.z += o.z, h.w += o.w;
      h4[i] = h;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; k++) {
    const uint32_t digit = dr[k] & 0xFFFFu;
    if (FULL || digit < RADIX) s_keys[my_hist[digit] + (dr[k] >> 16)] = key[k];
  }
  // chained scan over tiles (decoupled look-back), one descriptor per digit; LBK predecessors are fetched
  // per step so that a walk of depth D costs D / LBK memory round trips
  for (int d = tid; d < RADIX; d += SORT_THREADS) {
    unsigned long long excl = 0;
    if (tile != 0) {
      int64_t left = (int64_t)tile;  // predecessors not yet looked at
      const unsigned long long* p = desc + (size_t)(tile - 1) * RADIX + d;
      bool done = false;
      while (!done) {
        unsigned long long w[LBK];
#pragma unroll
        for (int i = 0; i < LBK; i++) w[i] = i < left ? lb_load(p - (size_t)i * RADIX) : (LB_PREFIX << 62);
#pragma unroll
        for (int i = 0; i < LBK; i++) {
          if (!done) {
            unsigned long long x = w[i];
            uint32_t spins = 0;
            while ((x >> 62) == LB_INVALID) {
              if (++spins > LB_SPIN_LIMIT) {  // watchdog, see common.cuh
                lb_raise_timeout();
                x = LB_PREFIX << 62;
                break;
              }
              __nanosleep(40);
              x = lb_load(p - (size_t)i * RADIX);
            }
            excl += x & 0x3FFFFFFFFFFFFFFFull;
            done = (x >> 62) == LB_PREFIX;
          }
        }
        p -= (size_t)LBK * RADIX;
        left -= LBK;
      }
      lb_store(desc + (size_t)tile * RADIX + d, LB_PREFIX, excl + (unsigned long long)s_bin_cnt[d]);
    }
    s_delta[d] = bin_base[d] + excl - (unsigned long long)s_bin_off[d];  // global position of s_keys[p] = delta + p
  }
  __syncthreads();
  // write out: consecutive threads write consecutive keys of a digit run
#pragma unroll 4
  for (int p = tid; p < (FULL ? SORT_TILE : cnt); p += SORT_THREADS) {
    const unsigned long long k = s_keys[p];
    const uint32_t digit = (uint32_t)((k >> shift) & (RADIX - 1));
    out[s_delta[digit] + (unsigned long long)p] = k;
  }
}

template <int RADIX_BITS, int RANK, int WARPS>
constexpr size_t onesweep_smem(int tile) {
  return (size_t)tile * 8 + (size_t)(RANK == 2 ? 2 : 1) * WARPS * (1 << RADIX_BITS) * 4 + (size_t)(1 << RADIX_BITS) * 16;
}

template <int RADIX_BITS, int SORT_THREADS, int SORT_ITEMS, int MIN_BLOCKS, bool BALLOT, int RANK, int LBK>
__global__ void __launch_bounds__(SORT_THREADS, MIN_BLOCKS) radix_onesweep_kernel(
    const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, uint64_t n, int shift,
    const unsigned long long* __restrict__ bin_base, unsigned long long* __restrict__ desc,
    uint32_t* __restrict__ ticket) {
  constexpr int RADIX = 1 << RADIX_BITS;
  constexpr int WARPS = SORT_THREADS / 32;
  constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_warp_hist = reinterpret_cast<uint32_t*>(smem_raw + SORT_TILE * 8);
  const int tid = threadIdx.x;
  for (int i = tid; i < WARPS * RADIX / 4; i += SORT_THREADS)
    reinterpret_cast<uint4*>(s_warp_hist)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (RANK == 2) {
    uint32_t* s_warp_mask = s_warp_hist + WARPS * RADIX + RADIX * 4;
    for (int i = tid; i < WARPS * RADIX / 4; i += SORT_THREADS)
      reinterpret_cast<uint4*>(s_warp_mask)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const uint32_t tile = blockIdx.x;  // blocks are dispatched in index order, see acquire_tile() in common.cuh
  (void)ticket;
  const uint64_t tile_first = (uint64_t)tile * SORT_TILE;
  const int cnt = (n - tile_first) < (uint64_t)SORT_TILE ? (int)(n - tile_first) : SORT_TILE;
  if (cnt == SORT_TILE)
    onesweep_tile<RADIX_BITS, SORT_THREADS, SORT_ITEMS, true, BALLOT, RANK, LBK>(in, out, cnt, tile_first, tile, shift, bin_base, desc, smem_raw);
  else
    onesweep_tile<RADIX_BITS, SORT_THREADS, SORT_ITEMS, false, BALLOT, RANK, LBK>(in, out, cnt, tile_first, tile, shift, bin_base, desc, smem_raw);
}

}  // namespace

// look-back watchdog flag of this translation unit: copied to *host_out on the stream (and cleared)
void sort_lb_flag_fetch(unsigned int* host_out, cudaStream_t st) {
  cudaMemcpyFromSymbolAsync(host_out, lb_timeout_flag, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st);
}
void sort_lb_flag_clear(cudaStream_t st) {
  static const unsigned int zero = 0;
  cudaMemcpyToSymbolAsync(lb_timeout_flag, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st);
}

size_t sort_temp_bytes(uint64_t n) {
  uint64_t tiles = (n + MIN_TILE - 1) / MIN_TILE;
  // histograms + per-pass descriptors + tickets
  return (size_t)MAX_PASSES * MAX_RADIX * 8 + (size_t)(tiles + 1) * MAX_RADIX * 8 + 256;
}

// digit width and pass count for the bits [begin_bit, end_bit) (CRGPU_SORT_BITS=8|9 forces a width)
static int plan_passes(int end_bit, int begin_bit, int* radix_bits) {
  const int bits = std::max(end_bit - begin_bit, 1);
  int n8 = std::min((bits + 7) / 8, MAX_PASSES), n9 = std::min((bits + 8) / 9, MAX_PASSES);
  // measured on B200 (r02): a 9-bit pass costs 1.03-1.18 ms against 0.84 ms at 168 M keys (twice the bins to
  // rank, publish and look back on; runs half as long in the scatter), so 7 x 9 loses to 8 x 8: default 8
  int rb = 8;
  (void)n9;
  if (const char* e = getenv("CRGPU_SORT_BITS")) {
    const int v = atoi(e);
    if (v == 8 || v == 9) rb = v;
  }
  if (radix_bits) *radix_bits = rb;
  return rb == 9 ? n9 : n8;
}

int sort_num_passes(int end_bit) { return plan_passes(end_bit, 0, nullptr); }

// step 1: digit histograms of every pass + their exclusive scans
int sort_histograms(const unsigned long long* keys, uint64_t n, int end_bit, void* temp, cudaStream_t st,
                    int begin_bit) {
  if (n <= 1) return 0;
  int rb = 8;
  const int n_passes = plan_passes(end_bit, begin_bit, &rb);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(temp);
  cudaMemsetAsync(hist, 0, (size_t)MAX_PASSES * MAX_RADIX * 8, st);
  int hgrid = (int)std::min<uint64_t>((n + 511) / 512, (uint64_t)sm_count() * 4);
  if (rb == 9) {
    radix_hist_kernel<9><<<hgrid, 512, 0, st>>>(keys, n, n_passes, begin_bit, hist);
    radix_scan_hist_kernel<512><<<n_passes, 512, 0, st>>>(hist);
  } else {
    radix_hist_kernel<8><<<hgrid, 512, 0, st>>>(keys, n, n_passes, begin_bit, hist);
    radix_scan_hist_kernel<256><<<n_passes, 256, 0, st>>>(hist);
  }
  return 2;
}

template <int RB, int THREADS, int ITEMS, int MIN_BLOCKS, bool BALLOT, int RANK, int LBK>
static int run_passes(unsigned long long* keys, unsigned long long* alt, uint64_t n, int n_passes, int begin_bit,
                      void* temp, unsigned long long** out, cudaStream_t st) {
  constexpr int TILE = THREADS * ITEMS;
  constexpr int RADIX = 1 << RB;
  static_assert(TILE >= MIN_TILE, "descriptor scratch is sized for tiles of at least MIN_TILE keys");
  uint64_t tiles = (n + TILE - 1) / TILE;
  uint64_t max_tiles = (n + MIN_TILE - 1) / MIN_TILE;
  unsigned char* t = static_cast<unsigned char*>(temp);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(t);
  unsigned long long* desc = reinterpret_cast<unsigned long long*>(t + (size_t)MAX_PASSES * MAX_RADIX * 8);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(t + (size_t)MAX_PASSES * MAX_RADIX * 8 + (size_t)(max_tiles + 1) * MAX_RADIX * 8);
  const size_t smem = onesweep_smem<RB, RANK, THREADS / 32>(TILE);
  auto kern = radix_onesweep_kernel<RB, THREADS, ITEMS, MIN_BLOCKS, BALLOT, RANK, LBK>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int launches = 0;
  unsigned long long* src = keys;
  unsigned long long* dst = alt;
  for (int p = 0; p < n_passes; p++) {
    cudaMemsetAsync(desc, 0, (size_t)tiles * RADIX * 8, st);
    cudaMemsetAsync(ticket, 0, 4, st);
    kern<<<(unsigned)tiles, THREADS, smem, st>>>(src, dst, n, begin_bit + p * RB, hist + (size_t)p * RADIX, desc, ticket);
    launches++;
    std::swap(src, dst);
  }
  *out = src;
  return launches;
}

// step 2: the onesweep passes; result in *out
int sort_passes(unsigned long long* keys, unsigned long long* alt, uint64_t n, int end_bit, void* temp,
                unsigned long long** out, cudaStream_t st, int begin_bit) {
  *out = keys;
  if (n <= 1) return 0;
  int rb = 8;
  const int np = plan_passes(end_bit, begin_bit, &rb);
  const int cfg = getenv("CRGPU_SORT_CFG") ? atoi(getenv("CRGPU_SORT_CFG")) : 0;
  if (rb == 9) {
    switch (cfg) {  // 9-bit digits
      case 1: return run_passes<9, 512, 16, 2, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
      case 2: return run_passes<9, 512, 12, 2, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
      case 3: return run_passes<9, 384, 16, 3, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
      case 7: return run_passes<9, 256, 16, 4, true, 2, 4>(keys, alt, n, np, begin_bit, temp, out, st);
      case 8: return run_passes<9, 512, 16, 2, true, 2, 4>(keys, alt, n, np, begin_bit, temp, out, st);
      default: return run_passes<9, 256, 16, 4, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
    }
  }
  switch (cfg) {  // CRGPU_SORT_CFG: variants kept for profiling; 0 = the measured best on B200
    case 1: return run_passes<8, 256, 16, 4, false, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);  // match.any, not ballots
    case 2: return run_passes<8, 512, 12, 2, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
    case 3: return run_passes<8, 384, 16, 3, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
    case 4: return run_passes<8, 256, 16, 4, true, 1, 4>(keys, alt, n, np, begin_bit, temp, out, st);  // peers re-read the counter
    case 5: return run_passes<8, 256, 16, 4, true, 0, 1>(keys, alt, n, np, begin_bit, temp, out, st);  // one descriptor per step
    case 6: return run_passes<8, 256, 16, 4, true, 0, 8>(keys, alt, n, np, begin_bit, temp, out, st);
    case 7: return run_passes<8, 256, 16, 4, true, 2, 4>(keys, alt, n, np, begin_bit, temp, out, st);  // ATOMS.OR peer masks
    default: return run_passes<8, 256, 16, 4, true, 0, 4>(keys, alt, n, np, begin_bit, temp, out, st);
  }
}

int sort_keys(unsigned long long* keys, unsigned long long* alt, uint64_t n, int end_bit, void* temp, size_t temp_bytes,
              unsigned long long** out, cudaStream_t st, int begin_bit) {
  (void)temp_bytes;
  int launches = sort_histograms(keys, n, end_bit, temp, st, begin_bit);
  launches += sort_passes(keys, alt, n, end_bit, temp, out, st, begin_bit);
  return launches;
}
