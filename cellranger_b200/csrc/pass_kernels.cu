// pass_kernels.cu — pass 1 (pack + exact whitelist match + priors) and pass 2 (Hamming-1 posterior
// correction) of the barcode path, plus the feature-barcode kernels. sm_100a.
//
// Reference semantics restated here (never copied): lib/rust/cr_types/src/rna_read.rs:285-368
// (barcode / UMI extraction), lib/rust/barcode/src/whitelist.rs:494-516 (check_and_update),
// lib/rust/cr_lib/src/make_shard_metrics.rs:171-188 (priors), lib/rust/umi/src/info.rs:20-74 (UMI
// validity), lib/rust/barcode/src/corrector.rs:111-171 (posterior),
// lib/rust/cr_types/src/reference/feature_extraction.rs:34-117,447-471 (feature barcodes).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

__constant__ double c_bc_prob[256];  // probability(q) = pow(10, -(q-33)/10), filled by host libm
__constant__ double c_fb_prob[64];   // pow(10, -qv/10), qv = 0..33

void upload_prob_luts(const double* bc_lut256, const double* fb_lut64, cudaStream_t st) {
  cudaMemcpyToSymbolAsync(c_bc_prob, bc_lut256, 256 * sizeof(double), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_fb_prob, fb_lut64, 64 * sizeof(double), 0, cudaMemcpyHostToDevice, st);
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA, 1-D) wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------
// per-read logic shared by the staged and the generic kernel
// ---------------------------------------------------------------------------
struct ReadResult {
  uint32_t bc_word;   // state | rank
  uint32_t umi_word;  // flags | packed
  bool emit_key;
  unsigned long long key;
  bool invalid;
};

__device__ __forceinline__ unsigned long long make_key(const KeyLayout& kl, uint32_t rank, uint32_t feature,
                                                       uint32_t lib, uint32_t umi) {
  return ((unsigned long long)rank << kl.rank_shift) | ((unsigned long long)feature << kl.feature_shift) |
         ((unsigned long long)lib << kl.lib_shift) | (unsigned long long)umi;
}


// a feature index the matrix has no row for (and that is not the "unmapped" marker): counted as an error of
// the batch and treated as unmapped, so that it can never spill into the barcode-rank bits of a key
__device__ __forceinline__ uint32_t checked_feature(uint32_t f, uint32_t n_features, unsigned long long* bad) {
  if (f != NO_FEATURE && f >= n_features) {
    if (bad) atomicAdd(bad, 1ull);
    return NO_FEATURE;
  }
  return f;
}

__device__ __forceinline__ void classify_read(const Pass1Args& a, uint32_t bc, uint32_t nmask, uint32_t umi,
                                              bool umi_has_n, bool umi_lowq, uint32_t feature, ReadResult* r) {
  // UmiInfo::new: valid = !(has_n || is_homopolymer || low_min_qual)
  uint32_t rep = 0x55555555u & mask_bits(2 * a.umi_len);
  bool homopolymer = umi == (umi & 3u) * rep;
  bool umi_valid = !(umi_has_n || homopolymer || umi_lowq) && a.umi_len > 0;
  r->umi_word = (umi & UMI_SEQ_MASK) | (umi_valid ? UMI_VALID_BIT : 0u) | (umi_has_n ? UMI_HASN_BIT : 0u);
  int idx = nmask ? -1 : wl_find(a.wl, bc);
  r->emit_key = false;
  r->key = 0ull;
  if (idx >= 0) {
    uint32_t rank = wl_rank_of(a.wl, idx);
    if (a.prior) atomicAdd(a.prior + rank, 1u);
    r->bc_word = (ST_VALID_BEFORE << BC_STATE_SHIFT) | rank;
    r->invalid = false;
    if (a.emit_keys && umi_valid && feature != NO_FEATURE) {
      r->emit_key = true;
      r->key = make_key(a.kl, rank, feature, a.lib, umi);
    }
  } else {
    r->bc_word = (ST_INVALID << BC_STATE_SHIFT) | BC_RANK_MASK;
    r->invalid = true;
  }
}

// ---------------------------------------------------------------------------
// Pass 1, staged: persistent CTAs, tiles of THREADS*RPT reads brought into shared memory by 1-D bulk
// copies (TMA) behind a STAGES-deep mbarrier ring, so HBM sees only full-line coalesced requests.
// Layout assumed: barcode = R1[0:16], UMI = R1[16:16+UMI_LEN], R1_LEN >= 16+UMI_LEN.
// ---------------------------------------------------------------------------
template <int R1_LEN>
__device__ __forceinline__ void load_record(const uint8_t* base, int j, uint32_t (&w)[(R1_LEN + 3) / 4]) {
  constexpr int NW = (R1_LEN + 3) / 4;
  if constexpr (R1_LEN % 4 == 0) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (size_t)j * R1_LEN);
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = p[i];
  } else {
    uint32_t o = (uint32_t)j * R1_LEN;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (o & ~3u));
    uint32_t sh = (o & 3u) * 8u;
    uint32_t x[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) x[i] = p[i];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(x[i], x[i + 1], sh);
  }
}

// result of packing one R1 record
struct Packed {
  uint32_t bc, nmask, umi;
  bool umi_has_n, umi_lowq;
  uint4 bcq;
};

// stage 1 of the packing: the barcode (so that its whitelist lookup can start), stage 2: UMI and qualities
template <int R1_LEN>
__device__ __forceinline__ void pack_barcode(const uint32_t (&ws)[(R1_LEN + 3) / 4], Packed* o) {
  uint32_t bad0, bad1, bad2, bad3;
  o->bc = (pack4(ws[0], &bad0) << 24) | (pack4(ws[1], &bad1) << 16) | (pack4(ws[2], &bad2) << 8) | pack4(ws[3], &bad3);
  uint32_t nmask = 0;
  if (bad0 | bad1 | bad2 | bad3) {  // bit `pos` for every non-ACGT base (rare path)
    uint32_t bb[4] = {bad0, bad1, bad2, bad3};
#pragma unroll
    for (int wd = 0; wd < 4; wd++)
#pragma unroll
      for (int by = 0; by < 4; by++)
        if (bb[wd] & (0x80u << (8 * by))) nmask |= 1u << (wd * 4 + by);
  }
  o->nmask = nmask;
}
template <int R1_LEN, int UMI_LEN>
__device__ __forceinline__ void pack_umi(const uint32_t (&ws)[(R1_LEN + 3) / 4], const uint32_t (&wq)[(R1_LEN + 3) / 4],
                                         Packed* o) {
  uint32_t umi = 0, ubad = 0, ulow = 0;
  constexpr int UW = (UMI_LEN + 3) / 4;
#pragma unroll
  for (int u = 0; u < UW; u++) {
    uint32_t w = ws[4 + u], q = wq[4 + u];
    constexpr int full = UMI_LEN / 4;
    uint32_t bmask = 0xFFFFFFFFu;
    if (u >= full) {  // partial last word: keep UMI_LEN % 4 bytes
      bmask = (1u << (8 * (UMI_LEN % 4))) - 1u;
      w = (w & bmask) | (0x41414141u & ~bmask);
      q = (q & bmask) | (0x49494949u & ~bmask);
    }
    uint32_t bad;
    uint32_t p = pack4(w, &bad);
    ubad |= bad & bmask;
    ulow |= lowqual4(q) & bmask;
    umi = (umi << 8) | p;
  }
  if constexpr (UMI_LEN % 4 != 0) umi >>= 2 * (4 - UMI_LEN % 4);
  o->umi = umi;
  o->umi_has_n = ubad != 0;
  o->umi_lowq = ulow != 0;
  o->bcq = make_uint4(wq[0], wq[1], wq[2], wq[3]);
}
template <int R1_LEN, int UMI_LEN>
__device__ __forceinline__ void pack_record(const uint8_t* s_seq, const uint8_t* s_qual, int j, Packed* o) {
  constexpr int NW = (R1_LEN + 3) / 4;
  uint32_t ws[NW], wq[NW];
  load_record<R1_LEN>(s_seq, j, ws);
  load_record<R1_LEN>(s_qual, j, wq);
  pack_barcode<R1_LEN>(ws, o);
  pack_umi<R1_LEN, UMI_LEN>(ws, wq, o);
}

__device__ __forceinline__ unsigned long long make_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

template <int R1_LEN, int UMI_LEN, int THREADS, int RPT, int STAGES>
__global__ void __launch_bounds__(THREADS) pass1_staged_kernel(const Pass1Args a) {
  constexpr int TILE = THREADS * RPT;
  constexpr int SEQ_BYTES = TILE * R1_LEN;
  constexpr int REC_PAD = 16;  // the unaligned record loader reads one word past a record
  constexpr int STAGE_BYTES = 2 * (SEQ_BYTES + REC_PAD) + TILE * 4;
  constexpr int NWARPS = THREADS / 32;
  static_assert(SEQ_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[STAGES];       // "full": bulk copies of a stage have landed
  __shared__ __align__(8) uint64_t empty_bar[STAGES];  // "empty": every warp has read the stage
  __shared__ uint32_t warp_tot[2][NWARPS];
  __shared__ unsigned long long base_bcast[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t n_tiles = (a.n + TILE - 1) / TILE;
  const bool have_feat = a.feature != nullptr;
  const unsigned long long policy = make_evict_first_policy();  // the reads stream through L2 once

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&mbar[s], 1);
      mbar_init(&empty_bar[s], NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto stage_seq = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };
  auto stage_qual = [&](int s) { return smem + (size_t)s * STAGE_BYTES + SEQ_BYTES + REC_PAD; };
  auto stage_feat = [&](int s) {
    return reinterpret_cast<uint32_t*>(smem + (size_t)s * STAGE_BYTES + 2 * (SEQ_BYTES + REC_PAD));
  };
  // thread 0: start the copies of a FULL tile into stage s
  auto issue = [&](uint64_t tile, int s) {
    uint64_t first = tile * TILE;
    if (first + TILE <= a.n) {
      uint32_t bytes = 2 * SEQ_BYTES + (have_feat ? TILE * 4 : 0);
      mbar_expect_tx(&mbar[s], bytes);
      bulk_g2s_hint(stage_seq(s), a.seq + first * R1_LEN, SEQ_BYTES, &mbar[s], policy);
      bulk_g2s_hint(stage_qual(s), a.qual + first * R1_LEN, SEQ_BYTES, &mbar[s], policy);
      if (have_feat) bulk_g2s_hint(stage_feat(s), a.feature + first, TILE * 4, &mbar[s], policy);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      uint64_t tile = blockIdx.x + (uint64_t)s * gridDim.x;
      if (tile < n_tiles) issue(tile, s);
    }
  }

  // Output placement is deferred by one tile: the keys / invalid entries of tile i wait in registers until the
  // block scan of tile i+1 has passed its barrier, by which time the global atomic that reserved their space
  // has long returned - its round trip is off the critical path and one of the two block barriers is gone.
  bool have_p = false;
  unsigned long long p_key[RPT];
  uint32_t p_bc[RPT], p_nmask[RPT], p_flags = 0, p_excl = 0;  // p_flags: bit k = key, bit 8+k = invalid entry
  uint4 p_bcq[RPT];
  uint64_t p_first = 0;
  int p_par = 0;
  unsigned long long pend_base = 0ull;  // thread 0: the atomic's result, stored to shared memory one tile later
  bool pend_store = false;
#pragma unroll
  for (int k = 0; k < RPT; k++) {
    p_key[k] = 0ull;
    p_bc[k] = p_nmask[k] = 0u;
    p_bcq[k] = make_uint4(0u, 0u, 0u, 0u);
  }
  auto flush_pending = [&]() {
    const unsigned long long base = base_bcast[p_par];
    uint64_t kpos = (base & 0xFFFFFFFFull) + (p_excl & 0xFFFFu);
    uint64_t ipos = (base >> 32) + (p_excl >> 16);
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      if ((p_flags >> k) & 1u) __stcs(a.keys + kpos++, p_key[k]);
      if ((p_flags >> (8 + k)) & 1u) {
        a.inv_idx[ipos] = (uint32_t)(a.idx_base + p_first + tid + k * THREADS);
        a.inv_bc[ipos] = p_bc[k];
        a.inv_nmask[ipos] = p_nmask[k];
        a.inv_qual[ipos] = p_bcq[k];
        ipos++;
      }
    }
  };

  uint32_t it = 0;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int s = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1u;
    const uint64_t first = tile * TILE;
    const int cnt = (int)((a.n - first) < (uint64_t)TILE ? (a.n - first) : (uint64_t)TILE);
    if (cnt == TILE) {
      mbar_wait(&mbar[s], parity);
    } else {
      // the partial last tile: plain cooperative loads
      const uint8_t* gs = a.seq + first * R1_LEN;
      const uint8_t* gq = a.qual + first * R1_LEN;
      uint8_t* ss = stage_seq(s);
      uint8_t* sq = stage_qual(s);
      for (int b = tid; b < cnt * R1_LEN; b += THREADS) {
        ss[b] = gs[b];
        sq[b] = gq[b];
      }
      if (have_feat)
        for (int b = tid; b < cnt; b += THREADS) stage_feat(s)[b] = a.feature[first + b];
      __syncthreads();
    }

    // ---- phase 1: shared memory -> registers, 2-bit packing, UMI checks. The whitelist slot of each read is
    // requested as soon as its barcode is packed, so the rest of the packing runs under the load ----
    Packed pk[RPT];
    uint32_t feat[RPT];
    bool live[RPT];
    WlProbe pr[RPT];
    {
      constexpr int NW = (R1_LEN + 3) / 4;
      uint32_t ws[RPT][NW];
#pragma unroll
      for (int k = 0; k < RPT; k++) {
        const int j = tid + k * THREADS;
        live[k] = j < cnt;
        load_record<R1_LEN>(stage_seq(s), live[k] ? j : 0, ws[k]);
        pack_barcode<R1_LEN>(ws[k], &pk[k]);
        pr[k] = wl_find_probe(a.wl, wl_find_begin(a.wl, pk[k].bc));
      }
#pragma unroll
      for (int k = 0; k < RPT; k++) {
        const int j = tid + k * THREADS;
        uint32_t wq[NW];
        load_record<R1_LEN>(stage_qual(s), live[k] ? j : 0, wq);
        pack_umi<R1_LEN, UMI_LEN>(ws[k], wq, &pk[k]);
        feat[k] = (have_feat && live[k]) ? checked_feature(stage_feat(s)[j], a.n_features, a.bad_feature) : NO_FEATURE;
        if (!live[k]) {
          pk[k].bc = 0;
          pk[k].nmask = 1;
          pk[k].umi = 0;
          pk[k].umi_has_n = true;
          pk[k].umi_lowq = false;
          pk[k].bcq = make_uint4(0, 0, 0, 0);
        }
      }
    }
    // Stage s is free once every warp has read its records: each warp arrives on the stage's "empty"
    // mbarrier, thread 0 waits for that phase and only then lets the bulk copies (async proxy) overwrite
    // the stage. A plain bar.sync is NOT enough here: it does not order the other warps' generic-proxy
    // reads before an async-proxy write issued right behind it (seen on B200 as stale feature words).
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (tid == 0) {
      uint64_t next = tile + (uint64_t)STAGES * gridDim.x;
      if (next < n_tiles) {
        mbar_wait(&empty_bar[s], parity);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(next, s);
      }
    }

    // ---- phase 2: exact whitelist lookups, all loads of the RPT reads in flight together ----
    uint32_t bcw[RPT], umw[RPT];
    unsigned long long key[RPT];
    bool emit[RPT], inval[RPT];
    uint32_t n_key = 0, n_inv = 0;
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      int idx = pk[k].nmask ? -1 : wl_find_end(a.wl, pr[k], pk[k].bc);
      const uint32_t rep = 0x55555555u & mask_bits(2 * UMI_LEN);
      bool homopolymer = pk[k].umi == (pk[k].umi & 3u) * rep;
      bool umi_valid = !(pk[k].umi_has_n || homopolymer || pk[k].umi_lowq);
      umw[k] = (pk[k].umi & UMI_SEQ_MASK) | (umi_valid ? UMI_VALID_BIT : 0u) | (pk[k].umi_has_n ? UMI_HASN_BIT : 0u);
      emit[k] = false;
      inval[k] = false;
      key[k] = 0ull;
      if (idx >= 0) {
        uint32_t rank = wl_rank_of(a.wl, idx);
        if (a.prior) atomicAdd(a.prior + rank, 1u);
        bcw[k] = (ST_VALID_BEFORE << BC_STATE_SHIFT) | rank;
        if (a.emit_keys && umi_valid && feat[k] != NO_FEATURE) {
          emit[k] = true;
          key[k] = make_key(a.kl, rank, feat[k], a.lib, pk[k].umi);
        }
      } else {
        bcw[k] = (ST_INVALID << BC_STATE_SHIFT) | BC_RANK_MASK;
        inval[k] = live[k];
      }
      n_key += emit[k];
      n_inv += inval[k];
    }

    // the per-read words need no placement
#pragma unroll
    for (int k = 0; k < RPT; k++)
      if (live[k]) {
        const uint64_t gi = first + tid + k * THREADS;
        __stcs(a.bc_out + gi, bcw[k]);
        __stcs(a.umi_out + gi, umw[k]);
      }
    // ---- phase 3: one block scan places the tile's keys and invalid entries (counts packed 16+16) ----
    const uint32_t v = n_key | (n_inv << 16);
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    const int par = it & 1;
    if (lane == 31) warp_tot[par][warp] = inc;
    if (tid == 0 && pend_store) {  // the previous tile's base: its atomic was issued a whole tile ago
      base_bcast[par ^ 1] = pend_base;
      pend_store = false;
    }
    __syncthreads();
    uint32_t wsum = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
      uint32_t c = warp_tot[par][w];
      if (w < warp) wsum += c;
      tot += c;
    }
    if (tid == 0) {
      pend_base = tot ? atomicAdd(a.counters, (unsigned long long)(tot & 0xFFFFu) | ((unsigned long long)(tot >> 16) << 32))
                      : 0ull;
      pend_store = true;
    }
    if (have_p) flush_pending();
    // this tile becomes the pending one
    p_flags = 0u;
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      p_key[k] = key[k];
      p_bc[k] = pk[k].bc;
      p_nmask[k] = pk[k].nmask;
      p_bcq[k] = pk[k].bcq;
      p_flags |= (emit[k] ? 1u : 0u) << k;
      p_flags |= (inval[k] ? 1u : 0u) << (8 + k);
    }
    p_excl = wsum + inc - v;
    p_first = first;
    p_par = par;
    have_p = true;
  }
  if (tid == 0 && pend_store) base_bcast[p_par] = pend_base;
  __syncthreads();
  if (have_p) flush_pending();
}

// counters layout for the staged kernel: one packed 64-bit word (keys in the low half). The host keeps
// the key counter and the per-batch invalid counter apart, so the kernel works on a scratch word that is
// split afterwards.
__global__ void split_counter_kernel(unsigned long long* packed, unsigned long long* keys_total,
                                     unsigned long long* inv_total) {
  unsigned long long v = *packed;
  *keys_total = v & 0xFFFFFFFFull;
  *inv_total = v >> 32;
}
__global__ void merge_counter_kernel(unsigned long long* packed, const unsigned long long* keys_total) {
  *packed = *keys_total;  // invalid count of a new batch starts at 0
}

// ---------------------------------------------------------------------------
// Pass 1, generic layout (any offsets / lengths <= 16): one thread per read, direct loads.
// Used for unusual chemistries and for the corrector batch seam.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pass1_generic_kernel(const Pass1Args a) {
  __shared__ uint32_t scan_a[9], scan_b[9];
  __shared__ unsigned long long base_bcast;
  const uint64_t n_blocks_work = (a.n + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < n_blocks_work; blk += gridDim.x) {
    uint64_t gi = blk * 256 + threadIdx.x;
    ReadResult res;
    res.emit_key = false;
    res.invalid = false;
    uint32_t bc = 0, nmask = 0;
    uint8_t q16[16];
#pragma unroll
    for (int i = 0; i < 16; i++) q16[i] = 'I';
    if (gi < a.n) {
      const uint8_t* s = a.seq + gi * a.r1_len;
      const uint8_t* q = a.have_qual ? a.qual + gi * a.r1_len : nullptr;
      for (int i = 0; i < a.bc_len; i++) {
        uint8_t c = s[a.bc_off + i];
        uint32_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
        if (code == 4u) {
          nmask |= 1u << i;
          code = 0u;
        }
        bc = (bc << 2) | code;
        if (q) q16[i] = q[a.bc_off + i];
      }
      uint32_t umi = 0;
      bool uhasn = false, ulow = false;
      for (int i = 0; i < a.umi_len; i++) {
        uint8_t c = s[a.umi_off + i];
        uint32_t code = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
        if (code == 4u) {
          uhasn = true;
          code = 0u;
        }
        umi = (umi << 2) | code;
        if (q) ulow |= (uint8_t)(q[a.umi_off + i] - 33) < 10;
      }
      uint32_t feature = a.feature ? checked_feature(a.feature[gi], a.n_features, a.bad_feature) : NO_FEATURE;
      classify_read(a, bc, nmask, umi, uhasn, ulow, feature, &res);
    }
    uint32_t tot_key, tot_inv;
    uint32_t off_key = block_exclusive_scan<256>(res.emit_key ? 1u : 0u, &tot_key, scan_a);
    uint32_t off_inv = block_exclusive_scan<256>(res.invalid ? 1u : 0u, &tot_inv, scan_b);
    if (threadIdx.x == 0)
      base_bcast = (tot_key | tot_inv)
                       ? atomicAdd(a.counters, (unsigned long long)tot_key | ((unsigned long long)tot_inv << 32))
                       : 0ull;
    __syncthreads();
    const unsigned long long base = base_bcast;
    if (gi < a.n) {
      a.bc_out[gi] = res.bc_word;
      a.umi_out[gi] = res.umi_word;
      if (res.emit_key) a.keys[(base & 0xFFFFFFFFull) + off_key] = res.key;
      if (res.invalid) {
        uint64_t ipos = (base >> 32) + off_inv;
        a.inv_idx[ipos] = (uint32_t)(a.idx_base + gi);
        a.inv_bc[ipos] = bc;
        a.inv_nmask[ipos] = nmask;
        uint4 qq;
        qq.x = q16[0] | (q16[1] << 8) | (q16[2] << 16) | ((uint32_t)q16[3] << 24);
        qq.y = q16[4] | (q16[5] << 8) | (q16[6] << 16) | ((uint32_t)q16[7] << 24);
        qq.z = q16[8] | (q16[9] << 8) | (q16[10] << 16) | ((uint32_t)q16[11] << 24);
        qq.w = q16[12] | (q16[13] << 8) | (q16[14] << 16) | ((uint32_t)q16[15] << 24);
        a.inv_qual[ipos] = qq;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Pass 1 split in two (standard layouts): pack_kernel streams the reads once at HBM speed (no table
// look-ups, no atomics), match_kernel works on the 8-byte packed records with many look-ups in flight.
// ---------------------------------------------------------------------------
#define UMI_BCN_BIT 0x20000000u  // umi_out word, between the two kernels: the barcode holds a non-ACGT base

template <int R1_LEN, int UMI_LEN, int THREADS, int RPT, int STAGES>
__global__ void __launch_bounds__(THREADS) pack_kernel(const Pass1Args a) {
  constexpr int TILE = THREADS * RPT;
  constexpr int SEQ_BYTES = TILE * R1_LEN;
  constexpr int REC_PAD = 16;
  constexpr int STAGE_BYTES = 2 * (SEQ_BYTES + REC_PAD);
  constexpr int NWARPS = THREADS / 32;
  static_assert(SEQ_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t n_tiles = (a.n + TILE - 1) / TILE;
  const unsigned long long policy = make_evict_first_policy();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&mbar[s], 1);
      mbar_init(&empty_bar[s], NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto stage_seq = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };
  auto stage_qual = [&](int s) { return smem + (size_t)s * STAGE_BYTES + SEQ_BYTES + REC_PAD; };
  auto issue = [&](uint64_t tile, int s) {
    uint64_t first = tile * TILE;
    if (first + TILE <= a.n) {
      mbar_expect_tx(&mbar[s], 2 * SEQ_BYTES);
      bulk_g2s_hint(stage_seq(s), a.seq + first * R1_LEN, SEQ_BYTES, &mbar[s], policy);
      bulk_g2s_hint(stage_qual(s), a.qual + first * R1_LEN, SEQ_BYTES, &mbar[s], policy);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      uint64_t tile = blockIdx.x + (uint64_t)s * gridDim.x;
      if (tile < n_tiles) issue(tile, s);
    }
  }
  uint32_t it = 0;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int s = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1u;
    const uint64_t first = tile * TILE;
    const int cnt = (int)((a.n - first) < (uint64_t)TILE ? (a.n - first) : (uint64_t)TILE);
    if (cnt == TILE) {
      mbar_wait(&mbar[s], parity);
    } else {
      const uint8_t* gs = a.seq + first * R1_LEN;
      const uint8_t* gq = a.qual + first * R1_LEN;
      for (int b = tid; b < cnt * R1_LEN; b += THREADS) {
        stage_seq(s)[b] = gs[b];
        stage_qual(s)[b] = gq[b];
      }
      __syncthreads();
    }
    uint32_t bcw[RPT], umw[RPT];
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      const int j = tid + k * THREADS;
      if (j < cnt) {
        Packed pk;
        pack_record<R1_LEN, UMI_LEN>(stage_seq(s), stage_qual(s), j, &pk);
        const uint32_t rep = 0x55555555u & mask_bits(2 * UMI_LEN);
        bool homopolymer = pk.umi == (pk.umi & 3u) * rep;
        bool umi_valid = !(pk.umi_has_n || homopolymer || pk.umi_lowq);
        bcw[k] = pk.bc;
        umw[k] = (pk.umi & 0x1FFFFFFFu) | (umi_valid ? UMI_VALID_BIT : 0u) | (pk.umi_has_n ? UMI_HASN_BIT : 0u) |
                 (pk.nmask ? UMI_BCN_BIT : 0u);
      }
    }
    // the stage may be refilled once every warp has read it (see pass1_staged_kernel for why not bar.sync)
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
    if (tid == 0) {
      uint64_t next = tile + (uint64_t)STAGES * gridDim.x;
      if (next < n_tiles) {
        mbar_wait(&empty_bar[s], parity);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(next, s);
      }
    }
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      const int j = tid + k * THREADS;
      if (j < cnt) {
        a.bc_out[first + j] = bcw[k];
        a.umi_out[first + j] = umw[k];
      }
    }
  }
}

// 16 bytes at an arbitrary address (needs 4 readable bytes past the end at most)
__device__ __forceinline__ uint4 load_bytes16(const uint8_t* p) {
  const uintptr_t addr = (uintptr_t)p;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(addr & 3) * 8u;
  uint32_t x0 = __ldg(q), x1 = __ldg(q + 1), x2 = __ldg(q + 2), x3 = __ldg(q + 3);
  uint32_t x4 = sh ? __ldg(q + 4) : 0u;
  return make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh),
                    __funnelshift_r(x3, x4, sh));
}

// the same for the last records of a caller-owned buffer: `avail` readable bytes from p on (the plain loader
// reads up to 20); bytes past the end read as 0
__device__ __forceinline__ uint4 load_bytes16_bounded(const uint8_t* p, uint64_t avail) {
  if (avail >= 20) return load_bytes16(p);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  for (int i = 0; i < 16 && (uint64_t)i < avail; i++) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int THREADS, int RPT>
__global__ void __launch_bounds__(THREADS) match_kernel(const Pass1Args a) {
  constexpr int TILE = THREADS * RPT;
  constexpr int NWARPS = THREADS / 32;
  __shared__ uint32_t warp_tot[2][NWARPS];
  __shared__ unsigned long long base_bcast[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t n_tiles = (a.n + TILE - 1) / TILE;
  const bool have_feat = a.feature != nullptr;
  const uint32_t umask = mask_bits(a.kl.umi_bits);
  uint32_t it = 0;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const uint64_t first = tile * TILE;
    uint32_t bc[RPT], uw[RPT], feat[RPT];
    bool live[RPT];
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      const uint64_t gi = first + tid + k * THREADS;
      live[k] = gi < a.n;
      bc[k] = live[k] ? a.bc_out[gi] : 0u;
      uw[k] = live[k] ? a.umi_out[gi] : UMI_BCN_BIT;
      feat[k] = (live[k] && have_feat) ? checked_feature(__ldcs(a.feature + gi), a.n_features, a.bad_feature) : NO_FEATURE;
    }
    uint32_t st0[RPT];
#pragma unroll
    for (int k = 0; k < RPT; k++) st0[k] = wl_find_begin(a.wl, bc[k]);
    WlProbe pr[RPT];
#pragma unroll
    for (int k = 0; k < RPT; k++) pr[k] = wl_find_probe(a.wl, st0[k]);
    uint32_t bcw[RPT];
    unsigned long long key[RPT];
    bool emit[RPT], inval[RPT];
    uint32_t n_key = 0, n_inv = 0;
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      const bool bcn = (uw[k] & UMI_BCN_BIT) != 0;
      int idx = bcn ? -1 : wl_find_end(a.wl, pr[k], bc[k]);
      emit[k] = false;
      inval[k] = false;
      key[k] = 0ull;
      if (idx >= 0) {
        uint32_t rank = wl_rank_of(a.wl, idx);
        if (a.prior && !(a.debug_flags & 1)) atomicAdd(a.prior + rank, 1u);
        bcw[k] = (ST_VALID_BEFORE << BC_STATE_SHIFT) | rank;
        if (a.emit_keys && (uw[k] & UMI_VALID_BIT) && feat[k] != NO_FEATURE) {
          emit[k] = true;
          key[k] = make_key(a.kl, rank, feat[k], a.lib, uw[k] & umask);
        }
      } else {
        bcw[k] = (ST_INVALID << BC_STATE_SHIFT) | BC_RANK_MASK;
        inval[k] = live[k];
      }
      n_key += emit[k];
      n_inv += inval[k];
    }
    const uint32_t v = n_key | (n_inv << 16);
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    const int par = it & 1;
    if (lane == 31) warp_tot[par][warp] = inc;
    __syncthreads();
    uint32_t wsum = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NWARPS; w++) {
      uint32_t c = warp_tot[par][w];
      if (w < warp) wsum += c;
      tot += c;
    }
    const uint32_t excl = wsum + inc - v;
    if (tid == 0)
      base_bcast[par] = tot ? atomicAdd(a.counters, (unsigned long long)(tot & 0xFFFFu) |
                                                        ((unsigned long long)(tot >> 16) << 32))
                            : 0ull;
    __syncthreads();
    const unsigned long long base = base_bcast[par];
    uint64_t kpos = (base & 0xFFFFFFFFull) + (excl & 0xFFFFu);
    uint64_t ipos = (base >> 32) + (excl >> 16);
#pragma unroll
    for (int k = 0; k < RPT; k++) {
      if (live[k]) {
        const uint64_t gi = first + tid + k * THREADS;
        a.bc_out[gi] = bcw[k];
        if (uw[k] & UMI_BCN_BIT) a.umi_out[gi] = uw[k] & ~UMI_BCN_BIT;
        if (emit[k] && !(a.debug_flags & 2)) __stcs(a.keys + kpos++, key[k]);
        if (inval[k]) {
          uint32_t nmask = 0;
          if (uw[k] & UMI_BCN_BIT) {  // rare: recompute which bases are not A,C,G,T
            uint4 sq = load_bytes16_bounded(a.seq + gi * a.r1_len + a.bc_off, (a.n - gi) * a.r1_len - a.bc_off);
            const uint32_t w4[4] = {sq.x, sq.y, sq.z, sq.w};
#pragma unroll
            for (int wd = 0; wd < 4; wd++) {
              uint32_t bad;
              pack4(w4[wd], &bad);
#pragma unroll
              for (int by = 0; by < 4; by++)
                if (bad & (0x80u << (8 * by))) nmask |= 1u << (wd * 4 + by);
            }
          }
          a.inv_idx[ipos] = (uint32_t)(a.idx_base + gi);
          a.inv_bc[ipos] = bc[k];
          a.inv_nmask[ipos] = nmask;
          a.inv_qual[ipos] = load_bytes16_bounded(a.qual + gi * a.r1_len + a.bc_off, (a.n - gi) * a.r1_len - a.bc_off);
          ipos++;
        }
      }
    }
  }
}

template <int R1_LEN, int UMI_LEN>
static int launch_split(const Pass1Args& a_in, int n_sms, cudaStream_t st) {
  Pass1Args a = a_in;
  a.debug_flags = getenv("CRGPU_P1_DBG") ? atoi(getenv("CRGPU_P1_DBG")) : 0;
  const bool only_pack = a.debug_flags & 4, only_match = a.debug_flags & 8;
  if (!only_match) {
    constexpr int THREADS = 256, RPT = 2, STAGES = 3, TILE = THREADS * RPT;
    auto kern = pack_kernel<R1_LEN, UMI_LEN, THREADS, RPT, STAGES>;
    size_t smem = (size_t)STAGES * 2 * (TILE * R1_LEN + 16);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    uint64_t tiles = (a.n + TILE - 1) / TILE;
    int grid = (int)std::min<uint64_t>(tiles, (uint64_t)n_sms * (per_sm > 0 ? per_sm : 1));
    kern<<<grid, THREADS, smem, st>>>(a);
  }
  if (!only_pack) {
    constexpr int THREADS = 256, RPT = 4, TILE = THREADS * RPT;
    auto kern = match_kernel<THREADS, RPT>;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, 0);
    uint64_t tiles = (a.n + TILE - 1) / TILE;
    int grid = (int)std::min<uint64_t>(tiles, (uint64_t)n_sms * (per_sm > 0 ? per_sm : 1));
    kern<<<grid, THREADS, 0, st>>>(a);
  }
  return 2;
}

template <int R1_LEN, int UMI_LEN, int THREADS, int RPT, int STAGES>
static int launch_staged(const Pass1Args& a, int n_sms, cudaStream_t st) {
  constexpr int TILE = THREADS * RPT;
  auto kern = pass1_staged_kernel<R1_LEN, UMI_LEN, THREADS, RPT, STAGES>;
  size_t smem = (size_t)STAGES * (2 * (TILE * R1_LEN + 16) + TILE * 4);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  uint64_t tiles = (a.n + TILE - 1) / TILE;
  int grid = (int)std::min<uint64_t>(tiles, (uint64_t)n_sms * (per_sm > 0 ? per_sm : 1));
  kern<<<grid, THREADS, smem, st>>>(a);
  return 1;
}

template <int R1_LEN, int UMI_LEN>
static int launch_staged_cfg(const Pass1Args& a, int n_sms, cudaStream_t st) {
  const int cfg = getenv("CRGPU_P1_CFG") ? atoi(getenv("CRGPU_P1_CFG")) : 0;
  switch (cfg) {
    case 1: return launch_staged<R1_LEN, UMI_LEN, 128, 2, 2>(a, n_sms, st);
    case 2: return launch_staged<R1_LEN, UMI_LEN, 128, 4, 2>(a, n_sms, st);
    case 3: return launch_staged<R1_LEN, UMI_LEN, 256, 1, 2>(a, n_sms, st);
    case 4: return launch_staged<R1_LEN, UMI_LEN, 128, 2, 3>(a, n_sms, st);
    case 5: return launch_staged<R1_LEN, UMI_LEN, 256, 2, 3>(a, n_sms, st);
    default: return launch_staged<R1_LEN, UMI_LEN, 256, 2, 2>(a, n_sms, st);
  }
}

int launch_pass1(const Pass1Args& a, int n_sms, cudaStream_t st) {
  if (a.n == 0) return 0;
  const bool std_layout = a.bc_off == 0 && a.bc_len == 16 && a.umi_off == 16 && a.have_qual &&
                          ((uintptr_t)a.seq % 16 == 0) && ((uintptr_t)a.qual % 16 == 0) &&
                          (a.feature == nullptr || (uintptr_t)a.feature % 16 == 0);
  const bool split = getenv("CRGPU_P1_MODE") && atoi(getenv("CRGPU_P1_MODE")) == 2;  // default: fused kernel
  if (std_layout && a.r1_len == 28 && a.umi_len == 12)
    return split ? launch_split<28, 12>(a, n_sms, st) : launch_staged_cfg<28, 12>(a, n_sms, st);
  if (std_layout && a.r1_len == 26 && a.umi_len == 10)
    return split ? launch_split<26, 10>(a, n_sms, st) : launch_staged_cfg<26, 10>(a, n_sms, st);
  uint64_t blocks = (a.n + 255) / 256;
  int grid = (int)std::min<uint64_t>(blocks, (uint64_t)n_sms * 8);
  pass1_generic_kernel<<<grid, 256, 0, st>>>(a);
  return 1;
}

// ---------------------------------------------------------------------------
// Pass 2: Posterior::correct_barcode for every invalid read (corrector.rs:111-165).
// One thread per invalid read: the neighbour mask comes from n_ord bucket scans, then the likelihoods are
// accumulated strictly in the reference's (position ascending, base A,C,G,T) order with separate f64
// multiply and add (no FMA), so every accept/reject decision is bit-identical.
// The kernel is bound by memory latency (about 17 L2 sectors per read, a third of them DRAM misses, 95 % of the
// warp slots occupied at 32 registers). Measured on B200 and rejected (profiles/r02_pass2_sort_experiments.txt):
// reading the buckets as 16-byte windows, or as one 32-byte slot per ordering, with a SWAR test and all
// orderings in flight (fewer loads and sectors but 64 registers: 3.0-3.1 ms against 2.37); one key-space atomic
// per warp instead of per block (3.2 ms: the counter's L2 slice serialises a million returning atomics); the
// side list ordered by barcode prefix (2.99 -> 2.84 ms for four bases: locality of the scans is not the limit).
// This form of the source (posterior state in a struct) also schedules better than the one it replaces
// (2.37 against 2.60 ms at the same 32 registers).
// ---------------------------------------------------------------------------
struct Posterior2 {
  bool have_best = false;
  double best = 0.0, total = 0.0;
  uint32_t best_rank = 0;
  __device__ __forceinline__ void add(uint32_t rank, uint32_t raw, uint32_t qv) {
    if (qv > 66u) qv = 66u;  // BC_MAX_QV
    const double lik = __dmul_rn(c_bc_prob[qv], (double)(1ull + (unsigned long long)raw));
    if (!have_best || lik > best || (lik == best && rank >= best_rank)) {
      have_best = true;
      best = lik;
      best_rank = rank;
    }
    total = __dadd_rn(total, lik);
  }
};

__global__ void __launch_bounds__(256, 6) pass2_kernel(const Pass2Args a) {
  __shared__ uint32_t scan_a[9];
  __shared__ unsigned long long base_bcast;
  const uint64_t n_invalid = a.n_invalid_dev ? (*a.n_invalid_dev >> a.n_invalid_dev_shift) : a.n_invalid;
  const uint64_t n_blocks_work = (n_invalid + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < n_blocks_work; blk += gridDim.x) {
    const uint64_t e = blk * 256 + threadIdx.x;
    bool emit = false;
    unsigned long long key = 0ull;
    if (e < n_invalid) {
      // the side list streams through once: keep it from displacing the whitelist tables in L2
      const uint32_t idx = __ldcs(a.inv_idx + e);
      const uint32_t q = __ldcs(a.inv_bc + e);
      const uint32_t nmask = __ldcs(a.inv_nmask + e);
      const uint4 qq = __ldcs(a.inv_qual + e);
      const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w};
      const int L = a.wl.L;
      uint32_t uw = 0u, feat_raw = NO_FEATURE;
      Posterior2 post;
      unsigned long long m;
      if (nmask == 0u) {
        m = wl_neighbor_mask(a.wl, q);
      } else if ((nmask & (nmask - 1u)) == 0u) {
        // a single non-ACGT base: only the four trials at that position can be whitelist sequences
        int pn = __ffs(nmask) - 1;
        m = 0xFull << (4 * pn);
      } else {
        m = 0ull;  // every trial still holds a non-ACGT base
      }
      while (m) {
        int bit = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        int pos = bit >> 2;
        uint32_t base = bit & 3;
        int sh = 2 * (L - 1 - pos);
        uint32_t trial = (q & ~(3u << sh)) | (base << sh);
        int widx = wl_find(a.wl, trial);
        if (widx < 0) continue;
        uint32_t rank = wl_rank_of(a.wl, widx);
        uint32_t raw = __ldg(a.prior + rank);
        uint32_t qv = a.have_qual ? ((qw[pos >> 2] >> (8 * (pos & 3))) & 0xFFu) : 66u;
        post.add(rank, raw, qv);
      }
      bool accept = false;
      if (post.have_best) {
        bool ee_ok = true;
        if (a.check_expected_errors) {
          double ee = 0.0;
          if (a.have_qual)
            for (int i = 0; i < L; i++) ee = __dadd_rn(ee, c_bc_prob[(qw[i >> 2] >> (8 * (i & 3))) & 0xFFu]);
          ee_ok = ee < a.max_expected_errors;
        }
        accept = ee_ok && (__ddiv_rn(post.best, post.total) >= a.threshold);
      }
      if (accept) {
        const uint32_t best_rank = post.best_rank;
        __stcs(a.bc_out + idx, (ST_VALID_AFTER << BC_STATE_SHIFT) | best_rank);
        if (a.corrected) atomicAdd(a.corrected + best_rank, 1u);
        if (a.emit_keys) {
          uw = __ldcs(a.umi_out + idx);
          if (a.feature) feat_raw = __ldcs(a.feature + idx);
          const uint32_t feature = a.feature ? checked_feature(feat_raw, a.n_features, nullptr) : NO_FEATURE;
          if ((uw & UMI_VALID_BIT) && feature != NO_FEATURE) {
            emit = true;
            key = make_key(a.kl, best_rank, feature, a.lib, uw & UMI_SEQ_MASK);
          }
        }
      }
    }
    if (a.emit_keys) {  // one atomic per block (per warp: 3.2 ms, the counter's L2 slice serialises them)
      uint32_t tot;
      uint32_t off = block_exclusive_scan<256>(emit ? 1u : 0u, &tot, scan_a);
      if (threadIdx.x == 0) base_bcast = tot ? atomicAdd(a.counters, (unsigned long long)tot) : 0ull;
      __syncthreads();
      if (emit) __stcs(a.keys + base_bcast + off, key);
      __syncthreads();
    }
  }
}

int launch_pass2(const Pass2Args& a, cudaStream_t st) {
  if (a.n_invalid == 0) return 0;
  uint64_t blocks = (a.n_invalid + 255) / 256;
  int grid = (int)std::min<uint64_t>(blocks, (uint64_t)sm_count() * 64);
  pass2_kernel<<<grid, 256, 0, st>>>(a);
  return 1;
}

// ---------------------------------------------------------------------------
// total_barcode_counts (barcode_correction.rs:327-362): the reads of the side list that pass 2 left invalid, as
// keys (non-ACGT mask << 32 | packed sequence) for a sort + run-length count.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) collect_invalid_kernel(const uint32_t* __restrict__ inv_idx,
                                                              const uint32_t* __restrict__ inv_bc,
                                                              const uint32_t* __restrict__ inv_nmask,
                                                              const unsigned long long* __restrict__ n_invalid_dev,
                                                              const uint32_t* __restrict__ bc_out,
                                                              unsigned long long* __restrict__ out,
                                                              unsigned long long* __restrict__ counter) {
  const uint64_t n = *n_invalid_dev;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (n + stride - 1) / stride;  // warp-uniform trip count: ballots inside
  for (uint64_t r = 0; r < rounds; r++) {
    const uint64_t e = r * stride + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool take = false;
    unsigned long long key = 0ull;
    if (e < n) {
      take = (bc_out[inv_idx[e]] >> BC_STATE_SHIFT) == ST_INVALID;
      key = ((unsigned long long)(inv_nmask[e] & 0xFFFFu) << 32) | inv_bc[e];
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, take);
    if (m) {
      unsigned long long base = 0ull;
      if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (take) out[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
  }
}

int launch_collect_invalid(const uint32_t* inv_idx, const uint32_t* inv_bc, const uint32_t* inv_nmask,
                           const unsigned long long* n_invalid_dev, uint64_t n_max, const uint32_t* bc_out,
                           unsigned long long* out, unsigned long long* counter, cudaStream_t st) {
  if (!n_max) return 0;
  const int grid = (int)std::min<uint64_t>((n_max + 255) / 256, (uint64_t)sm_count() * 16);
  collect_invalid_kernel<<<grid, 256, 0, st>>>(inv_idx, inv_bc, inv_nmask, n_invalid_dev, bc_out, out, counter);
  return 1;
}

// ---------------------------------------------------------------------------
// Feature-barcode libraries: tethered fixed-offset capture R2[fb_off : fb_off+fb_len].
// exact_counts != nullptr: MAKE_SHARD pre-count of exact captures (make_shard_metrics.rs:337-345).
// feat_dist   != nullptr: ALIGN_AND_COUNT extraction with Hamming-1 posterior correction
// (feature_extraction.rs:34-117,447-471).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int fb_find(const uint32_t* __restrict__ keys, int n, uint32_t q) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    uint32_t e = keys[mid];
    if (e == q) return mid;
    if (e < q)
      lo = mid + 1;
    else
      hi = mid;
  }
  return -1;
}

constexpr int FB_THREADS = 256;
constexpr int FB_RPT = 8;
constexpr int FB_TILE = FB_THREADS * FB_RPT;

// Phase A of a tile: every read's capture is packed and looked up exactly; the captures that miss (and may
// be corrected) are compacted into a shared-memory work list. Phase B corrects the list densely - otherwise
// every warp would walk the 3L mutants because one of its lanes has to.
__global__ void __launch_bounds__(FB_THREADS) fb_kernel(const FbArgs a) {
  extern __shared__ __align__(16) uint32_t s_fb[];
  uint32_t* s_keys = s_fb;                   // n_fb sorted packed feature sequences
  uint32_t* s_index = s_fb + a.n_fb;         // n_fb feature indices
  uint32_t* s_cnt = s_fb + 2 * a.n_fb;       // n_fb exact-hit counters
  uint32_t* w_q = s_fb + 3 * a.n_fb;         // FB_TILE: packed capture of a work item
  uint32_t* w_meta = w_q + FB_TILE;          // FB_TILE: index in tile | (non-ACGT position + 1) << 16
  uint4* w_qual = reinterpret_cast<uint4*>(s_fb + ((3 * a.n_fb + 2 * FB_TILE + 3) & ~3));  // FB_TILE quality words
  __shared__ uint32_t s_nwork;
  for (int i = threadIdx.x; i < a.n_fb; i += FB_THREADS) {
    s_keys[i] = a.fb_keys[i];
    s_index[i] = a.fb_index[i];
    s_cnt[i] = 0;
  }
  const uint64_t n_tiles = (a.n + FB_TILE - 1) / FB_TILE;
  const bool fits = a.r2_len >= a.fb_off + a.fb_len;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) s_nwork = 0;
    __syncthreads();
    const uint64_t first = tile * FB_TILE;
#pragma unroll 2
    for (int k = 0; k < FB_RPT; k++) {
      const uint64_t gi = first + (uint64_t)k * FB_THREADS + threadIdx.x;
      if (gi >= a.n) break;
      uint32_t out = NO_FEATURE;
      if (fits) {
        const uint64_t avail = (a.n - gi) * (uint64_t)a.r2_len - (uint64_t)a.fb_off;  // bytes left in the buffer
        const uint4 sq = load_bytes16_bounded(a.r2_seq + gi * a.r2_len + a.fb_off, avail);
        const uint32_t sw[4] = {sq.x, sq.y, sq.z, sq.w};
        uint32_t q = 0, nmask = 0;
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
          const int left = a.fb_len - 4 * wd;  // bases of the capture in this word
          if (left <= 0) break;
          uint32_t w = sw[wd];
          const uint32_t bmask = left >= 4 ? 0xFFFFFFFFu : ((1u << (8 * left)) - 1u);
          w = (w & bmask) | (0x41414141u & ~bmask);
          uint32_t bad;
          uint32_t pk = pack4(w, &bad);
          bad &= bmask;
          if (bad) {
#pragma unroll
            for (int by = 0; by < 4; by++)
              if (bad & (0x80u << (8 * by))) nmask |= 1u << (wd * 4 + by);
          }
          q = left >= 4 ? ((q << 8) | pk) : ((q << (2 * left)) | (pk >> (2 * (4 - left))));
        }
        int hit = nmask ? -1 : fb_find(s_keys, a.n_fb, q);
        if (hit >= 0) {
          out = s_index[hit];
          if (a.exact_counts) atomicAdd(&s_cnt[hit], 1u);  // a few hundred hot features: count per block
        } else if (a.feat_dist && (nmask & (nmask - 1u)) == 0u) {
          const uint32_t slot = atomicAdd(&s_nwork, 1u);
          w_q[slot] = q;
          w_meta[slot] = (uint32_t)(k * FB_THREADS + threadIdx.x) | ((nmask ? (uint32_t)__ffs(nmask) : 0u) << 16);
          w_qual[slot] = load_bytes16_bounded(a.r2_qual + gi * a.r2_len + a.fb_off, avail);
        }
      }
      if (a.feature_out) a.feature_out[gi] = out;
    }
    __syncthreads();
    // phase B: correct_feature_barcode for a single candidate, trials in (position, A,C,G,T) order
    const uint32_t nwork = s_nwork;
    for (uint32_t wi = threadIdx.x; wi < nwork; wi += FB_THREADS) {
      const uint32_t q = w_q[wi];
      const uint32_t meta = w_meta[wi];
      const int npos = (int)(meta >> 16) - 1;  // position of the single non-ACGT base, or -1
      const uint4 qq = w_qual[wi];
      const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w};
      double sum = 0.0, best = -1.0;
      uint32_t best_f = NO_FEATURE;
      for (int i = 0; i < a.fb_len; i++) {
        if (npos >= 0 && i != npos) continue;  // the other positions keep the N and cannot match
        const int sh = 2 * (a.fb_len - 1 - i);
        const uint32_t orig = npos >= 0 ? 4u : ((q >> sh) & 3u);
        for (uint32_t b = 0; b < 4; b++) {
          if (b == orig) continue;
          const uint32_t trial = (q & ~(3u << sh)) | (b << sh);
          const int h = fb_find(s_keys, a.n_fb, trial);
          if (h < 0) continue;
          const uint32_t f = s_index[h];
          uint32_t qv = (uint8_t)((uint8_t)(qw[i >> 2] >> (8 * (i & 3))) - 33);
          if (qv > 33u) qv = 33u;  // FEATURE_MAX_QV
          const double lik = __dmul_rn(a.feat_dist[f], c_fb_prob[qv]);
          sum = __dadd_rn(sum, lik);
          if (lik > best) {
            best = lik;
            best_f = f;
          }
        }
      }
      if (best_f != NO_FEATURE && __ddiv_rn(best, sum) >= a.threshold && a.feature_out)
        a.feature_out[first + (meta & 0xFFFFu)] = best_f;
    }
    __syncthreads();
  }
  if (a.exact_counts) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_fb; i += FB_THREADS)
      if (s_cnt[i]) atomicAdd(a.exact_counts + s_index[i], (unsigned long long)s_cnt[i]);
  }
}

int launch_fb(const FbArgs& a, cudaStream_t st) {
  if (a.n == 0) return 0;
  uint64_t tiles = (a.n + FB_TILE - 1) / FB_TILE;
  int grid = (int)std::min<uint64_t>(tiles, (uint64_t)sm_count() * 3);
  size_t smem = (((size_t)3 * a.n_fb + 2 * FB_TILE + 3) & ~(size_t)3) * 4 + (size_t)FB_TILE * 16;
  cudaFuncSetAttribute(fb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fb_kernel<<<grid, FB_THREADS, smem, st>>>(a);
  return 1;
}

// keys of a whole batch after its features are known (feature-barcode libraries)
__global__ void __launch_bounds__(256) emit_keys_kernel(const EmitArgs a) {
  __shared__ uint32_t scan_a[9];
  __shared__ unsigned long long base_bcast;
  const uint64_t n_blocks_work = (a.n + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < n_blocks_work; blk += gridDim.x) {
    uint64_t gi = blk * 256 + threadIdx.x;
    bool emit = false;
    unsigned long long key = 0;
    if (gi < a.n) {
      uint32_t bw = a.bc_out[gi], uw = a.umi_out[gi], f = a.feature[gi];
      uint32_t st = bw >> BC_STATE_SHIFT;
      if ((st == ST_VALID_BEFORE || st == ST_VALID_AFTER) && (uw & UMI_VALID_BIT) && f != NO_FEATURE) {
        emit = true;
        key = make_key(a.kl, bw & BC_RANK_MASK, f, a.lib, uw & UMI_SEQ_MASK);
      }
    }
    uint32_t tot;
    uint32_t off = block_exclusive_scan<256>(emit ? 1u : 0u, &tot, scan_a);
    if (threadIdx.x == 0) base_bcast = tot ? atomicAdd(a.counters, (unsigned long long)tot) : 0ull;
    __syncthreads();
    if (emit) a.keys[base_bcast + off] = key;
    __syncthreads();
  }
}

int launch_emit_keys(const EmitArgs& a, cudaStream_t st) {
  if (a.n == 0) return 0;
  uint64_t blocks = (a.n + 255) / 256;
  int grid = (int)std::min<uint64_t>(blocks, (uint64_t)sm_count() * 16);
  emit_keys_kernel<<<grid, 256, 0, st>>>(a);
  return 1;
}

__global__ void valid_counts_kernel(const uint32_t* prior, const uint32_t* corrected, uint32_t* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = prior[i] + corrected[i];
}
int launch_valid_counts(const uint32_t* prior, const uint32_t* corrected, uint32_t* out, uint64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count() * 16);
  valid_counts_kernel<<<grid, 256, 0, st>>>(prior, corrected, out, n);
  return 1;
}

// exported for crgpu.cu
void launch_split_counter(unsigned long long* packed, unsigned long long* keys_total, unsigned long long* inv_total,
                          cudaStream_t st) {
  split_counter_kernel<<<1, 1, 0, st>>>(packed, keys_total, inv_total);
}
void launch_merge_counter(unsigned long long* packed, const unsigned long long* keys_total, cudaStream_t st) {
  merge_counter_kernel<<<1, 1, 0, st>>>(packed, keys_total);
}
