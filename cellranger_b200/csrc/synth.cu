// synth.cu — device implementation of the synthetic read generator of cellranger_b200/synth.py.
// Same counter-based integer arithmetic as the numpy version, so both produce identical bytes
// (checked by tests/test_gpu_parity.py::test_synth_device_matches_numpy).
#include <algorithm>

#include "../../include/crgpu.h"
#include "kernels.h"

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// first index with cdf[i] > u (numpy searchsorted side='right'), clamped
__device__ __forceinline__ uint32_t pick(const uint32_t* __restrict__ cdf, uint32_t n, uint32_t u) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (cdf[mid] <= u)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo < n ? lo : n - 1;
}

struct SynthDev {
  unsigned long long seed_mix, seed_mol;
  uint32_t n_whitelist, n_cells, n_genes, n_fb;
  int bc_len, umi_len, fb_offset, fb_len, is_fb;
  uint32_t ambient_thr, unmapped_thr, bc_err_thr, umi_err_thr, n_thr, fb_err_thr, homopolymer_thr;
  int n_qual_ascii, n_qual_classes, n_equal_classes;
  uint32_t qual_thr[8], equal_thr[8];
  uint8_t qual_val[8], equal_val[8];
  const uint32_t *wl, *cell_rank, *cell_cdf, *n_mol, *gene_cdf, *fb_cdf, *fb_packed;
};

// first index with thr[i] >= u (numpy searchsorted side='left'), clamped
__device__ __forceinline__ uint8_t qual_of(const uint32_t* thr, const uint8_t* val, int k, uint32_t u) {
  int i = 0;
  while (i < k - 1 && thr[i] < u) i++;
  return val[i];
}

__global__ void __launch_bounds__(256) synth_kernel(const SynthDev p, uint64_t start, uint64_t n, uint8_t* r1_seq,
                                                    uint8_t* r1_qual, uint32_t* feature, uint8_t* r2_seq,
                                                    uint8_t* r2_qual) {
  const char bases[4] = {'A', 'C', 'G', 'T'};
  const int L1 = p.bc_len + p.umi_len;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long gi = start + i;
    const unsigned long long ctr = p.seed_mix + gi * 64ull;
    unsigned long long w0 = splitmix64(ctr + 0);
    bool ambient = (uint32_t)(w0 & 0xFFFFFFFFull) < p.ambient_thr;
    uint32_t cell = pick(p.cell_cdf, p.n_cells, (uint32_t)(w0 >> 32));
    unsigned long long w1 = splitmix64(ctr + 1);
    unsigned long long amb_rank = (w1 >> 11) % (unsigned long long)p.n_whitelist;
    unsigned long long rank = ambient ? amb_rank : (unsigned long long)p.cell_rank[cell];
    unsigned long long cid = ambient ? (unsigned long long)p.n_cells + amb_rank : (unsigned long long)cell;
    unsigned long long w2 = splitmix64(ctr + 2);
    unsigned long long nm = ambient ? (1ull << 20) : (unsigned long long)p.n_mol[cell];
    unsigned long long mol = (w2 >> 11) % nm;
    unsigned long long mkey = (cid << 27) + mol + (p.is_fb ? (1ull << 26) : 0ull);
    unsigned long long m0 = splitmix64(p.seed_mol + mkey * 4ull);
    unsigned long long m1 = splitmix64(p.seed_mol + mkey * 4ull + 1ull);
    unsigned long long umi = m1 & ((1ull << (2 * p.umi_len)) - 1ull);
    if (p.homopolymer_thr) {
      bool homo = (uint32_t)((m1 >> 32) & 0xFFFFFFFFull) < p.homopolymer_thr;
      unsigned long long hb = (m1 >> 30) & 3ull;
      unsigned long long rep = 0x5555555555555555ull & ((1ull << (2 * p.umi_len)) - 1ull);
      if (homo) umi = hb * rep;
    }
    uint32_t feat_i = 0;
    if (p.is_fb) {
      feat_i = pick(p.fb_cdf, p.n_fb, (uint32_t)(m0 >> 32));
    } else {
      uint32_t gene = pick(p.gene_cdf, p.n_genes, (uint32_t)(m0 >> 32));
      unsigned long long w3 = splitmix64(ctr + 3);
      bool unmapped = (uint32_t)(w3 & 0xFFFFFFFFull) < p.unmapped_thr;
      feature[i] = unmapped ? 0xFFFFFFFFu : gene;
    }
    unsigned long long bc = p.wl[rank];
    unsigned long long wn = splitmix64(ctr + 5);
    bool has_n = (uint32_t)(wn & 0xFFFFFFFFull) < p.n_thr;
    int npos = (int)((wn >> 32) % (unsigned long long)L1);
    uint8_t* so = r1_seq + i * (uint64_t)L1;
    uint8_t* qo = r1_qual + i * (uint64_t)L1;
    for (int pos = 0; pos < L1; pos++) {
      uint32_t base, thr;
      if (pos < p.bc_len) {
        base = (uint32_t)(bc >> (2 * (p.bc_len - 1 - pos))) & 3u;
        thr = p.bc_err_thr;
      } else {
        int q = pos - p.bc_len;
        base = (uint32_t)(umi >> (2 * (p.umi_len - 1 - q))) & 3u;
        thr = p.umi_err_thr;
      }
      unsigned long long wb = splitmix64(ctr + 8 + pos);
      bool err = (uint32_t)(wb & 0xFFFFFFFFull) < thr;
      uint32_t sub = (base + 1u + (uint32_t)((wb >> 32) & 0xFFull) % 3u) & 3u;
      if (err) base = sub;
      uint32_t uq = (uint32_t)((wb >> 40) & 0xFFFFull);
      uint8_t qv = err ? qual_of(p.equal_thr, p.equal_val, p.n_equal_classes, uq)
                       : qual_of(p.qual_thr, p.qual_val, p.n_qual_classes, uq);
      uint8_t c = (uint8_t)bases[base];
      if (has_n && pos == npos) {
        c = 'N';
        qv = (uint8_t)(p.n_qual_ascii);
      }
      so[pos] = c;
      qo[pos] = qv;
    }
    if (p.is_fb) {
      const int L2 = p.fb_offset + p.fb_len;
      unsigned long long fbp = p.fb_packed[feat_i];
      uint8_t* s2 = r2_seq + i * (uint64_t)L2;
      uint8_t* q2 = r2_qual + i * (uint64_t)L2;
      for (int pos = 0; pos < L2; pos++) {
        unsigned long long wb = splitmix64(ctr + 40 + pos);
        uint32_t base;
        bool err = false;
        if (pos < p.fb_offset) {
          base = (uint32_t)(wb >> 34) & 3u;
        } else {
          int q = pos - p.fb_offset;
          base = (uint32_t)(fbp >> (2 * (p.fb_len - 1 - q))) & 3u;
          err = (uint32_t)(wb & 0xFFFFFFFFull) < p.fb_err_thr;
          uint32_t sub = (base + 1u + (uint32_t)((wb >> 32) & 0xFFull) % 3u) & 3u;
          if (err) base = sub;
        }
        uint32_t uq = (uint32_t)((wb >> 40) & 0xFFFFull);
        uint8_t qv = err ? qual_of(p.equal_thr, p.equal_val, p.n_equal_classes, uq)
                         : qual_of(p.qual_thr, p.qual_val, p.n_qual_classes, uq);
        s2[pos] = (uint8_t)bases[base];
        q2[pos] = qv;
      }
    }
  }
}

}  // namespace

int launch_synth(const crgpu_synth_params* hp, const uint32_t* d_wl, const uint32_t* d_cell_rank,
                 const uint32_t* d_cell_cdf, const uint32_t* d_n_mol, const uint32_t* d_gene_cdf,
                 const uint32_t* d_fb_cdf, const uint32_t* d_fb_packed, uint64_t start, uint64_t n, uint8_t* r1_seq,
                 uint8_t* r1_qual, uint32_t* feature, uint8_t* r2_seq, uint8_t* r2_qual, cudaStream_t st) {
  if (n == 0) return 0;
  SynthDev p;
  p.seed_mix = hp->seed_mix;
  p.seed_mol = hp->seed_mol;
  p.n_whitelist = hp->n_whitelist;
  p.n_cells = hp->n_cells;
  p.n_genes = hp->n_genes;
  p.n_fb = hp->n_fb;
  p.bc_len = hp->bc_len;
  p.umi_len = hp->umi_len;
  p.fb_offset = hp->fb_offset;
  p.fb_len = hp->fb_len;
  p.is_fb = hp->is_fb;
  p.ambient_thr = hp->ambient_thr;
  p.unmapped_thr = hp->unmapped_thr;
  p.bc_err_thr = hp->bc_err_thr;
  p.umi_err_thr = hp->umi_err_thr;
  p.n_thr = hp->n_thr;
  p.fb_err_thr = hp->fb_err_thr;
  p.homopolymer_thr = hp->homopolymer_thr;
  p.n_qual_ascii = hp->n_qual_ascii;
  p.n_qual_classes = hp->n_qual_classes;
  p.n_equal_classes = hp->n_equal_classes;
  for (int i = 0; i < 8; i++) {
    p.qual_thr[i] = hp->qual_thr[i];
    p.equal_thr[i] = hp->equal_thr[i];
    p.qual_val[i] = hp->qual_val[i];
    p.equal_val[i] = hp->equal_val[i];
  }
  p.wl = d_wl;
  p.cell_rank = d_cell_rank;
  p.cell_cdf = d_cell_cdf;
  p.n_mol = d_n_mol;
  p.gene_cdf = d_gene_cdf;
  p.fb_cdf = d_fb_cdf;
  p.fb_packed = d_fb_packed;
  int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count() * 16);
  synth_kernel<<<grid, 256, 0, st>>>(p, start, n, r1_seq, r1_qual, feature, r2_seq, r2_qual);
  return 1;
}
