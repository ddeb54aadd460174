// compat.cu — the other users of the resident whitelist table: whitelist match rates and the barcode
// compatibility check between library types. sm_100a.
//
//   sample_valid_barcodes      lib/rust/cr_lib/src/stages/check_barcodes_compatibility.rs:98-120: a histogram of the
//                              reads' barcodes that are on the whitelist, "robust to a single N cycle"
//                              (Whitelist::match_to_whitelist, lib/rust/barcode/src/whitelist.rs:526-545)
//   whitelist match fraction   lib/rust/cr_lib/src/detect_chemistry/whitelist_filter.rs:61-110,162-193: the same
//                              match per read, reads_with_bc_in_wl / reads_with_bc
//   robust_cosine_similarity   check_barcodes_compatibility.rs:122-158 over two such histograms, counts capped at
//                              their N92.5 (stats::nx::nx, lib/rust/stats/src/nx.rs:6-38)
//
// All sums are integers (exact on the device, in any order); the three f64 operations of the reference's last
// line run on the host in its order, so the similarity is bit-identical whenever the reference's own f64 sums are
// exact (always: a histogram of at most 10^6 sampled reads has sum of squares < 2^53).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ctx.h"

namespace {

__global__ void __launch_bounds__(256) wl_match_kernel(const uint8_t* __restrict__ seqs, uint64_t n, int stride, int bc_off,
                                                       DevWhitelist wl, uint32_t* __restrict__ hist,
                                                       unsigned long long* __restrict__ matched) {
  unsigned long long hits = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t* s = seqs + i * (uint64_t)stride + bc_off;
    uint32_t q = 0;
    int n_bad = 0, pos_n = -1;
    bool other = false;  // a byte that is neither A,C,G,T nor N can never be rescued
    for (int p = 0; p < wl.L; p++) {
      const uint8_t ch = s[p];
      uint32_t code = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
      if (code == 4u) {
        if (ch == 'N' && pos_n < 0) pos_n = p;  // position of the FIRST N (whitelist.rs:536)
        else if (ch != 'N') other = true;
        n_bad++;
        code = 0u;
      }
      q = (q << 2) | code;
    }
    int idx = -1;
    if (n_bad == 0) {
      idx = wl_find(wl, q);
    } else if (n_bad == 1 && !other && pos_n >= 0) {
      // only the first N is replaced: with a second non-ACGT base no trial can be on the whitelist
      const int sh = 2 * (wl.L - 1 - pos_n);
      for (uint32_t b = 0; b < 4 && idx < 0; b++) idx = wl_find(wl, (q & ~(3u << sh)) | (b << sh));  // first hit in A,C,G,T
    }
    if (idx >= 0) {
      atomicAdd(hist + idx, 1u);
      hits++;
    }
  }
  for (int d = 16; d > 0; d >>= 1) hits += __shfl_xor_sync(0xFFFFFFFFu, hits, d);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(matched, hits);
}

// out[0] = sum of the counts >= t, out[1] = the largest count, out[2] = sum of all counts, out[3] = entries > 0
__global__ void __launch_bounds__(256) hist_tail_kernel(const uint32_t* __restrict__ h, uint64_t n, uint32_t t,
                                                        unsigned long long* __restrict__ out) {
  unsigned long long ge = 0, sum = 0, nz = 0;
  uint32_t mx = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = h[i];
    if (c >= t) ge += c;
    sum += c;
    nz += c != 0u;
    mx = c > mx ? c : mx;
  }
  for (int d = 16; d > 0; d >>= 1) {
    ge += __shfl_xor_sync(0xFFFFFFFFu, ge, d);
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    nz += __shfl_xor_sync(0xFFFFFFFFu, nz, d);
    const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, mx, d);
    mx = o > mx ? o : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    if (ge) atomicAdd(out + 0, ge);
    atomicMax(out + 1, (unsigned long long)mx);
    if (sum) atomicAdd(out + 2, sum);
    if (nz) atomicAdd(out + 3, nz);
  }
}

// histogram with its keys mapped (SimpleHistogram::map_key): out[map[i]] += h[i]
__global__ void __launch_bounds__(256) hist_map_kernel(const uint32_t* __restrict__ h, const uint32_t* __restrict__ map,
                                                       uint64_t n, uint32_t* __restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (h[i]) atomicAdd(out + map[i], h[i]);
}

// out[0] = sum min(a, ta)^2, out[1] = sum min(b, tb)^2, out[2] = sum min(a, ta) * min(b, tb)
__global__ void __launch_bounds__(256) capped_products_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                              uint64_t n, uint32_t ta, uint32_t tb,
                                                              unsigned long long* __restrict__ out) {
  unsigned long long aa = 0, bb = 0, ab = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long x = a[i] < ta ? a[i] : ta, y = b[i] < tb ? b[i] : tb;
    aa += x * x;
    bb += y * y;
    ab += x * y;
  }
  for (int d = 16; d > 0; d >>= 1) {
    aa += __shfl_xor_sync(0xFFFFFFFFu, aa, d);
    bb += __shfl_xor_sync(0xFFFFFFFFu, bb, d);
    ab += __shfl_xor_sync(0xFFFFFFFFu, ab, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (aa) atomicAdd(out + 0, aa);
    if (bb) atomicAdd(out + 1, bb);
    if (ab) atomicAdd(out + 2, ab);
  }
}

int grid_of(uint64_t n) { return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)sm_count() * 8)); }

// stats::nx::nx over the non-zero counts of a device histogram: the count at which the running sum of the counts,
// taken in descending order, first reaches fraction * total. That is the largest t with (sum of counts >= t) >=
// cutoff - found by bisection on t, one reduction per step. Returns 0 for an empty histogram (nx: None).
int hist_nx(crgpu_ctx* c, const uint32_t* d_hist, uint64_t n, double fraction, uint32_t* out_t, unsigned long long* d4) {
  unsigned long long h[4];
  auto tail = [&](uint32_t t) -> int {
    CU(cudaMemsetAsync(d4, 0, 32, c->stream));
    hist_tail_kernel<<<grid_of(n), 256, 0, c->stream>>>(d_hist, n, t, d4);
    c->launches++;
    CU(cudaMemcpyAsync(h, d4, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CRGPU_OK;
  };
  int rc;
  if ((rc = tail(1))) return rc;
  if (h[3] == 0) {
    *out_t = 0;
    return CRGPU_OK;
  }
  const double cutoff = (double)h[2] * fraction;  // sum * fraction, nx.rs:30
  uint32_t lo = 1, hi = (uint32_t)h[1];           // invariant: S_ge(lo) >= cutoff
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo + 1) / 2;
    if ((rc = tail(mid))) return rc;
    if ((double)h[0] >= cutoff)
      lo = mid;
    else
      hi = mid - 1;
  }
  *out_t = lo;
  return CRGPU_OK;
}

}  // namespace

extern "C" {

int crgpu_whitelist_entries(crgpu_ctx* c, int wl, uint64_t* n_entries) {
  if (!c || !n_entries || wl < 0 || wl >= (int)c->wls.size()) return fail(CRGPU_E_INVALID, "bad argument");
  *n_entries = c->wls[wl]->W;
  return CRGPU_OK;
}

int crgpu_dev_memset(crgpu_ctx* c, void* dev, int value, uint64_t bytes) {
  if (!c || (!dev && bytes)) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(dev, value, bytes, c->stream));
  return CRGPU_OK;
}

int crgpu_sample_valid_barcodes(crgpu_ctx* c, int wl, const uint8_t* seqs, uint64_t n, int32_t stride, int32_t bc_offset,
                                int on_device, uint32_t* dev_hist, uint64_t* n_in_whitelist) {
  if (!c || wl < 0 || wl >= (int)c->wls.size() || (!seqs && n) || !dev_hist) return fail(CRGPU_E_INVALID, "bad argument");
  HostWhitelist* w = c->wls[wl];
  if (stride < bc_offset + w->L || bc_offset < 0) return fail(CRGPU_E_INVALID, "stride shorter than the barcode range");
  CU(cudaSetDevice(c->device));
  const uint8_t* d_seq = seqs;
  DevBuf tmp;
  int rc;
  if (!on_device && n) {
    if ((rc = tmp.ensure(n * (uint64_t)stride + 16))) return rc;
    cudaError_t e = cudaMemcpyAsync(tmp.p, seqs, n * (uint64_t)stride, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      tmp.release();
      return fail(CRGPU_E_CUDA, cudaGetErrorString(e));
    }
    d_seq = tmp.as<uint8_t>();
  }
  unsigned long long* d_m = c->scalars.as<unsigned long long>() + 56;
  unsigned long long h = 0;
  cudaError_t e = cudaMemsetAsync(d_m, 0, 8, c->stream);
  if (e == cudaSuccess && n) {
    wl_match_kernel<<<grid_of(n), 256, 0, c->stream>>>(d_seq, n, stride, bc_offset, w->dev, dev_hist, d_m);
    c->launches++;
    e = cudaPeekAtLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_m, 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  tmp.release();
  if (e != cudaSuccess) return fail(CRGPU_E_CUDA, std::string("sample_valid_barcodes: ") + cudaGetErrorString(e));
  if (n_in_whitelist) *n_in_whitelist = h;
  return CRGPU_OK;
}

int crgpu_hist_nx(crgpu_ctx* c, const uint32_t* dev_hist, uint64_t n, double fraction, uint32_t* out) {
  if (!c || !dev_hist || !out) return fail(CRGPU_E_INVALID, "bad argument");
  if (!(fraction > 0.0 && fraction < 1.0)) return fail(CRGPU_E_INVALID, "fraction must lie in (0, 1)");  // nx.rs:11
  CU(cudaSetDevice(c->device));
  return hist_nx(c, dev_hist, n, fraction, out, c->scalars.as<unsigned long long>() + 56);
}

int crgpu_robust_cosine_similarity(crgpu_ctx* c, const uint32_t* dev_hist_a, const uint32_t* dev_hist_b, uint64_t n,
                                   int translate_whitelist, double* out) {
  if (!c || !dev_hist_a || !dev_hist_b || !out) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  const uint32_t* b = dev_hist_b;
  DevBuf mapped;
  int rc;
  if (translate_whitelist >= 0) {
    // this_hist.map_key(|key| translate[&key]), check_barcodes_compatibility.rs:241-242: the translation whitelist's
    // raw -> content map takes entry i of the raw whitelist to the entry of its translated sequence
    if (translate_whitelist >= (int)c->wls.size()) return fail(CRGPU_E_INVALID, "unknown whitelist");
    HostWhitelist* w = c->wls[translate_whitelist];
    if (!w->is_trans) return fail(CRGPU_E_INVALID, "translate_whitelist is not a translation whitelist");
    if (w->W != n || c->content.size() != n)
      return fail(CRGPU_E_INVALID, "the translation whitelist and the histograms must cover the same entries");
    if ((rc = mapped.ensure(n * 4))) return rc;
    CU(cudaMemsetAsync(mapped.p, 0, n * 4, c->stream));
    hist_map_kernel<<<grid_of(n), 256, 0, c->stream>>>(dev_hist_b, w->dev.vals[0], n, mapped.as<uint32_t>());
    c->launches++;
    b = mapped.as<uint32_t>();
  }
  unsigned long long* d4 = c->scalars.as<unsigned long long>() + 56;
  uint32_t ta = 0, tb = 0;
  const double ROBUST_FRACTION_THRESHOLD = 0.925;  // check_barcodes_compatibility.rs:80
  rc = hist_nx(c, dev_hist_a, n, ROBUST_FRACTION_THRESHOLD, &ta, d4);
  if (!rc) rc = hist_nx(c, b, n, ROBUST_FRACTION_THRESHOLD, &tb, d4);
  if (rc) {
    mapped.release();
    return rc;
  }
  if (ta == 0 || tb == 0) {  // an empty histogram: 0 similarity (:130-137)
    mapped.release();
    *out = 0.0;
    return CRGPU_OK;
  }
  unsigned long long h[3] = {0, 0, 0};
  cudaError_t e = cudaMemsetAsync(d4, 0, 32, c->stream);
  if (e == cudaSuccess) {
    capped_products_kernel<<<grid_of(n), 256, 0, c->stream>>>(dev_hist_a, b, n, ta, tb, d4);
    c->launches++;
    e = cudaMemcpyAsync(h, d4, 24, cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  mapped.release();
  if (e != cudaSuccess) return fail(CRGPU_E_CUDA, std::string("robust_cosine_similarity: ") + cudaGetErrorString(e));
  const double mag1 = std::sqrt((double)h[0]);  // :139-149
  const double mag2 = std::sqrt((double)h[1]);
  const double dot_prod = (double)h[2];          // :151-155
  *out = dot_prod / (mag1 * mag2);               // :157
  return CRGPU_OK;
}

}  // extern "C"
