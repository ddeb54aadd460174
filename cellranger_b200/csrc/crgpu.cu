// crgpu.cu — the C ABI of include/crgpu.h: context, whitelist tables, stage drivers, results.
// Host-side logic only; every kernel lives in pass_kernels.cu / sort.cu / dedup_kernels.cu / synth.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.h"

void launch_split_counter(unsigned long long* packed, unsigned long long* keys_total, unsigned long long* inv_total,
                          cudaStream_t st);
void launch_merge_counter(unsigned long long* packed, const unsigned long long* keys_total, cudaStream_t st);

static thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

static int phase_event(crgpu_ctx* c, cudaEvent_t* e) {
  if (!c->event_pool.empty()) {
    *e = c->event_pool.back();
    c->event_pool.pop_back();
    return CRGPU_OK;
  }
  CU(cudaEventCreate(e));
  return CRGPU_OK;
}
int phase_begin(crgpu_ctx* c, const char* name) {
  cudaEvent_t a, b;
  int rc;
  if ((rc = phase_event(c, &a))) return rc;
  if ((rc = phase_event(c, &b))) return rc;
  CU(cudaEventRecord(a, c->stream));
  c->phases.push_back({name, {a, b}});
  return CRGPU_OK;
}
int phase_end(crgpu_ctx* c) {
  CU(cudaEventRecord(c->phases.back().second.second, c->stream));
  return CRGPU_OK;
}
void phases_clear(crgpu_ctx* c, const char* prefix) {
  // drop earlier records of the same stage so repeated calls do not accumulate
  std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> keep;
  for (auto& p : c->phases) {
    if (p.first.rfind(prefix, 0) == 0) {
      c->event_pool.push_back(p.second.first);
      c->event_pool.push_back(p.second.second);
    } else {
      keep.push_back(p);
    }
  }
  c->phases.swap(keep);
}

namespace {

bool pack_ascii(const uint8_t* s, int L, uint32_t* out) {
  uint32_t r = 0;
  for (int i = 0; i < L; i++) {
    uint32_t v;
    switch (s[i]) {
      case 'A': v = 0; break;
      case 'C': v = 1; break;
      case 'G': v = 2; break;
      case 'T': v = 3; break;
      default: return false;
    }
    r = (r << 2) | v;
  }
  *out = r;
  return true;
}

uint32_t rotr_host(uint32_t q, int r, int nbits) {
  if (r == 0) return q;
  uint64_t m = nbits >= 32 ? 0xFFFFFFFFull : ((1ull << nbits) - 1ull);
  uint64_t v = q;
  return (uint32_t)(((v >> r) | (v << (nbits - r))) & m);
}

int ensure_layout(crgpu_ctx* c) {
  if (c->layout_ready) return CRGPU_OK;
  if (c->libs.empty()) return fail(CRGPU_E_INVALID, "no library defined");
  if (c->content.empty()) return fail(CRGPU_E_INVALID, "no whitelist defined");
  int umi_len = c->libs[0]->def.umi_length;
  for (auto* l : c->libs)
    if (l->def.umi_length != umi_len) return fail(CRGPU_E_LIMIT, "all libraries must share one UMI length");
  int nf = c->n_features > 0 ? c->n_features : 1;
  int ubits = 2 * umi_len;
  int lbits = bits_for(c->libs.size());
  int fbits = std::max(1, bits_for((uint64_t)nf));
  int rbits = std::max(1, bits_for(c->content.size()));
  if (ubits > 30) return fail(CRGPU_E_LIMIT, "UMI longer than 15 bases");
  if (ubits + lbits + fbits + rbits > 64)
    return fail(CRGPU_E_LIMIT, "barcode rank + feature + library + UMI need more than 64 key bits");
  c->kl.umi_bits = ubits;
  c->kl.lib_shift = ubits;
  c->kl.feature_shift = ubits + lbits;
  c->kl.rank_shift = ubits + lbits + fbits;
  c->kl.total_bits = ubits + lbits + fbits + rbits;
  c->layout_ready = true;
  return CRGPU_OK;
}

int alloc_library_tables(crgpu_ctx* c, Library* l) {
  size_t bytes = c->content.size() * sizeof(uint32_t);
  int rc;
  if ((rc = l->prior.ensure(bytes))) return rc;
  if ((rc = l->corrected.ensure(bytes))) return rc;
  if ((rc = l->valid.ensure(bytes))) return rc;
  CU(cudaMemsetAsync(l->prior.p, 0, bytes, c->stream));
  CU(cudaMemsetAsync(l->corrected.p, 0, bytes, c->stream));
  CU(cudaMemsetAsync(l->valid.p, 0, bytes, c->stream));
  return CRGPU_OK;
}

int build_fb_tables(crgpu_ctx* c) {
  c->have_fb = false;
  for (auto* l : c->libs) {
    l->fb_keys.clear();
    l->fb_index.clear();
    if (!l->def.is_feature_barcode) continue;
    c->have_fb = true;
    if (l->def.fb_length < 1 || l->def.fb_length > 16) return fail(CRGPU_E_LIMIT, "feature barcode length must be 1..16");
    std::vector<std::pair<uint32_t, uint32_t>> kv;
    for (int f = 0; f < c->n_features; f++) {
      if (c->feature_type[f] != l->def.feature_type) continue;
      uint32_t pk;
      if (!pack_ascii(c->fb_seqs.data() + (size_t)f * c->fb_stride, l->def.fb_length, &pk))
        return fail(CRGPU_E_LIMIT, "feature barcode sequences must be A,C,G,T");
      kv.emplace_back(pk, (uint32_t)f);
    }
    std::sort(kv.begin(), kv.end());
    for (size_t i = 1; i < kv.size(); i++)
      if (kv[i].first == kv[i - 1].first) return fail(CRGPU_E_INVALID, "duplicate feature barcode sequence in one pattern");
    if (kv.size() > 4096) return fail(CRGPU_E_LIMIT, "more than 4096 feature barcodes in one library");
    for (auto& p : kv) {
      l->fb_keys.push_back(p.first);
      l->fb_index.push_back(p.second);
    }
    int rc;
    if ((rc = l->d_fb_keys.ensure(std::max<size_t>(4, kv.size() * 4)))) return rc;
    if ((rc = l->d_fb_index.ensure(std::max<size_t>(4, kv.size() * 4)))) return rc;
    if (!kv.empty()) {
      CU(cudaMemcpyAsync(l->d_fb_keys.p, l->fb_keys.data(), kv.size() * 4, cudaMemcpyHostToDevice, c->stream));
      CU(cudaMemcpyAsync(l->d_fb_index.p, l->fb_index.data(), kv.size() * 4, cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
  }
  return CRGPU_OK;
}

// compute_feature_dist — lib/rust/cr_types/src/reference/feature_checker.rs:8-50 (tiny, host side)
std::vector<double> feature_dist_host(const std::vector<unsigned long long>& raw, const std::vector<int32_t>& ftype) {
  size_t n = raw.size();
  std::vector<double> p(n, 0.0);
  std::vector<std::pair<int32_t, long long>> sums;
  auto sum_of = [&](int32_t t) -> long long& {
    for (auto& s : sums)
      if (s.first == t) return s.second;
    sums.emplace_back(t, 0);
    return sums.back().second;
  };
  for (size_t i = 0; i < n; i++) sum_of(ftype[i]) += (long long)raw[i];
  for (size_t i = 0; i < n; i++) {
    long long s = sum_of(ftype[i]);
    if (s > 0) p[i] = (double)(long long)raw[i] / (double)s;
  }
  bool all_zero = true;
  for (double x : p) all_zero = all_zero && (x == 0.0);
  if (all_zero)
    for (double& x : p) x = 1.0 / (double)n;
  return p;
}

}  // namespace

extern "C" {

int crgpu_version(void) { return 100; }
const char* crgpu_last_error(void) { return g_err.c_str(); }

int crgpu_ctx_create(int device, crgpu_ctx** out) {
  if (!out) return fail(CRGPU_E_INVALID, "out is NULL");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(CRGPU_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                  " — libcrgpu has no CPU fallback");
  if (device < 0 || device >= n_dev) return fail(CRGPU_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  crgpu_ctx* c = new crgpu_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->n_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  // probability LUTs from the host libm, as the reference's f64::powf resolves to (corrector.rs:167-171,
  // feature_extraction.rs:43-44)
  double bc_lut[256], fb_lut[64];
  for (int q = 0; q < 256; q++) bc_lut[q] = pow(10.0, -((double)q - 33.0) / 10.0);
  for (int q = 0; q < 64; q++) fb_lut[q] = pow(10.0, -(double)q / 10.0);
  upload_prob_luts(bc_lut, fb_lut, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  int rc;
  if ((rc = c->counters.ensure(8 * 1024))) return rc;
  if ((rc = c->scalars.ensure(8 * 64))) return rc;
  if ((rc = c->tickets.ensure(64))) return rc;
  *out = c;
  return CRGPU_OK;
}

void crgpu_ctx_destroy(crgpu_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  shard_release(c);
  auto free_batch = [](Batch* b) {
    b->own_seq.release(); b->own_qual.release(); b->own_feat.release(); b->own_r2s.release(); b->own_r2q.release();
    b->own_select.release();
    b->feature_res.release(); b->inv_idx.release(); b->inv_bc.release(); b->inv_nmask.release(); b->inv_qual.release();
    delete b;
  };
  for (auto* b : c->batches) free_batch(b);
  for (auto* b : c->batch_pool) free_batch(b);
  for (auto* w : c->wls) {
    for (auto& b : w->bufs) b.release();
    delete w;
  }
  for (auto* l : c->libs) {
    l->prior.release(); l->corrected.release(); l->valid.release(); l->d_fb_keys.release(); l->d_fb_index.release();
    delete l;
  }
  DevBuf* all[] = {&c->d_fb_counts, &c->d_feat_dist, &c->bc_out, &c->umi_out, &c->umi_proc, &c->flags, &c->keys,
                   &c->keys_alt, &c->sort_temp, &c->counters, &c->dkeys, &c->c0, &c->best, &c->inc, &c->low, &c->key2,
                   &c->key2_alt, &c->lb_desc, &c->tickets, &c->scalars, &c->ent_rank, &c->ent_feature, &c->ent_count,
                   &c->mol, &c->col_of_rank, &c->barcode_rank, &c->indptr, &c->mol_rows, &c->min_read, &c->rep_raw, &c->summary, &c->fastq_text, &c->fastq_tmp,
                   &c->ls_slots, &c->mol_idx, &c->mol_sort, &c->mol_sort_alt, &c->on_target};
  for (auto* b : all) b->release();
  for (auto& p : c->phases) {
    cudaEventDestroy(p.second.first);
    cudaEventDestroy(p.second.second);
  }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  {
    DevBuf* cb[] = {&c->cb_seq, &c->cb_qual, &c->cb_bc, &c->cb_umi, &c->cb_keys, &c->cb_ctr, &c->cb_idx, &c->cb_ibc,
                    &c->cb_inm, &c->cb_iq};
    for (auto* b : cb) b->release();
  }
  for (int r = 0; r < CRGPU_MAX_PARTS; r++)
    if (c->peer_opened[r]) {
      cudaIpcCloseMemHandle(c->peer_buf[r]);
      cudaIpcCloseMemHandle(c->peer_cursor[r]);
    }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto e : c->copy_ev)
    if (e) cudaEventDestroy(e);
  if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
  if (c->xchg_ready) cudaEventDestroy(c->xchg_ready);
  if (c->xchg_done) cudaEventDestroy(c->xchg_done);
  if (c->xchg_buf) cudaFree(c->xchg_buf);
  if (c->xchg_cursor) cudaFree(c->xchg_cursor);
  cudaStreamDestroy(c->stream);
  delete c;
}

int crgpu_set_params(crgpu_ctx* c, double thr, double max_ee, int filter_umis) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  c->threshold = thr;
  c->max_expected_errors = max_ee;
  c->filter_umis = filter_umis;
  return CRGPU_OK;
}

int crgpu_set_target_filter(crgpu_ctx* c, const uint8_t* on_target, int32_t n, uint64_t min_reads) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (n < 0) return fail(CRGPU_E_INVALID, "n_features < 0");
  CU(cudaSetDevice(c->device));
  c->n_on_target = 0;
  c->target_min_reads = 0;
  if (!on_target || n == 0 || min_reads == 0) return CRGPU_OK;
  int rc;
  if ((rc = c->on_target.ensure((size_t)n))) return rc;
  CU(cudaMemcpyAsync(c->on_target.p, on_target, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // on_target is the caller's
  c->n_on_target = (uint32_t)n;
  c->target_min_reads = min_reads;
  return CRGPU_OK;
}

int crgpu_whitelist_add(crgpu_ctx* c, const uint8_t* seqs, uint64_t n, int L, const uint8_t* translated, int* out_id) {
  if (!c || !seqs || !out_id) return fail(CRGPU_E_INVALID, "NULL argument");
  if (L < 1 || L > 16) return fail(CRGPU_E_LIMIT, "whitelist sequence length must be 1..16");
  if (n == 0 || n >= CRGPU_NO_RANK) return fail(CRGPU_E_LIMIT, "whitelist size out of range");
  if (!c->content.empty() && L != c->L) return fail(CRGPU_E_LIMIT, "all whitelists must share one sequence length");
  CU(cudaSetDevice(c->device));
  std::vector<std::pair<uint32_t, uint32_t>> kv(n);  // raw key, content key
  for (uint64_t i = 0; i < n; i++) {
    uint32_t k, t;
    if (!pack_ascii(seqs + i * L, L, &k)) return fail(CRGPU_E_INVALID, "whitelist sequence with a base outside A,C,G,T");
    t = k;
    if (translated && !pack_ascii(translated + i * L, L, &t))
      return fail(CRGPU_E_INVALID, "translated sequence with a base outside A,C,G,T");
    kv[i] = {k, t};
  }
  std::sort(kv.begin(), kv.end());
  kv.erase(std::unique(kv.begin(), kv.end(), [](auto& a, auto& b) { return a.first == b.first; }), kv.end());
  const uint32_t W = (uint32_t)kv.size();
  if (c->content.empty()) {
    c->L = L;
    c->content.resize(W);
    for (uint32_t i = 0; i < W; i++) c->content[i] = kv[i].second;
    std::sort(c->content.begin(), c->content.end());
    c->content.erase(std::unique(c->content.begin(), c->content.end()), c->content.end());
  }
  // content ranks
  std::vector<uint32_t> rank(W);
  bool identity = true;
  for (uint32_t i = 0; i < W; i++) {
    auto it = std::lower_bound(c->content.begin(), c->content.end(), kv[i].second);
    if (it == c->content.end() || *it != kv[i].second)
      return fail(CRGPU_E_INVALID, "whitelist content outside the content space of the first whitelist");
    rank[i] = (uint32_t)(it - c->content.begin());
    identity = identity && rank[i] == i;
  }
  // bucket geometry: smallest even prefix p with <= 16 entries per bucket on average
  const int nbits = 2 * L;
  int p = 0;
  while (p < nbits - 2 && ((uint64_t)W >> p) > 16) p += 2;
  int s, g, n_ord;
  while (true) {
    s = nbits - p;
    g = s / 2;
    n_ord = (L + g - 1) / g;
    if (n_ord <= CRGPU_MAX_ORD || p == 0) break;
    p -= 2;
  }
  HostWhitelist* w = new HostWhitelist();
  w->L = L;
  w->W = W;
  w->is_trans = translated != nullptr;
  memset(&w->dev, 0, sizeof(w->dev));
  w->dev.L = L;
  w->dev.s = s;
  w->dev.n_ord = n_ord;
  w->dev.W = W;
  uint32_t covered = 0;
  const uint64_t n_buckets = 1ull << p;
  for (int o = 0; o < n_ord; o++) {
    int a = std::max(0, L - (o + 1) * g);  // leftmost base of this ordering's group
    int r = 2 * (L - a - g);
    uint32_t group = 0;
    for (int pos = a; pos < a + g; pos++) group |= 1u << pos;
    w->dev.rot[o] = r;
    w->dev.resp[o] = group & ~covered;
    covered |= group;
    std::vector<std::pair<uint32_t, uint32_t>> rk(W);
    for (uint32_t i = 0; i < W; i++) rk[i] = {rotr_host(kv[i].first, r, nbits), rank[i]};
    if (r != 0) std::sort(rk.begin(), rk.end());
    std::vector<uint32_t> keys(W + 4, 0xFFFFFFFFu), vals(W), offs(n_buckets + 1, 0);  // four sentinels
    for (uint32_t i = 0; i < W; i++) {
      keys[i] = rk[i].first;
      vals[i] = rk[i].second;
      uint64_t b = s >= 32 ? 0 : (rk[i].first >> s);
      offs[b + 1]++;
    }
    for (uint64_t b = 0; b < n_buckets; b++) offs[b + 1] += offs[b];
    DevBuf dk, dv, dof;
    int rc;
    if ((rc = dk.ensure((size_t)(W + 4) * 4))) return rc;
    if ((rc = dof.ensure((n_buckets + 1) * 4))) return rc;
    CU(cudaMemcpy(dk.p, keys.data(), (size_t)(W + 4) * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dof.p, offs.data(), (n_buckets + 1) * 4, cudaMemcpyHostToDevice));
    if (o == 0) {
      // exact-membership slots: about 6.5 entries per bucket on average, fourteen at most inline; the suffix
      // kept in a slot must fit 16 bits
      if ((uint64_t)W >= (1ull << 27)) return fail(CRGPU_E_INVALID, "whitelist too large for the slot table");
      int pe = 0;
      while (pe < nbits && (1ull << pe) < (uint64_t)W) pe++;  // ceil(log2 W)
      pe = pe > 3 ? pe - 3 : 0;  // 32 bytes per ~6.5 entries: the table stays small enough for L2
      if (pe < nbits - 16) pe = nbits - 16;
      if (pe > 26) pe = 26;
      if (pe > nbits) pe = nbits;
      const int sshift = nbits - pe;
      const uint64_t nb = 1ull << pe;
      constexpr uint32_t CAP = 14;
      std::vector<uint4> slots(2 * nb, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
      for (uint64_t b = 0; b < nb; b++) slots[2 * b].x = 0u;
      {
        const uint32_t smask = (sshift >= 32) ? 0xFFFFFFFFu : ((1u << sshift) - 1u);
        uint32_t i = 0;
        while (i < W) {
          const uint64_t b = sshift >= 32 ? 0 : (keys[i] >> sshift);
          uint32_t j = i;
          while (j < W && (sshift >= 32 ? 0 : (keys[j] >> sshift)) == b) j++;
          const uint32_t cnt = j - i;
          uint16_t suf[CAP];
          for (uint32_t k = 0; k < CAP; k++) suf[k] = k < cnt ? (uint16_t)(keys[i + k] & smask) : (uint16_t)0xFFFF;
          uint32_t wd[8];
          wd[0] = i | ((cnt <= CAP ? cnt : 31u) << 27);
          for (uint32_t k = 0; k < CAP / 2; k++) wd[1 + k] = suf[2 * k] | ((uint32_t)suf[2 * k + 1] << 16);
          slots[2 * b] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          slots[2 * b + 1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
          i = j;
        }
      }
      DevBuf de;
      if ((rc = de.ensure(2 * nb * sizeof(uint4)))) return rc;
      CU(cudaMemcpy(de.p, slots.data(), 2 * nb * sizeof(uint4), cudaMemcpyHostToDevice));
      w->dev.slots = de.as<uint4>();
      w->dev.slot_shift = sshift;
      w->bufs.push_back(de);
    }
    w->dev.keys[o] = dk.as<uint32_t>();
    w->dev.offs[o] = dof.as<uint32_t>();
    w->bufs.push_back(dk);
    w->bufs.push_back(dof);
    w->dev.sfx[o] = nullptr;
    if (s <= 16) {  // 16-bit suffix copy for the neighbour scans
      std::vector<uint16_t> sfx(W + 8, (uint16_t)0xFFFF);
      const uint32_t sm = s >= 32 ? 0xFFFFFFFFu : ((1u << s) - 1u);
      for (uint32_t i = 0; i < W; i++) sfx[i] = (uint16_t)(keys[i] & sm);
      DevBuf ds;
      if ((rc = ds.ensure((size_t)(W + 8) * 2))) return rc;
      CU(cudaMemcpy(ds.p, sfx.data(), (size_t)(W + 8) * 2, cudaMemcpyHostToDevice));
      w->dev.sfx[o] = ds.as<uint16_t>();
      w->bufs.push_back(ds);
    }
    if (o == 0 && identity) {
      w->dev.vals[o] = nullptr;
    } else {
      if ((rc = dv.ensure((size_t)W * 4))) return rc;
      CU(cudaMemcpy(dv.p, vals.data(), (size_t)W * 4, cudaMemcpyHostToDevice));
      w->dev.vals[o] = dv.as<uint32_t>();
      w->bufs.push_back(dv);
    }
  }
  c->wls.push_back(w);
  *out_id = (int)c->wls.size() - 1;
  return CRGPU_OK;
}

int crgpu_whitelist_size(crgpu_ctx* c, uint64_t* out_n, int* out_L) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (out_n) *out_n = c->content.size();
  if (out_L) *out_L = c->L;
  return CRGPU_OK;
}

int crgpu_barcode_seqs(crgpu_ctx* c, const uint32_t* ranks, uint64_t n, uint8_t* out) {
  if (!c || !ranks || !out) return fail(CRGPU_E_INVALID, "NULL argument");
  static const char B[4] = {'A', 'C', 'G', 'T'};
  for (uint64_t i = 0; i < n; i++) {
    if (ranks[i] >= c->content.size()) {
      memset(out + i * c->L, 'N', c->L);
      continue;
    }
    uint32_t k = c->content[ranks[i]];
    for (int p = 0; p < c->L; p++) out[i * c->L + p] = B[(k >> (2 * (c->L - 1 - p))) & 3u];
  }
  return CRGPU_OK;
}

// BarcodeCorrection::main, bc_counts_total (cr_lib/src/stages/barcode_correction.rs:327-362): every read the stage
// sees (the reads that were not valid before correction) is observed under its barcode after correction - a
// whitelist barcode when the correction succeeded, else the raw sequence with valid = false - and the entries seen
// fewer than min_reads_to_report_bc times are dropped. One histogram over all library types; the whole read set
// counts as one chunk (the reference's per-chunk `retain` makes its own output depend on the chunking).
int crgpu_total_barcode_counts(crgpu_ctx* c, uint64_t min_reads, uint64_t* out_n) {
  if (!c || !out_n) return fail(CRGPU_E_INVALID, "NULL argument");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  const int L = c->L;
  const uint64_t W = c->content.size();
  c->tbc_seqs.clear();
  c->tbc_valid.clear();
  c->tbc_counts.clear();
  int rc;
  // (1) still-invalid reads by raw sequence: collect, sort, count runs
  uint64_t n_side = 0;
  for (auto* b : c->batches) n_side += b->n;
  DevBuf keys, alt, temp, ctr_buf;
  if ((rc = keys.ensure(std::max<uint64_t>(n_side, 1) * 8))) return rc;
  if ((rc = alt.ensure(std::max<uint64_t>(n_side, 1) * 8))) return rc;
  if ((rc = ctr_buf.ensure(8))) return rc;
  auto release = [&]() { keys.release(); alt.release(); temp.release(); ctr_buf.release(); };
  CU(cudaMemsetAsync(ctr_buf.p, 0, 8, c->stream));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  for (size_t bi = 0; bi < c->batches.size(); bi++) {
    Batch* b = c->batches[bi];
    if (!b->n) continue;
    c->launches += launch_collect_invalid(b->inv_idx.as<uint32_t>(), b->inv_bc.as<uint32_t>(), b->inv_nmask.as<uint32_t>(),
                                          ctr + 2 + bi, b->n, c->bc_out.as<uint32_t>() + b->base,
                                          keys.as<unsigned long long>(), ctr_buf.as<unsigned long long>(), c->stream);
  }
  unsigned long long n_inv = 0;
  CU(cudaMemcpyAsync(&n_inv, ctr_buf.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<unsigned long long> sorted_h(n_inv);
  if (n_inv) {
    if ((rc = temp.ensure(sort_temp_bytes(n_inv)))) { release(); return rc; }
    unsigned long long* sorted = nullptr;
    c->launches += sort_keys(keys.as<unsigned long long>(), alt.as<unsigned long long>(), n_inv, 48, temp.p, temp.cap,
                             &sorted, c->stream, 0);
    CU(cudaMemcpyAsync(sorted_h.data(), sorted, n_inv * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  release();
  static const char B[4] = {'A', 'C', 'G', 'T'};
  struct Entry {
    std::string seq;
    uint8_t valid;
    uint64_t count;
  };
  std::vector<Entry> ent;
  for (uint64_t i = 0; i < n_inv;) {
    uint64_t j = i;
    while (j < n_inv && sorted_h[j] == sorted_h[i]) j++;
    if (j - i >= min_reads) {
      const uint32_t k = (uint32_t)sorted_h[i], nm = (uint32_t)(sorted_h[i] >> 32);
      std::string s((size_t)L, 'N');
      for (int p = 0; p < L; p++)
        if (!((nm >> p) & 1u)) s[p] = B[(k >> (2 * (L - 1 - p))) & 3u];
      ent.push_back(Entry{s, 0, j - i});
    }
    i = j;
  }
  std::sort(ent.begin(), ent.end(), [](const Entry& a, const Entry& b) { return a.seq < b.seq; });
  // (2) corrected reads by whitelist barcode, summed over the library types
  std::vector<uint64_t> corr(W, 0);
  std::vector<uint32_t> h(W);
  for (auto* l : c->libs) {
    if (!W) break;
    CU(cudaMemcpyAsync(h.data(), l->corrected.p, W * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint64_t r = 0; r < W; r++) corr[r] += h[r];
  }
  for (uint64_t r = 0; r < W; r++)  // content ranks ascend with the sequence
    if (corr[r] && corr[r] >= min_reads) {
      const uint32_t k = c->content[r];
      std::string s((size_t)L, 'A');
      for (int p = 0; p < L; p++) s[p] = B[(k >> (2 * (L - 1 - p))) & 3u];
      ent.push_back(Entry{s, 1, corr[r]});
    }
  for (const auto& e : ent) {
    c->tbc_seqs.insert(c->tbc_seqs.end(), e.seq.begin(), e.seq.end());
    c->tbc_valid.push_back(e.valid);
    c->tbc_counts.push_back(e.count);
  }
  *out_n = ent.size();
  return CRGPU_OK;
}

int crgpu_total_barcode_counts_get(crgpu_ctx* c, uint8_t* seqs, uint8_t* valid, uint64_t* counts) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (seqs && !c->tbc_seqs.empty()) memcpy(seqs, c->tbc_seqs.data(), c->tbc_seqs.size());
  if (valid && !c->tbc_valid.empty()) memcpy(valid, c->tbc_valid.data(), c->tbc_valid.size());
  if (counts && !c->tbc_counts.empty()) memcpy(counts, c->tbc_counts.data(), c->tbc_counts.size() * 8);
  return CRGPU_OK;
}

int crgpu_library_add(crgpu_ctx* c, const crgpu_library_def* def, int* out_lib) {
  if (!c || !def || !out_lib) return fail(CRGPU_E_INVALID, "NULL argument");
  if (def->whitelist < 0 || def->whitelist >= (int)c->wls.size()) return fail(CRGPU_E_INVALID, "unknown whitelist");
  if ((int)c->libs.size() >= CRGPU_MAX_LIBS) return fail(CRGPU_E_LIMIT, "too many library types");
  if (def->bc_length != c->wls[def->whitelist]->L) return fail(CRGPU_E_INVALID, "bc_length differs from the whitelist's");
  if (def->umi_length < 0 || def->umi_length > 15) return fail(CRGPU_E_LIMIT, "UMI length must be 0..15");
  if (c->layout_ready) return fail(CRGPU_E_INVALID, "libraries must be added before the first pass");
  CU(cudaSetDevice(c->device));
  Library* l = new Library();
  l->def = *def;
  int rc = alloc_library_tables(c, l);
  if (rc) {
    delete l;
    return rc;
  }
  c->libs.push_back(l);
  *out_lib = (int)c->libs.size() - 1;
  if (c->n_features) return build_fb_tables(c);
  return CRGPU_OK;
}

int crgpu_features_set(crgpu_ctx* c, int32_t n_features, const int32_t* feature_type, const uint8_t* fb_seqs,
                       int32_t fb_stride) {
  if (!c || n_features < 0) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->layout_ready) return fail(CRGPU_E_INVALID, "features must be set before the first pass");
  CU(cudaSetDevice(c->device));
  c->n_features = n_features;
  c->feature_type.assign(n_features, 0);
  if (feature_type) c->feature_type.assign(feature_type, feature_type + n_features);
  c->fb_stride = fb_stride;
  c->fb_seqs.clear();
  if (fb_seqs && fb_stride > 0) c->fb_seqs.assign(fb_seqs, fb_seqs + (size_t)n_features * fb_stride);
  int rc;
  if ((rc = c->d_fb_counts.ensure(std::max<size_t>(8, (size_t)n_features * 8)))) return rc;
  if ((rc = c->d_feat_dist.ensure(std::max<size_t>(8, (size_t)n_features * 8)))) return rc;
  return build_fb_tables(c);
}

int crgpu_reads_add(crgpu_ctx* c, int lib, const crgpu_read_batch* rb, int* out_batch) {
  if (!c || !rb) return fail(CRGPU_E_INVALID, "NULL argument");
  if (lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "unknown library");
  const crgpu_library_def& d = c->libs[lib]->def;
  if (rb->n >= (1ull << 31)) return fail(CRGPU_E_LIMIT, "a batch holds at most 2^31-1 reads");
  if (rb->n && (!rb->r1_seq || !rb->r1_qual)) return fail(CRGPU_E_INVALID, "r1_seq / r1_qual is NULL");
  if (rb->r1_len < d.bc_offset + d.bc_length || rb->r1_len < d.umi_offset + d.umi_length)
    return fail(CRGPU_E_INVALID, "r1_len shorter than the barcode / UMI ranges");
  if (d.is_feature_barcode && rb->n && (!rb->r2_seq || !rb->r2_qual))
    return fail(CRGPU_E_INVALID, "feature-barcode library needs r2_seq / r2_qual");
  if (!d.is_feature_barcode && rb->n && !rb->feature) return fail(CRGPU_E_INVALID, "feature array is NULL");
  CU(cudaSetDevice(c->device));
  Batch* b;
  if (!c->batch_pool.empty()) {
    b = c->batch_pool.back();
    c->batch_pool.pop_back();
  } else {
    b = new Batch();
  }
  b->lib = lib;
  b->n = rb->n;
  b->r1_len = rb->r1_len;
  b->r2_len = d.is_feature_barcode ? rb->r2_len : 0;
  b->on_device = rb->on_device != 0;
  b->host_pending = false;
  b->base = c->n_reads;
  b->n_invalid = 0;
  if (b->on_device) {
    b->r1_seq = rb->r1_seq;
    b->r1_qual = rb->r1_qual;
    b->feature = d.is_feature_barcode ? nullptr : rb->feature;
    b->r2_seq = d.is_feature_barcode ? rb->r2_seq : nullptr;
    b->r2_qual = d.is_feature_barcode ? rb->r2_qual : nullptr;
    b->select = reinterpret_cast<const unsigned long long*>(rb->select_key);
    b->h_select = nullptr;
  } else {
    int rc;
    size_t sb = (size_t)rb->n * rb->r1_len;
    // a batch that cannot be set up goes back to the pool (its buffers are reused, nothing leaks)
#define ENSURE_OR_RETURN(x)          \
  if ((rc = (x))) {                  \
    c->batch_pool.push_back(b);      \
    return rc;                       \
  }
    // device copies are allocated now; the bytes cross in crgpu_pass1, chunk by chunk under the kernels
    ENSURE_OR_RETURN(b->own_seq.ensure(sb + 16));
    ENSURE_OR_RETURN(b->own_qual.ensure(sb + 16));
    b->r1_seq = b->own_seq.as<uint8_t>();
    b->r1_qual = b->own_qual.as<uint8_t>();
    b->h_r1_seq = rb->r1_seq;
    b->h_r1_qual = rb->r1_qual;
    b->feature = nullptr;
    b->h_feature = nullptr;
    b->r2_seq = b->r2_qual = nullptr;
    b->h_r2_seq = b->h_r2_qual = nullptr;
    if (!d.is_feature_barcode) {
      ENSURE_OR_RETURN(b->own_feat.ensure((size_t)rb->n * 4 + 16));
      b->feature = b->own_feat.as<uint32_t>();
      b->h_feature = rb->feature;
    } else {
      size_t rb2 = (size_t)rb->n * rb->r2_len;
      ENSURE_OR_RETURN(b->own_r2s.ensure(rb2 + 16));
      ENSURE_OR_RETURN(b->own_r2q.ensure(rb2 + 16));
      b->r2_seq = b->own_r2s.as<uint8_t>();
      b->r2_qual = b->own_r2q.as<uint8_t>();
      b->h_r2_seq = rb->r2_seq;
      b->h_r2_qual = rb->r2_qual;
    }
    b->select = nullptr;
    b->h_select = nullptr;
    if (rb->select_key) {
      ENSURE_OR_RETURN(b->own_select.ensure((size_t)rb->n * 8 + 16));
      b->select = b->own_select.as<unsigned long long>();
      b->h_select = reinterpret_cast<const unsigned long long*>(rb->select_key);
    }
#undef ENSURE_OR_RETURN
    b->host_pending = rb->n > 0;
  }
  if (b->select) c->have_select = true;
  c->n_reads += rb->n;
  c->batches.push_back(b);
  c->stage = 0;
  if (out_batch) *out_batch = (int)c->batches.size() - 1;
  return CRGPU_OK;
}

int crgpu_fastq_extract(crgpu_ctx* c, const void* text, uint64_t n_bytes, int on_device, int read_len,
                        uint8_t* dev_seq, uint8_t* dev_qual, uint64_t capacity, uint64_t* n_records, uint64_t* n_short,
                        uint64_t* n_malformed) {
  if (!c || (!text && n_bytes) || !dev_seq || !dev_qual) return fail(CRGPU_E_INVALID, "NULL argument");
  if (read_len < 1 || read_len > 4096) return fail(CRGPU_E_INVALID, "read_len out of range");
  CU(cudaSetDevice(c->device));
  int rc;
  const uint8_t* d_text = static_cast<const uint8_t*>(text);
  if (!on_device) {
    if ((rc = c->fastq_text.ensure(n_bytes + 16))) return rc;
    CU(cudaMemcpyAsync(c->fastq_text.p, text, n_bytes, cudaMemcpyHostToDevice, c->stream));
    d_text = c->fastq_text.as<uint8_t>();
  } else if ((uintptr_t)text % 16 != 0) {
    return fail(CRGPU_E_INVALID, "device FASTQ text must be 16-byte aligned");
  }
  const size_t tb = fastq_temp_bytes(n_bytes);
  if ((rc = c->fastq_tmp.ensure(tb + 64))) return rc;
  unsigned long long* counters = reinterpret_cast<unsigned long long*>(c->fastq_tmp.as<unsigned char>() + ((tb + 7) & ~(size_t)7));
  c->launches += launch_fastq_extract(d_text, n_bytes, read_len, dev_seq, dev_qual, capacity, c->fastq_tmp.p, counters,
                                      c->stream);
  CHECK_KERNEL();
  unsigned long long h[3] = {0, 0, 0};
  uint8_t last = '\n';
  unsigned int lb_flag = 0u;
  CU(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  if (n_bytes) CU(cudaMemcpyAsync(&last, d_text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
  fastq_lb_flag_fetch(&lb_flag, c->stream);
  CU(cudaStreamSynchronize(c->stream));
  if (lb_flag) {
    fastq_lb_flag_clear(c->stream);
    return fail(CRGPU_E_CUDA, "chained-scan watchdog fired in crgpu_fastq_extract (see crgpu_count)");
  }
  const uint64_t lines = h[0] + (n_bytes && last != '\n' ? 1 : 0);
  if (lines % 4 != 0) return fail(CRGPU_E_INVALID, "FASTQ text does not hold a whole number of 4-line records");
  if (lines / 4 > capacity) return fail(CRGPU_E_LIMIT, "more FASTQ records than the output arrays hold");
  if (n_records) *n_records = lines / 4;
  if (n_short) *n_short = h[1];
  if (n_malformed) *n_malformed = h[2];
  return CRGPU_OK;
}

// error reporting for the host-only translation units of the library (mex_writer.cpp)
extern "C" int crgpu_set_error_(int code, const char* msg) { return fail(code, msg ? msg : ""); }

int crgpu_reads_clear(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto* b : c->batches) c->batch_pool.push_back(b);
  c->batches.clear();
  c->n_reads = 0;
  c->stage = 0;
  c->annotated = false;
  c->have_select = false;
  return CRGPU_OK;
}

int crgpu_pass1(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_layout(c))) return rc;
  if (c->n_reads >= (1ull << 32)) return fail(CRGPU_E_LIMIT, "a context holds at most 2^32-1 reads");
  if (c->batches.size() > 1000) return fail(CRGPU_E_LIMIT, "too many batches");
  for (auto* b : c->batches)
    if (!c->libs[b->lib]->def.is_feature_barcode && b->n && c->n_features <= 0)
      return fail(CRGPU_E_INVALID, "crgpu_features_set must be called before crgpu_pass1: gene-expression reads carry "
                                   "feature indices and the matrix has no rows yet");
  phases_clear(c, "pass1");
  if ((rc = phase_begin(c, "pass1"))) return rc;
  const size_t N = c->n_reads;
  if ((rc = c->bc_out.ensure(N * 4 + 16))) return rc;
  if ((rc = c->umi_out.ensure(N * 4 + 16))) return rc;
  if ((rc = c->keys.ensure(N * 8 + 16))) return rc;
  if ((rc = c->keys_alt.ensure(N * 8 + 16))) return rc;
  CU(cudaMemsetAsync(c->counters.p, 0, 8 * 1024, c->stream));
  for (auto* l : c->libs) {
    size_t bytes = c->content.size() * 4;
    CU(cudaMemsetAsync(l->prior.p, 0, bytes, c->stream));
    CU(cudaMemsetAsync(l->corrected.p, 0, bytes, c->stream));
  }
  if (c->n_features) CU(cudaMemsetAsync(c->d_fb_counts.p, 0, (size_t)c->n_features * 8, c->stream));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  for (size_t bi = 0; bi < c->batches.size(); bi++) {
    Batch* b = c->batches[bi];
    Library* l = c->libs[b->lib];
    if ((rc = b->inv_idx.ensure(b->n * 4 + 16))) return rc;
    if ((rc = b->inv_bc.ensure(b->n * 4 + 16))) return rc;
    if ((rc = b->inv_nmask.ensure(b->n * 4 + 16))) return rc;
    if ((rc = b->inv_qual.ensure(b->n * 16 + 16))) return rc;
    Pass1Args a;
    memset(&a, 0, sizeof(a));
    a.n = b->n;
    a.r1_len = b->r1_len;
    a.seq = b->r1_seq;
    a.qual = b->r1_qual;
    a.feature = l->def.is_feature_barcode ? nullptr : b->feature;
    a.bc_off = l->def.bc_offset;
    a.bc_len = l->def.bc_length;
    a.umi_off = l->def.umi_offset;
    a.umi_len = l->def.umi_length;
    a.wl = c->wls[l->def.whitelist]->dev;
    a.prior = l->prior.as<uint32_t>();
    a.bc_out = c->bc_out.as<uint32_t>() + b->base;
    a.umi_out = c->umi_out.as<uint32_t>() + b->base;
    a.keys = c->keys.as<unsigned long long>();
    a.counters = ctr;  // packed scratch
    a.inv_idx = b->inv_idx.as<uint32_t>();
    a.inv_bc = b->inv_bc.as<uint32_t>();
    a.inv_nmask = b->inv_nmask.as<uint32_t>();
    a.inv_qual = b->inv_qual.as<uint4>();
    a.kl = c->kl;
    a.lib = (uint32_t)b->lib;
    a.emit_keys = l->def.is_feature_barcode ? 0 : 1;
    a.have_qual = 1;
    a.n_features = (uint32_t)c->n_features;
    a.bad_feature = ctr + CTR_BAD_FEATURE;
    launch_merge_counter(ctr, ctr + 1, c->stream);
    c->launches += 1;
    if (!b->host_pending) {
      c->launches += launch_pass1(a, c->n_sms, c->stream);
    } else {
      // A host batch crosses PCIe here, in chunks on the copy stream; the pass-1 kernel of a chunk waits for its
      // event only, so all but the last chunk's kernel run under the copies that follow. The copy of the whole
      // batch is what bounds an end-to-end step (12 GB at ~55 GB/s for 200 M reads); the kernels hide behind it.
      if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming));
      }
      uint64_t chunk = getenv("CRGPU_H2D_CHUNK") ? strtoull(getenv("CRGPU_H2D_CHUNK"), nullptr, 10) : (4ull << 20);
      chunk = std::max<uint64_t>(4096, chunk & ~4095ull);  // keeps every chunk 16-byte aligned for any read length
      // the device buffers were last read by kernels of the context stream (an earlier step): order the copies
      // behind them
      CU(cudaEventRecord(c->copy_ev[0], c->stream));
      CU(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[0], 0));
      const bool is_fb = l->def.is_feature_barcode != 0;
      if (is_fb) {
        const size_t rb2 = (size_t)b->n * b->r2_len;
        CU(cudaMemcpyAsync(b->own_r2s.p, b->h_r2_seq, rb2, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaMemcpyAsync(b->own_r2q.p, b->h_r2_qual, rb2, cudaMemcpyHostToDevice, c->copy_stream));
      }
      int k = 0;
      for (uint64_t first = 0; first < b->n; first += chunk, k++) {
        const uint64_t cn = std::min<uint64_t>(chunk, b->n - first);
        const size_t off = (size_t)first * b->r1_len, bytes = (size_t)cn * b->r1_len;
        CU(cudaMemcpyAsync(b->own_seq.as<uint8_t>() + off, b->h_r1_seq + off, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaMemcpyAsync(b->own_qual.as<uint8_t>() + off, b->h_r1_qual + off, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        if (!is_fb)
          CU(cudaMemcpyAsync(b->own_feat.as<uint32_t>() + first, b->h_feature + first, (size_t)cn * 4,
                             cudaMemcpyHostToDevice, c->copy_stream));
        if (b->h_select)
          CU(cudaMemcpyAsync(b->own_select.as<unsigned long long>() + first, b->h_select + first, (size_t)cn * 8,
                             cudaMemcpyHostToDevice, c->copy_stream));
        cudaEvent_t ev = c->copy_ev[k & 1];
        CU(cudaEventRecord(ev, c->copy_stream));
        CU(cudaStreamWaitEvent(c->stream, ev, 0));
        Pass1Args ac = a;
        ac.n = cn;
        ac.idx_base = first;
        ac.seq = a.seq + off;
        ac.qual = a.qual + off;
        ac.feature = a.feature ? a.feature + first : nullptr;
        ac.bc_out = a.bc_out + first;
        ac.umi_out = a.umi_out + first;
        c->launches += launch_pass1(ac, c->n_sms, c->stream);
        CHECK_KERNEL();
      }
      // the caller's buffers have been read once this returns (the contract of crgpu_reads_add)
      CU(cudaStreamSynchronize(c->copy_stream));
      b->host_pending = false;
    }
    launch_split_counter(ctr, ctr + 1, ctr + 2 + bi, c->stream);
    c->launches += 1;
    CHECK_KERNEL();
    if (l->def.is_feature_barcode) {
      FbArgs f;
      memset(&f, 0, sizeof(f));
      f.n = b->n;
      f.r2_len = b->r2_len;
      f.fb_off = l->def.fb_offset;
      f.fb_len = l->def.fb_length;
      f.r2_seq = b->r2_seq;
      f.r2_qual = b->r2_qual;
      f.fb_keys = l->d_fb_keys.as<uint32_t>();
      f.fb_index = l->d_fb_index.as<uint32_t>();
      f.n_fb = (int)l->fb_keys.size();
      f.exact_counts = c->d_fb_counts.as<unsigned long long>();
      f.threshold = 0.975;
      c->launches += launch_fb(f, c->stream);
      CHECK_KERNEL();
    }
  }
  // what pass 1 leaves behind, so that crgpu_pass2 can be repeated
  CU(cudaMemcpyAsync(ctr + CTR_KEYS_PASS1, ctr + 1, 8, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = phase_end(c))) return rc;
  c->stage = 1;
  c->keys_external = false;
  c->key_src = nullptr;
  c->annotated = false;
  return CRGPU_OK;
}

int crgpu_prior_dev(crgpu_ctx* c, int lib, uint32_t** out, uint64_t* n) {
  if (!c || lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (out) *out = c->libs[lib]->prior.as<uint32_t>();
  if (n) *n = c->content.size();
  return CRGPU_OK;
}

int crgpu_prior_set(crgpu_ctx* c, int lib, const uint32_t* host, uint64_t n) {
  if (!c || lib < 0 || lib >= (int)c->libs.size() || !host) return fail(CRGPU_E_INVALID, "bad argument");
  if (n != c->content.size()) return fail(CRGPU_E_INVALID, "prior length differs from the content size");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->libs[lib]->prior.p, host, n * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}

int crgpu_fb_counts_dev(crgpu_ctx* c, unsigned long long** out, int32_t* n) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (out) *out = c->d_fb_counts.as<unsigned long long>();
  if (n) *n = c->n_features;
  return CRGPU_OK;
}

int crgpu_pass2(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 1) return fail(CRGPU_E_INVALID, "crgpu_pass1 must run first");
  if (c->stage >= 3)
    return fail(CRGPU_E_INVALID, "crgpu_pass2 after crgpu_count needs crgpu_pass1 first: the count stage sorts the key "
                                 "buffer in place, the keys pass 1 emitted are gone");
  CU(cudaSetDevice(c->device));
  int rc;
  phases_clear(c, "pass2");
  if ((rc = phase_begin(c, "pass2"))) return rc;
  const size_t nb = c->batches.size();
  if (c->stage >= 2) {
    // a repeated pass 2 (a retry after a failed collective, or a second call by mistake) starts from the state
    // pass 1 left: no corrected reads yet, the key list ends where pass 1 ended
    for (auto* l : c->libs) CU(cudaMemsetAsync(l->corrected.p, 0, c->content.size() * 4, c->stream));
    CU(cudaMemcpyAsync(c->counters.as<unsigned long long>() + 1, c->counters.as<unsigned long long>() + CTR_KEYS_PASS1, 8,
                       cudaMemcpyDeviceToDevice, c->stream));
  }
  // The per-batch invalid counts stay on the device (the kernels read them there); only a feature-barcode
  // library needs a host round trip here, for compute_feature_dist over its exact counts.
  if (c->have_fb && c->n_features) {
    std::vector<unsigned long long> fbc(c->n_features);
    CU(cudaMemcpyAsync(fbc.data(), c->d_fb_counts.p, (size_t)c->n_features * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<double> dist = feature_dist_host(fbc, c->feature_type);
    CU(cudaMemcpyAsync(c->d_feat_dist.p, dist.data(), dist.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // dist is a local
  }
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  for (size_t bi = 0; bi < nb; bi++) {
    Batch* b = c->batches[bi];
    Library* l = c->libs[b->lib];
    const bool is_fb = l->def.is_feature_barcode != 0;
    if (is_fb) {
      if ((rc = b->feature_res.ensure(b->n * 4 + 16))) return rc;
      FbArgs f;
      memset(&f, 0, sizeof(f));
      f.n = b->n;
      f.r2_len = b->r2_len;
      f.fb_off = l->def.fb_offset;
      f.fb_len = l->def.fb_length;
      f.r2_seq = b->r2_seq;
      f.r2_qual = b->r2_qual;
      f.fb_keys = l->d_fb_keys.as<uint32_t>();
      f.fb_index = l->d_fb_index.as<uint32_t>();
      f.n_fb = (int)l->fb_keys.size();
      f.feat_dist = c->d_feat_dist.as<double>();
      f.feature_out = b->feature_res.as<uint32_t>();
      f.threshold = 0.975;  // FEATURE_CONF_THRESHOLD, feature_extraction.rs:21
      c->launches += launch_fb(f, c->stream);
      CHECK_KERNEL();
    }
    Pass2Args a;
    memset(&a, 0, sizeof(a));
    a.n_invalid = b->n;  // upper bound: the grid is sized for it, the count itself is read on the device
    a.n_invalid_dev = ctr + 2 + bi;
    a.n_invalid_dev_shift = 0;
    a.inv_idx = b->inv_idx.as<uint32_t>();
    a.inv_bc = b->inv_bc.as<uint32_t>();
    a.inv_nmask = b->inv_nmask.as<uint32_t>();
    a.inv_qual = b->inv_qual.as<uint4>();
    a.wl = c->wls[l->def.whitelist]->dev;
    a.prior = l->prior.as<uint32_t>();
    a.corrected = l->corrected.as<uint32_t>();
    a.bc_out = c->bc_out.as<uint32_t>() + b->base;
    a.umi_out = c->umi_out.as<uint32_t>() + b->base;
    a.feature = is_fb ? nullptr : b->feature;
    a.keys = c->keys.as<unsigned long long>();
    a.counters = ctr + 1;
    a.kl = c->kl;
    a.lib = (uint32_t)b->lib;
    a.emit_keys = is_fb ? 0 : 1;
    a.have_qual = 1;
    a.threshold = c->threshold;
    a.max_expected_errors = c->max_expected_errors;
    a.check_expected_errors = c->max_expected_errors < 1.7976931348623157e308 ? 1 : 0;
    a.n_features = (uint32_t)c->n_features;
    c->launches += launch_pass2(a, c->stream);
    CHECK_KERNEL();
    if (is_fb) {
      EmitArgs e;
      memset(&e, 0, sizeof(e));
      e.n = b->n;
      e.bc_out = c->bc_out.as<uint32_t>() + b->base;
      e.umi_out = c->umi_out.as<uint32_t>() + b->base;
      e.feature = b->feature_res.as<uint32_t>();
      e.keys = c->keys.as<unsigned long long>();
      e.counters = ctr + 1;
      e.kl = c->kl;
      e.lib = (uint32_t)b->lib;
      c->launches += launch_emit_keys(e, c->stream);
      CHECK_KERNEL();
    }
  }
  for (auto* l : c->libs) {
    c->launches += launch_valid_counts(l->prior.as<uint32_t>(), l->corrected.as<uint32_t>(), l->valid.as<uint32_t>(),
                                       c->content.size(), c->stream);
    CHECK_KERNEL();
  }
  if ((rc = phase_end(c))) return rc;
  c->stage = 2;
  c->keys_external = false;
  return CRGPU_OK;
}

int crgpu_correct_barcodes(crgpu_ctx* c, int lib, const uint8_t* bc_ascii, const uint8_t* qual, uint64_t n,
                           uint32_t* out_rank, uint8_t* out_state) {
  if (!c || !bc_ascii || lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (n >= (1ull << 31)) return fail(CRGPU_E_LIMIT, "too many segments in one call");
  if (n == 0) return CRGPU_OK;
  CU(cudaSetDevice(c->device));
  Library* l = c->libs[lib];
  const int L = l->def.bc_length;
  // grow-only scratch kept in the context: no allocation per call once it has reached its size
  DevBuf &d_seq = c->cb_seq, &d_qual = c->cb_qual, &d_bc = c->cb_bc, &d_umi = c->cb_umi, &d_keys = c->cb_keys,
         &d_ctr = c->cb_ctr, &i_idx = c->cb_idx, &i_bc = c->cb_ibc, &i_nm = c->cb_inm, &i_q = c->cb_iq;
  int rc = 0;
#define TRY(x) \
  if ((rc = (x))) return rc;
#define CUX(call) CU(call)
  TRY(d_seq.ensure(n * L + 16));
  TRY(d_qual.ensure(n * L + 16));
  TRY(d_bc.ensure(n * 4));
  TRY(d_umi.ensure(n * 4));
  TRY(d_keys.ensure(16));
  TRY(d_ctr.ensure(32));
  TRY(i_idx.ensure(n * 4));
  TRY(i_bc.ensure(n * 4));
  TRY(i_nm.ensure(n * 4));
  TRY(i_q.ensure(n * 16));
  CUX(cudaMemcpyAsync(d_seq.p, bc_ascii, n * L, cudaMemcpyHostToDevice, c->stream));
  if (qual) CUX(cudaMemcpyAsync(d_qual.p, qual, n * L, cudaMemcpyHostToDevice, c->stream));
  CUX(cudaMemsetAsync(d_ctr.p, 0, 32, c->stream));
  KeyLayout kl{};
  Pass1Args a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  a.r1_len = L;
  a.seq = d_seq.as<uint8_t>();
  a.qual = d_qual.as<uint8_t>();
  a.bc_off = 0;
  a.bc_len = L;
  a.umi_off = 0;
  a.umi_len = 0;
  a.wl = c->wls[l->def.whitelist]->dev;
  a.prior = nullptr;
  a.bc_out = d_bc.as<uint32_t>();
  a.umi_out = d_umi.as<uint32_t>();
  a.keys = d_keys.as<unsigned long long>();
  a.counters = d_ctr.as<unsigned long long>();
  a.inv_idx = i_idx.as<uint32_t>();
  a.inv_bc = i_bc.as<uint32_t>();
  a.inv_nmask = i_nm.as<uint32_t>();
  a.inv_qual = i_q.as<uint4>();
  a.kl = kl;
  a.emit_keys = 0;
  a.have_qual = qual ? 1 : 0;
  c->launches += launch_pass1(a, c->n_sms, c->stream);
  // pass 2 takes its entry count from the device (the grid covers all n segments; the blocks past the
  // invalid count exit at once), so the two kernels run back to back without a host round trip
  Pass2Args b2;
  memset(&b2, 0, sizeof(b2));
  b2.n_invalid = n;
  b2.n_invalid_dev = d_ctr.as<unsigned long long>();
  b2.n_invalid_dev_shift = 32;
  b2.inv_idx = i_idx.as<uint32_t>();
  b2.inv_bc = i_bc.as<uint32_t>();
  b2.inv_nmask = i_nm.as<uint32_t>();
  b2.inv_qual = i_q.as<uint4>();
  b2.wl = a.wl;
  b2.prior = l->prior.as<uint32_t>();
  b2.corrected = nullptr;
  b2.bc_out = d_bc.as<uint32_t>();
  b2.umi_out = d_umi.as<uint32_t>();
  b2.keys = d_keys.as<unsigned long long>();
  b2.counters = d_ctr.as<unsigned long long>() + 1;
  b2.kl = kl;
  b2.emit_keys = 0;
  b2.have_qual = qual ? 1 : 0;
  b2.threshold = c->threshold;
  b2.max_expected_errors = c->max_expected_errors;
  b2.check_expected_errors = c->max_expected_errors < 1.7976931348623157e308 ? 1 : 0;
  c->launches += launch_pass2(b2, c->stream);
  c->cb_host.resize(n);
  std::vector<uint32_t>& h = c->cb_host;
  CUX(cudaMemcpyAsync(h.data(), d_bc.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUX(cudaStreamSynchronize(c->stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRGPU_E_CUDA, std::string("correct_barcodes: ") + cudaGetErrorString(e));
  for (uint64_t i = 0; i < n; i++) {
    uint32_t st = h[i] >> BC_STATE_SHIFT;
    if (out_state) out_state[i] = (uint8_t)st;
    if (out_rank) out_rank[i] = (st == ST_VALID_BEFORE || st == ST_VALID_AFTER) ? (h[i] & BC_RANK_MASK) : CRGPU_NO_RANK;
  }
#undef TRY
#undef CUX
  return CRGPU_OK;
}

int crgpu_key_layout(crgpu_ctx* c, int32_t* rank_shift, int32_t* feature_shift, int32_t* lib_shift, int32_t* umi_bits) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  int rc;
  if ((rc = ensure_layout(c))) return rc;
  if (rank_shift) *rank_shift = c->kl.rank_shift;
  if (feature_shift) *feature_shift = c->kl.feature_shift;
  if (lib_shift) *lib_shift = c->kl.lib_shift;
  if (umi_bits) *umi_bits = c->kl.umi_bits;
  return CRGPU_OK;
}

// the key count of this context's own reads (one host round trip); reports the reads pass 1 found with a
// feature index outside the matrix
static int fetch_n_keys(crgpu_ctx* c) {
  if (c->keys_external) return CRGPU_OK;
  unsigned long long h = 0, bad = 0;
  CU(cudaMemcpyAsync(&h, c->counters.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&bad, c->counters.as<unsigned long long>() + CTR_BAD_FEATURE, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->n_keys = h;
  if (bad)
    return fail(CRGPU_E_INVALID, std::to_string(bad) + " reads carry a feature index >= n_features (" +
                                     std::to_string(c->n_features) + "): the matrix has no such row");
  return CRGPU_OK;
}

int crgpu_keys_dev(crgpu_ctx* c, unsigned long long** out, uint64_t* n) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = fetch_n_keys(c))) return rc;
  if (out) *out = c->keys.as<unsigned long long>();
  if (n) *n = c->n_keys;
  return CRGPU_OK;
}

int crgpu_keys_partition(crgpu_ctx* c, int32_t n_parts, const uint32_t* bounds, uint64_t* out_counts) {
  if (!c || n_parts < 1 || !bounds || !out_counts) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = fetch_n_keys(c))) return rc;
  if (n_parts > CRGPU_MAX_PARTS) return fail(CRGPU_E_LIMIT, "too many parts");
  for (int p = 0; p < n_parts; p++)
    if (bounds[p] > bounds[p + 1]) return fail(CRGPU_E_INVALID, "bounds must be non-decreasing");
  c->launches += run_owner_partition(c->keys.as<unsigned long long>(), c->n_keys, c->kl.rank_shift, bounds, n_parts,
                                     c->keys_alt.as<unsigned long long>(), c->scalars.as<unsigned long long>() + 16,
                                     out_counts, c->stream);
  CHECK_KERNEL();
  std::swap(c->keys, c->keys_alt);
  return CRGPU_OK;
}

int crgpu_keys_set(crgpu_ctx* c, const unsigned long long* dev_keys, uint64_t n) {
  if (!c || (!dev_keys && n)) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_layout(c))) return rc;
  if ((rc = c->keys.ensure(n * 8 + 16))) return rc;
  if ((rc = c->keys_alt.ensure(n * 8 + 16))) return rc;
  if (n) CU(cudaMemcpyAsync(c->keys.p, dev_keys, n * 8, cudaMemcpyDeviceToDevice, c->stream));
  c->n_keys = n;
  c->key_src = nullptr;
  c->keys_external = true;
  if (c->stage < 2) c->stage = 2;
  return CRGPU_OK;
}

int crgpu_valid_counts_dev(crgpu_ctx* c, int lib, uint32_t** out, uint64_t* n) {
  if (!c || lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (out) *out = c->libs[lib]->valid.as<uint32_t>();
  if (n) *n = c->content.size();
  return CRGPU_OK;
}

int crgpu_corrected_dev(crgpu_ctx* c, int lib, uint32_t** out, uint64_t* n) {
  if (!c || lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (out) *out = c->libs[lib]->corrected.as<uint32_t>();
  if (n) *n = c->content.size();
  return CRGPU_OK;
}

int crgpu_valid_counts_refresh(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  for (auto* l : c->libs) {
    c->launches += launch_valid_counts(l->prior.as<uint32_t>(), l->corrected.as<uint32_t>(), l->valid.as<uint32_t>(),
                                       c->content.size(), c->stream);
    CHECK_KERNEL();
  }
  return CRGPU_OK;
}

int crgpu_exchange_init(crgpu_ctx* c, uint64_t capacity_keys, void* out_handle) {
  if (!c || !out_handle || capacity_keys == 0) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->xchg_buf) return fail(CRGPU_E_INVALID, "exchange already initialised");
  CU(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(&c->xchg_buf, capacity_keys * 8);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("exchange buffer: ") + cudaGetErrorString(e));
  e = cudaMalloc(&c->xchg_cursor, 16);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("exchange cursor: ") + cudaGetErrorString(e));
  CU(cudaMemset(c->xchg_cursor, 0, 16));
  c->xchg_capacity = capacity_keys;
  cudaIpcMemHandle_t h[2];
  CU(cudaIpcGetMemHandle(&h[0], c->xchg_buf));
  CU(cudaIpcGetMemHandle(&h[1], c->xchg_cursor));
  static_assert(sizeof(h) == CRGPU_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(out_handle, h, sizeof(h));
  return CRGPU_OK;
}

int crgpu_exchange_connect(crgpu_ctx* c, int32_t n_ranks, int32_t my_rank, const void* handles) {
  if (!c || !handles || n_ranks < 1 || n_ranks > CRGPU_MAX_PARTS || my_rank < 0 || my_rank >= n_ranks)
    return fail(CRGPU_E_INVALID, "bad argument");
  if (!c->xchg_buf) return fail(CRGPU_E_INVALID, "crgpu_exchange_init must run first");
  CU(cudaSetDevice(c->device));
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int r = 0; r < n_ranks; r++) {
    if (r == my_rank) {
      c->peer_buf[r] = static_cast<unsigned long long*>(c->xchg_buf);
      c->peer_cursor[r] = static_cast<unsigned long long*>(c->xchg_cursor);
      continue;
    }
    void *pb = nullptr, *pc = nullptr;
    CU(cudaIpcOpenMemHandle(&pb, h[2 * r], cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&pc, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
    c->peer_buf[r] = static_cast<unsigned long long*>(pb);
    c->peer_cursor[r] = static_cast<unsigned long long*>(pc);
    c->peer_opened[r] = true;
  }
  c->xchg_ranks = n_ranks;
  c->xchg_rank = my_rank;
  return CRGPU_OK;
}

int crgpu_exchange_reset(crgpu_ctx* c) {
  if (!c || !c->xchg_cursor) return fail(CRGPU_E_INVALID, "exchange not initialised");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->xchg_cursor, 0, 16, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->xchg_early = false;
  return CRGPU_OK;
}

int crgpu_keys_scatter_peers_begin(crgpu_ctx* c, int32_t n_parts, const uint32_t* bounds) {
  if (!c || !bounds || n_parts != c->xchg_ranks) return fail(CRGPU_E_INVALID, "bad argument / exchange not connected");
  if (c->stage < 1) return fail(CRGPU_E_INVALID, "crgpu_pass1 must run first");
  if (c->keys_external) return fail(CRGPU_E_INVALID, "the keys of this context were replaced");
  CU(cudaSetDevice(c->device));
  for (int p = 0; p < n_parts; p++)
    if (bounds[p] > bounds[p + 1]) return fail(CRGPU_E_INVALID, "bounds must be non-decreasing");
  if (!c->xchg_stream) {
    CU(cudaStreamCreateWithFlags(&c->xchg_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->xchg_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->xchg_done, cudaEventDisableTiming));
  }
  // the keys emitted so far (pass 1 is complete on the context stream once the counter has been read)
  unsigned long long n1 = 0;
  CU(cudaMemcpyAsync(&n1, c->counters.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventRecord(c->xchg_ready, c->stream));
  CU(cudaStreamWaitEvent(c->xchg_stream, c->xchg_ready, 0));
  unsigned long long* d_sent = c->scalars.as<unsigned long long>() + 32;
  c->launches += run_owner_scatter_peers(c->keys.as<unsigned long long>(), n1, c->kl.rank_shift, bounds, n_parts,
                                         c->peer_buf, c->peer_cursor, c->xchg_capacity, d_sent, c->xchg_stream);
  CHECK_KERNEL();
  CU(cudaEventRecord(c->xchg_done, c->xchg_stream));
  c->xchg_early = true;
  c->xchg_early_keys = n1;
  memcpy(c->xchg_early_bounds, bounds, (size_t)(n_parts + 1) * 4);
  return CRGPU_OK;
}

int crgpu_keys_scatter_peers(crgpu_ctx* c, int32_t n_parts, const uint32_t* bounds, uint64_t* out_sent) {
  if (!c || !bounds || n_parts != c->xchg_ranks) return fail(CRGPU_E_INVALID, "bad argument / exchange not connected");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = fetch_n_keys(c))) return rc;
  for (int p = 0; p < n_parts; p++)
    if (bounds[p] > bounds[p + 1]) return fail(CRGPU_E_INVALID, "bounds must be non-decreasing");
  uint64_t first = 0;
  if (c->xchg_early) {
    if (memcmp(c->xchg_early_bounds, bounds, (size_t)(n_parts + 1) * 4) != 0)
      return fail(CRGPU_E_INVALID, "bounds differ from the ones given to crgpu_keys_scatter_peers_begin");
    first = c->xchg_early_keys;  // already on their way
    if (first > c->n_keys) return fail(CRGPU_E_INVALID, "internal: early key count exceeds the key count");
  }
  unsigned long long* d_sent = c->scalars.as<unsigned long long>() + 16;
  c->launches += run_owner_scatter_peers(c->keys.as<unsigned long long>() + first, c->n_keys - first, c->kl.rank_shift,
                                         bounds, n_parts, c->peer_buf, c->peer_cursor, c->xchg_capacity, d_sent,
                                         c->stream);
  CHECK_KERNEL();
  unsigned long long h[CRGPU_MAX_PARTS] = {0}, h_early[CRGPU_MAX_PARTS] = {0};
  if (c->xchg_early) {
    CU(cudaStreamWaitEvent(c->stream, c->xchg_done, 0));
    CU(cudaMemcpyAsync(h_early, c->scalars.as<unsigned long long>() + 32, n_parts * 8, cudaMemcpyDeviceToHost, c->stream));
    c->xchg_early = false;
  }
  CU(cudaMemcpyAsync(h, d_sent, n_parts * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // this rank's peer stores (both parts) are complete
  if (out_sent)
    for (int p = 0; p < n_parts; p++) out_sent[p] = h[p] + h_early[p];
  return CRGPU_OK;
}

int crgpu_exchange_finish(crgpu_ctx* c, uint64_t* out_received) {
  if (!c || !c->xchg_cursor) return fail(CRGPU_E_INVALID, "exchange not initialised");
  CU(cudaSetDevice(c->device));
  unsigned long long h[2] = {0, 0};
  CU(cudaMemcpyAsync(h, c->xchg_cursor, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (h[1] || h[0] > c->xchg_capacity)
    return fail(CRGPU_E_LIMIT, "exchange buffer overflow: " + std::to_string(h[0]) + " keys for a capacity of " +
                                   std::to_string(c->xchg_capacity));
  int rc;
  if ((rc = ensure_layout(c))) return rc;
  if ((rc = c->keys_alt.ensure(h[0] * 8 + 16))) return rc;
  // the receive buffer itself is the sort input (and one of its two ping-pong buffers): no copy. Peers
  // write into it again only after the next step's collectives, i.e. after this context's count.
  c->key_src = static_cast<unsigned long long*>(c->xchg_buf);
  c->n_keys = h[0];
  c->keys_external = true;
  if (c->stage < 2) c->stage = 2;
  if (out_received) *out_received = h[0];
  return CRGPU_OK;
}

int crgpu_set_owned_range(crgpu_ctx* c, uint32_t lo, uint32_t hi) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  c->own_lo = lo;
  c->own_hi = hi;
  return CRGPU_OK;
}

int crgpu_count(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = fetch_n_keys(c))) return rc;
  phases_clear(c, "count");
  const uint64_t nk = c->n_keys;
  const uint64_t cap = std::max<uint64_t>(nk, 1);
  if ((rc = c->sort_temp.ensure(sort_temp_bytes(cap)))) return rc;
  // Default: the radix sort covers every key bit (8 passes for a 62-bit key) and rle_kernel encodes the runs.
  // CRGPU_FINISH_SORT=1 sorts only the bits above the UMI (5 passes) and leaves the UMI bits to the per-segment
  // shared-memory kernels of finish_kernels.cu, fused with the run-length encoding: bit-exact, but measured slower
  // on B200 (3 passes + RLE = 3.7 ms against 7.6 ms at 168 M keys; DESIGN.md section 4 has the ncu numbers).
  const bool finish_umi = finish_supported(c->kl.umi_bits) && getenv("CRGPU_FINISH_SORT") && atoi(getenv("CRGPU_FINISH_SORT"));
  const int sort_begin = finish_umi ? c->kl.umi_bits : 0;
  if ((rc = phase_begin(c, "count.sort.hist"))) return rc;
  unsigned long long* key_src = c->key_src ? c->key_src : c->keys.as<unsigned long long>();
  c->launches += sort_histograms(key_src, nk, c->kl.total_bits, c->sort_temp.p, c->stream, sort_begin);
  CHECK_KERNEL();
  if ((rc = phase_end(c))) return rc;
  {
    std::string nm = "count.sort.onesweep_x" + std::to_string(sort_num_passes(c->kl.total_bits - sort_begin));
    if ((rc = phase_begin(c, nm.c_str()))) return rc;
  }
  unsigned long long* sorted = key_src;
  c->launches += sort_passes(key_src, c->keys_alt.as<unsigned long long>(), nk,
                             c->kl.total_bits, c->sort_temp.p, &sorted, c->stream, sort_begin);
  CHECK_KERNEL();
  if ((rc = phase_end(c))) return rc;
  if ((rc = phase_begin(c, "count.dedup.alloc"))) return rc;
  if ((rc = c->dkeys.ensure(cap * 8))) return rc;
  if ((rc = c->c0.ensure(cap * 4))) return rc;
  if ((rc = c->best.ensure(cap * 4))) return rc;
  if ((rc = c->inc.ensure(cap * 8))) return rc;
  if ((rc = c->low.ensure(cap))) return rc;
  if ((rc = c->key2.ensure(cap * 8))) return rc;
  if ((rc = c->key2_alt.ensure(cap * 8))) return rc;
  if ((rc = c->mol.ensure(cap * 4))) return rc;
  if ((rc = c->mol_idx.ensure(cap * 4))) return rc;
  if ((rc = c->ent_rank.ensure(cap * 4))) return rc;
  if ((rc = c->ent_feature.ensure(cap * 4))) return rc;
  if ((rc = c->ent_count.ensure(cap * 4))) return rc;
  const uint64_t scan_items = std::max<uint64_t>(cap, c->content.size() + 2);
  if ((rc = c->lb_desc.ensure((scan_items / 2048 + 2) * 8))) return rc;
  DedupBuffers b;
  memset(&b, 0, sizeof(b));
  b.sorted = sorted;
  b.sorted_alt = sorted == key_src ? c->keys_alt.as<unsigned long long>() : key_src;
  b.finish_umi = finish_umi ? 1 : 0;
  b.n_keys = nk;
  b.kl = c->kl;
  b.umi_correction_mask = 0;
  for (size_t l = 0; l < c->libs.size(); l++)
    if (c->libs[l]->def.umi_correction) b.umi_correction_mask |= 1u << l;
  b.filter_umis = c->filter_umis;
  b.on_target = c->target_min_reads ? c->on_target.as<uint8_t>() : nullptr;
  b.n_on_target = c->n_on_target;
  b.target_min_reads = c->target_min_reads;
  b.verify = getenv("CRGPU_VERIFY") != nullptr;
  b.mark = [](void* user, const char* name) {
    crgpu_ctx* cc = static_cast<crgpu_ctx*>(user);
    phase_end(cc);
    phase_begin(cc, name);
  };
  b.mark_user = c;
  b.dkeys = c->dkeys.as<unsigned long long>();
  b.c0 = c->c0.as<uint32_t>();
  b.best = c->best.as<uint32_t>();
  b.inc = c->inc.as<unsigned long long>();
  b.low = c->low.as<uint8_t>();
  b.key2 = c->key2.as<unsigned long long>();
  b.key2_alt = c->key2_alt.as<unsigned long long>();
  b.lb_desc = c->lb_desc.as<unsigned long long>();
  b.tickets = c->tickets.as<uint32_t>();
  b.scalars = c->scalars.as<unsigned long long>();
  b.sort_temp = c->sort_temp.p;
  b.sort_temp_bytes = c->sort_temp.cap;
  if (getenv("CRGPU_LS_GLOBAL") && atoi(getenv("CRGPU_LS_GLOBAL"))) {  // the global slot table of the round-1 pre-filter
    int slot_bits = 16;
    while (slot_bits < 30 && (1ull << slot_bits) < 8 * cap) slot_bits++;
    if ((rc = c->ls_slots.ensure(((size_t)1 << slot_bits) / 4))) return rc;
    b.slots = c->ls_slots.as<uint32_t>();
    b.slots_bytes = c->ls_slots.cap;
  }
  b.ent_rank = c->ent_rank.as<uint32_t>();
  b.ent_feature = c->ent_feature.as<uint32_t>();
  b.ent_count = c->ent_count.as<uint32_t>();
  b.mol = c->mol.as<uint32_t>();
  b.mol_idx = c->mol_idx.as<uint32_t>();
  b.cap = cap;
  uint64_t m = 0;
  {
    int nl = run_dedup(b, &m, c->stream);
    if (nl < 0) return fail(CRGPU_E_INVALID, "internal: low-support slot table too small");
    c->launches += nl;
  }
  CHECK_KERNEL();
  c->n_distinct = m;
  if ((rc = phase_end(c))) return rc;
  if ((rc = phase_begin(c, "count.matrix"))) return rc;
  unsigned long long hs[16];
  CU(cudaMemcpyAsync(hs, c->scalars.p, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->n_mol = m ? hs[2] : 0;
  const unsigned long long hs_dedup12 = hs[12];
  const uint32_t nc = (uint32_t)c->content.size();
  if ((rc = c->col_of_rank.ensure((size_t)(nc + 1) * 4))) return rc;
  if ((rc = c->barcode_rank.ensure((size_t)(nc + 1) * 4))) return rc;
  if ((rc = c->indptr.ensure((size_t)(nc + 2) * 8))) return rc;
  MatrixArgs ma;
  memset(&ma, 0, sizeof(ma));
  ma.n_libs = (int)c->libs.size();
  for (int l = 0; l < ma.n_libs; l++) ma.valid_counts[l] = c->libs[l]->valid.as<uint32_t>();
  ma.n_content = nc;
  ma.own_lo = c->own_lo;
  ma.own_hi = std::min<uint32_t>(c->own_hi, nc);
  ma.col_of_rank = c->col_of_rank.as<uint32_t>();
  ma.barcode_rank = c->barcode_rank.as<uint32_t>();
  ma.indptr = c->indptr.as<long long>();
  uint64_t nbc = 0;
  c->launches += run_matrix(b, ma, 0, c->n_mol, &nbc, c->stream);
  CHECK_KERNEL();
  c->n_barcodes = nbc;
  CU(cudaMemcpyAsync(hs, c->scalars.p, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->nnz = c->n_mol ? hs[1] : 0;
  if ((rc = phase_end(c))) return rc;
  unsigned int lb_flags[3] = {0u, 0u, 0u};
  sort_lb_flag_fetch(&lb_flags[0], c->stream);
  dedup_lb_flag_fetch(&lb_flags[1], c->stream);
  finish_lb_flag_fetch(&lb_flags[2], c->stream);
  CU(cudaStreamSynchronize(c->stream));
  if (lb_flags[0] | lb_flags[1] | lb_flags[2]) {
    sort_lb_flag_clear(c->stream);
    dedup_lb_flag_clear(c->stream);
    finish_lb_flag_clear(c->stream);
    return fail(CRGPU_E_CUDA, "chained-scan watchdog: a tile waited for a predecessor that never became resident "
                              "(blocks were not dispatched in index order); results of this call are invalid - "
                              "set CRGPU_TICKETS=1 for ticket-ordered tiles");
  }
  c->stats[CRGPU_STAT_KEYS] = nk;
  c->stats[CRGPU_STAT_DISTINCT_KEYS] = m;
  c->stats[CRGPU_STAT_UMI_CORRECTED_KEYS] = m ? hs[3] : 0;
  c->stats[CRGPU_STAT_LOW_SUPPORT_KEYS] = m ? hs[4] : 0;
  c->stats[CRGPU_STAT_UMI_CORRECTED_READS] = m ? hs[5] : 0;
  c->stats[CRGPU_STAT_LOW_SUPPORT_READS] = m ? hs[6] : 0;
  c->stats[CRGPU_STAT_MOLECULES] = c->n_mol;
  c->stats[CRGPU_STAT_SORT_VIOLATIONS] = m ? hs[10] : 0;
  c->stats[CRGPU_STAT_RLE_VIOLATIONS] = m ? hs[11] : 0;
  c->stats[CRGPU_STAT_NNZ] = c->nnz;
  c->stats[CRGPU_STAT_FILTERED_TARGET_UMIS] = m ? hs_dedup12 : 0;
  c->stats[CRGPU_STAT_BARCODES] = c->n_barcodes;
  c->stage = 3;
  c->annotated = false;
  return CRGPU_OK;
}

int crgpu_annotate_reads(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  if (c->keys_external) return fail(CRGPU_E_INVALID, "per-read annotation needs the keys of this context's own reads");
  CU(cudaSetDevice(c->device));
  int rc;
  phases_clear(c, "annotate");
  if ((rc = phase_begin(c, "annotate"))) return rc;
  const uint64_t m = c->n_distinct;
  if ((rc = c->umi_proc.ensure(c->n_reads * 4 + 16))) return rc;
  if ((rc = c->flags.ensure(c->n_reads + 16))) return rc;
  if ((rc = c->min_read.ensure(std::max<uint64_t>(m, 1) * 8))) return rc;  // UmiSelectKey word per distinct key
  if ((rc = c->rep_raw.ensure(std::max<uint64_t>(m, 1) * 4))) return rc;
  DedupBuffers b;
  memset(&b, 0, sizeof(b));
  b.kl = c->kl;
  b.dkeys = c->dkeys.as<unsigned long long>();
  b.best = c->best.as<uint32_t>();
  b.low = c->low.as<uint8_t>();
  c->launches += run_annotate_prepare(b, m, c->min_read.as<unsigned long long>(), c->rep_raw.as<uint32_t>(), c->stream);
  for (int pass = 0; pass < 2; pass++) {
    for (auto* bt : c->batches) {
      Library* l = c->libs[bt->lib];
      AnnotateArgs a;
      memset(&a, 0, sizeof(a));
      a.n = bt->n;
      a.bc_out = c->bc_out.as<uint32_t>() + bt->base;
      a.umi_out = c->umi_out.as<uint32_t>() + bt->base;
      a.umi_proc = c->umi_proc.as<uint32_t>() + bt->base;
      a.feature = l->def.is_feature_barcode ? bt->feature_res.as<uint32_t>() : bt->feature;
      a.flags_out = c->flags.as<uint8_t>() + bt->base;
      a.lib = (uint32_t)bt->lib;
      a.read_base = bt->base;
      a.select = bt->select;
      if (pass == 0)
        c->launches += run_annotate_min(b, m, a, c->min_read.as<unsigned long long>(), c->stream);
      else
        c->launches += run_annotate_final(b, m, a, c->min_read.as<unsigned long long>(), c->rep_raw.as<uint32_t>(),
                                          nullptr, c->stream);
      CHECK_KERNEL();
    }
  }
  if ((rc = phase_end(c))) return rc;
  CU(cudaStreamSynchronize(c->stream));
  c->annotated = true;
  return CRGPU_OK;
}

int crgpu_run(crgpu_ctx* c) {
  int rc;
  if ((rc = crgpu_pass1(c))) return rc;
  if ((rc = crgpu_pass2(c))) return rc;
  return crgpu_count(c);
}

int crgpu_sync(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return CRGPU_OK;
}

int crgpu_stats(crgpu_ctx* c, uint64_t out[CRGPU_STAT_COUNT]) {
  if (!c || !out) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  if (c->stage >= 1 && c->n_reads) {
    // barcode states of this context's own reads (the histograms may hold all-reduced, global counts)
    unsigned long long* d4 = c->scalars.as<unsigned long long>() + 48;
    unsigned long long h4[4];
    CU(cudaMemsetAsync(d4, 0, 32, c->stream));
    c->launches += launch_state_counts(c->bc_out.as<uint32_t>(), c->n_reads, d4, c->stream);
    CU(cudaMemcpyAsync(h4, d4, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats[CRGPU_STAT_VALID_BEFORE] = h4[1];
    c->stats[CRGPU_STAT_CORRECTED] = h4[2];
    c->stats[CRGPU_STAT_INVALID] = h4[3];
  }
  c->stats[CRGPU_STAT_READS] = c->n_reads;
  c->stats[CRGPU_STAT_KERNEL_LAUNCHES] = c->launches;
  memcpy(out, c->stats, sizeof(c->stats));
  return CRGPU_OK;
}

int crgpu_reads_get(crgpu_ctx* c, int batch, uint32_t* bc_rank, uint8_t* bc_state, uint32_t* umi, uint8_t* flags,
                    uint32_t* feature) {
  if (!c || batch < 0 || batch >= (int)c->batches.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 1) return fail(CRGPU_E_INVALID, "crgpu_pass1 must run first");
  CU(cudaSetDevice(c->device));
  Batch* b = c->batches[batch];
  const uint64_t n = b->n;
  std::vector<uint32_t> h(n);
  if (bc_rank || bc_state) {
    CU(cudaMemcpyAsync(h.data(), c->bc_out.as<uint32_t>() + b->base, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint64_t i = 0; i < n; i++) {
      uint32_t st = h[i] >> BC_STATE_SHIFT;
      if (bc_state) bc_state[i] = (uint8_t)st;
      if (bc_rank) bc_rank[i] = (st == ST_VALID_BEFORE || st == ST_VALID_AFTER) ? (h[i] & BC_RANK_MASK) : CRGPU_NO_RANK;
    }
  }
  if (umi || flags) {
    const uint32_t* src = c->annotated ? c->umi_proc.as<uint32_t>() : c->umi_out.as<uint32_t>();
    CU(cudaMemcpyAsync(h.data(), src + b->base, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (umi)
      for (uint64_t i = 0; i < n; i++) umi[i] = h[i] & UMI_SEQ_MASK;
    if (flags) {
      if (c->annotated) {
        CU(cudaMemcpyAsync(flags, c->flags.as<uint8_t>() + b->base, n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
      } else {
        for (uint64_t i = 0; i < n; i++) flags[i] = (h[i] & UMI_VALID_BIT) ? CRGPU_F_UMI_VALID : 0;
      }
    }
  }
  if (feature) {
    Library* l = c->libs[b->lib];
    const uint32_t* src = l->def.is_feature_barcode ? b->feature_res.as<uint32_t>() : b->feature;
    if (l->def.is_feature_barcode && c->stage < 2) return fail(CRGPU_E_INVALID, "features are resolved by crgpu_pass2");
    CU(cudaMemcpyAsync(feature, src, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return CRGPU_OK;
}

int crgpu_bc_counts_get(crgpu_ctx* c, int lib, int which, uint32_t* out, uint64_t n) {
  if (!c || lib < 0 || lib >= (int)c->libs.size() || !out) return fail(CRGPU_E_INVALID, "bad argument");
  if (n != c->content.size()) return fail(CRGPU_E_INVALID, "length differs from the content size");
  CU(cudaSetDevice(c->device));
  const void* src = which == 0 ? c->libs[lib]->prior.p : c->libs[lib]->corrected.p;
  CU(cudaMemcpyAsync(out, src, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}

int crgpu_fb_counts_get(crgpu_ctx* c, int64_t* out, int32_t n) {
  if (!c || !out || n != c->n_features) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  if (n) {
    CU(cudaMemcpyAsync(out, c->d_fb_counts.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return CRGPU_OK;
}

int crgpu_matrix_dims(crgpu_ctx* c, uint64_t* n_barcodes, uint64_t* nnz, uint64_t* n_features) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  if (n_barcodes) *n_barcodes = c->n_barcodes;
  if (nnz) *nnz = c->nnz;
  if (n_features) *n_features = (uint64_t)c->n_features;
  return CRGPU_OK;
}

int crgpu_matrix_get(crgpu_ctx* c, uint32_t* barcode_rank, int64_t* indptr, uint32_t* indices, int32_t* data) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  CU(cudaSetDevice(c->device));
  if (barcode_rank && c->n_barcodes)
    CU(cudaMemcpyAsync(barcode_rank, c->barcode_rank.p, c->n_barcodes * 4, cudaMemcpyDeviceToHost, c->stream));
  if (indptr) CU(cudaMemcpyAsync(indptr, c->indptr.p, (c->n_barcodes + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (indices && c->nnz) CU(cudaMemcpyAsync(indices, c->ent_feature.p, c->nnz * 4, cudaMemcpyDeviceToHost, c->stream));
  if (data && c->nnz) CU(cudaMemcpyAsync(data, c->ent_count.p, c->nnz * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}

int crgpu_barcode_summary(crgpu_ctx* c, int lib, uint32_t* out) {
  if (!c || !out) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  if (lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "no such library");
  if (c->n_barcodes == 0) return CRGPU_OK;
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->summary.ensure(c->n_barcodes * 16))) return rc;
  DedupBuffers b;
  memset(&b, 0, sizeof(b));
  b.kl = c->kl;
  b.dkeys = c->dkeys.as<unsigned long long>();
  b.c0 = c->c0.as<uint32_t>();
  b.best = c->best.as<uint32_t>();
  b.inc = c->inc.as<unsigned long long>();
  b.low = c->low.as<uint8_t>();
  c->launches += run_barcode_summary(b, c->n_distinct, (uint32_t)lib, c->barcode_rank.as<uint32_t>(),
                                     c->libs[lib]->valid.as<uint32_t>(), c->col_of_rank.as<uint32_t>(), c->n_barcodes,
                                     c->summary.as<uint32_t>(), c->stream);
  CHECK_KERNEL();
  CU(cudaMemcpyAsync(out, c->summary.p, c->n_barcodes * 16, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}

int crgpu_barcode_diversity(crgpu_ctx* c, int lib, uint64_t* barcodes_detected, double* effective_diversity) {
  if (!c || lib < 0 || lib >= (int)c->libs.size()) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 2) return fail(CRGPU_E_INVALID, "crgpu_pass2 must run first");
  CU(cudaSetDevice(c->device));
  unsigned long long* d4 = c->scalars.as<unsigned long long>() + 52;
  unsigned long long h[4] = {0, 0, 0, 0};
  c->launches += launch_diversity(c->libs[lib]->valid.as<uint32_t>(), c->content.size(), d4, c->stream);
  CHECK_KERNEL();
  CU(cudaMemcpyAsync(h, d4, 32, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (barcodes_detected) *barcodes_detected = h[0];
  if (effective_diversity) {
    // s.powi(2) / s2 in f64 as the reference computes it; the integer sums are exact, so this equals the
    // reference's value whenever its own f64 accumulation is exact (sum c^2 < 2^53), in any iteration order
    const double s = (double)h[1];
    const double s2 = (double)h[2] + (double)h[3] * 18446744073709551616.0;
    *effective_diversity = (s * s) / s2;  // NaN for an empty histogram, like 0/0 in the reference
  }
  return CRGPU_OK;
}

int crgpu_molecules_count(crgpu_ctx* c, uint64_t* n) {
  if (!c || !n) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  *n = c->n_mol;
  return CRGPU_OK;
}

int crgpu_molecules_get(crgpu_ctx* c, uint32_t* out6) {
  if (!c || !out6) return fail(CRGPU_E_INVALID, "bad argument");
  if (c->stage < 3) return fail(CRGPU_E_INVALID, "crgpu_count must run first");
  if (!c->n_mol) return CRGPU_OK;
  if (c->have_select && !c->annotated)
    return fail(CRGPU_E_INVALID, "a batch carries select keys: UmiCount::utype is the type of the representative read, "
                                 "crgpu_annotate_reads must run before crgpu_molecules_get");
  CU(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->mol_rows.ensure(c->n_mol * 24))) return rc;
  // UmiCount sorts by (library_idx, feature_idx, umi) inside a barcode; the keys carry the feature above the
  // library. The two orders agree when every library's features lie above those of the libraries before it (one
  // library; GEX genes first and a feature-barcode library behind them): otherwise the rows are re-sorted.
  int reorder = 0;
  for (size_t l1 = 0; l1 < c->libs.size(); l1++)  // two libraries over the same features: feature-major != library-major
    for (size_t l2 = l1 + 1; l2 < c->libs.size(); l2++) {
      const crgpu_library_def &d1 = c->libs[l1]->def, &d2 = c->libs[l2]->def;
      if ((d1.is_feature_barcode ? d1.feature_type : 0) == (d2.is_feature_barcode ? d2.feature_type : 0)) reorder = 1;
    }
  if (c->libs.size() > 1 && !reorder) {
    int last_lib = -1;
    for (int f = 0; f < c->n_features && !reorder; f++) {
      int lib_of = -1;
      for (size_t l = 0; l < c->libs.size(); l++) {
        const crgpu_library_def& d = c->libs[l]->def;
        if ((d.is_feature_barcode ? d.feature_type : 0) == c->feature_type[f]) lib_of = (int)l;
      }
      if (lib_of < 0) continue;
      if (lib_of < last_lib) reorder = 1;
      last_lib = std::max(last_lib, lib_of);
    }
  }
  if (reorder) {
    if ((rc = c->mol_sort.ensure(c->n_mol * 8 + 16))) return rc;
    if ((rc = c->mol_sort_alt.ensure(c->n_mol * 8 + 16))) return rc;
    if ((rc = c->sort_temp.ensure(sort_temp_bytes(c->n_mol)))) return rc;
  }
  DedupBuffers b;
  memset(&b, 0, sizeof(b));
  b.kl = c->kl;
  b.key2 = c->key2.as<unsigned long long>();
  b.mol = c->mol.as<uint32_t>();
  b.mol_idx = c->mol_idx.as<uint32_t>();
  const bool typed = c->have_select && c->annotated;
  c->launches += run_molecule_rows(b, c->col_of_rank.as<uint32_t>(), c->n_mol,
                                   typed ? c->min_read.as<unsigned long long>() : nullptr,
                                   typed ? c->rep_raw.as<uint32_t>() : nullptr, reorder,
                                   c->mol_sort.as<unsigned long long>(), c->mol_sort_alt.as<unsigned long long>(),
                                   c->sort_temp.p, c->sort_temp.cap, c->mol_rows.as<uint32_t>(), c->stream);
  CHECK_KERNEL();
  CU(cudaMemcpyAsync(out6, c->mol_rows.p, c->n_mol * 24, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}

int crgpu_synth_generate(crgpu_ctx* c, const crgpu_synth_params* p, uint64_t start, uint64_t n, uint8_t* r1_seq,
                         uint8_t* r1_qual, uint32_t* feature, uint8_t* r2_seq, uint8_t* r2_qual) {
  if (!c || !p) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  DevBuf wl, cr, cc, nm, gc, fc, fp;
  int rc = 0;
  auto up = [&](DevBuf& d, const uint32_t* h, size_t cnt) -> int {
    int r = d.ensure(std::max<size_t>(4, cnt * 4));
    if (r) return r;
    if (cnt && h) {
      cudaError_t e = cudaMemcpyAsync(d.p, h, cnt * 4, cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) return fail(CRGPU_E_CUDA, cudaGetErrorString(e));
    }
    return 0;
  };
  rc = up(wl, p->wl_packed, p->n_whitelist);
  if (!rc) rc = up(cr, p->cell_rank, p->n_cells);
  if (!rc) rc = up(cc, p->cell_cdf, p->n_cells);
  if (!rc) rc = up(nm, p->n_mol, p->n_cells);
  if (!rc) rc = up(gc, p->gene_cdf, p->n_genes);
  if (!rc) rc = up(fc, p->fb_cdf, p->n_fb);
  if (!rc) rc = up(fp, p->fb_packed, p->n_fb);
  if (!rc) {
    c->launches += launch_synth(p, wl.as<uint32_t>(), cr.as<uint32_t>(), cc.as<uint32_t>(), nm.as<uint32_t>(),
                                gc.as<uint32_t>(), fc.as<uint32_t>(), fp.as<uint32_t>(), start, n, r1_seq, r1_qual,
                                feature, r2_seq, r2_qual, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(CRGPU_E_CUDA, std::string("synth: ") + cudaGetErrorString(e));
  }
  wl.release(); cr.release(); cc.release(); nm.release(); gc.release(); fc.release(); fp.release();
  return rc;
}

int crgpu_dev_alloc(crgpu_ctx* c, uint64_t bytes, void** out) {
  if (!c || !out) return fail(CRGPU_E_INVALID, "bad argument");
  CU(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return CRGPU_OK;
}
int crgpu_dev_free(crgpu_ctx* c, void* dev) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaFree(dev));
  return CRGPU_OK;
}
int crgpu_memcpy_d2h(crgpu_ctx* c, void* host, const void* dev, uint64_t bytes) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}
int crgpu_memcpy_h2d(crgpu_ctx* c, void* dev, const void* host, uint64_t bytes) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CRGPU_OK;
}
int crgpu_host_alloc_pinned(uint64_t bytes, void** out) {
  if (!out) return fail(CRGPU_E_INVALID, "out is NULL");
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return CRGPU_OK;
}
int crgpu_host_free_pinned(void* host) {
  CU(cudaFreeHost(host));
  return CRGPU_OK;
}

int crgpu_phase_times(crgpu_ctx* c, float* out_ms, int32_t cap, int32_t* out_n, const char** out_names) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  c->phase_names.clear();
  int n = 0;
  for (auto& p : c->phases) {
    if (n < cap && out_ms) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, p.second.first, p.second.second));
      out_ms[n] = ms;
    }
    c->phase_names += p.first;
    c->phase_names.push_back('\0');
    n++;
  }
  c->phase_names.push_back('\0');
  if (out_n) *out_n = n;
  if (out_names) *out_names = c->phase_names.c_str();
  return CRGPU_OK;
}

int crgpu_stream(crgpu_ctx* c, void** out) {
  if (!c || !out) return fail(CRGPU_E_INVALID, "bad argument");
  *out = (void*)c->stream;
  return CRGPU_OK;
}

}  // extern "C"
