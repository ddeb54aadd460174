// shard.cu — the sharded run behind the C ABI: communicator, owner ranges, fused key exchange, count.
//
// The reference shards this path by barcode range through files (ShardReader::make_chunks,
// lib/rust/cr_lib/src/stages/barcode_correction.rs:252-262, stages/align_and_count.rs:519-524) and sums the
// priors over all chunks in a join (stages/make_shard.rs:303-358). Here the same structure runs on the GPUs of
// one box with nothing but the context stream between the stages: NCCL all-reduces for the global count vectors,
// owner ranges from a device scan, keys stored straight into their owner's memory over NVLink, an on-stream
// barrier, shard-local count. One host round trip per step (the received key count sizes the sort).
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"
#include "nccl_dl.h"

namespace {

#define NC(call)                                                                                              \
  do {                                                                                                        \
    int r__ = (call);                                                                                         \
    if (r__ != nccl_dl::kSuccess)                                                                             \
      return fail(CRGPU_E_CUDA, std::string(#call) + ": " + (api->GetErrorString ? api->GetErrorString(r__) : "NCCL error")); \
  } while (0)

// layout of the device scratch of a sharded context
constexpr size_t SH_BOUNDS = 0;       // u32[CRGPU_MAX_PARTS + 1]
constexpr size_t SH_BARRIER = 128;    // u32: the word of the on-stream barrier
constexpr size_t SH_MY_RECORD = 256;  // PeerRecord of this rank
constexpr size_t SH_RECORDS = 512;    // PeerRecord[CRGPU_MAX_PARTS]
constexpr size_t SH_PARTIAL = 8192;   // u64 partial sums of the owner-bounds scan

struct PeerRecord {  // what every rank tells the others about its receive buffer
  unsigned long long pid;
  unsigned long long host_hash;
  unsigned long long buf, cursor;  // device addresses, usable by ranks of the same process
  int device;
  int pad;
  cudaIpcMemHandle_t h_buf, h_cursor;  // for ranks in other processes
  char fill[256 - 8 * 4 - 8 - 2 * sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(PeerRecord) == 256, "PeerRecord is exchanged as 256 bytes");

unsigned long long host_hash() {
  char name[256] = {0};
  gethostname(name, sizeof(name) - 1);
  unsigned long long h = 1469598103934665603ull;
  for (const char* p = name; *p; p++) h = (h ^ (unsigned char)*p) * 1099511628211ull;
  return h;
}

}  // namespace

// called by crgpu_ctx_destroy
void shard_release(crgpu_ctx* c) {
  if (c->nccl_comm) {
    const nccl_dl::Api* api = nccl_dl::load(nullptr);
    if (api) api->CommDestroy(static_cast<nccl_dl::comm_t>(c->nccl_comm));
    c->nccl_comm = nullptr;
  }
  c->shard_buf.release();
}

extern "C" {

int crgpu_comm_unique_id(void* out_id) {
  if (!out_id) return fail(CRGPU_E_INVALID, "out_id is NULL");
  std::string why;
  const nccl_dl::Api* api = nccl_dl::load(&why);
  if (!api) return fail(CRGPU_E_CUDA, why);
  nccl_dl::unique_id id;
  NC(api->GetUniqueId(&id));
  static_assert(sizeof(id) == CRGPU_COMM_ID_BYTES, "unique id size");
  memcpy(out_id, &id, sizeof(id));
  return CRGPU_OK;
}

int crgpu_comm_init(crgpu_ctx* c, const void* id_bytes, int32_t n_ranks, int32_t rank, uint64_t capacity_keys) {
  if (!c || !id_bytes) return fail(CRGPU_E_INVALID, "NULL argument");
  if (n_ranks < 1 || n_ranks > CRGPU_MAX_PARTS || rank < 0 || rank >= n_ranks)
    return fail(CRGPU_E_INVALID, "rank / n_ranks out of range (at most " + std::to_string(CRGPU_MAX_PARTS) + " ranks)");
  if (c->nccl_comm) return fail(CRGPU_E_INVALID, "this context already has a communicator");
  if (c->xchg_buf) return fail(CRGPU_E_INVALID, "this context already has an exchange buffer (crgpu_exchange_init)");
  if (capacity_keys == 0) return fail(CRGPU_E_INVALID, "exchange_capacity_keys must be > 0");
  std::string why;
  const nccl_dl::Api* api = nccl_dl::load(&why);
  if (!api) return fail(CRGPU_E_CUDA, why);
  CU(cudaSetDevice(c->device));
  nccl_dl::unique_id id;
  memcpy(&id, id_bytes, sizeof(id));
  nccl_dl::comm_t comm = nullptr;
  NC(api->CommInitRank(&comm, n_ranks, id, rank));
  c->nccl_comm = comm;
  c->comm_n = n_ranks;
  c->comm_rank = rank;
  int rc;
  if ((rc = c->shard_buf.ensure(SH_PARTIAL + (size_t)(1u << 20)))) return rc;
  CU(cudaMemsetAsync(c->shard_buf.p, 0, SH_PARTIAL, c->stream));
  // this rank's receive buffer and cursor
  cudaError_t e = cudaMalloc(&c->xchg_buf, capacity_keys * 8);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("exchange buffer: ") + cudaGetErrorString(e));
  e = cudaMalloc(&c->xchg_cursor, 16);
  if (e != cudaSuccess) return fail(CRGPU_E_NOMEM, std::string("exchange cursor: ") + cudaGetErrorString(e));
  CU(cudaMemsetAsync(c->xchg_cursor, 0, 16, c->stream));
  c->xchg_capacity = capacity_keys;
  // tell everybody where it is: one NCCL all-gather of 256-byte records
  PeerRecord mine;
  memset(&mine, 0, sizeof(mine));
  mine.pid = (unsigned long long)getpid();
  mine.host_hash = host_hash();
  mine.buf = (unsigned long long)(uintptr_t)c->xchg_buf;
  mine.cursor = (unsigned long long)(uintptr_t)c->xchg_cursor;
  mine.device = c->device;
  CU(cudaIpcGetMemHandle(&mine.h_buf, c->xchg_buf));
  CU(cudaIpcGetMemHandle(&mine.h_cursor, c->xchg_cursor));
  unsigned char* sb = c->shard_buf.as<unsigned char>();
  CU(cudaMemcpyAsync(sb + SH_MY_RECORD, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  NC(api->AllGather(sb + SH_MY_RECORD, sb + SH_RECORDS, sizeof(PeerRecord), nccl_dl::kUint8, comm, c->stream));
  std::vector<PeerRecord> all(n_ranks);
  CU(cudaMemcpyAsync(all.data(), sb + SH_RECORDS, sizeof(PeerRecord) * n_ranks, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < n_ranks; r++) {
    if (r == rank) {
      c->peer_buf[r] = static_cast<unsigned long long*>(c->xchg_buf);
      c->peer_cursor[r] = static_cast<unsigned long long*>(c->xchg_cursor);
      continue;
    }
    const PeerRecord& p = all[r];
    if (p.host_hash != mine.host_hash)
      return fail(CRGPU_E_INVALID, "rank " + std::to_string(r) + " runs on another host: the key exchange goes over "
                                       "peer memory (NVLink) and needs every rank on one box");
    if (p.pid == mine.pid) {
      // same process: plain peer access, the other context's pointers are valid here
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, c->device, p.device));
      if (!can)
        return fail(CRGPU_E_CUDA, "device " + std::to_string(c->device) + " cannot access the memory of device " +
                                      std::to_string(p.device) + " (no NVLink / PCIe peer path)");
      cudaError_t pe = cudaDeviceEnablePeerAccess(p.device, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
        return fail(CRGPU_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
      cudaGetLastError();  // clear the "already enabled" status
      c->peer_buf[r] = reinterpret_cast<unsigned long long*>((uintptr_t)p.buf);
      c->peer_cursor[r] = reinterpret_cast<unsigned long long*>((uintptr_t)p.cursor);
    } else {
      void *pb = nullptr, *pc = nullptr;
      cudaError_t oe = cudaIpcOpenMemHandle(&pb, p.h_buf, cudaIpcMemLazyEnablePeerAccess);
      if (oe == cudaSuccess) oe = cudaIpcOpenMemHandle(&pc, p.h_cursor, cudaIpcMemLazyEnablePeerAccess);
      if (oe != cudaSuccess)
        return fail(CRGPU_E_CUDA, std::string("cannot map the receive buffer of rank ") + std::to_string(r) + ": " +
                                      cudaGetErrorString(oe));
      c->peer_buf[r] = static_cast<unsigned long long*>(pb);
      c->peer_cursor[r] = static_cast<unsigned long long*>(pc);
      c->peer_opened[r] = true;
    }
  }
  c->xchg_ranks = n_ranks;
  c->xchg_rank = rank;
  if (!c->xchg_stream) {
    CU(cudaStreamCreateWithFlags(&c->xchg_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->xchg_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->xchg_done, cudaEventDisableTiming));
  }
  return CRGPU_OK;
}

int crgpu_sharded_run(crgpu_ctx* c) {
  if (!c) return fail(CRGPU_E_INVALID, "ctx is NULL");
  if (!c->nccl_comm) return fail(CRGPU_E_INVALID, "crgpu_comm_init must run first");
  std::string why;
  const nccl_dl::Api* api = nccl_dl::load(&why);
  if (!api) return fail(CRGPU_E_CUDA, why);
  nccl_dl::comm_t comm = static_cast<nccl_dl::comm_t>(c->nccl_comm);
  CU(cudaSetDevice(c->device));
  const int G = c->comm_n;
  const bool early = getenv("CRGPU_EARLY_SCATTER") && atoi(getenv("CRGPU_EARLY_SCATTER")) != 0;
  const uint32_t W = (uint32_t)c->content.size();
  int rc;
  if (c->shard_buf.cap < SH_PARTIAL + owner_bounds_scratch_bytes(W)) {  // a whitelist larger than the scratch foresaw
    if ((rc = c->shard_buf.ensure(SH_PARTIAL + owner_bounds_scratch_bytes(W)))) return rc;
    CU(cudaMemsetAsync(c->shard_buf.p, 0, SH_PARTIAL, c->stream));
  }
  unsigned char* sb = c->shard_buf.as<unsigned char>();
  uint32_t* d_bounds = reinterpret_cast<uint32_t*>(sb + SH_BOUNDS);
  uint32_t* d_barrier = reinterpret_cast<uint32_t*>(sb + SH_BARRIER);
  unsigned long long* d_partial = reinterpret_cast<unsigned long long*>(sb + SH_PARTIAL);
  phases_clear(c, "shard");

  // The receive cursor of this rank: cleared before this rank contributes to the first all-reduce, hence before
  // any peer (which scatters only after that all-reduce has completed) stores a key of this step.
  CU(cudaMemsetAsync(c->xchg_cursor, 0, 16, c->stream));
  if ((rc = crgpu_pass1(c))) return rc;

  if ((rc = phase_begin(c, "shard.allreduce.priors"))) return rc;
  for (auto* l : c->libs) NC(api->AllReduce(l->prior.p, l->prior.p, W, nccl_dl::kUint32, nccl_dl::kSum, comm, c->stream));
  if (c->have_fb && c->n_features)
    NC(api->AllReduce(c->d_fb_counts.p, c->d_fb_counts.p, (size_t)c->n_features, nccl_dl::kUint64, nccl_dl::kSum, comm,
                      c->stream));
  if ((rc = phase_end(c))) return rc;

  unsigned long long* ctr = c->counters.as<unsigned long long>();
  unsigned long long* d_sent = c->scalars.as<unsigned long long>() + 16;
  unsigned long long* d_sent_early = c->scalars.as<unsigned long long>() + 32;
  const uint32_t* vec[CRGPU_MAX_LIBS] = {nullptr};
  if (early) {
    // owner ranges from the global valid-before counts alone: known one pass earlier, so the keys of pass 1
    // (about 85 % of all keys) cross NVLink on a second stream while pass 2 runs
    for (size_t l = 0; l < c->libs.size(); l++) vec[l] = c->libs[l]->prior.as<uint32_t>();
    c->launches += run_owner_bounds(vec, (int)c->libs.size(), W, G, d_partial, d_bounds, c->stream);
    CHECK_KERNEL();
    CU(cudaEventRecord(c->xchg_ready, c->stream));
    CU(cudaStreamWaitEvent(c->xchg_stream, c->xchg_ready, 0));
    c->launches += run_owner_scatter_peers_dev(c->keys.as<unsigned long long>(), nullptr, ctr + CTR_KEYS_PASS1, c->n_reads,
                                               c->kl.rank_shift, d_bounds, G, c->peer_buf, c->peer_cursor,
                                               c->xchg_capacity, d_sent_early, c->xchg_stream);
    CHECK_KERNEL();
    CU(cudaEventRecord(c->xchg_done, c->xchg_stream));
  } else {
    CU(cudaMemsetAsync(d_sent_early, 0, CRGPU_MAX_PARTS * 8, c->stream));
  }

  if ((rc = crgpu_pass2(c))) return rc;

  if ((rc = phase_begin(c, "shard.allreduce.corrected"))) return rc;
  for (auto* l : c->libs)
    NC(api->AllReduce(l->corrected.p, l->corrected.p, W, nccl_dl::kUint32, nccl_dl::kSum, comm, c->stream));
  // prior + corrected, both global now = the valid-barcode counts (barcode index of the count stage)
  for (auto* l : c->libs) {
    c->launches += launch_valid_counts(l->prior.as<uint32_t>(), l->corrected.as<uint32_t>(), l->valid.as<uint32_t>(), W,
                                       c->stream);
    CHECK_KERNEL();
  }
  if ((rc = phase_end(c))) return rc;

  if ((rc = phase_begin(c, "shard.exchange"))) return rc;
  if (!early) {
    for (size_t l = 0; l < c->libs.size(); l++) vec[l] = c->libs[l]->valid.as<uint32_t>();
    c->launches += run_owner_bounds(vec, (int)c->libs.size(), W, G, d_partial, d_bounds, c->stream);
    CHECK_KERNEL();
  }
  c->launches += run_owner_scatter_peers_dev(c->keys.as<unsigned long long>(), early ? ctr + CTR_KEYS_PASS1 : nullptr,
                                             ctr + 1, c->n_reads, c->kl.rank_shift, d_bounds, G, c->peer_buf,
                                             c->peer_cursor, c->xchg_capacity, d_sent, c->stream);
  CHECK_KERNEL();
  if (early) CU(cudaStreamWaitEvent(c->stream, c->xchg_done, 0));
  // On-stream barrier: this all-reduce completes on a rank only after every rank has contributed, and a rank
  // contributes only after its scatter kernels have retired, so behind it every key of the step has landed.
  NC(api->AllReduce(d_barrier, d_barrier, 1, nccl_dl::kUint32, nccl_dl::kSum, comm, c->stream));
  if ((rc = phase_end(c))) return rc;

  // the one host round trip of the step: received key count, bounds, statistics
  unsigned long long h_cursor[2] = {0, 0}, h_sent[CRGPU_MAX_PARTS] = {0}, h_sent_early[CRGPU_MAX_PARTS] = {0}, h_bad = 0;
  CU(cudaMemcpyAsync(h_cursor, c->xchg_cursor, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_bounds, d_bounds, (size_t)(G + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(h_sent, d_sent, (size_t)G * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(h_sent_early, d_sent_early, (size_t)G * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&h_bad, ctr + CTR_BAD_FEATURE, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (h_bad)
    return fail(CRGPU_E_INVALID, std::to_string(h_bad) + " reads carry a feature index >= n_features (" +
                                     std::to_string(c->n_features) + "): the matrix has no such row");
  if (h_cursor[1] || h_cursor[0] > c->xchg_capacity)
    return fail(CRGPU_E_LIMIT, "exchange buffer overflow: " + std::to_string(h_cursor[0]) + " keys for a capacity of " +
                                   std::to_string(c->xchg_capacity) + " (exchange_capacity_keys of crgpu_comm_init)");
  c->shard_sent_remote_keys = 0;
  for (int r = 0; r < G; r++)
    if (r != c->comm_rank) c->shard_sent_remote_keys += h_sent[r] + h_sent_early[r];
  c->shard_received_keys = h_cursor[0];
  if ((rc = c->keys_alt.ensure(h_cursor[0] * 8 + 16))) return rc;
  // the receive buffer is the sort input (and one of its ping-pong buffers): peers store into it again only
  // behind the first all-reduce of the next step, which this rank joins after its count stage
  c->key_src = static_cast<unsigned long long*>(c->xchg_buf);
  c->n_keys = h_cursor[0];
  c->keys_external = true;
  c->own_lo = c->h_bounds[c->comm_rank];
  c->own_hi = c->h_bounds[c->comm_rank + 1];
  return crgpu_count(c);
}

int crgpu_owner_bounds_get(crgpu_ctx* c, uint32_t* out, int32_t cap) {
  if (!c || !out) return fail(CRGPU_E_INVALID, "NULL argument");
  if (!c->nccl_comm) return fail(CRGPU_E_INVALID, "no sharded run on this context");
  if (cap < c->comm_n + 1) return fail(CRGPU_E_INVALID, "out_bounds holds fewer than n_ranks + 1 values");
  memcpy(out, c->h_bounds, (size_t)(c->comm_n + 1) * 4);
  return CRGPU_OK;
}

int crgpu_shard_stats(crgpu_ctx* c, uint64_t out4[4]) {
  if (!c || !out4) return fail(CRGPU_E_INVALID, "NULL argument");
  out4[0] = c->shard_sent_remote_keys;
  out4[1] = c->shard_received_keys;
  out4[2] = (uint64_t)c->comm_n;
  out4[3] = (uint64_t)(c->comm_rank < 0 ? 0 : c->comm_rank);
  return CRGPU_OK;
}

int crgpu_owner_bounds_compute(crgpu_ctx* c, const uint32_t* host_counts, uint64_t n, int32_t n_parts, uint32_t* out) {
  if (!c || !out || (!host_counts && n)) return fail(CRGPU_E_INVALID, "NULL argument");
  if (n_parts < 1 || n_parts > CRGPU_MAX_PARTS || n >= (1ull << 32)) return fail(CRGPU_E_INVALID, "bad n / n_parts");
  CU(cudaSetDevice(c->device));
  DevBuf v, scratch;
  int rc;
  if ((rc = v.ensure(std::max<size_t>(n, 1) * 4))) return rc;
  if ((rc = scratch.ensure(owner_bounds_scratch_bytes((uint32_t)n) + 256))) {
    v.release();
    return rc;
  }
  cudaError_t e = cudaSuccess;
  if (n) e = cudaMemcpyAsync(v.p, host_counts, n * 4, cudaMemcpyHostToDevice, c->stream);
  const uint32_t* vec[1] = {v.as<uint32_t>()};
  uint32_t* d_bounds = scratch.as<uint32_t>();
  unsigned long long* d_partial = reinterpret_cast<unsigned long long*>(scratch.as<unsigned char>() + 256);
  if (e == cudaSuccess) {
    c->launches += run_owner_bounds(vec, 1, (uint32_t)n, n_parts, d_partial, d_bounds, c->stream);
    e = cudaMemcpyAsync(out, d_bounds, (size_t)(n_parts + 1) * 4, cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  v.release();
  scratch.release();
  if (e != cudaSuccess) return fail(CRGPU_E_CUDA, std::string("owner bounds: ") + cudaGetErrorString(e));
  return CRGPU_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// single-process group: one context and one host thread per device
// ---------------------------------------------------------------------------
struct crgpu_group {
  std::vector<crgpu_ctx*> ctx;
  uint64_t capacity_keys = 0;
};

namespace {

// run f(i) for every member on its own thread; the first failure (lowest index) becomes this thread's error
template <typename F>
int group_parallel(crgpu_group* g, F f) {
  const int n = (int)g->ctx.size();
  std::vector<int> rc(n, 0);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++)
    th.emplace_back([&, i]() {
      rc[i] = f(i);
      if (rc[i]) msg[i] = crgpu_last_error();
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i]) return fail(rc[i], "device " + std::to_string(g->ctx[i]->device) + ": " + msg[i]);
  return CRGPU_OK;
}

}  // namespace

extern "C" {

int crgpu_group_create(const int32_t* devices, int32_t n, uint64_t capacity_keys, crgpu_group** out) {
  if (!devices || !out || n < 1 || n > CRGPU_MAX_PARTS) return fail(CRGPU_E_INVALID, "bad argument");
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) return fail(CRGPU_E_INVALID, "a device appears twice in the group");
  crgpu_group* g = new crgpu_group();
  for (int i = 0; i < n; i++) {
    crgpu_ctx* c = nullptr;
    int rc = crgpu_ctx_create(devices[i], &c);
    if (rc) {
      for (auto* x : g->ctx) crgpu_ctx_destroy(x);
      delete g;
      return rc;
    }
    g->ctx.push_back(c);
  }
  g->capacity_keys = capacity_keys;  // the communicator and the receive buffers come with the first crgpu_group_run
  *out = g;
  return CRGPU_OK;
}

crgpu_ctx* crgpu_group_ctx(crgpu_group* g, int32_t i) {
  if (!g || i < 0 || i >= (int)g->ctx.size()) return nullptr;
  return g->ctx[i];
}

int crgpu_group_size(crgpu_group* g) { return g ? (int)g->ctx.size() : 0; }

int crgpu_group_run(crgpu_group* g) {
  if (!g) return fail(CRGPU_E_INVALID, "group is NULL");
  int rc;
  if (!g->ctx[0]->nccl_comm) {
    const uint64_t capacity = g->capacity_keys;
    unsigned char id[CRGPU_COMM_ID_BYTES];
    if ((rc = crgpu_comm_unique_id(id))) return rc;
    const int n = (int)g->ctx.size();
    // ncclCommInitRank is a collective over the ranks: one thread per device
    if ((rc = group_parallel(g, [&](int i) { return crgpu_comm_init(g->ctx[i], id, n, i, capacity); }))) return rc;
  }
  return group_parallel(g, [&](int i) { return crgpu_sharded_run(g->ctx[i]); });
}

int crgpu_group_matrix_dims(crgpu_group* g, uint64_t* n_barcodes, uint64_t* nnz, uint64_t* n_features) {
  if (!g) return fail(CRGPU_E_INVALID, "group is NULL");
  uint64_t nb = 0, nz = 0, nf = 0;
  for (auto* c : g->ctx) {
    uint64_t a = 0, b = 0, f = 0;
    int rc = crgpu_matrix_dims(c, &a, &b, &f);
    if (rc) return rc;
    nb += a;
    nz += b;
    nf = f;
  }
  if (n_barcodes) *n_barcodes = nb;
  if (nnz) *nnz = nz;
  if (n_features) *n_features = nf;
  return CRGPU_OK;
}

int crgpu_group_matrix_get(crgpu_group* g, uint32_t* barcode_rank, int64_t* indptr, uint32_t* indices, int32_t* data) {
  if (!g) return fail(CRGPU_E_INVALID, "group is NULL");
  uint64_t col = 0, ent = 0;
  for (auto* c : g->ctx) {
    uint64_t nb = 0, nz = 0;
    int rc = crgpu_matrix_dims(c, &nb, &nz, nullptr);
    if (rc) return rc;
    // a device's indptr starts at 0: shift it by the entries of the devices before it (the last offset of one
    // block is the first of the next, so the blocks overlap by one element)
    std::vector<int64_t> ip(indptr ? nb + 1 : 0);
    rc = crgpu_matrix_get(c, barcode_rank ? barcode_rank + col : nullptr, indptr ? ip.data() : nullptr,
                          indices ? indices + ent : nullptr, data ? data + ent : nullptr);
    if (rc) return rc;
    if (indptr)
      for (uint64_t i = 0; i <= nb; i++) indptr[col + i] = ip[i] + (int64_t)ent;
    col += nb;
    ent += nz;
  }
  return CRGPU_OK;
}

void crgpu_group_destroy(crgpu_group* g) {
  if (!g) return;
  for (auto* c : g->ctx) crgpu_ctx_destroy(c);
  delete g;
}

}  // extern "C"
