/* host_c_sharded.c — the sharded run driven from plain C through include/crgpu.h: no Python, no torch, nothing
 * but the C ABI between the caller and the GPUs. This is the shape of the call a Rust stage `main`
 * (lib/rust/cr_lib/src/stages/align_and_count.rs:552-789 in the reference) makes through its `extern "C"` block.
 *
 *   host_c_sharded [n_devices [n_reads]]
 *
 * A synthetic GEM well (random 16-mer whitelist, cells with PCR duplicates, 1 % substitutions in barcode and UMI)
 * is counted (a) on device 0 alone and (b) by a crgpu_group over n_devices with the reads split evenly. The two
 * matrices must be identical - barcode index, indptr, indices, data: a barcode-owner sharded run is a partition
 * of the single-device run (the reference's chunk outputs concatenate the same way,
 * lib/rust/cr_lib/src/stages/barcode_correction.rs:252-262). Prints OK and exits 0 on success; exits 2 when
 * fewer than n_devices GPUs are present.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crgpu.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { /* xorshift64* */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (uint32_t)((rng_state * 2685821657736338717ull) >> 32);
}

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != CRGPU_OK) {                                                  \
      printf("%s failed (%d): %s\n", #call, rc__, crgpu_last_error());       \
      return rc__ == CRGPU_E_CUDA ? 2 : 1;                                   \
    }                                                                        \
  } while (0)

enum { BC = 16, UMI = 12, R1 = 28, N_WL = 50000, N_CELLS = 300, N_GENES = 2000 };

static int setup(crgpu_ctx* c, const uint8_t* wl) {
  int wl_id = -1, lib = -1;
  crgpu_library_def def;
  CHECK(crgpu_whitelist_add(c, wl, N_WL, BC, NULL, &wl_id));
  memset(&def, 0, sizeof(def));
  def.whitelist = wl_id;
  def.bc_offset = 0, def.bc_length = BC, def.umi_offset = BC, def.umi_length = UMI;
  def.umi_correction = 1;
  CHECK(crgpu_library_add(c, &def, &lib));
  CHECK(crgpu_features_set(c, N_GENES, NULL, NULL, 0));
  return 0;
}

static int add(crgpu_ctx* c, const uint8_t* seq, const uint8_t* qual, const uint32_t* feat, uint64_t lo, uint64_t hi) {
  crgpu_read_batch rb;
  memset(&rb, 0, sizeof(rb));
  rb.n = hi - lo;
  rb.r1_len = R1;
  rb.r1_seq = seq + lo * R1;
  rb.r1_qual = qual + lo * R1;
  rb.feature = feat + lo;
  CHECK(crgpu_reads_add(c, 0, &rb, NULL));
  return 0;
}

int main(int argc, char** argv) {
  const int n_dev = argc > 1 ? atoi(argv[1]) : 2;
  const uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 400000;
  static const char B[4] = {'A', 'C', 'G', 'T'};
  uint8_t* wl = (uint8_t*)malloc((size_t)N_WL * BC);
  uint8_t* seq = (uint8_t*)malloc(n * R1);
  uint8_t* qual = (uint8_t*)malloc(n * R1);
  uint32_t* feat = (uint32_t*)malloc(n * 4);
  uint64_t i;
  int k, d, rc;
  if (n_dev < 1 || n_dev > 16) return 1;
  for (i = 0; i < (uint64_t)N_WL * BC; i++) wl[i] = (uint8_t)B[rnd() & 3];
  for (i = 0; i < n; i++) {
    /* 90 % of the reads from N_CELLS cells, the rest ambient; ~4 reads per molecule */
    const uint32_t cell = (rnd() % 10) ? rnd() % N_CELLS : rnd() % N_WL;
    const uint32_t mol = rnd() % 1500, gene = (mol * 2654435761u) % N_GENES;
    uint64_t u = ((uint64_t)cell * 1315423911u + mol) * 0x9E3779B97F4A7C15ull;
    memcpy(seq + i * R1, wl + (size_t)cell * BC, BC);
    for (k = 0; k < UMI; k++, u >>= 2) seq[i * R1 + BC + k] = (uint8_t)B[u & 3];
    for (k = 0; k < R1; k++) {
      qual[i * R1 + k] = 'I';
      if (rnd() % 100 == 0) { /* a substitution, with a low quality value */
        seq[i * R1 + k] = (uint8_t)B[rnd() & 3];
        qual[i * R1 + k] = '5';
      }
    }
    feat[i] = (rnd() % 8) ? gene : CRGPU_NO_FEATURE;
  }

  /* (a) one device */
  crgpu_ctx* one = NULL;
  rc = crgpu_ctx_create(0, &one);
  if (rc != CRGPU_OK) {
    printf("crgpu_ctx_create failed (%d): %s\n", rc, crgpu_last_error());
    return 2;
  }
  if ((rc = setup(one, wl)) || (rc = add(one, seq, qual, feat, 0, n))) return rc;
  CHECK(crgpu_run(one));
  uint64_t nb1 = 0, nnz1 = 0, nf = 0;
  CHECK(crgpu_matrix_dims(one, &nb1, &nnz1, &nf));
  uint32_t* rank1 = (uint32_t*)malloc((nb1 + 1) * 4);
  int64_t* ptr1 = (int64_t*)malloc((nb1 + 1) * 8);
  uint32_t* idx1 = (uint32_t*)malloc((nnz1 + 1) * 4);
  int32_t* dat1 = (int32_t*)malloc((nnz1 + 1) * 4);
  CHECK(crgpu_matrix_get(one, rank1, ptr1, idx1, dat1));
  printf("1 device : %llu barcodes, %llu entries\n", (unsigned long long)nb1, (unsigned long long)nnz1);
  crgpu_ctx_destroy(one);

  /* (b) the group: the same reads split evenly, the whole step inside the library */
  int32_t devices[16];
  for (d = 0; d < n_dev; d++) devices[d] = d;
  crgpu_group* g = NULL;
  rc = crgpu_group_create(devices, n_dev, n, &g);
  if (rc != CRGPU_OK) {
    printf("crgpu_group_create failed (%d): %s\n", rc, crgpu_last_error());
    return rc == CRGPU_E_CUDA || rc == CRGPU_E_INVALID ? 2 : 1;
  }
  for (d = 0; d < n_dev; d++) {
    crgpu_ctx* c = crgpu_group_ctx(g, d);
    if ((rc = setup(c, wl)) || (rc = add(c, seq, qual, feat, n * d / n_dev, n * (d + 1) / n_dev))) return rc;
  }
  for (k = 0; k < 2; k++) CHECK(crgpu_group_run(g)); /* twice: the second step reuses every buffer */
  uint64_t nb2 = 0, nnz2 = 0;
  CHECK(crgpu_group_matrix_dims(g, &nb2, &nnz2, NULL));
  printf("%d devices: %llu barcodes, %llu entries\n", n_dev, (unsigned long long)nb2, (unsigned long long)nnz2);
  if (nb2 != nb1 || nnz2 != nnz1) {
    printf("MISMATCH: matrix dimensions differ\n");
    return 1;
  }
  uint32_t* rank2 = (uint32_t*)malloc((nb2 + 1) * 4);
  int64_t* ptr2 = (int64_t*)malloc((nb2 + 1) * 8);
  uint32_t* idx2 = (uint32_t*)malloc((nnz2 + 1) * 4);
  int32_t* dat2 = (int32_t*)malloc((nnz2 + 1) * 4);
  CHECK(crgpu_group_matrix_get(g, rank2, ptr2, idx2, dat2));
  if (memcmp(rank1, rank2, nb1 * 4) || memcmp(ptr1, ptr2, (nb1 + 1) * 8) || memcmp(idx1, idx2, nnz1 * 4) ||
      memcmp(dat1, dat2, nnz1 * 4)) {
    printf("MISMATCH: the sharded matrix differs from the single-device matrix\n");
    return 1;
  }
  for (d = 0; d < n_dev; d++) {
    uint64_t st[4];
    CHECK(crgpu_shard_stats(crgpu_group_ctx(g, d), st));
    printf("  device %d: received %llu keys, stored %llu into peers\n", d, (unsigned long long)st[1],
           (unsigned long long)st[0]);
  }
  crgpu_group_destroy(g);
  printf("OK\n");
  return 0;
}
