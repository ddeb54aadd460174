// mex_writer.cpp — Matrix Market output of the count matrix (host code on top of the public C ABI).
//
// Restates MtxWriter::{write_matrix_mtx, write_barcodes_tsv, write_features_tsv}
// (cr_lib/src/stages/write_matrix_market.rs:41-120): `matrix.mtx.gz` with the two header lines and the
// dimension line followed by one `feature+1 barcode+1 count` line per entry in (barcode, feature) order,
// `barcodes.tsv.gz` with one `SEQ-<gem group>` line per column (Barcode's Display form) and `features.tsv.gz`
// with the caller's rows. gzip at the fast level, like the reference's flate2::Compression::fast().
#include <sys/stat.h>
#include <zlib.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/crgpu.h"

extern "C" int crgpu_set_error_(int code, const char* msg);  // crgpu.cu

namespace {

struct GzOut {
  gzFile f = nullptr;
  std::vector<char> buf;
  size_t n = 0;
  explicit GzOut(const std::string& path) : buf(1 << 20) { f = gzopen(path.c_str(), "wb1"); }
  ~GzOut() { close(); }
  bool ok() const { return f != nullptr; }
  void flush() {
    if (n) gzwrite(f, buf.data(), (unsigned)n);
    n = 0;
  }
  char* room(size_t need) {
    if (n + need > buf.size()) flush();
    return buf.data() + n;
  }
  void put(const char* s, size_t len) {
    if (len > buf.size()) {
      flush();
      gzwrite(f, s, (unsigned)len);
      return;
    }
    memcpy(room(len), s, len);
    n += len;
  }
  bool close() {
    if (!f) return true;
    flush();
    int rc = gzclose(f);
    f = nullptr;
    return rc == Z_OK;
  }
};

inline char* put_u64(char* p, uint64_t v) {
  char tmp[24];
  int k = 0;
  do {
    tmp[k++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (k) *p++ = tmp[--k];
  return p;
}

}  // namespace

extern "C" int crgpu_matrix_write_mex(crgpu_ctx* ctx, const char* folder, const char* software_version, int gem_group,
                                      const char* features_tsv) {
  if (!ctx || !folder || !software_version) return crgpu_set_error_(CRGPU_E_INVALID, "NULL argument");
  uint64_t n_bc = 0, nnz = 0, n_feat = 0;
  int rc;
  if ((rc = crgpu_matrix_dims(ctx, &n_bc, &nnz, &n_feat))) return rc;
  int L = 0;
  uint64_t n_content = 0;
  if ((rc = crgpu_whitelist_size(ctx, &n_content, &L))) return rc;
  std::vector<uint32_t> rank(n_bc), indices(nnz);
  std::vector<int64_t> indptr(n_bc + 1);
  std::vector<int32_t> data(nnz);
  if ((rc = crgpu_matrix_get(ctx, rank.data(), indptr.data(), indices.data(), data.data()))) return rc;
  std::vector<uint8_t> seqs(n_bc * (size_t)L);
  if (n_bc && (rc = crgpu_barcode_seqs(ctx, rank.data(), n_bc, seqs.data()))) return rc;

  if (mkdir(folder, 0777) != 0 && errno != EEXIST)
    return crgpu_set_error_(CRGPU_E_INVALID, (std::string("cannot create ") + folder + ": " + strerror(errno)).c_str());
  const std::string dir(folder);
  {
    GzOut out(dir + "/matrix.mtx.gz");
    if (!out.ok()) return crgpu_set_error_(CRGPU_E_INVALID, "cannot open matrix.mtx.gz");
    std::string head = "%%MatrixMarket matrix coordinate integer general\n%metadata_json: {\"software_version\": \"";
    head += software_version;
    head += "\", \"format_version\": 2}\n";
    head += std::to_string(n_feat) + " " + std::to_string(n_bc) + " " + std::to_string(nnz) + "\n";
    out.put(head.data(), head.size());
    for (uint64_t c = 0; c < n_bc; c++)
      for (int64_t e = indptr[c]; e < indptr[c + 1]; e++) {  // indices are 1-based
        char* p = out.room(72);
        char* q = put_u64(p, (uint64_t)indices[e] + 1);
        *q++ = ' ';
        q = put_u64(q, c + 1);
        *q++ = ' ';
        q = put_u64(q, (uint64_t)data[e]);
        *q++ = '\n';
        out.n += (size_t)(q - p);
      }
    if (!out.close()) return crgpu_set_error_(CRGPU_E_INVALID, "write error on matrix.mtx.gz");
  }
  {
    GzOut out(dir + "/barcodes.tsv.gz");
    if (!out.ok()) return crgpu_set_error_(CRGPU_E_INVALID, "cannot open barcodes.tsv.gz");
    const std::string suffix = "-" + std::to_string(gem_group) + "\n";
    for (uint64_t c = 0; c < n_bc; c++) {
      out.put(reinterpret_cast<const char*>(seqs.data() + c * (size_t)L), (size_t)L);
      out.put(suffix.data(), suffix.size());
    }
    if (!out.close()) return crgpu_set_error_(CRGPU_E_INVALID, "write error on barcodes.tsv.gz");
  }
  if (features_tsv) {
    GzOut out(dir + "/features.tsv.gz");
    if (!out.ok()) return crgpu_set_error_(CRGPU_E_INVALID, "cannot open features.tsv.gz");
    out.put(features_tsv, strlen(features_tsv));
    if (!out.close()) return crgpu_set_error_(CRGPU_E_INVALID, "write error on features.tsv.gz");
  }
  return CRGPU_OK;
}
