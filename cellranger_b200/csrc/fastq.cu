// fastq.cu — FASTQ text -> fixed-stride read arrays, on the device.
//
// The front end of MAKE_SHARD reads FASTQ records and slices barcode / UMI ranges out of R1
// (RnaProcessor::process_read, cr_types/src/rna_read.rs:363-467; the ranges are chemistry constants,
// extract_barcode :285-368). The kernels of pass 1 want the sequence and quality lines as fixed-stride arrays.
// One kernel does the conversion in a single pass over the text: a tile counts its newlines, a chained scan
// (decoupled look-back) turns that into the line number at the start of the tile, and every line start of kind
// "sequence" (line % 4 == 1) or "quality" (line % 4 == 3) found in the tile is copied - its first `read_len`
// bytes, one warp per line - to the record's slot. No global array of line offsets is ever materialised.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int FQ_THREADS = 256;
constexpr int FQ_BYTES = 32;  // bytes per thread: 8 KB tiles, staged in shared memory for the line copies
constexpr int FQ_TILE = FQ_THREADS * FQ_BYTES;

// bit j set iff byte j of the 16-byte chunk is '\n'
__device__ __forceinline__ uint32_t newline_mask(const uint4 v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t x = w[i] ^ 0x0A0A0A0Au;
    // exact zero-byte detector: bit 7 of a byte is set iff the byte is zero
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
    m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * i);
  }
  return m;
}

__global__ void __launch_bounds__(FQ_THREADS) fastq_extract_kernel(
    const uint8_t* __restrict__ text, uint64_t n_bytes, int read_len, uint8_t* __restrict__ out_seq,
    uint8_t* __restrict__ out_qual, uint64_t capacity, unsigned long long* desc, uint32_t* ticket,
    unsigned long long* counters /* [0] lines, [1] short reads, [2] malformed records */) {
  __shared__ uint32_t scan_s[FQ_THREADS / 32 + 1];
  __shared__ unsigned long long bcast;
  __shared__ uint32_t tile_s;
  __shared__ uint16_t nl_s[FQ_TILE];  // offsets of the tile's newlines, in order
  const uint32_t tile = acquire_tile(ticket, &tile_s);
  const uint64_t tile_base = (uint64_t)tile * FQ_TILE;
  const uint64_t p0 = tile_base + (uint64_t)threadIdx.x * FQ_BYTES;
  // stage the thread's bytes in shared memory and find its newlines
  __shared__ __align__(16) uint8_t text_s[FQ_TILE];
  unsigned long long m = 0ull;
  if (p0 + FQ_BYTES <= n_bytes) {
    const uint4* src = reinterpret_cast<const uint4*>(text + p0);
    uint4* dst = reinterpret_cast<uint4*>(text_s + threadIdx.x * FQ_BYTES);
#pragma unroll
    for (int c = 0; c < FQ_BYTES / 16; c++) {
      const uint4 v = __ldg(src + c);
      dst[c] = v;
      m |= (unsigned long long)newline_mask(v) << (16 * c);
    }
  } else {
    for (int j = 0; j < FQ_BYTES; j++) {
      const uint8_t ch = p0 + j < n_bytes ? text[p0 + j] : (uint8_t)0;
      text_s[threadIdx.x * FQ_BYTES + j] = ch;
      if (ch == '\n') m |= 1ull << j;
    }
  }
  uint32_t n_nl;
  uint32_t off = block_exclusive_scan<FQ_THREADS>((uint32_t)__popcll(m), &n_nl, scan_s);
  while (m) {
    const int j = __ffsll((long long)m) - 1;
    m &= m - 1ull;
    nl_s[off++] = (uint16_t)(threadIdx.x * FQ_BYTES + j);
  }
  const uint64_t n_tiles = (n_bytes + FQ_TILE - 1) / FQ_TILE;
  const unsigned long long before = lookback_exclusive(desc, tile, (unsigned long long)n_nl, &bcast);  // syncs
  if (tile == n_tiles - 1 && threadIdx.x == 0) counters[0] = before + n_nl;
  // byte at tile offset o: from the staged copy, or from global memory past the end of the tile
  auto byte_at = [&](uint32_t o) -> uint8_t { return o < (uint32_t)FQ_TILE ? text_s[o] : text[tile_base + o]; };
  // The line that starts after newline i of the tile has index before + i + 1. Eight lanes per line (four
  // lines per warp instruction - with a whole warp per line the kernel was bound by the scalar per-line logic,
  // 94 warp instructions per line): the group works out where the line goes, its lanes copy the bytes.
  constexpr int GROUP = 8;
  const int lane = threadIdx.x & 31, sub = lane & (GROUP - 1);
  const uint32_t group = threadIdx.x / GROUP, n_groups = FQ_THREADS / GROUP;
  for (uint32_t i = group; i < n_nl; i += n_groups) {
    const uint64_t line = before + i + 1;
    const uint32_t kind = (uint32_t)(line & 3u);
    const uint32_t so = (uint32_t)nl_s[i] + 1u;  // tile offset of the line start
    const uint64_t start = tile_base + so;
    if (start >= n_bytes) continue;  // the newline that ends the text
    if (!(kind & 1u)) {              // header / separator line: only its first byte is looked at
      if (sub == 0 && byte_at(so) != (kind == 0u ? '@' : '+')) atomicAdd(counters + 2, 1ull);
      continue;
    }
    const uint64_t rec = line >> 2;
    if (rec >= capacity) continue;
    // length of the line, without its line terminator
    int len;
    if (i + 1 < n_nl) {
      const uint32_t eo = nl_s[i + 1];
      len = (int)(eo - so);
      if (len > 0 && text_s[eo - 1] == '\r') len--;
    } else {  // the line runs into the next tile: look for its end, read_len bytes at most
      len = 0;
      while (len < read_len && start + len < n_bytes) {
        const uint8_t ch = byte_at(so + len);
        if (ch == '\n' || ch == '\r') break;
        len++;
      }
    }
    uint8_t* dst = (kind == 1u ? out_seq : out_qual) + rec * (uint64_t)read_len;
    const uint8_t pad = kind == 1u ? (uint8_t)'N' : (uint8_t)'#';
    for (int k = sub; k < read_len; k += GROUP) dst[k] = k < len ? byte_at(so + k) : pad;
    if (sub == 0 && kind == 1u && len < read_len) atomicAdd(counters + 1, 1ull);
  }
  // line 0 starts at byte 0 without a newline in front of it
  if (tile == 0 && threadIdx.x == 0 && n_bytes && text[0] != '@') atomicAdd(counters + 2, 1ull);
}

}  // namespace

void fastq_lb_flag_fetch(unsigned int* host_out, cudaStream_t st) {
  cudaMemcpyFromSymbolAsync(host_out, lb_timeout_flag, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st);
}
void fastq_lb_flag_clear(cudaStream_t st) {
  static const unsigned int zero = 0;
  cudaMemcpyToSymbolAsync(lb_timeout_flag, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st);
}

size_t fastq_temp_bytes(uint64_t n_bytes) { return ((n_bytes + FQ_TILE - 1) / FQ_TILE + 1) * 8 + 64; }

// temp: fastq_temp_bytes(n_bytes); counters: 3 x u64 (zeroed here)
int launch_fastq_extract(const uint8_t* text, uint64_t n_bytes, int read_len, uint8_t* out_seq, uint8_t* out_qual,
                         uint64_t capacity, void* temp, unsigned long long* counters, cudaStream_t st) {
  cudaMemsetAsync(counters, 0, 3 * 8, st);
  if (n_bytes == 0) return 0;
  const uint64_t tiles = (n_bytes + FQ_TILE - 1) / FQ_TILE;
  unsigned long long* desc = static_cast<unsigned long long*>(temp);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(desc + tiles + 1);
  cudaMemsetAsync(temp, 0, (tiles + 1) * 8 + 8, st);
  fastq_extract_kernel<<<(unsigned)tiles, FQ_THREADS, 0, st>>>(text, n_bytes, read_len, out_seq, out_qual, capacity, desc,
                                                              tile_ticket(ticket), counters);
  return 1;
}
