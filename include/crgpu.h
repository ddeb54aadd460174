/* crgpu.h — C ABI of the B200-native barcode / UMI correction and counting path.
 *
 * This is the drop-in boundary: a Rust (cgo / JNI / ctypes ...) host binds these
 * symbols and nothing else. Plain pointers and sizes only; the library owns all
 * device memory behind the opaque context; every entry point returns 0 on
 * success or a negative CRGPU_E_* code and never unwinds across the boundary
 * (crgpu_last_error() returns the message of the calling thread's last failure).
 * Calls on one context are serialised by the caller, as the reference's stage
 * `main`s are (lib/rust/cr_lib/src/stages/barcode_correction.rs:265-370 is
 * single threaded). There is no CPU fallback: without a CUDA device every
 * compute entry point fails with CRGPU_E_CUDA.
 *
 * Each group below names the reference interface it replaces (paths relative
 * to the reference checkout, lib/rust/<crate>/src/...). INTEGRATION.md shows
 * the Rust `extern "C"` block and the stage adapters that call it.
 *
 * Sequence conventions: sequences cross the boundary as fixed-width ASCII
 * (A,C,G,T,N), as the reference's SSeqGen does, or packed 2 bits per base with
 * the first base most significant and A=0 C=1 G=2 T=3 (SSeqGen::encode_2bit_u32,
 * used at lib/rust/tx_annotation/src/mark_dups.rs:347). A barcode is reported as
 * its *content rank*: the index of its (translated) sequence in the sorted list
 * of whitelist sequences, which is also the reference's Barcode sort order.
 */
#ifndef CRGPU_H
#define CRGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRGPU_OK 0
#define CRGPU_E_INVALID -1 /* bad argument / call order */
#define CRGPU_E_CUDA -2    /* CUDA runtime failure (or no device) */
#define CRGPU_E_NOMEM -3
#define CRGPU_E_LIMIT -4 /* a documented capacity limit was exceeded */

#define CRGPU_NO_FEATURE 0xFFFFFFFFu
#define CRGPU_NO_RANK 0x3FFFFFFFu

/* BarcodeSegmentState — lib/rust/barcode/src/lib.rs:270-283 (same numbering) */
#define CRGPU_BC_NOT_CHECKED 0
#define CRGPU_BC_VALID_BEFORE_CORRECTION 1
#define CRGPU_BC_VALID_AFTER_CORRECTION 2
#define CRGPU_BC_INVALID 3

/* per-read flag bits returned by crgpu_reads_get(); DupInfo — tx_annotation/src/mark_dups.rs:61-72 */
#define CRGPU_F_UMI_VALID 1u      /* UmiInfo::is_valid, umi/src/info.rs:20-37 */
#define CRGPU_F_HAS_DUPINFO 2u    /* BarcodeDupMarker::process returned Some */
#define CRGPU_F_UMI_CORRECTED 4u  /* DupInfo::is_corrected */
#define CRGPU_F_LOW_SUPPORT 8u    /* DupInfo::is_low_support_umi */
#define CRGPU_F_UMI_COUNT 16u     /* DupInfo::is_umi_count (the representative read) */

typedef struct crgpu_ctx crgpu_ctx;

/* ---- lifecycle ---------------------------------------------------------- */
int crgpu_version(void);
const char* crgpu_last_error(void);
int crgpu_ctx_create(int device, crgpu_ctx** out);
void crgpu_ctx_destroy(crgpu_ctx* ctx);

/* Posterior{max_expected_barcode_errors, bc_confidence_threshold} — barcode/src/corrector.rs:93-109;
 * filter_umis — cr_lib/src/aligner.rs:270. Defaults: 0.975, f64::MAX, 1. */
int crgpu_set_params(crgpu_ctx* ctx, double bc_confidence_threshold, double max_expected_barcode_errors,
                     int filter_umis);
/* Targeted-panel UMI filter: DupBuilder::build(.., targeted_umi_min_read_count) with FeatureReference::target_set
 * (tx_annotation/src/mark_dups.rs:156-170,189-191,311-320; cr_lib/src/aligner.rs:319). A molecule whose feature
 * is on target (on_target[feature] != 0), whose read count after UMI correction is below min_read_count and
 * that is not low support is no UMI count: it enters neither the matrix nor the UmiCount rows, and its reads
 * carry flag bit 5 (is_filtered_target_umi). on_target NULL or min_read_count 0 (the default): no filter. */
int crgpu_set_target_filter(crgpu_ctx* ctx, const uint8_t* on_target, int32_t n_features, uint64_t min_read_count);

/* ---- Whitelist — barcode/src/whitelist.rs:452-525 (Whitelist::{Plain,Trans}) ----
 * seqs: n*L ASCII. translated: NULL for a plain whitelist, else n*L ASCII (raw -> translated).
 * The first whitelist added defines the content space (its content sequences, sorted, define
 * the barcode ranks); later whitelists must map into it. L <= 16. */
int crgpu_whitelist_add(crgpu_ctx* ctx, const uint8_t* seqs, uint64_t n, int L, const uint8_t* translated,
                        int* out_whitelist);
int crgpu_whitelist_size(crgpu_ctx* ctx, uint64_t* out_n_content, int* out_L);
/* content rank -> ASCII sequence (n*L bytes) */
int crgpu_barcode_seqs(crgpu_ctx* ctx, const uint32_t* ranks, uint64_t n, uint8_t* out_ascii);

/* ---- Library type: chemistry read layout + per-type behaviour ----
 * ChemistryDef barcode/umi components — cr_types/src/chemistry/mod.rs:718-751,866-884;
 * umi_correction = 0 for Multiplexing Capture (cr_lib/src/aligner.rs:313-318);
 * a feature-barcode library takes its feature from a tethered fixed-offset capture in R2
 * (cr_types/src/reference/feature_extraction.rs:358-471), a GEX library from the `feature` array. */
typedef struct crgpu_library_def {
  int32_t whitelist;
  int32_t bc_offset, bc_length;   /* barcode = R1[bc_offset : +bc_length] */
  int32_t umi_offset, umi_length; /* UMI     = R1[umi_offset : +umi_length] */
  int32_t umi_correction;
  int32_t is_feature_barcode;
  int32_t feature_type;           /* id of the feature type this library's features carry */
  int32_t fb_offset, fb_length;   /* capture = R2[fb_offset : +fb_length] */
} crgpu_library_def;
int crgpu_library_add(crgpu_ctx* ctx, const crgpu_library_def* def, int* out_library);

/* FeatureReference (only what the path needs): feature_type[f] = 0 for genes, else the
 * feature_type id of the owning library; fb_seqs[f*fb_stride ..] = capture sequence. */
int crgpu_features_set(crgpu_ctx* ctx, int32_t n_features, const int32_t* feature_type, const uint8_t* fb_seqs,
                       int32_t fb_stride);

/* ---- Reads (what RnaRead carries onto the path — cr_types/src/rna_read.rs:379-467) ----
 * Fixed-stride SoA. With on_device != 0 the pointers are device pointers (16-byte aligned),
 * borrowed until crgpu_reads_clear(); otherwise host pointers, copied H2D on the context stream. */
typedef struct crgpu_read_batch {
  uint64_t n;
  int32_t r1_len;
  const uint8_t* r1_seq;
  const uint8_t* r1_qual;
  const uint32_t* feature; /* gene index or CRGPU_NO_FEATURE; NULL for feature-barcode libraries */
  int32_t r2_len;
  const uint8_t* r2_seq;
  const uint8_t* r2_qual;
  int32_t on_device;
  /* optional (NULL: every read is UmiType::Txomic and its qname orders like its index in the context):
   * UmiSelectKey{utype, qname} of each read (tx_annotation/src/mark_dups.rs:110-114) as one order-preserving word -
   * bit 63 = 1 for UmiType::NonTxomic (Txomic < NonTxomic, umi/src/lib.rs:101-107), bits 0..62 = the rank of the
   * read's qname among the qnames of the GEM well (any strictly order-preserving map of the header bytes; equal
   * headers - the mates of one pair - get equal words). The read with the smallest word stands for its
   * (UMI, feature): DupInfo::is_umi_count, and UmiCount::utype of the molecule. Same memory space as the other
   * arrays of the batch. */
  const uint64_t* select_key;
} crgpu_read_batch;
int crgpu_reads_add(crgpu_ctx* ctx, int library, const crgpu_read_batch* batch, int* out_batch);
/* FASTQ front end: the read loop of MAKE_SHARD slices barcode / UMI ranges out of FASTQ records
 * (RnaProcessor::process_read, cr_types/src/rna_read.rs:363-467; ranges from the chemistry, extract_barcode
 * :285-368). crgpu_fastq_extract turns uncompressed 4-line FASTQ text into the fixed-stride arrays of
 * crgpu_read_batch, on the device, in one pass: record r gets bytes [0, read_len) of its sequence and quality
 * lines at dev_seq / dev_qual + r * read_len (device memory, capacity records each). A sequence line shorter
 * than read_len is padded with 'N' (quality '#') and counted in n_short - the reference fails such a read with
 * "Barcode range out of bounds" (ReadPair::check_range), here its barcode simply cannot be valid. n_malformed
 * counts records whose header / separator line does not start with '@' / '+'. text: host memory, or device
 * memory (16-byte aligned) when on_device != 0; a chunk must hold whole records. Feed the arrays to
 * crgpu_reads_add with on_device = 1. */
int crgpu_fastq_extract(crgpu_ctx* ctx, const void* text, uint64_t n_bytes, int on_device, int read_len,
                        uint8_t* dev_seq, uint8_t* dev_qual, uint64_t capacity, uint64_t* n_records, uint64_t* n_short,
                        uint64_t* n_malformed);

int crgpu_reads_clear(crgpu_ctx* ctx);

/* ---- Stage MAKE_SHARD hot loop: exact whitelist check + priors ----
 * RnaProcessor::process_read → Whitelist::check_and_update (cr_types/src/rna_read.rs:285-368),
 * MakeShardHistograms::observe (cr_lib/src/make_shard_metrics.rs:171-188),
 * exact feature-barcode counts (make_shard_metrics.rs:337-345). Resets and fills the priors. */
int crgpu_pass1(crgpu_ctx* ctx);

/* Cross-chunk / cross-GPU state: priors are global per library type
 * (cr_lib/src/stages/make_shard.rs:303-358). Device pointers so the host can all-reduce them. */
int crgpu_prior_dev(crgpu_ctx* ctx, int library, uint32_t** out_dev_u32, uint64_t* out_n);
int crgpu_prior_set(crgpu_ctx* ctx, int library, const uint32_t* host_counts, uint64_t n);
int crgpu_fb_counts_dev(crgpu_ctx* ctx, unsigned long long** out_dev_u64, int32_t* out_n);

/* ---- Stage BARCODE_CORRECTION: Posterior::correct_barcode over the invalid reads ----
 * barcode/src/corrector.rs:111-165 driven as cr_lib/src/stages/barcode_correction.rs:76-99,327-345;
 * feature-barcode correction with feat_dist (feature_extraction.rs:34-117, feature_checker.rs:8-50).
 * May be repeated (a retry): it restarts from the state crgpu_pass1 left - no corrected reads, the key list ending
 * where pass 1 ended. After crgpu_count it is refused (CRGPU_E_INVALID) until crgpu_pass1 has run again, because
 * the count stage sorts the key buffer in place. */
int crgpu_pass2(crgpu_ctx* ctx);

/* Corrector plugin seam, batch form of trait CorrectBarcode (barcode/src/corrector.rs:73-81):
 * n segments of the library's bc_length (ASCII) with qualities (or NULL = no qualities) against the
 * library's whitelist and its CURRENT priors. out_rank[i] = content rank or CRGPU_NO_RANK;
 * out_state[i] = ValidBeforeCorrection (exact hit), ValidAfterCorrection or Invalid.
 * Divergence from the trait: Posterior::correct_barcode asserts that the caller only hands over segments in state
 * Invalid (corrector.rs:120) and never re-checks membership; the batch form has no state on its input, so a
 * segment that IS on the whitelist comes back as an exact hit (ValidBeforeCorrection) instead of being treated
 * as Invalid. A caller that follows the reference (correct_barcode_in_read, barcode_correction.rs:95-97: invalid
 * segments only) never sees the difference. */
int crgpu_correct_barcodes(crgpu_ctx* ctx, int library, const uint8_t* bc_ascii, const uint8_t* qual,
                           uint64_t n, uint32_t* out_rank, uint8_t* out_state);

/* ---- Barcode-owner sharding (replaces ShardReader::make_chunks ranges,
 * cr_lib/src/stages/align_and_count.rs:519-524) ----
 * After pass2 each context holds the packed keys of its local reads. */
int crgpu_key_layout(crgpu_ctx* ctx, int32_t* rank_shift, int32_t* feature_shift, int32_t* lib_shift,
                     int32_t* umi_bits);
int crgpu_keys_dev(crgpu_ctx* ctx, unsigned long long** out_dev, uint64_t* out_n);
/* reorder the local keys into n_parts contiguous groups by rank range [bounds[p], bounds[p+1]);
 * out_counts[p] = keys in group p */
int crgpu_keys_partition(crgpu_ctx* ctx, int32_t n_parts, const uint32_t* bounds, uint64_t* out_counts);
/* replace the key set by n keys at a device pointer (copied) */
int crgpu_keys_set(crgpu_ctx* ctx, const unsigned long long* dev_keys, uint64_t n);
/* per-library valid-barcode counts (raw valid + corrected) as a device vector to all-reduce:
 * corrected_barcode_counts of BARCODE_CORRECTION join (barcode_correction.rs:401-407) */
int crgpu_valid_counts_dev(crgpu_ctx* ctx, int library, uint32_t** out_dev_u32, uint64_t* out_n);
/* per-library counts of the reads corrected by this context (device vector, to all-reduce), and the
 * recomputation valid = prior + corrected after both have been made global */
int crgpu_corrected_dev(crgpu_ctx* ctx, int library, uint32_t** out_dev_u32, uint64_t* out_n);
int crgpu_valid_counts_refresh(crgpu_ctx* ctx);
/* Fused key exchange over peer memory (NVLink / NVSwitch) instead of partition + all-to-all: every context
 * owns a receive buffer that its peers map through CUDA IPC, and crgpu_keys_scatter_peers() writes each key
 * straight into the buffer of the rank that owns its barcode (one pass over the keys, block-aggregated remote
 * cursor claims, coalesced peer stores). Protocol per step, on every rank:
 *   crgpu_exchange_reset            (before the first collective of the step)
 *   ... pass1, all-reduce of the priors [, crgpu_keys_scatter_peers_begin], pass2, all-reduce ...
 *   crgpu_keys_scatter_peers        (returns when this rank's stores are complete)
 *   <any cross-rank barrier>
 *   crgpu_exchange_finish           (adopts the received keys as this context's key set)
 * handle: 2 * 64 bytes (cudaIpcMemHandle_t of the buffer and of its cursor). */
#define CRGPU_IPC_HANDLE_BYTES 128
int crgpu_exchange_init(crgpu_ctx* ctx, uint64_t capacity_keys, void* out_handle);
int crgpu_exchange_connect(crgpu_ctx* ctx, int32_t n_ranks, int32_t my_rank, const void* handles);
int crgpu_exchange_reset(crgpu_ctx* ctx);
int crgpu_keys_scatter_peers(crgpu_ctx* ctx, int32_t n_parts, const uint32_t* bounds, uint64_t* out_sent);
/* Optional early part: called between crgpu_pass1 and crgpu_pass2 (once the bounds are known - e.g. from the
 * all-reduced priors), it sends the keys pass 1 emitted on a second stream, so that the NVLink traffic runs
 * under pass 2; the later crgpu_keys_scatter_peers (same bounds) then sends only the keys of pass 2 and waits
 * for both parts. */
int crgpu_keys_scatter_peers_begin(crgpu_ctx* ctx, int32_t n_parts, const uint32_t* bounds);
int crgpu_exchange_finish(crgpu_ctx* ctx, uint64_t* out_received);
/* restrict the matrix columns this context owns to content ranks [lo, hi) */
int crgpu_set_owned_range(crgpu_ctx* ctx, uint32_t lo, uint32_t hi);

/* ---- The sharded run behind the boundary ----
 * What a stage `main` of the reference would call (one call per stage, no scripting language in the step:
 * cr_lib/src/stages/align_and_count.rs:552-789 is one Rust function). The library owns the communicator (NCCL,
 * loaded at run time; none of this is needed - or loaded - for a single GPU), the owner ranges and the key
 * exchange; everything of a step is enqueued on the context stream, with one host round trip (the received key
 * count) before the count stage:
 *
 *   pass 1 -> all-reduce(sum) of the priors and exact feature-barcode counts   (priors are global per library
 *             type, cr_lib/src/stages/make_shard.rs:303-358)
 *          -> pass 2 -> all-reduce(sum) of the corrected-read counts (barcode_correction.rs:401-407)
 *          -> owner ranges of contiguous content ranks balanced by valid reads, computed on the device (the
 *             analogue of ShardReader::make_chunks, align_and_count.rs:519-524)
 *          -> every key stored straight into its owner's receive buffer over NVLink (peer pointers: same
 *             process, or CUDA IPC across processes) -> on-stream barrier -> count on the owned range.
 * With CRGPU_EARLY_SCATTER=1 in the environment the owner ranges are taken from the (global) priors alone and
 * the keys of pass 1 travel on a second stream while pass 2 runs.
 *
 * Per-process form (one process per GPU, e.g. under torchrun or MPI): rank 0 calls crgpu_comm_unique_id and
 * hands the 128 bytes to every rank by whatever means the host has; every rank then calls crgpu_comm_init
 * (collective) once and crgpu_sharded_run (collective) per step. exchange_capacity_keys = the most keys this
 * rank may own (CRGPU_E_LIMIT from crgpu_sharded_run when exceeded). */
#define CRGPU_COMM_ID_BYTES 128
int crgpu_comm_unique_id(void* out_id);
int crgpu_comm_init(crgpu_ctx* ctx, const void* id, int32_t n_ranks, int32_t rank, uint64_t exchange_capacity_keys);
int crgpu_sharded_run(crgpu_ctx* ctx);
/* owner ranges of the last sharded run (n_ranks + 1 values) and its exchange statistics:
 * out4 = {keys this rank stored into other ranks' buffers, keys this rank received (its own included),
 *         n_ranks, rank} */
int crgpu_owner_bounds_get(crgpu_ctx* ctx, uint32_t* out_bounds, int32_t cap);
int crgpu_shard_stats(crgpu_ctx* ctx, uint64_t out4[4]);
/* owner ranges for a vector of per-rank read counts, on the device, with the arithmetic of the sharded run
 * (exposed for tests and for hosts that keep their own exchange): counts u32[n] host -> out_bounds[n_parts+1] */
int crgpu_owner_bounds_compute(crgpu_ctx* ctx, const uint32_t* host_counts, uint64_t n, int32_t n_parts,
                               uint32_t* out_bounds);

/* Single-process form: one context per device and one host thread per device inside the library. Set each
 * device's whitelist / libraries / features / reads through crgpu_group_ctx(group, i) exactly as for a single
 * context (the same whitelist and libraries on every device; the reads split in any way), then crgpu_group_run
 * does the whole step on all devices. The matrix is the concatenation of the devices' column blocks in device
 * order (barcode ranges are contiguous and ascending, as the reference's chunk outputs are when joined,
 * barcode_correction.rs:252-262): crgpu_group_matrix_get returns it as one CSC. */
typedef struct crgpu_group crgpu_group;
int crgpu_group_create(const int32_t* devices, int32_t n_devices, uint64_t exchange_capacity_keys, crgpu_group** out);
crgpu_ctx* crgpu_group_ctx(crgpu_group* group, int32_t i);
int crgpu_group_size(crgpu_group* group);
int crgpu_group_run(crgpu_group* group);
int crgpu_group_matrix_dims(crgpu_group* group, uint64_t* n_barcodes, uint64_t* nnz, uint64_t* n_features);
int crgpu_group_matrix_get(crgpu_group* group, uint32_t* barcode_rank, int64_t* indptr, uint32_t* indices, int32_t* data);
void crgpu_group_destroy(crgpu_group* group);

/* ---- Stage ALIGN_AND_COUNT (dedup part) + matrix ----
 * DupBuilder::observe, correct_umis, determine_low_support_umigenes, BarcodeDupMarker::{new,process}
 * (tx_annotation/src/mark_dups.rs:19-59,87-108,116-170,201-363), BcUmiInfo::feature_counts
 * (cr_types/src/types.rs:180-188), BarcodeIndex (cr_types/src/barcode_index.rs:39-53),
 * CSC assembly (cr_h5/src/count_matrix.rs:382-448). */
int crgpu_count(crgpu_ctx* ctx);
/* per-read DupInfo (corrected UMI, flags); optional, needs crgpu_count() first */
int crgpu_annotate_reads(crgpu_ctx* ctx);

/* pass1 + pass2 + count on one device */
int crgpu_run(crgpu_ctx* ctx);
int crgpu_sync(crgpu_ctx* ctx);

/* ---- Results ---- */
enum {
  CRGPU_STAT_READS = 0,
  CRGPU_STAT_VALID_BEFORE = 1,
  CRGPU_STAT_CORRECTED = 2,
  CRGPU_STAT_INVALID = 3,
  CRGPU_STAT_KEYS = 4,          /* reads entering dedup (valid barcode, valid UMI, feature) */
  CRGPU_STAT_DISTINCT_KEYS = 5, /* distinct (barcode, feature, library, raw UMI) */
  CRGPU_STAT_UMI_CORRECTED_KEYS = 6,
  CRGPU_STAT_LOW_SUPPORT_KEYS = 7,
  CRGPU_STAT_MOLECULES = 8, /* UmiCount rows = total UMIs in the matrix */
  CRGPU_STAT_NNZ = 9,
  CRGPU_STAT_BARCODES = 10,
  CRGPU_STAT_KERNEL_LAUNCHES = 11, /* kernels of this library launched so far */
  CRGPU_STAT_UMI_CORRECTED_READS = 12,
  CRGPU_STAT_LOW_SUPPORT_READS = 13,
  CRGPU_STAT_SORT_VIOLATIONS = 14, /* with CRGPU_VERIFY=1 in the environment: order violations found after */
  CRGPU_STAT_RLE_VIOLATIONS = 15,  /* the radix sort and after the run-length encoding (must be 0) */
  CRGPU_STAT_FILTERED_TARGET_UMIS = 16, /* molecules dropped by crgpu_set_target_filter */
  CRGPU_STAT_COUNT = 17
};
int crgpu_stats(crgpu_ctx* ctx, uint64_t out[CRGPU_STAT_COUNT]);

/* per-read results of one batch; any pointer may be NULL. bc_rank = content rank or CRGPU_NO_RANK;
 * umi = processed UMI, 2-bit packed (raw UMI when there is no DupInfo; undefined bits when the UMI
 * holds an N); feature = resolved feature index. */
int crgpu_reads_get(crgpu_ctx* ctx, int batch, uint32_t* bc_rank, uint8_t* bc_state, uint32_t* umi,
                    uint8_t* flags, uint32_t* feature);

/* total_barcode_counts of BARCODE_CORRECTION (cr_lib/src/stages/barcode_correction.rs:327-362,401-407): the reads
 * that were not valid before correction, counted under their barcode after correction - a whitelist barcode
 * (valid = 1) or, when the correction failed, the raw sequence (valid = 0, non-ACGT bases as 'N') - without the
 * entries seen fewer than min_reads_to_report_bc times. One histogram over all library types, the whole read set
 * as one chunk. Needs crgpu_pass2. Entries in Barcode order: the invalid sequences first, each group ascending.
 * crgpu_total_barcode_counts computes and returns the number of entries; _get copies them out
 * (seqs: n * L bytes). */
int crgpu_total_barcode_counts(crgpu_ctx* ctx, uint64_t min_reads_to_report_bc, uint64_t* out_n);
int crgpu_total_barcode_counts_get(crgpu_ctx* ctx, uint8_t* seqs, uint8_t* valid, uint64_t* counts);

/* which: 0 = prior (reads valid before correction), 1 = corrected reads; n = content size */
int crgpu_bc_counts_get(crgpu_ctx* ctx, int library, int which, uint32_t* out, uint64_t n);
int crgpu_fb_counts_get(crgpu_ctx* ctx, int64_t* out, int32_t n);

/* feature x barcode matrix in CSC (cr_h5/src/count_matrix.rs:24-55): barcode_rank[n_barcodes] sorted,
 * indptr[n_barcodes+1], indices[nnz] (feature index), data[nnz] */
int crgpu_matrix_dims(crgpu_ctx* ctx, uint64_t* n_barcodes, uint64_t* nnz, uint64_t* n_features);
int crgpu_matrix_get(crgpu_ctx* ctx, uint32_t* barcode_rank, int64_t* indptr, uint32_t* indices, int32_t* data);

/* raw_feature_bc_matrix in Matrix Market form: MtxWriter::{write_matrix_mtx, write_barcodes_tsv,
 * write_features_tsv} (cr_lib/src/stages/write_matrix_market.rs:41-120). Writes folder/matrix.mtx.gz
 * ("%%MatrixMarket matrix coordinate integer general", the %metadata_json line with software_version =
 * "<product> <version>" as the caller passes it, "features barcodes nnz", then "feature+1 barcode+1 count" per
 * entry in (barcode, feature) order), folder/barcodes.tsv.gz ("SEQ-<gem_group>" per column) and, when
 * features_tsv is not NULL, folder/features.tsv.gz with that text (id, name, feature type per row). */
int crgpu_matrix_write_mex(crgpu_ctx* ctx, const char* folder, const char* software_version, int gem_group,
                           const char* features_tsv);

/* BarcodeSummary rows (cr_lib/src/aligner.rs:33-68, accumulated per valid barcode and library type by
 * visit_read_annotation, cr_lib/src/align_metrics.rs:705-721; written as barcode_summary.csv by ALIGN_AND_COUNT,
 * stages/align_and_count.rs:806-817). out: uint32[n_barcodes][4] in matrix column order =
 * {reads, umis, candidate_dup_reads, umi_corrected_reads} of `library`; the reference emits a row only where
 * reads > 0. Needs crgpu_count() first. */
int crgpu_barcode_summary(crgpu_ctx* ctx, int library, uint32_t* out);

/* BarcodeDiversityMetrics of one library type (BARCODE_CORRECTION join, cr_lib/src/stages/barcode_correction.rs:
 * 428-441): barcodes_detected = valid barcodes with at least one read (raw valid + corrected),
 * effective_barcode_diversity = inverse Simpson index of their read counts (SimpleHistogram::effective_diversity,
 * metric/src/histogram.rs:161-171). After a sharded run the counts are the global ones on every rank. */
int crgpu_barcode_diversity(crgpu_ctx* ctx, int library, uint64_t* barcodes_detected, double* effective_diversity);

/* UmiCount rows (cr_types/src/types.rs:148-160) in the order ALIGN_AND_COUNT hands them to molecule_info: by
 * barcode, and inside a barcode as umi_counts.sort() leaves them (stages/align_and_count.rs:314) - library_idx,
 * feature_idx, umi, read_count. out6[6*i..] = {barcode column index, library_idx, feature_idx, umi 2-bit,
 * read_count, umi_type as molecule_info stores it (UmiType::to_u32: 1 = Txomic, 0 = NonTxomic)}; probe_idx is
 * None on this path (no probe alignments). umi_type is that of the representative read, so when a batch
 * carries select keys crgpu_annotate_reads must have run; without select keys every molecule is Txomic. */
int crgpu_molecules_count(crgpu_ctx* ctx, uint64_t* n);
int crgpu_molecules_get(crgpu_ctx* ctx, uint32_t* out6);

/* ---- Other users of the resident whitelist ----
 * CHECK_BARCODES_COMPATIBILITY (cr_lib/src/stages/check_barcodes_compatibility.rs:98-262) and the whitelist match
 * rate of chemistry detection (cr_lib/src/detect_chemistry/whitelist_filter.rs:61-110,162-193). Both look every
 * read's barcode up with Whitelist::match_to_whitelist (barcode/src/whitelist.rs:526-545): an exact hit, or - for a
 * barcode with exactly one N - the first of A,C,G,T at that position that gives a hit.
 *
 * crgpu_sample_valid_barcodes: sample_valid_barcodes(). seqs = n records of `stride` bytes with the barcode at
 * bc_offset (host, or device with on_device != 0); dev_hist = device uint32[crgpu_whitelist_entries] (from
 * crgpu_dev_alloc, zeroed with crgpu_dev_memset) indexed by the entry's position in the sorted raw sequences of
 * `whitelist`, ACCUMULATED into (several FASTQs of one library type merge, :197-204); *n_in_whitelist = matched
 * reads of this call, so n_in_whitelist / n is WhitelistMatchStats::fraction(). The caller stops at
 * MAX_READS_BARCODE_COMPATIBILITY = 1 000 000 reads as the reference does (:79,114).
 * crgpu_hist_nx: stats::nx::nx (stats/src/nx.rs:6-38) over the non-zero counts; 0 for an empty histogram.
 * crgpu_robust_cosine_similarity: robust_cosine_similarity(a, b) (:122-158) with counts capped at their N92.5;
 * translate_whitelist >= 0 maps b's keys through that translation whitelist first (this_hist.map_key(translate),
 * :241-242; the plain whitelist must be the context's first whitelist). A library is to be translated when the
 * translated similarity is the larger one (:243-246); the caller compares with min_barcode_similarity (0.1). */
int crgpu_whitelist_entries(crgpu_ctx* ctx, int whitelist, uint64_t* n_entries);
int crgpu_dev_memset(crgpu_ctx* ctx, void* dev, int value, uint64_t bytes);
int crgpu_sample_valid_barcodes(crgpu_ctx* ctx, int whitelist, const uint8_t* seqs, uint64_t n, int32_t stride,
                                int32_t bc_offset, int on_device, uint32_t* dev_hist, uint64_t* n_in_whitelist);
int crgpu_hist_nx(crgpu_ctx* ctx, const uint32_t* dev_hist, uint64_t n, double fraction, uint32_t* out);
int crgpu_robust_cosine_similarity(crgpu_ctx* ctx, const uint32_t* dev_hist_a, const uint32_t* dev_hist_b, uint64_t n,
                                   int translate_whitelist, double* out);

/* ---- Synthetic workload generator (bench / tests; see cellranger_b200/synth.py) ---- */
typedef struct crgpu_synth_params {
  uint64_t seed_mix, seed_mol;
  uint32_t n_whitelist, n_cells, n_genes, n_fb;
  int32_t bc_len, umi_len, fb_offset, fb_len, is_fb;
  uint32_t ambient_thr, unmapped_thr, bc_err_thr, umi_err_thr, n_thr, fb_err_thr, homopolymer_thr;
  int32_t n_qual_ascii;
  int32_t n_qual_classes, n_equal_classes;
  uint32_t qual_thr[8], equal_thr[8];
  uint8_t qual_val[8], equal_val[8];
  /* host tables, copied to the device by the call */
  const uint32_t* wl_packed;   /* [n_whitelist] true barcode sequence per content rank (raw for FB) */
  const uint32_t* cell_rank;   /* [n_cells] */
  const uint32_t* cell_cdf;    /* [n_cells] */
  const uint32_t* n_mol;       /* [n_cells] */
  const uint32_t* gene_cdf;    /* [n_genes] */
  const uint32_t* fb_cdf;      /* [n_fb] */
  const uint32_t* fb_packed;   /* [n_fb] */
} crgpu_synth_params;
/* fills device buffers (allocated by the caller through crgpu_dev_alloc) with reads [start, start+n) */
int crgpu_synth_generate(crgpu_ctx* ctx, const crgpu_synth_params* p, uint64_t start, uint64_t n, uint8_t* dev_r1_seq,
                         uint8_t* dev_r1_qual, uint32_t* dev_feature, uint8_t* dev_r2_seq, uint8_t* dev_r2_qual);
int crgpu_dev_alloc(crgpu_ctx* ctx, uint64_t bytes, void** out_dev);
int crgpu_dev_free(crgpu_ctx* ctx, void* dev);
int crgpu_memcpy_d2h(crgpu_ctx* ctx, void* host, const void* dev, uint64_t bytes);
int crgpu_memcpy_h2d(crgpu_ctx* ctx, void* dev, const void* host, uint64_t bytes);
int crgpu_host_alloc_pinned(uint64_t bytes, void** out_host);
int crgpu_host_free_pinned(void* host);

/* timing of the last crgpu_pass1 / pass2 / count calls, measured with CUDA events on the context
 * stream: out_ms[0..n) per named phase; names are returned as a NUL-separated list */
int crgpu_phase_times(crgpu_ctx* ctx, float* out_ms, int32_t cap, int32_t* out_n, const char** out_names);
/* the CUDA stream all work of this context is enqueued on (cudaStream_t as void*) */
int crgpu_stream(crgpu_ctx* ctx, void** out_stream);

#ifdef __cplusplus
}
#endif
#endif /* CRGPU_H */
