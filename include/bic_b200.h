/*
 * bic_b200.h -- C ABI of libbic_b200.so: the B200 (sm_100a) implementation of the encoder hot
 * path of nacho-pancho/binary-image-compression (bit-packed binary matrix factorisation
 * "bsvd" + Golomb / EG coding of the factors and the residual).
 *
 * The reference has no FFI (it is one C++ program): its plug points are the global
 * function pointers of src/bsvd.h:104-125 and the classes of src/binmat.h, src/GolombCoder.h,
 * src/eg.h. The host-side C++ shim that keeps those names (binary-image-compression_b200/host/)
 * is a thin layer over the entry points below; each entry point cites the reference interface
 * it stands in for (paths relative to /root/reference/).
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes; every call returns a bic_status.
 *   - one caller thread per context; a context owns one CUDA stream; calls are ordered on it.
 *     Calls that return values to the host synchronise that stream, the others are async.
 *   - host matrices use the reference's word layout (src/binmat.h:114-116, src/binmat.cpp:140-149):
 *     row-major uint64 words, ceil(cols/64) per row, bit j of a row at (1<<63) >> (j % 64).
 *     Pad bits are ignored on upload and zero on download.
 *   - device layout (private): row-major uint32 words, ceil(cols/32) per row, MSB first, pad
 *     bits zero; 8x8/16x16/32x32 patches therefore occupy 8/32/128 contiguous bytes and a
 *     32-atom coefficient row 4 bytes (the reference pads it to 8).
 *   - there is NO CPU fallback: without a CUDA device bic_ctx_create fails with
 *     BIC_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef BIC_B200_H
#define BIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int bic_status;
enum {
  BIC_OK = 0,
  BIC_ERR_INVALID = 1,   /* bad argument / shape mismatch (the reference asserts, src/binmat.cpp:465-469) */
  BIC_ERR_CUDA = 2,      /* a CUDA runtime call failed; see bic_ctx_last_error */
  BIC_ERR_NOMEM = 3,
  BIC_ERR_CAPACITY = 4,  /* caller buffer too small; the needed size is still reported */
  BIC_ERR_NO_DEVICE = 5,
  BIC_ERR_CORRUPT = 6,   /* undecodable stream / container */
  BIC_ERR_UNSUPPORTED = 7
};

typedef struct bic_ctx bic_ctx;       /* device + stream + scratch */
typedef struct bic_mat bic_mat;       /* device bit matrix (stands in for binary_matrix, src/binmat.h:29) */
typedef struct bic_stream bic_stream; /* device-resident coded bit stream + chunk index */

/* ---------------------------------------------------------------- context */
bic_status bic_ctx_create(int device, bic_ctx** out);
/* same, but enqueue on a stream the caller owns (a cudaStream_t), e.g. torch's current stream */
bic_status bic_ctx_create_on_stream(int device, void* cuda_stream, bic_ctx** out);
bic_status bic_ctx_destroy(bic_ctx* ctx);
bic_status bic_ctx_sync(bic_ctx* ctx);
const char* bic_ctx_last_error(bic_ctx* ctx);
const char* bic_status_string(bic_status s);
void* bic_ctx_cuda_stream(bic_ctx* ctx);
int bic_ctx_sm_count(bic_ctx* ctx);
/* make `waiter`'s stream wait for everything queued so far on `signal`'s stream (for callers
 * that run several contexts -- one per independent page -- concurrently) */
bic_status bic_ctx_wait_ctx(bic_ctx* waiter, bic_ctx* signal);
/* tuning switches (every setting gives the same bits).
 * "dict_algo": how update_dictionary_steepest walks the atoms. 2 (default) = all atom histograms in one pass, then the
 *   in-order atom chain inside ONE thread-block cluster (dict3.cu) where the histograms fit shared memory, else 1;
 *   1 = same histograms, one launch per atom that changes (dict2.cu); 0 = one grid barrier per atom (dict.cu).
 * "dict_update": which dictionary update bic_learn_model_traditional (and the MDL learners through it) calls: 0 (default)
 *   update_dictionary_steepest, 1 update_dictionary_proximus -- the reference's global update_dictionary pointer (-d 0 / -d 1,
 *   src/bsvd.cpp:1235). This one DOES change the result, exactly as the flag does in the reference.
 * "coef_algo": 1 (default) = dictionaries of >= 64 atoms use the weight-sorted warp-per-row coefficient kernel; 0 = always a
 *   lane per row.
 * "chain_cluster": CTAs in dict3.cu's cluster, 1/2/4/8/16 (default 16, 8 where 16 cannot be co-scheduled).
 * "chain_bucket_cap": entries of dict3.cu's per-atom row buckets, -1 (default) = 2 per row; 0 = always scan the list.
 * "gol_algo": 2 (default) = Golomb encoder with wide tiles, scans fused into the passes and register-assembled codewords
 *   (coding2.cu), 1 = the first formulation (coding.cu: counts / scan / lengths / scan / scatter).
 * "gol_list": coding2.cu codes a sparse tile (at most one bit in 64 set) from a list of its ones written by the count pass
 *   instead of re-reading its words: 0 never, 1 (default) for streams long enough for the wide tiles, 2 always.
 * "gol_scan": coding2.cu's scans over the tiles: 0 in the last CTA of the count / length pass, 1 (default) as their own
 *   1024-thread launch for long streams (>= 8192 tiles), 2 always.
 * "gol_presize_pct": the encoders that do not wait for the bit count size the code buffer to this percentage of the input bits
 *   (default 125); a code that does not fit is re-encoded by the exact-size path (values below 100 exist to test that).
 * "gol_onepass": 1 = single-pass Golomb encoder with decoupled look-back, 0 (default) = counts / lengths / scatter.
 * "wait_mode": how the calling thread waits for results: 0 (default) cudaStreamSynchronize, 1 poll + sched_yield
 *   (many contexts / several ranks per box), 2 blocking event. */
bic_status bic_ctx_set_option(bic_ctx* ctx, const char* name, int64_t value);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t bic_ctx_launch_count(bic_ctx* ctx);
/* device timers on the context's stream (cudaEvent pairs): ms between start and stop */
bic_status bic_timer_start(bic_ctx* ctx);
bic_status bic_timer_stop(bic_ctx* ctx, float* ms);
/* optional per-launch device timers: an event pair around every kernel this context launches,
 * accumulated per kernel (used by bench.py for the live roofline; off by default) */
bic_status bic_prof_enable(bic_ctx* ctx, int on);
bic_status bic_prof_reset(bic_ctx* ctx);
int bic_prof_kernel_count(void);
bic_status bic_prof_get(bic_ctx* ctx, int kernel, const char** name, uint64_t* launches, double* total_ms);
/* device-side work counters for the measurement harness (reading waits for the stream and resets the counter):
 * "coef_passes" = row passes over the dictionary done by update_coefficients since the last read (each is p * ceil(m/32)
 * XOR + POPC: the algorithmic popcount work of SURVEY 8d) */
bic_status bic_ctx_read_counter(bic_ctx* ctx, const char* name, uint64_t* value);
/* pinned host memory for callers that want async H2D/D2H */
bic_status bic_host_alloc(size_t bytes, void** out);
bic_status bic_host_free(void* p);

/* ---------------------------------------------------------------- matrices
 * binary_matrix(rows, cols) / allocate / destroy: src/binmat.cpp:140-163, src/binmat.h:178.
 * Unlike the reference (uninitialised words) a new matrix is all zero. */
bic_status bic_mat_create(bic_ctx* ctx, uint64_t rows, uint64_t cols, bic_mat** out);
bic_status bic_mat_destroy(bic_ctx* ctx, bic_mat* m);
uint64_t bic_mat_rows(const bic_mat* m);   /* get_rows, src/binmat.h:84 */
uint64_t bic_mat_cols(const bic_mat* m);   /* get_cols, src/binmat.h:87 */
void* bic_mat_device_ptr(const bic_mat* m);
uint64_t bic_mat_stride_words32(const bic_mat* m);
/* host words in the reference layout <-> device */
bic_status bic_mat_upload_words64(bic_ctx* ctx, bic_mat* m, const uint64_t* host_words);
bic_status bic_mat_download_words64(bic_ctx* ctx, const bic_mat* m, uint64_t* host_words);
/* P4 payload (rows of ceil(cols/8) bytes, MSB first; src/pbm.cpp:29-77) <-> device */
bic_status bic_mat_upload_pbm(bic_ctx* ctx, bic_mat* m, const uint8_t* payload);
bic_status bic_mat_download_pbm(bic_ctx* ctx, const bic_mat* m, uint8_t* payload);
bic_status bic_mat_clear(bic_ctx* ctx, bic_mat* m);                       /* clear(), src/binmat.cpp:165-168 */
bic_status bic_mat_copy(bic_ctx* ctx, const bic_mat* src, bic_mat* dst);  /* copy_to, src/binmat.cpp:195-197 */
/* rows [src_row0, src_row0+nrows) of src -> rows [dst_row0, ...) of dst; equal column counts
 * (row-wise set_submatrix / copy_submatrix_to, src/binmat.cpp:267-298, 373-414) */
bic_status bic_mat_copy_rows(bic_ctx* ctx, const bic_mat* src, uint64_t src_row0, uint64_t nrows, bic_mat* dst,
                             uint64_t dst_row0);
bic_status bic_mat_weight(bic_ctx* ctx, const bic_mat* m, uint64_t* w);   /* weight(), src/binmat.cpp:57-67 */
bic_status bic_mat_dist(bic_ctx* ctx, const bic_mat* a, const bic_mat* b, uint64_t* d); /* dist, :499-512 */
bic_status bic_mat_xor(bic_ctx* ctx, const bic_mat* a, const bic_mat* b, bic_mat* c);   /* add, :463-478 */

/* ---------------------------------------------------------------- bsvd hot path */

/* Patch extraction loop of the driver, src/bsvd_test.cpp:80-99 (copy_submatrix_to
 * src/binmat.cpp:267-298, copy_vectorized_to :306-320, set_row :362-371).
 * X must be ceil(rows/W)*ceil(cols/W) x W*W. Row-major patch order, zero padding, and the
 * reference's read-through into the next raster row for edge tiles when W does not divide 64. */
bic_status bic_extract_patches(bic_ctx* ctx, const bic_mat* raster, uint64_t W, bic_mat* X);
/* inverse (set_vectorized + set_submatrix, src/bsvd_test.cpp:128-139): patches -> raster */
bic_status bic_assemble_patches(bic_ctx* ctx, const bic_mat* X, uint64_t W, bic_mat* raster);

/* GSL rand48 + gsl_rng_uniform_int as used by get_rng / initialize_model_neighbor,
 * src/bsvd.cpp:8-15, :241. The 48-bit state lives in one uint64 owned by the caller. */
void bic_rand48_seed(uint64_t* state, unsigned long seed);
uint64_t bic_rand48_uniform_int(uint64_t* state, uint64_t n);

/* Pivot draw of initialize_model_neighbor (src/bsvd.cpp:239-243): draws until p non-zero rows
 * of X were accepted; consumes the generator exactly like the reference. The zero-row test
 * runs on the device, the (serial) draw on the host. BIC_ERR_INVALID if X is all zero (the
 * reference would never return). */
bic_status bic_draw_pivots(bic_ctx* ctx, const bic_mat* X, uint64_t p, uint64_t* rng_state,
                           uint64_t* pivots_out, uint64_t* ndraws_out);
/* Body of initialize_model_neighbor for a given pivot list (src/bsvd.cpp:237-238, :244-262). */
bic_status bic_initialize_model_neighbor_pivots(bic_ctx* ctx, const bic_mat* X, const uint64_t* pivots,
                                                uint64_t p, bic_mat* D, bic_mat* A);
/* initialize_model_neighbor(E, D, A), src/bsvd.cpp:227-267 (mi_algorithm_t, src/bsvd.h:104-106). */
bic_status bic_initialize_model_neighbor(bic_ctx* ctx, const bic_mat* X, bic_mat* D, bic_mat* A,
                                         uint64_t* rng_state);

/* update_coefficients_omp / _basic (cu_algorithm_t, src/bsvd.h:108-110; src/bsvd.cpp:1029-1107,
 * :399-460). E and A are updated in place; *changed = rows that changed (the exact count of the
 * serial variant; the OpenMP variant's own count is racy, src/bsvd.cpp:1096). */
bic_status bic_update_coefficients(bic_ctx* ctx, bic_mat* E, const bic_mat* D, bic_mat* A, uint64_t* changed);

/* update_dictionary_steepest (du_algorithm_t, src/bsvd.h:112-114; src/bsvd.cpp:463-527): atoms
 * strictly in order, E patched after every changed atom. *changed = atoms that changed. */
bic_status bic_update_dictionary_steepest(bic_ctx* ctx, bic_mat* E, bic_mat* D, const bic_mat* A, uint64_t* changed);

/* E = A*D xor X: mul(A,false,D,false,E); add(E,X,E)  (src/binmat.cpp:516-543, :463-478), as in
 * src/bsvd.cpp:1219-1220 and src/bsvd_test.cpp:153-154. */
bic_status bic_residual(bic_ctx* ctx, const bic_mat* X, const bic_mat* A, const bic_mat* D, bic_mat* E);

/* learn_model_traditional(X, E, D, A) (ml_algorithm_t, src/bsvd.h:116-119; src/bsvd.cpp:1215-1244).
 * trace (optional) receives {changed_coefs, changed_atoms} per iteration, up to trace_cap. */
bic_status bic_learn_model_traditional(bic_ctx* ctx, const bic_mat* X, bic_mat* E, bic_mat* D, bic_mat* A,
                                       uint64_t* iterations, uint64_t* trace, uint64_t trace_cap);

/* learn_model_traditional for a batch of independent fits of identical shape (the bitplanes of one grey
 * image, pages that each get their own dictionary, ...): each problem runs exactly the loop of
 * src/bsvd.cpp:1215-1244 and stops when one of its iterations changes nothing, but every kernel of an
 * iteration is launched once for all problems still running. iterations[b] = problem b's return value.
 * Limits: rows up to 1024 bits and p * ceil32(m) / 8 <= 200 KB (else BIC_ERR_UNSUPPORTED; use the
 * single-problem call). */
bic_status bic_learn_model_traditional_batched(bic_ctx* ctx, uint32_t nprob, const bic_mat* const* X, bic_mat* const* E,
                                               bic_mat* const* D, bic_mat* const* A, uint64_t* iterations);

/* update_dictionary_proximus (du_algorithm_t, catalog slot 1), src/bsvd.cpp:528-729: per atom, in order, a majority vote for
 * the atom over its users alternates with a majority vote for its coefficient column until neither changes. E, D and A are
 * updated in place; *changed = atoms whose row changed. A chain of small launches (parity path; the throughput path is the
 * steepest update). */
bic_status bic_update_dictionary_proximus(bic_ctx* ctx, bic_mat* E, bic_mat* D, bic_mat* A, uint64_t* changed);

/* ---- role-switched learners --------------------------------------------------------------
 * binary_matrix::transpose_to, src/binmat.cpp:199-208: dst (cols x rows) = src' */
bic_status bic_mat_transpose(bic_ctx* ctx, const bic_mat* src, bic_mat* dst);
/* variant 1 / 2 / 3: learn_model_alter1 (src/bsvd.cpp:1245-1311), learn_model_alter2 (:1314-1388), learn_model_alter3
 * (:1391-1434) with the default plug points: the fit's updates alternate with the same updates on the transposed problem
 * (E' = Et, D' = At, A' = Dt). D, A in/out, E out; returns the reference's iteration count. */
bic_status bic_learn_model_alter(bic_ctx* ctx, int variant, const bic_mat* X, bic_mat* E, bic_mat* D, bic_mat* A,
                                 uint64_t* iterations);

/* ---- bit planes of a grey image (bitplane_tool, src/bitplane_tool.cpp:24-39) -------------------------------
 * planes[bi](i, j) = gray(i, j) & (1 << bi) for every mask 1 << bi < maxval (bic_bitplane_count of them); the input is
 * the P5 payload as read_pgm_p5_data reads it (src/pnm.cpp:54-78): one byte per pixel if maxval < 256, else two, high
 * byte first. Header parsing and file I/O stay on the host. */
uint32_t bic_bitplane_count(uint32_t maxval);
bic_status bic_split_bitplanes(bic_ctx* ctx, const uint8_t* p5_payload, uint64_t rows, uint64_t cols, uint32_t maxval,
                               bic_mat* const* planes, uint32_t nplanes);

/* ---- MDL model selection (the learners that call the fit repeatedly) -----------------------
 * universal_codelength, src/coding.cpp:24-32 (host arithmetic: double log2 over integer counts) */
double bic_universal_codelength(unsigned n, unsigned r);
/* model_codelength, src/bsvd.cpp:1438-1461: |E|, the row weights of D and the column weights of A are reduced on
 * the device, the description length is the reference's host expression over them. D = A = NULL: the empty model. */
bic_status bic_model_codelength(bic_ctx* ctx, const bic_mat* E, const bic_mat* D, const bic_mat* A, uint64_t* L);
/* lm = 4: learn_model_mdl_forward_selection (src/bsvd.cpp:1463-1546), 5: learn_model_mdl_backward_selection (:1548-1660),
 * 6: learn_model_mdl_full_search (:1662-1717), with initialize_model_neighbor and learn_model_traditional as the
 * initialiser and inner learner (learn_model_setup(0,0,0,lm,0)). *D (p x m) and *A (n x p) hold the initialised model
 * on entry (for lm = 6 only the row count of *D is used: the largest dictionary tried) and are REPLACED by newly
 * created matrices of the selected size (the old ones are destroyed, as the reference destroy()s and allocate()s
 * them); both are NULL after a backward selection that ends with the empty model. E receives the selected model's
 * residual. rng_state: the rand48 stream the reference keeps in a function-static (may be NULL for lm = 5). */
bic_status bic_learn_model_mdl(bic_ctx* ctx, int lm, const bic_mat* X, bic_mat* E, bic_mat** D, bic_mat** A,
                               uint64_t* rng_state, uint64_t* best_codelength);

/* ---- template matching of the compress*_test experiments -----------------------------------
 * For every W x W patch of a raster (row-major patch order, as the drivers walk them): the earlier window of the image that is
 * closest in Hamming distance (dist(), src/binmat.cpp:499-512, over get_submatrix windows), and the drivers' costing of
 * "difference to that window" against "the patch itself": enumL (enumerative code length, src/compress_test.cpp:37-40) plus
 * one GolombCoder per branch over the weights (src/GolombCoder.cpp:13-34). 1 <= W <= 32. recs: one record per patch (host
 * memory, ceil(rows/W) * ceil(cols/W) entries). */
typedef struct {
  uint64_t besti, bestj, bestd;      /* top-left pixel of the best window and its distance ("besti= bestj= bestd=") */
  uint64_t weight;                   /* P.weight() */
  uint64_t match_len, nomatch_len;   /* "nomatch len= match_len=" */
  uint64_t use_match;                /* 1: the "USE MATCH!" branch */
} bic_match_rec;
typedef struct {
  uint64_t matches, weight_sum;      /* "MATCHES:", sum of bestd over the matches (average_weight before the division) */
  uint64_t bits_match, bits_nomatch; /* golomb_match.bitcount, golomb_nomatch.bitcount */
  double L;                          /* sum of the chosen lengths; the drivers print (L + both bit counts) / 8 */
} bic_match_totals;
double bic_enumL(uint64_t n, uint64_t r);
/* main loop of src/compress_test.cpp:73-141: every patch searches all earlier positions of the (unmodified) image; the
 * first smallest distance in scan order wins. Patches are independent. */
bic_status bic_match_patches_v1(bic_ctx* ctx, const bic_mat* raster, uint64_t W, bic_match_rec* recs, bic_match_totals* totals);
/* main loop of src/compress4_test.cpp:89-171 (flags W T R): window of radius R behind / above the patch scanned backwards,
 * stop at the first distance <= T, and a matched patch is replaced by its residual IN the raster (the driver's diff.pbm), so
 * later patches search the coded image. W must divide 32 and the number of columns. */
bic_status bic_match_patches_v4(bic_ctx* ctx, bic_mat* raster, uint64_t W, uint64_t T, uint64_t R, bic_match_rec* recs,
                                bic_match_totals* totals);

/* ---------------------------------------------------------------- several GPUs: rows sharded, D replicated
 * One process per GPU. Every rank holds a contiguous block of the patch rows (its X, E, A); D is
 * replicated. Integer statistics are combined with NCCL (loaded at run time: the libnccl.so.2 already in
 * the process, else the system one), so the result equals the single-GPU fit of the concatenated rows
 * bit for bit. The 128-byte id comes from bic_comm_unique_id on one rank and reaches the others through
 * whatever the caller uses for plumbing (torch.distributed broadcast, a file, MPI ...). */
typedef struct bic_comm bic_comm;
bic_status bic_comm_unique_id(uint8_t id[128]);
bic_status bic_comm_create(bic_ctx* ctx, int rank, int nranks, const uint8_t id[128], bic_comm** out);
bic_status bic_comm_destroy(bic_ctx* ctx, bic_comm* comm);
uint64_t bic_comm_collective_count(const bic_comm* comm);
/* initialize_model_neighbor over all ranks' rows (src/bsvd.cpp:227-267): allgather of the zero-row
 * bitmaps, the same rand48 replay on every rank (advance rng_state identically everywhere), pivot rows
 * and [column histogram | intersect counts] by allreduce */
bic_status bic_dist_initialize_model_neighbor(bic_ctx* ctx, bic_comm* comm, const bic_mat* X_local, bic_mat* D,
                                              bic_mat* A_local, uint64_t* rng_state);
/* update_dictionary_steepest over all ranks' rows (src/bsvd.cpp:463-527): one allreduce of the atom
 * histograms, then the in-order resolve replicated on every rank with one more allreduce per atom that
 * changes. *changed = changed atoms (same on every rank). */
bic_status bic_dist_update_dictionary_steepest(bic_ctx* ctx, bic_comm* comm, bic_mat* E_local, bic_mat* D,
                                               const bic_mat* A_local, uint64_t* changed);
/* learn_model_traditional over all ranks' rows (src/bsvd.cpp:1215-1244); trace[2i] = changed rows summed
 * over ranks, trace[2i+1] = changed atoms */
bic_status bic_dist_learn_model_traditional(bic_ctx* ctx, bic_comm* comm, const bic_mat* X_local, bic_mat* E_local,
                                            bic_mat* D, bic_mat* A_local, uint64_t* iterations, uint64_t* trace,
                                            uint64_t trace_cap);

/* Golomb coding of a row-sharded matrix (rank order = row order): every rank codes its rows as the exact
 * substring of the ONE global stream the serial coder would write for the whole matrix. `out` receives the
 * shard's bits starting at bit (code_bit_offset & 31) of its buffer, so the global stream is the word-wise OR
 * of the shards placed at 32-bit word (code_bit_offset >> 5); chunk-index entries hold global offsets and
 * `first_chunk` is the global number of the shard's first entry. */
typedef struct {
  uint64_t global_bitcount, global_nsamples;  /* of the whole matrix: GolombCoder::bitcount, popcount + 1 */
  uint64_t code_bit_offset, local_code_bits;  /* where this shard's codewords sit in the global stream */
  uint64_t first_chunk, local_chunks;
} bic_shard_info;
struct bic_stream;
bic_status bic_dist_golomb_encode(bic_ctx* ctx, bic_comm* comm, const bic_mat* M_local, uint32_t chunk_samples,
                                  struct bic_stream* out, bic_shard_info* shard);

/* The same coder for callers with their own plumbing: the prefix state of the shard is an explicit input (what the two
 * allgathers of bic_dist_golomb_encode provide). The rows of M follow `bits_before` bits of the one global matrix that hold
 * `ones_before` ones, the last of them at global bit `last_one_before` (-1: none), and produced `code_bits_before` code bits.
 * closing != 0: M holds the last rows, so the run closed by the virtual one at `total_bits` is written too. out may be NULL
 * (lengths only: shard->local_code_bits). The coder state at the seam is GolombCoder's (src/Golomb.h:12-29,
 * src/GolombCoder.cpp:29-34): samples = ones_before, accumulatedError = last_one_before + 1 - ones_before in uint32. */
bic_status bic_golomb_encode_shard(bic_ctx* ctx, const bic_mat* M_local, uint32_t chunk_samples, uint64_t ones_before,
                                   uint64_t bits_before, int64_t last_one_before, uint64_t code_bits_before, int closing,
                                   uint64_t total_bits, struct bic_stream* out, bic_shard_info* shard);

/* ---------------------------------------------------------------- entropy coding
 * A coded stream is a byte string: stream bit t is in byte t/8 at mask 0x80 >> (t%8)
 * (writeBits / readBits order, src/GolombCoder.cpp:22-25, src/GolombDecoder.cpp:15-23). */
enum { BIC_CODER_GOLOMB = 1, BIC_CODER_EG = 2 };

bic_status bic_stream_create(bic_ctx* ctx, bic_stream** out);
bic_status bic_stream_destroy(bic_ctx* ctx, bic_stream* s);
typedef struct {
  uint32_t coder;         /* BIC_CODER_* */
  uint32_t chunk_samples; /* samples per decoder chunk (Golomb) */
  uint64_t rows, cols;    /* shape of the coded matrix */
  uint64_t bitcount;      /* == GolombCoder::bitcount / EGCoder::bitcount over the same input */
  uint64_t nsamples;      /* Golomb: popcount + 1 */
  uint64_t nchunks;       /* entries of the chunk index (2 uint64 each) */
} bic_stream_info;
bic_status bic_stream_get_info(const bic_stream* s, bic_stream_info* info);
/* bytes: ceil(bitcount/8); index: 2*nchunks uint64 {code bit offset, decoded bit position} */
bic_status bic_stream_download(bic_ctx* ctx, const bic_stream* s, uint8_t* bytes, uint64_t cap_bytes,
                               uint64_t* index, uint64_t cap_index_entries);
bic_status bic_stream_upload(bic_ctx* ctx, bic_stream* s, const bic_stream_info* info,
                             const uint8_t* bytes, const uint64_t* index);

/* GolombCoder (src/GolombCoder.h:19-28, src/GolombCoder.cpp:13-34, state src/Golomb.h:12-29)
 * applied to the zero-run lengths of M read row-major (a virtual one closes the last run):
 * run-length extraction -> adaptive k per sample -> codeword lengths -> device-wide prefix sum
 * of bit offsets -> bit scatter. Byte-identical to the serial coder writing k remainder bits,
 * (x>>k) zeros and a one per sample. */
bic_status bic_golomb_encode(bic_ctx* ctx, const bic_mat* M, uint32_t chunk_samples, bic_stream* out);
/* count only: the reference's GolombCoder::bitcount for the same samples */
bic_status bic_golomb_bitcount(bic_ctx* ctx, const bic_mat* M, uint64_t* bitcount, uint64_t* nsamples);
/* chunk-parallel decoder (GolombDecoder::decodeSample, src/GolombDecoder.cpp:15-40, unsigned samples) */
bic_status bic_golomb_decode(bic_ctx* ctx, const bic_stream* s, bic_mat* M);

/* EGCoder::codeRun (src/eg.h:19-27, src/eg.cpp:20-37) over each row's zero runs: a run broken
 * by a one has eol=false, the row's last run eol=true. */
bic_status bic_eg_encode(bic_ctx* ctx, const bic_mat* M, bic_stream* out);
bic_status bic_eg_decode(bic_ctx* ctx, const bic_stream* s, bic_mat* M);

/* ---------------------------------------------------------------- whole encoder (what bsvd_test's main does,
 * src/bsvd_test.cpp:56-125, with the three PBM dumps replaced by Golomb streams) */
typedef struct {
  uint64_t rows, cols, W, K;  /* raster shape, patch width, atoms */
  uint64_t n, m;              /* patches, bits per patch */
  uint64_t iterations;        /* learn_model_traditional's return value */
  uint64_t weight_E, weight_A, weight_D;
  uint64_t bits_D, bits_A, bits_E; /* Golomb bit counts */
  uint64_t container_bytes;
} bic_encode_info;

/* raster: P4 payload on the host. out: container (header + 3 streams + chunk indexes). */
bic_status bic_encode_raster(bic_ctx* ctx, const uint8_t* pbm_payload, uint64_t rows, uint64_t cols,
                             uint64_t W, uint64_t K, unsigned long seed,
                             uint8_t* out, uint64_t cap_bytes, bic_encode_info* info);
/* bic_encode_raster for a raster that is already in device memory (one of bic_split_bitplanes' planes): same container */
bic_status bic_encode_raster_resident(bic_ctx* ctx, const bic_mat* raster, uint64_t W, uint64_t K, unsigned long seed,
                                      uint8_t* out, uint64_t cap_bytes, bic_encode_info* info);
/* inverse: container -> P4 payload (decode D, A, E; X = A*D xor E; patches -> raster) */
bic_status bic_decode_raster(bic_ctx* ctx, const uint8_t* container, uint64_t container_bytes,
                             uint8_t* pbm_payload, uint64_t cap_bytes, uint64_t* rows, uint64_t* cols);

/* ---------------------------------------------------------------- the encoder as a pipeline (one host thread, many rasters in flight)
 * A pool of `nslots` encoder slots (a CUDA stream and a workspace each) on one device. Rasters are queued with _submit and come
 * out exactly as bic_encode_raster / bic_encode_raster_resident would produce them (same container bytes), but no call ever
 * waits for the device: the pivot draw of initialize_model_neighbor (src/bsvd.cpp:239-243) runs on the device, the iterations of
 * learn_model_traditional (src/bsvd.cpp:1227-1242) are queued in small batches and a device flag turns the ones queued past the
 * loop's end into no-ops, and the stream sizes reach the host with one small copy. bic_pipeline_poll advances every slot whose
 * last event has fired and starts queued rasters on free slots; it returns at once. All calls on one pipeline come from ONE
 * thread. `out` / `info` of a job belong to the pipeline until the job is done; host buffers should be pinned (bic_host_alloc)
 * or the copies serialise. */
typedef struct bic_pipeline bic_pipeline;
bic_status bic_pipeline_create(int device, int nslots, bic_pipeline** out);
bic_status bic_pipeline_destroy(bic_pipeline* p);
/* "first_batch" / "next_batch": iterations queued before the loop flag is looked at (default 2 / 2); any bic_ctx_set_option
 * name is passed on to every slot */
bic_status bic_pipeline_set_option(bic_pipeline* p, const char* name, int64_t value);
/* bic_encode_raster, queued. *job (optional) receives the job's id (> 0). */
bic_status bic_pipeline_submit(bic_pipeline* p, const uint8_t* pbm_payload, uint64_t rows, uint64_t cols, uint64_t W, uint64_t K,
                               unsigned long seed, uint8_t* out, uint64_t cap_bytes, bic_encode_info* info, uint64_t* job);
/* bic_encode_raster_resident, queued. producer (optional): the context on whose stream the raster is being written (e.g. by
 * bic_split_bitplanes); the job is ordered after everything queued there so far. */
bic_status bic_pipeline_submit_resident(bic_pipeline* p, const bic_mat* raster, bic_ctx* producer, uint64_t W, uint64_t K,
                                        unsigned long seed, uint8_t* out, uint64_t cap_bytes, bic_encode_info* info, uint64_t* job);
bic_status bic_pipeline_poll(bic_pipeline* p, uint64_t* unfinished);
/* poll (yielding the core in between) until `job` is done; job = 0: until everything submitted is done */
bic_status bic_pipeline_wait(bic_pipeline* p, uint64_t job);
/* returns the job's own status once it is done (BIC_OK while it is still running); *error: its message */
bic_status bic_pipeline_job_status(bic_pipeline* p, uint64_t job, int* done, const char** error);
bic_status bic_pipeline_forget_finished(bic_pipeline* p);
bic_status bic_pipeline_stats(bic_pipeline* p, uint64_t* launches, uint64_t* polls, uint64_t* batches, uint64_t* sync_fallbacks,
                              uint64_t* recodes);
/* Sharded mode: every job is THIS RANK'S ROW SHARD (a band of the raster, rank order = row order) of one matrix that all ranks
 * fit together with ONE dictionary: bic_dist_* semantics (src/bsvd.cpp:227-267, 463-527, 1215-1244 over the concatenated rows),
 * with nothing waiting for the host -- the pivot draw replays rand48 on the device over the gathered zero-row bitmaps, the
 * statistics are combined by NCCL calls queued on the slot's stream, the per-changed-atom corrections go over NVLink peer memory
 * inside the chain kernel, the shard's Golomb prefix state is computed on the device from two small all-gathers. Slot i of every
 * rank shares communicator i: create them with bic_comm_create(bic_pipeline_slot_ctx(p, i), ...), attach, then submit the SAME
 * jobs in the SAME order on every rank (job q runs on slot q % nslots everywhere). Output per job: a "shard container" (header
 * with the bic_shard_info of A and E, then this rank's D stream, A shard and E shard; layout in csrc/pipeline.cu);
 * info->bits_A / bits_E / weight_* are those of the global streams, identical on every rank. */
struct bic_comm;
bic_ctx* bic_pipeline_slot_ctx(bic_pipeline* p, int slot);
bic_status bic_pipeline_attach_comms(bic_pipeline* p, struct bic_comm* const* comms, int n);
/* The shard containers of ONE sharded job, one per rank (gathered by the caller; any order), merged into the ordinary container
 * of the whole raster: byte for byte what bic_encode_raster gives for the concatenated bands, so bic_decode_raster reads it.
 * Host code only (no device, no context). Every band but the last must be a whole number of patch rows (rows % W == 0).
 * out = NULL: only *bytes is set. BIC_ERR_CORRUPT for inputs that are not the N shards of one job. Buffers 8-byte aligned. */
bic_status bic_merge_shard_containers(const uint8_t* const* shards, const uint64_t* shard_bytes, int nshards, uint8_t* out,
                                      uint64_t cap_bytes, uint64_t* bytes);
/* stream ordering against a context outside the pool: every slot after `signal` / `waiter` after every slot */
bic_status bic_pipeline_wait_ctx(bic_pipeline* p, bic_ctx* signal);
bic_status bic_ctx_wait_pipeline(bic_ctx* waiter, bic_pipeline* p);

#ifdef __cplusplus
}
#endif
#endif /* BIC_B200_H */
