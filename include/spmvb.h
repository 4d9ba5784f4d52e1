/*
 * spmvb.h - C ABI of the B200-native SpMV engine (drop-in boundary).
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * Every entry point names the interface of euroexa/spmv-fpga it replaces
 * (paths relative to the reference's src/).  The C++ host API of the reference
 * (create_csr_hw_matrix / create_csr_hw_x_vector / spmv_hw / delete_*) is
 * re-exported on top of this ABI by include/spmv_fpga_compat.h.
 *
 * All functions returning int return SPMVB_OK (0) or a negative SPMVB_E_* code;
 * spmvb_last_error() gives the message for the calling thread.  There is no CPU
 * fallback: engine calls fail with SPMVB_E_CUDA when no sm_100 device is usable.
 *
 * Threading: handles are not thread-safe; different handles may be used from
 * different threads.  One engine drives one GPU; multi-GPU = one engine (and
 * normally one process) per GPU over a row partition (spmvb_partition_rows).
 */
#ifndef SPMVB_H
#define SPMVB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPMVB_OK 0
#define SPMVB_E_ARG (-1)    /* invalid argument / unsupported CU, VF */
#define SPMVB_E_NOMEM (-2)  /* host allocation failed */
#define SPMVB_E_CUDA (-3)   /* CUDA error or no device */
#define SPMVB_E_IO (-4)     /* matrix-file error */
#define SPMVB_E_RANGE (-5)  /* a count does not fit the reference's 32-bit IndexType */

typedef struct spmvb_layout spmvb_layout; /* host hw_matrix layout: all (CU, block) pieces + bitmap + device aux */
typedef struct spmvb_engine spmvb_engine; /* device-resident copy of a layout + x / y + stream */

const char *spmvb_last_error(void);
int spmvb_version(void);

/* Process-wide tuning options for A/B experiments and tests; every one has a default chosen by the library (-1) and
 * applies to the layouts built / engines created after the call.  The library itself never reads the environment.
 *   run_log2, occ_run_log2, xs_run_log2  run lengths (log2 chunks) of the zero list / the two kernels, 1..8
 *   cu_major 0/1      device order of the pieces        zero_all 1   clear all of y before every SpMV
 *   tall 0/1          explicit L2 eviction policies     autotune 1   time both kernels at engine creation
 *   dev_tiles, dev_cdb, tile_mb, xs_pairs               the engine-private device layout
 *   e2e_tiles 0       spmv_host does not pipeline the row tiles (one launch, then the copy of y)
 *   xs_config 0/1/2   x-window kernel: 128 KB window x 1 CTA per SM / 64 KB x 2 / 32 KB x 3 (decides the device blocks' width)
 *   build_trace 1     print the time of every stage of the GPU layout builder
 * spmvb_options_from_env() applies SPMVB_<NAME>=<integer> for every option and returns how many it found: for
 * executables with no other means of configuration (the reference's run.elf is configured by -D macros only). */
int spmvb_set_option(const char *name, int64_t value);
int64_t spmvb_get_option(const char *name); /* INT64_MIN for an unknown name */
int spmvb_options_from_env(void);

/* ------------------------------------------------------------------ layout (host) */

/* Replaces create_csr_hw_matrix (csr_hw_wrapper.cpp:3-80 -> csr_hw.cpp:377-429, 496-554, ... 1277-1395):
 * scan_matrix (csr_hw.cpp:7-146), prepare_balanced_hw_matrix (:327-361, 432-484), hw_matrix_alloc (:151-183),
 * create_block_matrix (:190-265) and generate_balanced_hw_submatrix (:270-318), in O(nnz + rows) and in
 * parallel.  n_cu in {1,2,4,8,10,12} (any n_cu >= 1 is accepted), vf in {1,2,4,8}, is_double 1 = fp64
 * values, 0 = fp32.  cols_div_blocks 0 = reference default for that CU (util.h:41-59: 32768, or 16384 for
 * CU 10/12); otherwise a multiple of 4 that is <= 32768.  row_ptr has rows+1 entries. */
int spmvb_layout_build(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                       const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                       spmvb_layout **out);
/* Same with the reference's 32-bit row_ptr (csr.h:15-24 csr_matrix). */
int spmvb_layout_build_u32(uint32_t rows, uint32_t cols, const uint32_t *row_ptr, const uint32_t *col_ind,
                           const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                           spmvb_layout **out);
/* Replaces delete_csr_hw_matrix (csr_hw_wrapper.cpp:291-296 -> csr_hw.cpp:1414-1431). */
void spmvb_layout_free(spmvb_layout *l);

/* csr_hw_matrix fields (csr_hw.h:16-26). */
int spmvb_layout_blocks(const spmvb_layout *l);       /* hw_matrix[k]->blocks */
int spmvb_layout_n_cu(const spmvb_layout *l);
uint32_t spmvb_layout_rows(const spmvb_layout *l);
uint32_t spmvb_layout_cols(const spmvb_layout *l);
uint32_t spmvb_layout_expanded_cols(const spmvb_layout *l); /* csr_hw_header.expanded_nr_cols, csr_hw.cpp:30-33 */
uint64_t spmvb_layout_real_nnz(const spmvb_layout *l);      /* non-zeros of the input (no padding) */
uint64_t spmvb_layout_padded_nnz(const spmvb_layout *l);    /* sum of nr_nzeros over all pieces */
uint64_t spmvb_layout_pairs(const spmvb_layout *l);         /* non-empty (row, block) pairs = bitmap zeros */
uint64_t spmvb_layout_stream_bytes(const spmvb_layout *l);  /* bytes of the device image of all pieces */
/* rows cleared before each y = A x (rows updated with atomics or not at all); -1 = all of y is cleared */
int64_t spmvb_layout_zero_rows(const spmvb_layout *l);
/* out[5] = nr_rows, nr_cols, nr_nzeros, nr_ci, nr_val of hw_matrix[cu]->...[block]; nr_val floors like
 * hw_matrix_alloc (csr_hw.cpp:179) although ceil(nr_nzeros/RATIO_v) value words are stored. */
int spmvb_layout_piece_info(const spmvb_layout *l, int cu, int block, uint32_t *out);
/* hw_matrix[cu]->submatrix[block]: nr_ci + ceil(nr_nzeros/RATIO_v) 128-bit words, bit-exact with
 * generate_balanced_hw_submatrix (csr_hw.cpp:270-318); unused index slots are zero. */
const void *spmvb_layout_piece_words(const spmvb_layout *l, int cu, int block);
/* empty_rows_bitmap[block][0..rows) (csr_hw.cpp:340-345, 391-393) expanded from the compact row map. */
int spmvb_layout_bitmap_row(const spmvb_layout *l, int block, uint8_t *out_rows_bytes);
/* storage_overhead (csr_hw.cpp:1401-1409): MB of hw_matrix[cu]. */
double spmvb_layout_storage_mb(const spmvb_layout *l, int cu);
/* write_csr_hw_vector (csr_hw.cpp:1470-1488): x -> expanded_nr_cols zero-padded values. */
int spmvb_layout_pack_x(const spmvb_layout *l, const void *x, uint32_t n, void *out_expanded);

/* Work plan of the shared-memory-x kernel for a GPU with n_cta SMs (host only; for inspection and tests): items_out
 * receives up to max_items records of 8 x uint32 {chunk_begin, chunk_count, x_off, x_bytes, col_base, block, 0, 0}
 * (x_bytes == 0: the item gathers x from global memory), cta_first_out n_cta + 1 item indices.  Returns the number of
 * items, or a negative error. */
int64_t spmvb_layout_xs_plan(const spmvb_layout *l, int n_cta, int run_log2, uint32_t *items_out, uint64_t max_items,
                             uint32_t *cta_first_out);
/* Column ranges of x a SpMV with this layout can read: maximal runs of column blocks that hold at least one entry,
 * as [first, end) pairs in out_pairs (2 x uint64 each, at most max_ranges of them; NULL just counts).  Returns the
 * number of ranges.  spmvb_engine_set_x / spmv_host upload only these (a row shard of a banded matrix reads a band of
 * x, not all of it - the reference copies every block's slice to every compute unit, spmv.cpp:180-192). */
int64_t spmvb_layout_x_ranges(const spmvb_layout *l, uint64_t *out_pairs, uint64_t max_ranges);
/* number of 256-entry chunks of the device image, and the [lo, hi] column-in-block range of chunk c */
uint64_t spmvb_layout_chunks(const spmvb_layout *l);
int spmvb_layout_chunk_cols(const spmvb_layout *l, uint64_t c, uint32_t *lo, uint32_t *hi, uint32_t *block);

/* What the GPU streams for this layout: out[9] = {compute units (row tiles), VF, column-block width, 1 = CU-major
 * order, 1 = an engine-private device layout exists next to the API pieces, (row, block) pairs, chunks, rows cleared
 * per SpMV (UINT64_MAX = all), image bytes}.  The API-visible pieces (piece_info / piece_words) never change with it. */
int spmvb_layout_device_params(const spmvb_layout *l, uint64_t *out);
/* mean number of distinct 128-byte lines of x that the entries of a 256-entry chunk touch (API layout): the measure of
 * irregularity that picks the kernel and the device layout (5-point Laplacian 12, R-MAT ~150-200, uniform ~250) */
double spmvb_layout_x_lines_per_chunk(const spmvb_layout *l);

/* The sliced-ELLPACK image (engine-private candidate for regular matrices - rows of almost equal length whose columns
 * stay within 65 536 of each other per 32 rows): out[8] = {present, width = slots per row, slices of 32
 * rows, bytes per slice, slots, image bytes, real entries, 0}. */
int spmvb_layout_ell_params(const spmvb_layout *l, uint64_t *out);
/* The bytes of the (host-built) ELL image: copied to `out` when max_bytes suffices; returns their number, 0 without an image. */
int64_t spmvb_layout_ell_image(const spmvb_layout *l, void *out, uint64_t max_bytes);
/* Every slot of the ELL image, row-major (slices * 32 rows x width): absolute column and value bits; padding slots carry
 * the row's first column and value 0.  Returns the number of slots.  For tests. */
int64_t spmvb_layout_ell_decode(const spmvb_layout *l, uint32_t *cols_out, void *vals_out, uint64_t max_slots);

/* The wide image (a second engine-private candidate): out[8] = {present, column-block width, column
 * blocks, (row, block) pairs, chunks, rows cleared per SpMV (UINT64_MAX = all), image bytes, real entries}. */
int spmvb_layout_wide_params(const spmvb_layout *l, uint64_t *out);
/* Walks the wide image like the kernel does and returns its real entries in image order (row, column, value bits;
 * any output may be NULL; at most max_entries are stored).  Returns the number of entries, negative on an
 * inconsistent image.  For tests. */
int64_t spmvb_layout_wide_decode(const spmvb_layout *l, uint32_t *rows_out, uint32_t *cols_out, void *vals_out,
                                 uint64_t max_entries);

/* 1 if the two layouts are identical in every table and byte (pieces, row map, chunk metadata, rows to clear, column
 * ranges), 0 if not (why receives the first difference), negative on error.  Used to check the GPU builder against
 * the host builder. */
int spmvb_layout_equal(const spmvb_layout *a, const spmvb_layout *b, char *why, size_t why_len);

/* Row partition for multi-GPU (the CU dimension mapped to GPUs, SURVEY 8e mapping A): `parts` contiguous row
 * ranges balanced by non-zero count with the reference's split rule S1/S2/S3 (csr_hw.cpp:459-460) applied
 * to whole rows.  bounds has parts+1 entries. */
int spmvb_partition_rows(uint32_t rows, const uint64_t *row_ptr, int parts, int ratio_v, uint32_t *bounds);

/* ------------------------------------------------------------------ engine (device) */

/* Uploads the layout to GPU `device` (the analogue of sds_alloc_non_cacheable buffers, csr_hw.cpp:180).
 * variant: 0 = default, otherwise a kernel variant id. */
int spmvb_engine_create(const spmvb_layout *l, int device, int variant, spmvb_engine **out);
/* Replaces create_csr_hw_matrix (csr_hw_wrapper.cpp:3-80 -> csr_hw.cpp:7-429) ON the GPU: same
 * layout, bit for bit, as spmvb_layout_build, built by
 * CUDA kernels (flag / scan / stable radix sort by column block / scatter) straight into the engine's device image.
 * row_ptr (rows+1 x uint64), col_ind, values are host pointers (csr_on_device 0; uploaded inside) or device pointers
 * on `device` (csr_on_device 1).  Fewer than 2^31 rows and non-zeros; rows with unsorted columns cost one extra
 * sort pass.  *layout_out receives every host-side table (piece_info, chunk metadata, rows to clear, XS plan input);
 * its pieces and row map stay on the device until spmvb_engine_fetch_layout copies them back (piece_words and
 * bitmap_row fail before that).  There is no host fallback: invalid input is an error. */
int spmvb_engine_create_from_csr(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                                 const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                                 int device, int variant, int csr_on_device, spmvb_layout **layout_out,
                                 spmvb_engine **engine_out);
/* device image -> host: hw_matrix[k]->submatrix[b] words and the row map of a GPU-built layout */
int spmvb_engine_fetch_layout(spmvb_engine *e, spmvb_layout *l);
/* out3 = milliseconds of {CSR upload, build kernels (CUDA events), whole create_from_csr call (host clock)} */
int spmvb_engine_build_ms(const spmvb_engine *e, float *out3);
void spmvb_engine_free(spmvb_engine *e);
int spmvb_engine_set_variant(spmvb_engine *e, int variant);
int spmvb_engine_variant(const spmvb_engine *e);
/* number of kernel launches issued by this engine so far (bench.py's gpu_launches) */
uint64_t spmvb_engine_launches(const spmvb_engine *e);
/* algorithmic bytes of one SpMV with this engine: nnz*(2+vb) + rows*vb + x*vb, x = the columns
 * of the column blocks that hold at least one entry of this shard (= cols for a whole matrix) */
uint64_t spmvb_engine_algorithmic_bytes(const spmvb_engine *e);
/* engine-owned device vectors: x has expanded_nr_cols values, y has rows values */
void *spmvb_engine_x_dev(spmvb_engine *e);
uint64_t spmvb_engine_x_len(const spmvb_engine *e); /* values the device x holds (>= expanded_nr_cols, zero padded) */
void *spmvb_engine_y_dev(spmvb_engine *e);
void *spmvb_engine_stream(spmvb_engine *e); /* cudaStream_t */

/* Replaces create_csr_hw_x_vector (csr_hw_wrapper.cpp:187-191): host x[n] -> device hw_x (zero padded).  Only the
 * column ranges the matrix can read (spmvb_layout_x_ranges) are copied; the rest of the device vector keeps its old
 * contents, which no kernel of this engine reads. */
int spmvb_engine_set_x(spmvb_engine *e, const void *x_host, uint32_t n);
/* bytes one spmvb_engine_set_x of a full-length x moves to the device */
uint64_t spmvb_engine_x_upload_bytes(const spmvb_engine *e);
/* Replaces the per-block spmv() loop (csr_hw_wrapper.cpp:202-271; spmv.cpp:6-205) fused with accum_results
 * (csr_hw.cpp:1531-1565): y_dev (+)= A * x_dev on `stream` (NULL = engine stream).  x_dev must hold
 * expanded_nr_cols values (or NULL = engine x), y_dev rows values (or NULL = engine y).  accumulate 1 keeps
 * the reference's `+=` semantics, 0 zeroes y first.  Asynchronous. */
int spmvb_engine_spmv_dev(spmvb_engine *e, const void *x_dev, void *y_dev, int accumulate, void *stream);
int spmvb_engine_sync(spmvb_engine *e);
/* device y -> host; accumulate 1: y_host[i] += y_dev[i] (spmv_hw semantics), 0: overwrite */
int spmvb_engine_get_y(spmvb_engine *e, void *y_host, uint32_t n, int accumulate);
/* Replaces spmv_hw (csr_hw_wrapper.cpp:193-288) end to end with HOST buffers: H2D x, SpMV, D2H y,
 * y_host (+)= result.  Synchronous. */
int spmvb_engine_spmv_host(spmvb_engine *e, const void *x_host, uint32_t n, void *y_host, int accumulate);
/* The second half of spmvb_engine_spmv_host - kernel(s) and y to the host, pipelined by row tiles where the device layout
 * has them - for a caller that has put x on the device itself (spmvb_engine_x_dev; the multi-GPU group replicates x over
 * NVLink). */
int spmvb_engine_spmv_host_x_resident(spmvb_engine *e, void *y_host, int accumulate);
/* Times `iters` device SpMVs (y = A x, engine vectors) with CUDA events on the engine stream;
 * ms_out[iters] per-iteration milliseconds.  flush_l2 1 writes a >L2 scratch buffer between iterations. */
int spmvb_engine_time_spmv(spmvb_engine *e, int iters, int flush_l2, float *ms_out);
/* Asynchronous version for the benchmark: enqueues `steps` x (zero y; SpMV kernel) on the engine stream with
 * CUDA events around the whole region and around every kernel launch, and returns at once so the caller can
 * sample clocks while the GPU works.  collect waits and returns the region time and per-launch kernel times. */
int spmvb_engine_enqueue_steps(spmvb_engine *e, int steps, int flags); /* bit 0: flush L2 between steps; bit 1: no
                                                                          per-launch events (region time only) */
int spmvb_engine_steps_done(spmvb_engine *e);
int spmvb_engine_collect_steps(spmvb_engine *e, float *total_ms, float *kernel_ms);
/* Iterated SpMV on one GPU (square matrices): x <- A x / ||A x||_2, `iters` times, all on device.
 * Returns the last norm in *norm_out.  The multi-GPU version lives in the host driver (NCCL all-gather). */
int spmvb_engine_power_iter(spmvb_engine *e, int iters, double *norm_out);
/* Bounds-checked build (make -C spmv-fpga_b200 check: lib/libspmvb_check.so): counts of out-of-range indices the kernels
 * caught since the library was loaded, out5 = {chunk slot, row-map entry, y row, x element, x window offset}; returns 1.
 * A release build returns -1 and leaves out5 alone. */
int spmvb_debug_bounds_errors(uint64_t *out5);
/* For tests: copies the sliced-ELLPACK image the engine streams to `out` (when max_bytes suffices) and returns its size in
 * bytes; 0 when the engine streams something else. */
int64_t spmvb_engine_ell_image(spmvb_engine *e, void *out, uint64_t max_bytes);
/* device time per iteration (CUDA events around the loop) of the last spmvb_engine_power_iter / spmvb_engine_cg call */
float spmvb_engine_last_iter_ms(const spmvb_engine *e);
/* out[20] = the device layout in use: {compute units, VF, column-block width, CU-major, pairs, chunks, rows cleared per
 * SpMV (UINT64_MAX = all), image bytes, row tiles spmv_host pipelines (0 = none), 1 = explicit L2 policies, x-window kernel configuration (0 wide / 1 medium /
 * 2 narrow), microseconds of one SpMV measured at creation for {the API image with global gathers, the device layout - or, for an
 * irregular matrix without one, the x-window kernel on the same image} (0 = not measured), 1 = the wide image is what is streamed, microseconds of one SpMV over the wide image, column blocks,
 * 1 = the sliced-ELLPACK image is what is streamed, microseconds of one SpMV over it, its width (slots per row), row tiles of
 * its end-to-end pipeline} */
int spmvb_engine_device_layout(const spmvb_engine *e, uint64_t *out);
/* Conjugate gradients for A x = b on one GPU (A symmetric positive definite, e.g. the Laplacian of BASELINE config 2):
 * the second iterated caller of SURVEY 8(f) rank 3 (the reference's caller runs spmv_hw once, main.cpp:68-75; an
 * iterated caller keeps x / y on the device between the calls).  x0 = 0; per iteration one SpMV (the engine's kernel) and three
 * fused vector kernels (p.q; x += a p, r -= a q, r.r; p = r + b p) with all scalars on the device; the host looks at
 * ||r|| every 8 iterations and stops when ||r|| <= rel_tol * ||b|| (the recurrence residual) or after max_iters.
 * b_host / x_host hold rows values of the engine's type.  *relres_out = ||r|| / ||b|| at the last check. */
int spmvb_engine_cg(spmvb_engine *e, const void *b_host, void *x_host, int max_iters, double rel_tol, int *iters_out,
                    double *relres_out);
/* x_dev[i] = y_dev[i] * scale for i < n (the normalisation step of the iterated caller) */
int spmvb_engine_scale_copy(spmvb_engine *e, const void *src_dev, void *dst_dev, uint32_t n, double scale,
                            void *stream);
/* dst_dev[i] = src_dev[i] / sqrt(*sumsq_dev) for i < n (0 if the sum is 0): the same step with the norm's square
 * still on the device, e.g. straight after the all-reduce; dst may equal src */
int spmvb_engine_scale_rsqrt(spmvb_engine *e, const void *src_dev, void *dst_dev, uint32_t n, const double *sumsq_dev,
                             void *stream);
/* sum of squares of y_dev[0..n) into a device double (for the norm all-reduce) */
int spmvb_engine_sumsq(spmvb_engine *e, const void *src_dev, uint32_t n, double *out_dev, void *stream);

/* ------------------------------------------------------------------ multi-GPU (the CU dimension mapped to GPUs) */

/* A group of engines over one matrix: every GPU owns a contiguous row range balanced by non-zero count
 * (spmvb_partition_rows), the hw_matrix layout of its own rows, all of x and its slice of y.  Replaces the reference's
 * CU > 1 dispatch, where all compute units run inside one spmv() call with a private x each
 * (spmv.cpp:249-294; csr_hw_wrapper.cpp:3-80, 202-271).  A single SpMV needs no collective; the iterated caller
 * exchanges the y slices into every GPU's x over NVLink (NCCL, loaded on first use). */
typedef struct spmvb_group spmvb_group;
/* One process drives n_devices GPUs (devices NULL = 0 .. n_devices-1): what a -DCU=n program of the reference becomes. */
int spmvb_group_create(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                       const void *values, int is_double, int n_devices, const int *devices, int variant,
                       spmvb_group **out);
/* This process is rank `rank` of `world` (one process per GPU, e.g. under torchrun).  bounds[world + 1] = the row
 * ownership of ALL ranks, the CSR holds this rank's rows only (row_ptr_local rebased to 0, global column indices).
 * unique_id128 = the 128 bytes of spmvb_group_unique_id() made on one rank and handed to all by the launcher. */
int spmvb_group_unique_id(uint8_t *out128);
int spmvb_group_create_rank(uint32_t global_rows, uint32_t cols, const uint32_t *bounds, const uint64_t *row_ptr_local,
                            const uint32_t *col_ind, const void *values, int is_double, int device, int variant,
                            const uint8_t *unique_id128, int rank, int world, spmvb_group **out);
void spmvb_group_free(spmvb_group *g);
int spmvb_group_world(const spmvb_group *g);
int spmvb_group_local_count(const spmvb_group *g);            /* GPUs driven by this process */
int spmvb_group_bounds(const spmvb_group *g, uint32_t *out);  /* world + 1 */
spmvb_engine *spmvb_group_engine(spmvb_group *g, int local_index);
int spmvb_group_rank(const spmvb_group *g, int local_index);
/* spmv_hw (csr_hw_wrapper.cpp:193-288) over the group with HOST buffers: x (n values) to every GPU, all kernels
 * concurrently, every local GPU's rows of y into their place of y_host (global row index; accumulate like spmv_hw). */
int spmvb_group_spmv_host(spmvb_group *g, const void *x_host, uint32_t n, void *y_host, int accumulate);
/* The same with y_rows holding the rows of this process's GPUs only (y_rows[0] = first row of the first local GPU).
 * How x reaches the GPUs is decided once per group, collectively: when every GPU reads (almost) all of x - an irregular
 * matrix - each GPU uploads 1/world of x over its own PCIe link and one in-place ncclAllGather over NVLink fills every
 * GPU's x (world x fewer host bytes); row shards of a banded matrix keep uploading the band they read.
 * spmvb_group_x_over_links: 1 / 0 once decided, -1 before the first call.  Collective across a multi-process group. */
int spmvb_group_spmv_host_rows(spmvb_group *g, const void *x_host, uint32_t n, void *y_rows, int accumulate);
int spmvb_group_x_over_links(const spmvb_group *g);
/* One rank of a multi-process group around an engine the caller created and keeps (spmvb_group_free leaves it alone). */
int spmvb_group_adopt_engine(uint32_t global_rows, uint32_t cols, const uint32_t *bounds, spmvb_engine *engine, int is_double,
                             int device, const uint8_t *unique_id128, int rank, int world, spmvb_group **out);
int spmvb_group_set_x(spmvb_group *g, const void *x_host, uint32_t n);
int spmvb_group_get_x(spmvb_group *g, void *x_host, uint32_t n); /* from the first local GPU */
int spmvb_group_get_y(spmvb_group *g, void *y_host);             /* local GPUs' rows into their place of y_host */
/* x <- A x / ||A x||_2, `iters` times (square matrices), x replicated on every GPU, nothing leaves the devices.  Per
 * iteration: SpMV, sum of squares, all-reduce of one double, scale kernel writing the normalised slice into its place of
 * x, ONE grouped NCCL exchange (a broadcast per row owner, in place).  Collective across the ranks of a multi-process
 * group.  *norm_out = the last norm. */
int spmvb_group_power_iter(spmvb_group *g, int iters, double *norm_out);
float spmvb_group_last_iter_ms(const spmvb_group *g); /* device time per iteration of the last call, max over local GPUs */
/* device time of the four phases of the LAST iteration of the last call, max over the local GPUs: {clear rows + SpMV + sum
 * of squares, all-reduce of the norm, normalisation (+ peer stores + barrier), gather (NCCL broadcasts / all-gather)} */
int spmvb_group_phase_ms(const spmvb_group *g, float *out4);
/* How the power iteration moves the y slices into every GPU's x:
 *   0  NCCL only: one grouped call of a broadcast per row owner, in place in x
 *   1  the normalisation kernel stores its rows straight into EVERY GPU's x over NVLink (peer memory), then an 8-byte
 *      all-reduce as barrier
 *   2  the normalisation kernel stores its rows into x of the ONE GPU that forwards that part of the vector (x is also
 *      cut into `world` equal chunks), barrier, then an all-gather of the equal chunks in place (NCCL; NVLS multicast
 *      on NVSwitch): balanced whatever the row ownership looks like.  Default when the peers' x are mapped.
 * A group made by spmvb_group_create maps the peers itself (cudaDeviceEnablePeerAccess).  A multi-process group needs
 * the launcher's help: every rank publishes spmvb_group_ipc_handle (64 bytes = cudaIpcMemHandle_t of its x), all ranks
 * then receive all handles (world x 64 bytes, in rank order) through spmvb_group_set_peer_handles. */
int spmvb_group_ipc_handle(spmvb_group *g, uint8_t *out64);
int spmvb_group_set_peer_handles(spmvb_group *g, const uint8_t *handles, int mode);
int spmvb_group_set_exchange(spmvb_group *g, int mode);
int spmvb_group_exchange(const spmvb_group *g);

/* ------------------------------------------------------------------ matrix files and synthetic inputs */

/* A CSR matrix owned by the library (csr.h:15-24 csr_matrix with 64-bit row offsets). */
typedef struct spmvb_csr spmvb_csr;
void spmvb_csr_free(spmvb_csr *m);
uint32_t spmvb_csr_rows(const spmvb_csr *m);
uint32_t spmvb_csr_cols(const spmvb_csr *m);
uint64_t spmvb_csr_nnz(const spmvb_csr *m);
int spmvb_csr_is_double(const spmvb_csr *m);
const uint64_t *spmvb_csr_row_ptr(const spmvb_csr *m); /* rows + 1 */
const uint32_t *spmvb_csr_col_ind(const spmvb_csr *m);
const void *spmvb_csr_values(const spmvb_csr *m);       /* fp64 or fp32 */
/* create_csr_hw_matrix on a library-owned CSR */
int spmvb_layout_build_csr(const spmvb_csr *m, int n_cu, int vf, uint32_t cols_div_blocks, spmvb_layout **out);

/* Matrix-file format of the reference (read_csr_header / read_csr_matrix, csr.cpp:10-46, 87-136):
 * "rows cols nnz" then one "row col value" line per entry, 1-based, sorted by row.  Trailing empty rows get
 * row_ptr = nnz (the reference leaves them uninitialised). */
int spmvb_csr_read(const char *path, int is_double, spmvb_csr **out);
int spmvb_csr_write(const spmvb_csr *m, const char *path);
/* Binary form of a parsed matrix (header + row_ptr + col_ind + values as they sit in memory) and a reader that keeps
 * it next to the text file: spmvb_csr_read_cached parses `path` once, writes `path`.f64.spmvb / .f32.spmvb, and loads
 * that sidecar on later calls as long as it is not older than the text file (csr.cpp:87-136 parses the text with
 * fgets + sscanf on every run). */
int spmvb_csr_save(const spmvb_csr *m, const char *path);
int spmvb_csr_load(const char *path, spmvb_csr **out);
int spmvb_csr_read_cached(const char *path, int is_double, spmvb_csr **out);

/* Synthetic inputs of BASELINE.json's configs; values are U(-1,1) from a counter-based hash of (seed, row, col)
 * unless stated.  [row_begin,row_end) selects a row slice of the same global matrix (multi-GPU shards);
 * row_end == 0 means all rows. */
int spmvb_csr_gen_band(uint32_t n, int half_bandwidth, uint64_t seed, int is_double, spmvb_csr **out);
/* 5-point Laplacian on an nx x ny grid: 4 on the diagonal, -1 for the (up to) four neighbours */
int spmvb_csr_gen_laplacian2d(uint32_t nx, uint32_t ny, uint32_t row_begin, uint32_t row_end, int is_double,
                              spmvb_csr **out);
/* nnz_per_row distinct uniformly random columns per row, sorted */
int spmvb_csr_gen_uniform(uint32_t rows, uint32_t cols, int nnz_per_row, uint64_t seed, uint32_t row_begin,
                          uint32_t row_end, int is_double, spmvb_csr **out);
/* R-MAT with 2^scale rows/cols and edge_factor * 2^scale edges before de-duplication, probabilities
 * (a, b, c, 1-a-b-c), no vertex permutation (keeps the empty-row tail); the last row is made non-empty. */
int spmvb_csr_gen_rmat(int scale, int edge_factor, double a, double b, double c, uint64_t seed, uint32_t row_begin,
                       uint32_t row_end, int is_double, spmvb_csr **out);

#ifdef __cplusplus
}
#endif
#endif /* SPMVB_H */
