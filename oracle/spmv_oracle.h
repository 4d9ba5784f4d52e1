/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * Plain-C restatement of the euroexa/spmv-fpga hot path: the hw_matrix layout
 * builder, the HLS kernel arithmetic, the host accumulation and the gold CSR
 * SpMV.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this.  PARITY PINNED: checked word-for-word
 * against the unmodified reference compiled by oracle/Makefile (oracle/_ref)
 * and against the committed fixtures in tests/golden/ generated from it.
 */
#ifndef SPMV_ORACLE_H
#define SPMV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_layout orc_layout;

/* cols_div_blocks == 0 selects the reference default for that CU (src/util.h:41-59). */
orc_layout *orc_layout_build(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                             const void *values, int cu, int vf, int is_double, uint32_t cols_div_blocks);
void orc_layout_free(orc_layout *l);
int orc_blocks(const orc_layout *l);
uint32_t orc_expanded_cols(const orc_layout *l);
/* out[5] = nr_rows, nr_cols, nr_nzeros, nr_ci, nr_val (nr_val floors like the reference) */
void orc_piece_info(const orc_layout *l, int cu, int block, uint32_t *out);
const void *orc_piece_words(const orc_layout *l, int cu, int block);
/* number of 128-bit words actually backing the piece: nr_ci + ceil(nnz/RATIO_v) */
uint32_t orc_piece_alloc_words(const orc_layout *l, int cu, int block);
const uint8_t *orc_bitmap_row(const orc_layout *l, int block);
/* hw_x: expanded_nr_cols values, zero-padded (src/csr_hw.cpp:1470-1488) */
void orc_hw_x(const orc_layout *l, const void *x, uint32_t n, void *out);
/* emulated spmv_hw: y (rows values) is accumulated into (src/csr_hw_wrapper.cpp:193-288) */
int orc_spmv_emu(const orc_layout *l, const void *x, uint32_t n, void *y);

void orc_spmv_gold(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                   const void *x, void *y, int is_double);
/* same loop, rows split over all host cores with OpenMP (bit-identical result); returns the thread count */
int orc_spmv_gold_omp(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                      const void *x, void *y, int is_double);
void orc_abs_ax(uint32_t rows, const uint64_t *row_ptr, const uint32_t *col_ind, const void *values,
                const void *x, double *out, int is_double);
int orc_verification(uint32_t n, const void *sw, const void *hw, int is_double);

#ifdef __cplusplus
}
#endif
#endif
