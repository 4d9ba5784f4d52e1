// Internal C++ view of the hw_matrix layout shared by the host builder and the CUDA engine.
// The public boundary is include/spmvb.h; nothing here is exported.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace spmvb {

constexpr int kBusBytes = 16;        // BUS_BIT_WIDTH / 8, reference src/util.h:61
constexpr int kRatioCi = 8;          // 16-bit index slots per bus word, src/util.h:65
constexpr int kGroupsPerChunk = 32;  // one warp lane per 8-entry group
constexpr int kChunkEntries = kRatioCi * kGroupsPerChunk;  // 256 stream entries per chunk

// Per-chunk device metadata (16 bytes per 2560 B (fp64) / 1536 B (fp32) of stream).
struct ChunkMeta {
  uint32_t rank0;      // global row-map index of the segment that contains the chunk's first entry
  uint32_t block;      // bits 0..19 column block (x slice starts at block * cols_div_blocks); bits 20..28 row ends in the chunk
  uint32_t valid;      // real entries in this chunk (padding rows / tail padding excluded), 0..256
  uint32_t row_first;  // rowmap[rank0] (row of the first segment), for the consecutive-rows fast path
};

// flags in ChunkMeta::valid (the entry count lives in the low 10 bits)
constexpr uint32_t kMetaBlockMask = 0xFFFFFu;
constexpr uint32_t kMetaRowsShift = 20;
constexpr uint32_t kChunkRowsConsecutive = 0x80000000u;  // rows of the chunk's segments are row_first, row_first+1, ...
constexpr uint32_t kChunkSole = 0x40000000u;       // every row touched by the chunk lives in exactly one column block
constexpr uint32_t kChunkStartsMid = 0x20000000u;  // the chunk's first entry continues a row begun in the previous chunk

struct Layout {
  int cu = 1, vf = 1, is_double = 1, blocks = 0;
  uint32_t rows = 0, cols = 0, expanded_cols = 0, cdb = 32768;
  int ratio_v = 2, ratio_col_val = 5, vb = 8, group_bytes = 80, chunk_bytes = 2560;
  uint64_t real_nnz = 0, padded_nnz = 0, n_pairs = 0;

  // [k * blocks + b]
  std::vector<uint32_t> nr_rows, nr_nzeros, nr_ci, nr_val;
  std::vector<uint32_t> nr_cols;         // [b]
  // [b * cu + k]
  std::vector<uint64_t> piece_off;       // byte offset into stream
  std::vector<uint64_t> piece_chunk0;    // first chunk index
  std::vector<uint64_t> piece_chunk1;    // one past the last chunk index
  std::vector<uint32_t> dev_order;       // pieces (b * cu + k) in device order
  bool cu_major = false;                 // device order: CU-major instead of block-major
  std::vector<uint32_t> piece_real_nnz;  // entries before the padding rows

  uint8_t *stream = nullptr;  // all pieces, each zero-padded to whole chunks
  uint64_t stream_bytes = 0;
  std::vector<uint64_t> rank_base;  // [blocks + 1] into rowmap
  uint32_t *rowmap = nullptr;       // rank -> row id; the compact form of empty_rows_bitmap
  ChunkMeta *chunks = nullptr;
  uint64_t n_chunks = 0;
  // Warps walk runs of consecutive chunks (carry in registers inside a run, atomics across runs); a row split across
  // two runs must be in zero_rows.  2^run_log2 is the granularity the list is built for: every kernel's run length
  // is a multiple of it (and runs are aligned globally or to the block start).
  int run_log2 = 1;
  // Rows that must be zero before the kernel runs: rows no chunk writes (empty rows), rows updated with atomics
  // (several column blocks, or split across a run boundary).  zero_all: too many to list -> clear all of y.
  std::vector<uint32_t> zero_rows;
  bool zero_all = true;
  // smallest / largest column-in-block among the real entries of each chunk (for shared-memory x windows)
  std::vector<uint16_t> chunk_col_lo, chunk_col_hi;

  ~Layout();
};

// Work item of the XS kernel (x window in shared memory): a range of chunks of one column block.
struct XsItem {
  uint32_t chunk_begin, chunk_count;
  uint32_t x_off;     // first element of the window in x (16-byte aligned)
  uint32_t x_bytes;   // window size in bytes (multiple of 16); 0 = gather from global memory
  uint32_t col_base;  // column-in-block of the window's first element
  uint32_t block, pad0, pad1;
};
constexpr uint32_t kXsCap = 128 * 1024;  // bytes of shared memory for the x window of an XS work item
constexpr int kXsWarps = 16;             // warps per CTA of the XS kernel (one CTA per SM)
void build_xs_items(const Layout *L, int n_cta, uint32_t run_log2, std::vector<XsItem> &items,
                    std::vector<uint32_t> &cta_first);

// pieces shared by the host builder (layout_builder.cpp) and the GPU builder (layout_gpu.cuh)
int layout_init_header(Layout *L, uint32_t rows, uint32_t cols, uint64_t nnz, int cu, int vf, int is_double,
                       uint32_t cdb_in);
void layout_finish_pieces(Layout *L, const uint64_t *fp, const uint32_t *pad_rows, std::vector<uint64_t> &piece_last_rank);

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

}  // namespace spmvb
