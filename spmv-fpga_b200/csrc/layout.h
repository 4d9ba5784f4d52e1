// Internal C++ view of the hw_matrix layout shared by the host builder and the CUDA engine.
// The public boundary is include/spmvb.h; nothing here is exported.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "errors.h"

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

struct EllImage;

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
  int xs_cfg = 0;                        // configuration of the x-window kernel this layout is planned for (xs_config)
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
  // distinct 128-byte lines of x the real entries of each chunk touch: what a gather from global memory costs in L1
  // tag lookups, and the measure of "irregular" that picks the kernel and the device layout (plan_device_params)
  std::vector<uint16_t> chunk_x_lines;

  // Engine-private device layout (owned; nullptr = the device image is this layout itself).  The API-visible pieces
  // above are fixed by the caller's CU / VF / COLS_DIV_BLOCKS; what the GPU streams is the same matrix in the same
  // hw_matrix format under parameters the engine picks for the device (plan_device_params): column blocks whose x
  // slice fits the kernel's shared-memory window, row tiles whose y range stays in the L2 cache, no VF padding.
  Layout *dev = nullptr;

  // Second engine-private candidate: the "wide" image (owned; nullptr = none).  Same stream grammar (8-entry groups,
  // 15-bit column | end-of-row bit, chunks of 32 groups, row map, chunk flags) but column blocks of up to 2^23 columns:
  // an x range the L2 cache holds, gathered with ld.global, while the rows ascend through the range so that every
  // (row, range) partial sum is formed in registers and y is updated by coalesced requests.  The extra column bits
  // live in a byte plane of their own and a chunk is stored plane by plane (see kWide* below).
  Layout *wide = nullptr;
  bool is_wide = false;  // this layout IS a wide image
  // Third engine-private candidate, for regular matrices only (ell.h): sliced ELLPACK, one row per lane
  EllImage *ell = nullptr;

  ~Layout();
};

// Parameters of the engine-private device layout for the API layout L; false when the API layout is what should be
// streamed as it is.
bool plan_device_params(const Layout *L, int *cu_dev, int *vf_dev, uint32_t *cdb_dev);
// set (per thread) while an engine-private device layout is being built: its compute units are row tiles, walked CU-major
bool &building_device_layout();
bool layout_is_irregular(const Layout *L);
// set (per thread) while a wide image is being built: column blocks up to 2^23 columns, planar chunks
bool &building_wide_layout();
// Column-block width of the wide image for the API layout L (0 = none wanted)
uint32_t plan_wide_cdb(const Layout *L);

// Planar chunk of a wide image: 32 lanes x {16 B index word | 8 B high column bytes | 8 values in 16-byte words}.
//   [0, 512)            index words, 16 B per lane (8 x (column & 0x7FFF | end-of-row bit))
//   [512, 768)          high column bytes, 8 B per lane (column >> 15)
//   [768 + k * 512 ...) value word k of every lane (k = 0..3 fp64, 0..1 fp32), 16 B per lane
// so that every LDS of a warp is conflict-free.  2816 B (fp64) / 1792 B (fp32) per chunk.
constexpr uint32_t kWideHiOff = 512, kWideValOff = 768, kWidePlane = 512;
constexpr uint32_t kWideMaxCdb = 1u << 23;
inline uint32_t wide_chunk_bytes(int vb) { return kWideValOff + 32u * 8u * (uint32_t)vb; }

// Work item of the XS kernel (x window in shared memory): a range of chunks of one column block.
struct XsItem {
  uint32_t chunk_begin, chunk_count;
  uint32_t x_off;     // first element of the window in x (16-byte aligned)
  uint32_t x_bytes;   // window size in bytes (multiple of 16); 0 = gather from global memory
  uint32_t col_base;  // column-in-block of the window's first element
  uint32_t block, pad0, pad1;
};
// Configurations of the x-window kernel (XS): bytes of shared memory for the x window of a work item, warps per CTA,
// CTAs per SM.  Every warp owns two ring stages of one chunk slot (2576 B fp64 / 1552 B fp32); what is left of the
// 227 KB next to the window(s) and the register file bound the warps.  The kernel lives on warps - its per-chunk
// instruction chain is long and serial (measured on the Laplacian: 12 / 16 / 18 warps = 78 / 69 / 66 us) - so a narrower
// window buys parallelism: wide = the reference's whole x slice of a 16 384-column block on chip (fp64), medium / narrow =
// half / a quarter of it with two / three independent CTAs per SM.
struct XsConfig { uint32_t cap; int warps, ctas_per_sm; };
constexpr int kXsConfigs = 3;
inline XsConfig xs_config(int is_double, int cfg) {
  static const XsConfig f64[kXsConfigs] = {{128u << 10, 18, 1}, {64u << 10, 9, 2}, {32u << 10, 8, 3}};
  static const XsConfig f32[kXsConfigs] = {{128u << 10, 24, 1}, {64u << 10, 14, 2}, {32u << 10, 10, 3}};
  cfg = cfg < 0 ? 0 : (cfg >= kXsConfigs ? kXsConfigs - 1 : cfg);
  return is_double ? f64[cfg] : f32[cfg];
}
// CU-major layouts only: the same work dealt tile by tile (one kernel launch per row tile)
struct XsTilePlan {
  int n_tiles = 0;
  std::vector<XsItem> items;
  std::vector<uint32_t> cta_first;  // [n_tiles * (n_cta + 1)]
  std::vector<uint32_t> rows_end;   // [n_tiles] rows [0, rows_end[k]) are final once tiles 0..k are done
};
void build_xs_items(const Layout *L, int n_cta, uint32_t run_log2, std::vector<XsItem> &items,
                    std::vector<uint32_t> &cta_first, XsTilePlan *tiles);

// pieces shared by the host builder (layout_builder.cpp) and the GPU builder (layout_gpu.cuh)
int layout_init_header(Layout *L, uint32_t rows, uint32_t cols, uint64_t nnz, int cu, int vf, int is_double,
                       uint32_t cdb_in);
void layout_finish_pieces(Layout *L, const uint64_t *fp, const uint32_t *pad_rows, std::vector<uint64_t> &piece_last_rank);

// Process-wide tuning options (include/spmvb.h: spmvb_set_option).  -1 = let the library decide.  Nothing in the
// library reads the environment: a caller that wants SPMVB_* variables honoured calls spmvb_options_from_env() itself.
struct Options {
  int64_t run_log2 = -1;      // zero-list granularity of new layouts (1..8)
  int64_t cu_major = -1;      // device order of the pieces: 0 block-major, 1 CU-major
  int64_t zero_all = -1;      // 1: clear all of y before every SpMV instead of the listed rows
  int64_t tall = -1;          // explicit L2 eviction policies (x evict-first, y evict-last)
  int64_t occ_run_log2 = -1;  // run length of the OCC kernel
  int64_t xs_run_log2 = -1;   // run length of the XS kernel
  int64_t autotune = -1;      // 1: time the candidate kernels on the actual matrix at engine creation
  int64_t build_trace = 0;    // 1: print the time of every stage of the GPU layout builder
  int64_t dev_tiles = -1;     // row tiles of the engine-private device layout (0/1 = none)
  int64_t dev_cdb = -1;       // column-block width of the engine-private device layout (0 = same as the API layout)
  int64_t xs_pairs = -1;      // distinct x lines per 256-entry chunk above which the x-window kernel is preferred
  int64_t tile_mb = -1;       // target size of a row tile's y range in MB
  int64_t xs_config = -1;     // 0 wide / 1 medium / 2 narrow x window of the XS kernel (see XsConfig)
  int64_t l2_persist_mb = -1; // > 0: set aside that much L2 for persisting (evict-last) lines on tall matrices
  int64_t diag_flags = -1;    // diagnostics of the x-window kernel (WRONG results): 16 = no x window traffic, 32 = no y updates
  int64_t tile_launch = -1;   // 1: one kernel launch per row tile with the tile's y range as persisting L2 window
  int64_t e2e_tiles = -1;     // 0: spmv_host does not pipeline row tiles (one launch, then the copy of y)
  int64_t wide = -1;          // wide image: 1 build one (a third candidate for the engine), otherwise none
  int64_t ell = -1;           // sliced-ELLPACK image: 0 never; 1 build and use it whenever the format can hold the matrix;
                              // -1 build it when its padding is (almost) free and use it when it measures faster
  int64_t ell_tiles = -1;     // row tiles of the end-to-end pipeline over an ELL image (x up / kernel / y down overlapped)
  int64_t wide_range_log2 = -1;  // log2 of the column-block width of the wide image (2..23)
  int64_t wide_hints = -1;    // L2 policies of the wide kernel: bit 0 x gathers evict-last, bit 1 y updates / row map evict-first
};
Options &options();


}  // namespace spmvb
