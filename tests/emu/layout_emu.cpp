// TEST INFRASTRUCTURE ONLY.  Runs the step bodies of the GPU layout builder (spmv-fpga_b200/csrc/layout_gpu_steps.h)
// on the CPU so that the algorithm can be compared with the host builder and the oracle where there is no GPU.
// Built by tests/test_layout_gpu_emu.py into tests/_build/; never linked into libspmvb.so, never used by the product.
// "Kernels" are serial loops that visit the indices in DESCENDING order (any order must give the same result).
#include <algorithm>
#include <numeric>

#include "../../include/spmvb.h"
#include "../../spmv-fpga_b200/csrc/layout_gpu_steps.h"

using namespace spmvb;

namespace {

struct HostBackend {
  bool ok() const { return true; }
  int code() const { return SPMVB_OK; }
  std::string error() const { return ""; }
  void *alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
  void release(void *p) { free(p); }
  void *alloc_output(size_t bytes) { return malloc(bytes ? bytes : 1); }
  void release_output(void *p) { free(p); }
  void fill(void *p, int byte, size_t bytes) { memset(p, byte, bytes); }
  void to_host(void *dst, const void *src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); }
  void to_device(void *dst, const void *src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); }
  void to_host_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height) {
    for (size_t i = 0; i < height; i++) memcpy((uint8_t *)dst + i * dpitch, (const uint8_t *)src + i * spitch, width);
  }
  template <class Body>
  void launch(uint64_t n, const LbCtx &c) {
    for (uint64_t i = n; i-- > 0;) Body::run(i, c);
  }
  void reserve(uint64_t) {}
  void trace(const char *) {}
  void inclusive_sum_bit0_u32(const uint8_t *in, uint32_t *out, uint64_t n) {
    uint32_t s = 0;
    for (uint64_t i = 0; i < n; i++) { s += in[i] & 1; out[i] = s; }
  }
  void exclusive_sum_u32_u64(const uint32_t *in, uint64_t *out, uint64_t n) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) { out[i] = s; s += in[i]; }
  }
  void exclusive_sum_u8_u32(const uint8_t *in, uint32_t *out, uint64_t n) {
    uint32_t s = 0;
    for (uint64_t i = 0; i < n; i++) { out[i] = s; s += in[i]; }
  }
  void sort_pairs(const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint64_t n, int bits) {
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (uint64_t i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
  }
  void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, uint64_t n, int bits) {
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (uint64_t i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
  }
  uint64_t count_nonzero_u8(const uint8_t *in, uint64_t n) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) s += in[i] != 0;
    return s;
  }
  uint64_t select_flagged_iota(const uint8_t *flags, uint32_t *out, uint64_t n) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++) if (flags[i]) out[k++] = (uint32_t)i;
    return k;
  }
};

}  // namespace

extern "C" int emu_layout_build(uint32_t rows, uint32_t cols, const uint64_t *row_ptr, const uint32_t *col_ind,
                                const void *values, int n_cu, int vf, int is_double, uint32_t cols_div_blocks,
                                spmvb_layout **out) {
  if (!out || !row_ptr) return SPMVB_E_ARG;
  HostBackend be, be_dev;
  Layout *L = nullptr;
  LbImage img, img_dev;
  // the same two-step build as spmvb_engine_create_from_csr: API layout, then the engine-private device layout
  int rc = lb_build_pair(be, be_dev, rows, cols, row_ptr[rows], row_ptr, col_ind, values, n_cu, vf, is_double,
                         cols_div_blocks, &L, &img, &img_dev);
  if (rc) return rc;
  rc = lb_fetch_host(be, L, img);
  if (rc == SPMVB_OK && L->dev) rc = lb_fetch_host(be_dev, L->dev, img_dev);
  // the slot metadata inside the image must equal the compact copy
  for (int which = 0; which < 2 && rc == SPMVB_OK; which++) {
    const Layout *X = which ? L->dev : L;
    const LbImage &xi = which ? img_dev : img;
    if (!X) continue;
    const size_t slot = (size_t)X->chunk_bytes + sizeof(ChunkMeta);
    for (uint64_t c = 0; c < X->n_chunks && rc == SPMVB_OK; c++)
      if (memcmp(xi.image + c * slot + X->chunk_bytes, &X->chunks[c], sizeof(ChunkMeta)) != 0) rc = fail(SPMVB_E_ARG, "slot meta");
  }
  be.release_output(img.image); be.release_output(img.rowmap); be.release_output(img.zero_rows);
  be.release_output(img_dev.image); be.release_output(img_dev.rowmap); be.release_output(img_dev.zero_rows);
  if (rc) { delete L; return rc; }
  *out = (spmvb_layout *)L;
  return SPMVB_OK;
}
