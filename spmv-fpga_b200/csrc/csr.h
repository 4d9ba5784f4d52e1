// Library-owned CSR matrix (csr.h:15-24 csr_matrix with 64-bit row offsets) and the counter-based random numbers of the
// synthetic generators.  Shared by matrix_io.cpp (reader / writer) and matrix_gen.cpp (generators).
#pragma once
#include <cstdint>
#include <vector>

namespace spmvb {

struct Csr {
  uint32_t rows = 0, cols = 0;
  int is_double = 1;
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col_ind;
  std::vector<uint8_t> values;  // nnz * (8 | 4) bytes
  uint64_t nnz() const { return row_ptr.empty() ? 0 : row_ptr.back(); }
  void set_value(uint64_t i, double v) {
    if (is_double) ((double *)values.data())[i] = v;
    else ((float *)values.data())[i] = (float)v;
  }
};

// counter-based generator: splitmix64 finaliser over a 64-bit key
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double unit_pm1(uint64_t h) {  // U(-1, 1), never exactly 0
  double u = (double)((h >> 11) | 1ull) * (1.0 / 9007199254740992.0);
  return 2.0 * u - 1.0;
}
static inline double value_at(uint64_t seed, uint32_t r, uint32_t c) {
  return unit_pm1(mix64(mix64(seed ^ 0xA5A5A5A5ull) ^ (((uint64_t)r << 32) | c)));
}

}  // namespace spmvb
