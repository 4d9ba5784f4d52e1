// Library-owned CSR matrices: reader/writer for the reference's matrix-file format and the synthetic generators
// for BASELINE.json's configs.  Reference: read_csr_header / read_csr_matrix, src/csr.cpp:10-46, 87-136 (format
// "%u %u %u\n" then "%u %u %lf\n" / "%u %u %f\n", 1-based, sorted by row; src/util.h:28-29).
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "layout.h"

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

using namespace spmvb;

extern "C" {

void spmvb_csr_free(spmvb_csr *m) { delete (Csr *)m; }
uint32_t spmvb_csr_rows(const spmvb_csr *m) { return ((const Csr *)m)->rows; }
uint32_t spmvb_csr_cols(const spmvb_csr *m) { return ((const Csr *)m)->cols; }
uint64_t spmvb_csr_nnz(const spmvb_csr *m) { return ((const Csr *)m)->nnz(); }
int spmvb_csr_is_double(const spmvb_csr *m) { return ((const Csr *)m)->is_double; }
const uint64_t *spmvb_csr_row_ptr(const spmvb_csr *m) { return ((const Csr *)m)->row_ptr.data(); }
const uint32_t *spmvb_csr_col_ind(const spmvb_csr *m) { return ((const Csr *)m)->col_ind.data(); }
const void *spmvb_csr_values(const spmvb_csr *m) { return ((const Csr *)m)->values.data(); }

int spmvb_layout_build_csr(const spmvb_csr *m, int n_cu, int vf, uint32_t cols_div_blocks, spmvb_layout **out) {
  const Csr *A = (const Csr *)m;
  if (!A) return fail(SPMVB_E_ARG, "csr is NULL");
  return spmvb_layout_build(A->rows, A->cols, A->row_ptr.data(), A->col_ind.data(), A->values.data(), n_cu, vf,
                            A->is_double, cols_div_blocks, out);
}

// Parallel reader: the file is mapped, cut into one byte range per thread at line boundaries, lines are counted and
// then parsed in place (strtoul / strtod), so that a 1 B-line file is parsed by all cores instead of one fgets+sscanf
// loop (the reference: csr.cpp:87-136, minutes per 100 M lines).  Same acceptance rules as the reference reader plus
// range / order checks.
int spmvb_csr_read(const char *path, int is_double, spmvb_csr **out) {
  if (!path || !out) return fail(SPMVB_E_ARG, "csr_read");
  *out = nullptr;
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(SPMVB_E_IO, std::string("Could not open file ") + path);  // csr.cpp:15-18
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return fail(SPMVB_E_IO, "unexpected eof found"); }
  const size_t size = (size_t)st.st_size;
  const char *buf = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (buf == MAP_FAILED) return fail(SPMVB_E_IO, "mmap failed");
  auto done = [&](int code, const std::string &msg) { munmap((void *)buf, size); return fail(code, msg); };
  // header line (csr.cpp:21): "%u %u %u\n"
  const char *eol = (const char *)memchr(buf, '\n', size);
  const size_t hdr_len = eol ? (size_t)(eol - buf) + 1 : size;
  unsigned rows = 0, cols = 0;
  unsigned long long nnz = 0;
  {
    std::string h(buf, hdr_len);
    if (sscanf(h.c_str(), "%u %u %llu", &rows, &cols, &nnz) != 3) return done(SPMVB_E_IO, "parse error in header");
  }
  const int T = std::max(1, omp_get_max_threads());
  std::vector<size_t> cut(T + 1);
  cut[0] = hdr_len; cut[T] = size;
  for (int t = 1; t < T; t++) {
    size_t p = hdr_len + (size - hdr_len) / T * t;
    const char *nl = p < size ? (const char *)memchr(buf + p, '\n', size - p) : nullptr;
    cut[t] = nl ? (size_t)(nl - buf) + 1 : size;
    if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
  }
  // pass 1: non-blank lines per range
  std::vector<uint64_t> first(T + 1, 0);
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t n = 0;
    for (size_t p = cut[t]; p < cut[t + 1];) {
      const char *nl = (const char *)memchr(buf + p, '\n', cut[t + 1] - p);
      const size_t e = nl ? (size_t)(nl - buf) : cut[t + 1];
      size_t q = p;
      while (q < e && (buf[q] == ' ' || buf[q] == '\t' || buf[q] == '\r')) q++;
      if (q < e) n++;
      p = e + 1;
    }
    first[t + 1] = n;
  }
  for (int t = 0; t < T; t++) first[t + 1] += first[t];
  if (first[T] != nnz) return done(SPMVB_E_IO, "entry count differs from the header");
  Csr *A = new Csr();
  A->rows = rows; A->cols = cols; A->is_double = is_double ? 1 : 0;
  A->row_ptr.assign((size_t)rows + 1, 0);
  A->col_ind.resize(nnz);
  A->values.resize((size_t)nnz * (is_double ? 8 : 4));
  std::vector<uint32_t> row_of(nnz);
  int bad = 0;
  // pass 2: parse.  strtoul/strtod stop at the newline; the mapping ends with the file, so the last line is copied
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t i = first[t];
    char tail[256];
    for (size_t p = cut[t]; p < cut[t + 1] && !bad;) {
      const char *nl = (const char *)memchr(buf + p, '\n', cut[t + 1] - p);
      const size_t e = nl ? (size_t)(nl - buf) : cut[t + 1];
      size_t q = p;
      while (q < e && (buf[q] == ' ' || buf[q] == '\t' || buf[q] == '\r')) q++;
      if (q < e) {
        const char *line = buf + q;
        if (!nl) {  // last line of the file without a trailing newline: parse a NUL-terminated copy
          size_t len = std::min(e - q, sizeof tail - 1);
          memcpy(tail, buf + q, len); tail[len] = 0; line = tail;
        }
        char *c1, *c2, *c3;
        unsigned long r = strtoul(line, &c1, 10);
        unsigned long c = strtoul(c1, &c2, 10);
        double v = strtod(c2, &c3);
        if (c1 == line || c2 == c1 || c3 == c2 || r < 1 || r > rows || c < 1 || c > cols) { bad = 1; break; }
        row_of[i] = (uint32_t)(r - 1);
        A->col_ind[i] = (uint32_t)(c - 1);  // mmarket files are not zero-based, csr.cpp:118
        A->set_value(i, v);
        i++;
      }
      p = e + 1;
    }
  }
  munmap((void *)buf, size);
  if (bad) { delete A; return fail(SPMVB_E_IO, "parse error: malformed or out-of-range entry"); }
  // sorted by row? then row_ptr = counts, prefix-summed (empty and trailing rows come out right by construction, Q3)
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t i = 1; i < (int64_t)nnz; i++) bad |= row_of[i] < row_of[i - 1];
  if (bad) { delete A; return fail(SPMVB_E_IO, "entries are not sorted by row"); }
  for (uint64_t i = 0; i < nnz; i++) A->row_ptr[(size_t)row_of[i] + 1]++;
  for (uint32_t r = 0; r < rows; r++) A->row_ptr[r + 1] += A->row_ptr[r];
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

// Writer: every thread formats its own row range, the pieces are written in order.
// ---- binary sidecar of a matrix file: the parsed CSR as it sits in memory, so that a matrix is
// parsed from text once.  Layout: 8-byte magic, u32 version, u32 is_double, u32 rows, u32 cols, u64 nnz, then
// row_ptr[rows + 1] (u64), col_ind[nnz] (u32), values[nnz] (f64 | f32).  Native byte order (little endian here).
static const char kBinMagic[8] = {'S', 'P', 'M', 'V', 'B', 'C', 'S', 'R'};

int spmvb_csr_save(const spmvb_csr *m, const char *path) {
  const Csr *A = (const Csr *)m;
  if (!A || !path) return fail(SPMVB_E_ARG, "csr_save");
  const std::string tmp = std::string(path) + ".tmp";
  FILE *fp = fopen(tmp.c_str(), "wb");
  if (!fp) return fail(SPMVB_E_IO, std::string("cannot create ") + tmp);
  const uint32_t hdr[4] = {1u, (uint32_t)A->is_double, A->rows, A->cols};
  const uint64_t nnz = A->nnz();
  bool ok = fwrite(kBinMagic, 1, 8, fp) == 8 && fwrite(hdr, 4, 4, fp) == 4 && fwrite(&nnz, 8, 1, fp) == 1;
  ok = ok && fwrite(A->row_ptr.data(), 8, A->row_ptr.size(), fp) == A->row_ptr.size();
  ok = ok && fwrite(A->col_ind.data(), 4, (size_t)nnz, fp) == (size_t)nnz;
  ok = ok && fwrite(A->values.data(), 1, A->values.size(), fp) == A->values.size();
  ok = (fclose(fp) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(SPMVB_E_IO, std::string("write error on ") + path); }
  return SPMVB_OK;
}

int spmvb_csr_load(const char *path, spmvb_csr **out) {
  if (!path || !out) return fail(SPMVB_E_ARG, "csr_load");
  *out = nullptr;
  FILE *fp = fopen(path, "rb");
  if (!fp) return fail(SPMVB_E_IO, std::string("Could not open file ") + path);
  char magic[8];
  uint32_t hdr[4];
  uint64_t nnz = 0;
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, kBinMagic, 8) != 0 || fread(hdr, 4, 4, fp) != 4 ||
      fread(&nnz, 8, 1, fp) != 1 || hdr[0] != 1u || hdr[1] > 1u || hdr[2] == 0 || hdr[3] == 0) {
    fclose(fp);
    return fail(SPMVB_E_IO, std::string("not a spmvb binary matrix: ") + path);
  }
  struct stat st;
  const uint64_t vb = hdr[1] ? 8 : 4;
  const uint64_t want = 8 + 16 + 8 + ((uint64_t)hdr[2] + 1) * 8 + nnz * 4 + nnz * vb;
  if (fstat(fileno(fp), &st) != 0 || (uint64_t)st.st_size != want) {
    fclose(fp);
    return fail(SPMVB_E_IO, std::string("truncated binary matrix: ") + path);
  }
  Csr *A = new Csr();
  A->is_double = (int)hdr[1]; A->rows = hdr[2]; A->cols = hdr[3];
  A->row_ptr.resize((size_t)A->rows + 1);
  A->col_ind.resize((size_t)nnz);
  A->values.resize((size_t)(nnz * vb));
  bool ok = fread(A->row_ptr.data(), 8, A->row_ptr.size(), fp) == A->row_ptr.size();
  ok = ok && fread(A->col_ind.data(), 4, (size_t)nnz, fp) == (size_t)nnz;
  ok = ok && fread(A->values.data(), 1, A->values.size(), fp) == A->values.size();
  fclose(fp);
  // the file is only trusted as far as the builders need: monotone offsets that end at nnz, columns in range
  ok = ok && A->row_ptr[0] == 0 && A->row_ptr[A->rows] == nnz;
  for (uint32_t r = 0; ok && r < A->rows; r++) ok = A->row_ptr[r] <= A->row_ptr[r + 1];
  if (ok) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t j = 0; j < (int64_t)nnz; j++) bad |= A->col_ind[j] >= A->cols;
    ok = !bad;
  }
  if (!ok) { delete A; return fail(SPMVB_E_IO, std::string("corrupt binary matrix: ") + path); }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_read_cached(const char *path, int is_double, spmvb_csr **out) {
  if (!path || !out) return fail(SPMVB_E_ARG, "csr_read_cached");
  const std::string side = std::string(path) + (is_double ? ".f64.spmvb" : ".f32.spmvb");
  struct stat ts, bs;
  if (stat(path, &ts) == 0 && stat(side.c_str(), &bs) == 0 &&
      (bs.st_mtim.tv_sec > ts.st_mtim.tv_sec ||
       (bs.st_mtim.tv_sec == ts.st_mtim.tv_sec && bs.st_mtim.tv_nsec >= ts.st_mtim.tv_nsec))) {
    if (spmvb_csr_load(side.c_str(), out) == SPMVB_OK && spmvb_csr_is_double(*out) == (is_double ? 1 : 0)) return SPMVB_OK;
    if (*out) { spmvb_csr_free(*out); *out = nullptr; }  // stale or foreign sidecar: parse the text again
  }
  int rc = spmvb_csr_read(path, is_double, out);
  if (rc) return rc;
  spmvb_csr_save(*out, side.c_str());  // best effort: a read-only directory just means no cache
  return SPMVB_OK;
}

int spmvb_csr_write(const spmvb_csr *m, const char *path) {
  const Csr *A = (const Csr *)m;
  if (!A || !path) return fail(SPMVB_E_ARG, "csr_write");
  FILE *fp = fopen(path, "w");
  if (!fp) return fail(SPMVB_E_IO, std::string("cannot create ") + path);
  fprintf(fp, "%u %u %llu\n", A->rows, A->cols, (unsigned long long)A->nnz());
  const int T = std::max(1, omp_get_max_threads());
  const uint32_t batch = 1u << 20;  // rows per round: bounds the memory of the formatted text
  int err = 0;
  for (uint32_t r0 = 0; r0 < A->rows && !err; r0 += batch) {
    const uint32_t r1 = (uint32_t)std::min<uint64_t>((uint64_t)r0 + batch, A->rows);
    std::vector<std::string> part(T);
#pragma omp parallel num_threads(T)
    {
      const int t = omp_get_thread_num();
      const uint32_t a = r0 + (uint32_t)((uint64_t)(r1 - r0) * t / T), b = r0 + (uint32_t)((uint64_t)(r1 - r0) * (t + 1) / T);
      std::string &s = part[t];
      char line[96];
      for (uint32_t r = a; r < b; r++)
        for (uint64_t j = A->row_ptr[r]; j < A->row_ptr[r + 1]; j++) {
          int n = A->is_double
                      ? snprintf(line, sizeof line, "%u %u %.17g\n", r + 1, A->col_ind[j] + 1, ((const double *)A->values.data())[j])
                      : snprintf(line, sizeof line, "%u %u %.9g\n", r + 1, A->col_ind[j] + 1, (double)((const float *)A->values.data())[j]);
          s.append(line, (size_t)n);
        }
    }
    for (int t = 0; t < T && !err; t++)
      if (!part[t].empty() && fwrite(part[t].data(), 1, part[t].size(), fp) != part[t].size()) err = 1;
  }
  if (fclose(fp) != 0 || err) return fail(SPMVB_E_IO, "write error");
  return SPMVB_OK;
}

int spmvb_csr_gen_band(uint32_t n, int hb, uint64_t seed, int is_double, spmvb_csr **out) {
  if (!out || n == 0 || hb < 0) return fail(SPMVB_E_ARG, "gen_band");
  Csr *A = new Csr();
  A->rows = A->cols = n; A->is_double = is_double ? 1 : 0;
  A->row_ptr.assign((size_t)n + 1, 0);
  for (uint32_t r = 0; r < n; r++) {
    uint32_t lo = r >= (uint32_t)hb ? r - hb : 0, hi = std::min<uint64_t>((uint64_t)r + hb, n - 1);
    A->row_ptr[r + 1] = A->row_ptr[r] + (hi - lo + 1);
  }
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)n; r++) {
    uint32_t lo = r >= hb ? (uint32_t)r - hb : 0;
    uint64_t j = A->row_ptr[r];
    for (uint64_t k = 0; k < A->row_ptr[r + 1] - A->row_ptr[r]; k++, j++) {
      A->col_ind[j] = lo + (uint32_t)k;
      A->set_value(j, value_at(seed, (uint32_t)r, lo + (uint32_t)k));
    }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_laplacian2d(uint32_t nx, uint32_t ny, uint32_t row_begin, uint32_t row_end, int is_double,
                              spmvb_csr **out) {
  if (!out || nx == 0 || ny == 0 || (uint64_t)nx * ny > 0xFFFFFFFFull) return fail(SPMVB_E_ARG, "gen_laplacian2d");
  const uint32_t n = nx * ny;
  if (row_end == 0) row_end = n;
  if (row_begin >= row_end || row_end > n) return fail(SPMVB_E_ARG, "gen_laplacian2d: row range");
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = n; A->is_double = is_double ? 1 : 0;
  A->row_ptr.assign((size_t)A->rows + 1, 0);
  for (uint32_t i = 0; i < A->rows; i++) {
    uint32_t r = row_begin + i, ix = r % nx, iy = r / nx;
    uint32_t d = 1 + (iy > 0) + (ix > 0) + (ix + 1 < nx) + (iy + 1 < ny);
    A->row_ptr[i + 1] = A->row_ptr[i] + d;
  }
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)A->rows; i++) {
    uint32_t r = row_begin + (uint32_t)i, ix = r % nx, iy = r / nx;
    uint64_t j = A->row_ptr[i];
    if (iy > 0) { A->col_ind[j] = r - nx; A->set_value(j++, -1.0); }
    if (ix > 0) { A->col_ind[j] = r - 1; A->set_value(j++, -1.0); }
    A->col_ind[j] = r; A->set_value(j++, 4.0);
    if (ix + 1 < nx) { A->col_ind[j] = r + 1; A->set_value(j++, -1.0); }
    if (iy + 1 < ny) { A->col_ind[j] = r + nx; A->set_value(j++, -1.0); }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_uniform(uint32_t rows, uint32_t cols, int k, uint64_t seed, uint32_t row_begin, uint32_t row_end,
                          int is_double, spmvb_csr **out) {
  if (!out || rows == 0 || cols == 0 || k < 1 || (uint32_t)k > cols || k > 1024) return fail(SPMVB_E_ARG, "gen_uniform");
  if (row_end == 0) row_end = rows;
  if (row_begin >= row_end || row_end > rows) return fail(SPMVB_E_ARG, "gen_uniform: row range");
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = cols; A->is_double = is_double ? 1 : 0;
  A->row_ptr.resize((size_t)A->rows + 1);
  for (uint64_t i = 0; i <= A->rows; i++) A->row_ptr[i] = i * (uint64_t)k;
  const uint64_t nnz = A->nnz();
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)A->rows; i++) {
    const uint32_t r = row_begin + (uint32_t)i;
    uint32_t c[1024];
    uint64_t ctr = 0;
    const uint64_t key = mix64(seed * 0x100000001B3ull + r);
    int have = 0;
    while (have < k) {  // draw, sort, drop duplicates, top up
      while (have < k) c[have++] = (uint32_t)(((mix64(key + ctr++) >> 32) * (uint64_t)cols) >> 32);
      std::sort(c, c + k);
      have = (int)(std::unique(c, c + k) - c);
    }
    uint64_t j = (uint64_t)i * k;
    for (int q = 0; q < k; q++, j++) {
      A->col_ind[j] = c[q];
      A->set_value(j, value_at(seed, r, c[q]));
    }
  }
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

int spmvb_csr_gen_rmat(int scale, int ef, double a, double b, double c, uint64_t seed, uint32_t row_begin,
                       uint32_t row_end, int is_double, spmvb_csr **out) {
  if (!out || scale < 1 || scale > 31 || ef < 1 || a <= 0 || b < 0 || c < 0 || a + b + c >= 1.0)
    return fail(SPMVB_E_ARG, "gen_rmat");
  const uint32_t n = 1u << scale;
  if (row_end == 0) row_end = n;
  if (row_begin >= row_end || row_end > n) return fail(SPMVB_E_ARG, "gen_rmat: row range");
  const uint64_t edges = (uint64_t)ef << scale;
  // thresholds on a 32-bit uniform
  const uint64_t ta = (uint64_t)(a * 4294967296.0), tab = (uint64_t)((a + b) * 4294967296.0),
                 tabc = (uint64_t)((a + b + c) * 4294967296.0);
  // bucket by the top bits of the row so that buckets can be sorted independently
  const int bucket_bits = std::min(scale, 12);
  const uint32_t nb = 1u << bucket_bits;
  const int shift = scale - bucket_bits;
  const int T = std::max(1, omp_get_max_threads());
  std::vector<uint64_t> cnt((size_t)T * nb, 0);
  auto edge = [&](uint64_t e, uint32_t &r, uint32_t &cc) {
    uint64_t key = mix64(seed ^ (e * 0xD1342543DE82EF95ull));
    r = 0; cc = 0;
    for (int lvl = 0; lvl < scale; lvl += 2) {  // one 64-bit hash feeds two levels
      uint64_t h = mix64(key + (uint64_t)lvl);
      for (int half = 0; half < 2 && lvl + half < scale; half++) {
        uint64_t u = (half ? (h >> 32) : (h & 0xFFFFFFFFull));
        uint32_t rb = u >= tab, cb = (u >= ta && u < tab) || u >= tabc;
        r = (r << 1) | rb; cc = (cc << 1) | cb;
      }
    }
  };
  std::vector<uint64_t> tstart(T + 1);
  for (int t = 0; t <= T; t++) tstart[t] = edges / T * t;
  tstart[T] = edges;
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t *ct = &cnt[(size_t)t * nb];
    for (uint64_t e = tstart[t]; e < tstart[t + 1]; e++) {
      uint32_t r, cc;
      edge(e, r, cc);
      if (r >= row_begin && r < row_end) ct[r >> shift]++;
    }
  }
  std::vector<uint64_t> bstart(nb + 1, 0);
  for (uint32_t q = 0; q < nb; q++) {
    uint64_t s = 0;
    for (int t = 0; t < T; t++) { uint64_t v = cnt[(size_t)t * nb + q]; cnt[(size_t)t * nb + q] = bstart[q] + s; s += v; }
    bstart[q + 1] = bstart[q] + s;
  }
  std::vector<uint64_t> keys(bstart[nb] + 1);
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t *ct = &cnt[(size_t)t * nb];
    for (uint64_t e = tstart[t]; e < tstart[t + 1]; e++) {
      uint32_t r, cc;
      edge(e, r, cc);
      if (r >= row_begin && r < row_end) keys[ct[r >> shift]++] = ((uint64_t)r << 32) | cc;
    }
  }
  std::vector<uint64_t> uniq(nb + 1, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < (int64_t)nb; q++) {
    uint64_t *lo = keys.data() + bstart[q], *hi = keys.data() + bstart[q + 1];
    std::sort(lo, hi);
    uniq[q + 1] = (uint64_t)(std::unique(lo, hi) - lo);
  }
  for (uint32_t q = 0; q < nb; q++) uniq[q + 1] += uniq[q];
  Csr *A = new Csr();
  A->rows = row_end - row_begin; A->cols = n; A->is_double = is_double ? 1 : 0;
  // make the globally last row non-empty (reference reader defect Q3)
  const bool owns_last = row_end == n;
  bool need_last = false;
  if (owns_last) {
    const uint32_t q = (n - 1) >> shift;
    const uint64_t cntq = uniq[q + 1] - uniq[q];
    need_last = cntq == 0 || (keys[bstart[q] + cntq - 1] >> 32) != n - 1;
  }
  const uint64_t nnz = uniq[nb] + (need_last ? 1 : 0);
  A->row_ptr.assign((size_t)A->rows + 1, 0);
  A->col_ind.resize(nnz);
  A->values.resize(nnz * (is_double ? 8 : 4));
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < (int64_t)nb; q++) {
    const uint64_t *src = keys.data() + bstart[q];
    uint64_t dst = uniq[q];
    for (uint64_t i = 0; i < uniq[q + 1] - uniq[q]; i++, dst++) {
      const uint32_t r = (uint32_t)(src[i] >> 32), cc = (uint32_t)src[i];
      A->col_ind[dst] = cc;
      A->set_value(dst, value_at(seed, r, cc));
      A->row_ptr[(size_t)(r - row_begin) + 1]++;  // rows of a bucket belong to this thread only
    }
  }
  if (need_last) {
    A->col_ind[nnz - 1] = n - 1;
    A->set_value(nnz - 1, value_at(seed, n - 1, n - 1));
    A->row_ptr[A->rows]++;
  }
  for (uint32_t i = 0; i < A->rows; i++) A->row_ptr[i + 1] += A->row_ptr[i];
  *out = (spmvb_csr *)A;
  return SPMVB_OK;
}

}  // extern "C"
