// Library-owned CSR matrices: reader/writer for the reference's matrix-file format and the synthetic generators
// for BASELINE.json's configs.  Reference: read_csr_header / read_csr_matrix, src/csr.cpp:10-46, 87-136 (format
// "%u %u %u\n" then "%u %u %lf\n" / "%u %u %f\n", 1-based, sorted by row; src/util.h:28-29).
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spmvb.h"
#include "csr.h"
#include "layout.h"

using namespace spmvb;

// One "row col value" entry inside [p, e): both indices and the value must be complete tokens that end before e.
static bool parse_entry(const char *p, const char *e, unsigned long &r, unsigned long &c, double &v) {
  auto skip = [&]() { while (p < e && (*p == ' ' || *p == '\t')) p++; };
  auto index = [&](unsigned long &out) {
    skip();
    if (p >= e || *p < '0' || *p > '9') return false;
    unsigned long long x = 0;
    for (; p < e && *p >= '0' && *p <= '9'; p++) {
      x = x * 10 + (unsigned)(*p - '0');
      if (x > 0xFFFFFFFFull) return false;
    }
    out = (unsigned long)x;
    return true;
  };
  if (!index(r) || !index(c)) return false;
  skip();
  const char *t0 = p;
  while (p < e && *p != ' ' && *p != '\t' && *p != '\r') p++;
  const size_t n = (size_t)(p - t0);
  if (n == 0 || n + 1 >= 64) return false;
  // plain decimal tokens (all a matrix file normally holds) in place, correctly rounded like strtod; anything else
  // that sscanf's %lf accepts - a leading '+', hex floats, out-of-range magnitudes - through strtod on a bounded copy
  const std::from_chars_result fr = std::from_chars(t0, p, v);
  if (fr.ec == std::errc() && fr.ptr == p) return true;
  char tok[64];
  memcpy(tok, t0, n);
  tok[n] = 0;
  char *end = nullptr;
  v = strtod(tok, &end);
  return end == tok + n;  // what follows the value on the line is ignored, like sscanf does
}

extern "C" {

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
  std::atomic<int> bad_line(0);
  // pass 2: parse.  Every line is parsed inside its own [begin, end): a line with a missing field is an error (the
  // reference's per-line sscanf reports it too, csr.cpp:111-113) and nothing is ever read beyond the mapping.
#pragma omp parallel num_threads(T)
  {
    const int t = omp_get_thread_num();
    uint64_t i = first[t];
    for (size_t p = cut[t]; p < cut[t + 1] && !bad_line.load(std::memory_order_relaxed);) {
      const char *nl = (const char *)memchr(buf + p, '\n', cut[t + 1] - p);
      const size_t e = nl ? (size_t)(nl - buf) : cut[t + 1];
      size_t q = p;
      while (q < e && (buf[q] == ' ' || buf[q] == '\t' || buf[q] == '\r')) q++;
      if (q < e) {
        unsigned long r = 0, c = 0;
        double v = 0.0;
        if (!parse_entry(buf + q, buf + e, r, c, v) || r < 1 || r > rows || c < 1 || c > cols) {
          bad_line.store(1, std::memory_order_relaxed);
          break;
        }
        row_of[i] = (uint32_t)(r - 1);
        A->col_ind[i] = (uint32_t)(c - 1);  // mmarket files are not zero-based, csr.cpp:118
        A->set_value(i, v);
        i++;
      }
      p = e + 1;
    }
  }
  munmap((void *)buf, size);
  int bad = bad_line.load();
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
  // nnz is bounded by the file size before it enters any multiplication (an absurd header must not wrap `want`)
  const uint64_t fixed = 8 + 16 + 8 + ((uint64_t)hdr[2] + 1) * 8;
  const bool stat_ok = fstat(fileno(fp), &st) == 0;
  const uint64_t fsize = stat_ok ? (uint64_t)st.st_size : 0;
  const bool nnz_ok = stat_ok && fsize >= fixed && nnz <= (fsize - fixed) / (4 + vb);
  const uint64_t want = nnz_ok ? fixed + nnz * (4 + vb) : 0;
  if (!nnz_ok || fsize != want) {
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

}  // extern "C"
