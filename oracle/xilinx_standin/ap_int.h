// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for the Xilinx Vivado-HLS <ap_int.h> header, written from
// scratch so that the *unmodified* reference sources under /root/reference/src
// compile with plain g++ (the reference needs ap_uint<W> only as a bit-slicing
// container: src/util.h:9-16,69; .range() uses at src/csr_hw.cpp:288-310,
// src/spmv.cpp:45,58,86-88,115).  Requirements collected in SURVEY.md App. B:
//   * exact object size (ap_uint<32> is written through by fscanf("%u"),
//     src/csr.cpp:21; ap_uint<128> is sized by sizeof, src/csr_hw.cpp:180)
//   * one implicit conversion to a builtin integer so arithmetic, comparison,
//     indexing and iostream output work without operator overloads
//   * .range(hi,lo) proxy, readable and assignable
// The reference sources also rely on <ap_int.h> dragging in iostream/cstdio/cstdlib.
#ifndef ORACLE_STANDIN_AP_INT_H
#define ORACLE_STANDIN_AP_INT_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <type_traits>

namespace standin_detail {
template <int W> struct storage_for {
  typedef typename std::conditional<(W <= 8), uint8_t,
          typename std::conditional<(W <= 16), uint16_t,
          typename std::conditional<(W <= 32), uint32_t, uint64_t>::type>::type>::type type;
};
inline uint64_t low_mask(int nbits) { return nbits >= 64 ? ~uint64_t(0) : ((uint64_t(1) << nbits) - 1); }
}  // namespace standin_detail

template <int W> struct ap_uint;

// Proxy for bits [hi:lo] of a little-endian array of 64-bit limbs (or of a
// narrower scalar, addressed as one limb).  Field width never exceeds 64.
template <typename Limb, int NL>
struct ap_range_ref {
  Limb *limbs;
  int hi, lo;
  ap_range_ref(Limb *l, int h, int o) : limbs(l), hi(h), lo(o) {}
  uint64_t get() const {
    const int width = hi - lo + 1;
    const int lbits = int(sizeof(Limb)) * 8;
    const int li = lo / lbits, off = lo % lbits;
    uint64_t v = uint64_t(limbs[li]) >> off;
    if (off + width > lbits && li + 1 < NL) v |= uint64_t(limbs[li + 1]) << (lbits - off);
    return v & standin_detail::low_mask(width);
  }
  void set(uint64_t v) {
    const int width = hi - lo + 1;
    const int lbits = int(sizeof(Limb)) * 8;
    const int li = lo / lbits, off = lo % lbits;
    const uint64_t m = standin_detail::low_mask(width);
    v &= m;
    uint64_t cur = uint64_t(limbs[li]);
    cur = (cur & ~(m << off)) | (v << off);
    limbs[li] = Limb(cur);
    if (off + width > lbits && li + 1 < NL) {
      const int done = lbits - off;
      uint64_t cur2 = uint64_t(limbs[li + 1]);
      cur2 = (cur2 & ~(m >> done)) | (v >> done);
      limbs[li + 1] = Limb(cur2);
    }
  }
  operator uint64_t() const { return get(); }
  ap_range_ref &operator=(uint64_t v) { set(v); return *this; }
  ap_range_ref &operator=(const ap_range_ref &o) { set(o.get()); return *this; }
};

template <int W>
struct ap_uint {
  typedef typename standin_detail::storage_for<W>::type store_t;
  store_t v;
  ap_uint() : v(0) {}
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_uint(T x) : v(store_t(uint64_t(x) & standin_detail::low_mask(W))) {}
  template <typename L, int N>
  ap_uint(const ap_range_ref<L, N> &r) : v(store_t(r.get() & standin_detail::low_mask(W))) {}
  template <int W2>
  ap_uint(const ap_uint<W2> &o) : v(store_t(uint64_t(o) & standin_detail::low_mask(W))) {}
  operator uint64_t() const { return uint64_t(v); }
  ap_uint &operator++() { v = store_t((uint64_t(v) + 1) & standin_detail::low_mask(W)); return *this; }
  ap_uint operator++(int) { ap_uint t(*this); ++*this; return t; }
  ap_uint &operator--() { v = store_t((uint64_t(v) - 1) & standin_detail::low_mask(W)); return *this; }
  ap_uint operator--(int) { ap_uint t(*this); --*this; return t; }
  ap_uint &operator+=(uint64_t x) { v = store_t((uint64_t(v) + x) & standin_detail::low_mask(W)); return *this; }
  ap_uint &operator-=(uint64_t x) { v = store_t((uint64_t(v) - x) & standin_detail::low_mask(W)); return *this; }
  ap_range_ref<store_t, 1> range(int hi, int lo) { return ap_range_ref<store_t, 1>(&v, hi, lo); }
  ap_range_ref<const store_t, 1> range(int hi, int lo) const { return ap_range_ref<const store_t, 1>(&v, hi, lo); }
};

// 128-bit bus word: two little-endian 64-bit limbs, sizeof == 16.
template <>
struct ap_uint<128> {
  uint64_t limb[2];
  ap_uint() { limb[0] = limb[1] = 0; }
  template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_uint(T x) { limb[0] = uint64_t(x); limb[1] = 0; }
  operator uint64_t() const { return limb[0]; }
  ap_range_ref<uint64_t, 2> range(int hi, int lo) { return ap_range_ref<uint64_t, 2>(limb, hi, lo); }
  ap_range_ref<const uint64_t, 2> range(int hi, int lo) const { return ap_range_ref<const uint64_t, 2>(limb, hi, lo); }
};

static_assert(sizeof(ap_uint<1>) == 1, "ap_uint<1>");
static_assert(sizeof(ap_uint<16>) == 2, "ap_uint<16>");
static_assert(sizeof(ap_uint<32>) == 4, "ap_uint<32>");
static_assert(sizeof(ap_uint<128>) == 16, "ap_uint<128>");

#endif
