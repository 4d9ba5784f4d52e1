// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for Vivado-HLS <hls_stream.h>:
// an unbounded FIFO; reading an empty stream throws (this is how the
// reference's nr_val floor defect, SURVEY.md Q1, becomes visible instead of
// silently blocking like hardware would).  Used by src/spmv.cpp:6-120.
#ifndef ORACLE_STANDIN_HLS_STREAM_H
#define ORACLE_STANDIN_HLS_STREAM_H
#include <deque>
#include <stdexcept>
#include <string>
namespace hls {
template <typename T>
class stream {
  std::deque<T> q_;
  std::string name_;
 public:
  stream() {}
  explicit stream(const char *name) : name_(name ? name : "") {}
  stream &operator<<(const T &v) { q_.push_back(v); return *this; }
  void write(const T &v) { q_.push_back(v); }
  T read() {
    if (q_.empty()) throw std::runtime_error("hls::stream '" + name_ + "': read on empty stream");
    T v = q_.front();
    q_.pop_front();
    return v;
  }
  stream &operator>>(T &v) { v = read(); return *this; }
  bool empty() const { return q_.empty(); }
  size_t size() const { return q_.size(); }
};
}  // namespace hls
#endif
