#include "errors.h"

#include "../../include/spmvb.h"

namespace spmvb {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) { g_err = msg; return code; }
}  // namespace spmvb

extern "C" const char *spmvb_last_error(void) {
  static thread_local std::string copy;
  copy = spmvb::g_err;  // stable storage for the caller: later failures do not change what was returned
  return copy.c_str();
}
