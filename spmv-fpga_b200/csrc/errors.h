// Error state of the C ABI: the message of the last failure on the calling thread (spmvb_last_error).
#pragma once
#include <string>

namespace spmvb {
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);  // records msg, returns code
}  // namespace spmvb
