// ctc_internal.h -- shared between the translation units of libctc_b200.so (not part of the public ABI).
#pragma once
#include <string>

namespace ctcb200 {
void ctcb200_set_error(const std::string &msg);   // what ctc_b200_last_error() returns on this thread
void ctcb200_count_launch();                      // the per-thread launch counter of ctc_b200_info()
}  // namespace ctcb200
