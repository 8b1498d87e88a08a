// internal.h -- shared by the translation units of libntcard_b200 (not installed).
#pragma once
#include <stdint.h>

namespace ntc {
// Records the thread-local message returned by ntc_last_error() and returns `code`.
int set_err(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_err();
} // namespace ntc
