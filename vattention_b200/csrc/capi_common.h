// shared by the C-ABI translation units: thread-local last-error string and the
// exception -> status-code mapping.
#pragma once
#include <string>

namespace vattn {
extern thread_local std::string g_last_error;
// call inside a catch(...) block; rethrows and maps to a VATTN_ERR_* code
int translate_exception();
}  // namespace vattn
