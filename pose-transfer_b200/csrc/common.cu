#include <stdarg.h>
#include "common.cuh"

namespace ptk {
thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace ptk

extern "C" int ptk_version(void) { return 100; }
extern "C" const char* ptk_last_error(void) { return ptk::g_err; }
extern "C" int64_t ptk_launch_count(void) { return ptk::g_launches.load(); }
