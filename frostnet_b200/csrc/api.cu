// api.cu - error reporting, launch accounting, ABI version.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace frost {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace frost

extern "C" int frost_abi_version(void) { return 1; }
extern "C" const char* frost_last_error(void) { return frost::g_err; }
extern "C" int64_t frost_launch_count(void) { return frost::g_launches.load(); }
