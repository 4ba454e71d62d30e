// api.cu - error reporting, launch accounting, ABI version.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <unordered_map>

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

bool first_use_on_device(const void* key) {
  static std::mutex mu;
  static std::unordered_map<const void*, uint64_t> seen;   // kernel -> bit mask of devices
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  uint64_t& mask = seen[key];
  const uint64_t bit = 1ull << dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// launch-shape knobs (FROST_TUNE_*): defaults are the values measured best on B200 (tools/microbench_ops.py)
static std::atomic<int> g_tune[FROST_TUNE_COUNT] = {};
static const int g_tune_default[FROST_TUNE_COUNT] = {
    /* DW_FWD_CTAS_PER_SM */ 2, /* DW_WGRAD_CTAS_PER_SM */ 3, /* BN_RED_CTAS_PER_SM */ 3, /* BN_RED_MAX_CGB */ 16,
    /* STEM_FWD_CTAS_PER_SM */ 8, /* STEM_WGRAD_CTAS_PER_SM */ 4, /* DW_DGRAD_CTAS_PER_SM */ 16,
    /* PDL */ 1, /* BN_RED_UNROLL */ 4, /* BN_APPLY_UNROLL */ 1, /* BNQ_UNROLL */ 2,
    /* DW_FWD_TILED */ 1, /* DW_DGRAD_TILED */ 1};
int tunable(int which) {
  const int v = g_tune[which].load(std::memory_order_relaxed);
  return v > 0 ? v : g_tune_default[which];
}
}  // namespace frost

extern "C" int frost_set_tunable(int which, int value) {
  if (which < 0 || which >= FROST_TUNE_COUNT || value < 0) {
    frost::set_error("frost_set_tunable: unknown knob %d or negative value %d", which, value);
    return FROST_EINVAL;
  }
  frost::g_tune[which].store(value, std::memory_order_relaxed);
  return FROST_OK;
}
extern "C" int frost_get_tunable(int which) {
  if (which < 0 || which >= FROST_TUNE_COUNT) return FROST_EINVAL;
  return frost::tunable(which);
}

extern "C" int frost_abi_version(void) { return 2; }
extern "C" const char* frost_last_error(void) { return frost::g_err; }
extern "C" int64_t frost_launch_count(void) { return frost::g_launches.load(); }
