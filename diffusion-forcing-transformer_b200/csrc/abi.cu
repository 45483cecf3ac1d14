// C-ABI bookkeeping: version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dfot {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Latency mode (dfot_set_latency_mode / DFOT_LATENCY_MODE=1): see include/dfot_b200.h
static std::atomic<int> g_latency{-1};
bool latency_mode() {
  int v = g_latency.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("DFOT_LATENCY_MODE");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_latency.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}
}  // namespace dfot

extern "C" {
int dfot_abi_version(void) { return DFOT_ABI_VERSION; }
const char* dfot_last_error(void) { return dfot::g_err; }
int64_t dfot_launch_count(void) { return dfot::g_launches.load(std::memory_order_relaxed); }
int dfot_set_latency_mode(int on) {
  dfot::g_latency.store(on ? 1 : 0, std::memory_order_relaxed);
  return DFOT_OK;
}
int dfot_get_latency_mode(void) { return dfot::latency_mode() ? 1 : 0; }
}
