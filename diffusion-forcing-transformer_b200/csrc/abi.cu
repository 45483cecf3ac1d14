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
}  // namespace dfot

extern "C" {
int dfot_abi_version(void) { return DFOT_ABI_VERSION; }
const char* dfot_last_error(void) { return dfot::g_err; }
int64_t dfot_launch_count(void) { return dfot::g_launches.load(std::memory_order_relaxed); }
}
