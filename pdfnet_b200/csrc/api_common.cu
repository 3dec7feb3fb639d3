// Error reporting and bookkeeping shared by every entry point.
#include <atomic>
#include <stdlib.h>
#include "pdf_common.cuh"

namespace pdf {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = getenv("PDF_NO_PDL") == nullptr;
  return on;
}
}  // namespace pdf

extern "C" int pdf_version(void) { return 100; }
extern "C" const char* pdf_last_error(void) { return pdf::g_err; }
extern "C" int64_t pdf_launch_count(void) { return (int64_t)pdf::g_launches.load(); }
