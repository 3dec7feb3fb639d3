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

// Device-side alias of a page-locked, mapped HOST buffer (cudaHostAlloc / cudaHostRegister memory): kernels that only
// touch a sparse subset of a large host-resident input (pdf_pyramid_gather_bf16 reads 2 x 1664 pixels of a frame's
// feature pyramid) can read it in place over the PCIe link instead of having the whole tensor copied first.
extern "C" int pdf_host_device_pointer(const void* host, void** device_out) {
  PDF_REQUIRE(host && device_out, PDF_ERR_BAD_ARG, "pdf_host_device_pointer: null pointer");
  void* d = nullptr;
  const cudaError_t e = cudaHostGetDevicePointer(&d, const_cast<void*>(host), 0);
  if (e != cudaSuccess || d == nullptr) {
    cudaGetLastError();
    pdf::set_error("pdf_host_device_pointer: %p is not mapped page-locked host memory (%s)", host, cudaGetErrorString(e));
    return PDF_ERR_BAD_ARG;
  }
  *device_out = d;
  return PDF_OK;
}
