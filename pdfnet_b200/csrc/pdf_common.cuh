// Shared helpers for the pdfnet_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/pdfnet_b200.h"

namespace pdf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PDF_ERR_CUDA;
  }
  count_launch();
  return PDF_OK;
}

// cudaFuncSetAttribute is per device: returns true the first time it is called for the current
// device with this flag array (one process may drive several GPUs).
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// A step is a chain of ~200 short dependent kernels.  Launched with the programmatic-stream-serialization
// attribute, kernel N+1 is set up (launch processing, CTA scheduling, its on-chip prologue) while kernel N is
// still draining; every such kernel calls pdl_wait() before its first global-memory access (it returns once the
// preceding kernel has completed and its writes are visible) and pdl_trigger() right after (lets ITS successor
// start launching as soon as all of its CTAs are resident).  In a kernel launched the ordinary way both are
// no-ops, so every launch site may choose freely.  Captured into CUDA graphs as programmatic dependency edges.
// PDF_NO_PDL=1 in the environment turns the attribute off (plain stream order).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PDF_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      pdf::set_error(__VA_ARGS__);              \
      return (code);                            \
    }                                           \
  } while (0)

// fp32 squared distance exactly as the reference evaluates it:
// fl(fl(dx*dx + dy*dy) + dz*dz), no FMA contraction.
__device__ __forceinline__ float sqdist_rn(float px, float py, float pz, float cx, float cy, float cz) {
  float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace pdf
