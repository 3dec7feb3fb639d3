// Shared helpers for the pdfnet_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

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
