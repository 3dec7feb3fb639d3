// Pixel->point gather from a bf16 channels-last pyramid (what an autocast, channels-last RGB neck emits),
// writing the condition rows DIRECTLY as the bf16 tile images the SFT GEMMs read (gemm_bf16.cu): no fp32
// condition rows and no rows_to_image pass in between.  Algorithmic traffic per cloud (SURVEY 8d, bf16
// features): reads 1024*3*2 + 512*64*2 + 128*256*2 + 1024*8 (choose) + 1024*12 (xyz), writes
// 1024*12 + 65,536 + 65,536.
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {
using namespace umma;

__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
__device__ __forceinline__ float leaky01_b(float v) { return v > 0.f ? v : 0.1f * v; }

// sft0: 3 -> 3 -> 3 scale and shift branches (intaghand_encoder.py:205-219), fp32; same arithmetic as
// gather.cu::sft0_apply (the level-0 SFT decides the neighbour indices).
__device__ __forceinline__ void sft0_apply_b(const float* __restrict__ P, const float e[3], float xyz[3]) {
  float out[2][3];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const float* W0 = P + br * 24;
    const float* b0 = W0 + 9;
    const float* W1 = b0 + 3;
    const float* b1 = W1 + 9;
    float h[3];
#pragma unroll
    for (int o = 0; o < 3; ++o)
      h[o] = leaky01_b(fmaf(W0[o * 3 + 2], e[2], fmaf(W0[o * 3 + 1], e[1], fmaf(W0[o * 3], e[0], b0[o]))));
#pragma unroll
    for (int o = 0; o < 3; ++o)
      out[br][o] = fmaf(W1[o * 3 + 2], h[2], fmaf(W1[o * 3 + 1], h[1], fmaf(W1[o * 3], h[0], b1[o])));
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) xyz[o] = __fadd_rn(__fmul_rn(xyz[o], __fadd_rn(out[0][o], 1.f)), out[1][o]);
}

// rows [lo, hi) of one level: thread = one 16 B chunk (8 channels) of one point row; U independent loads in
// flight.  Destination: image row (cloud * n + i), k-block = chunk / 8, SW128 position inside the block.
__device__ __forceinline__ void gather_rows_to_image(const uint16_t* __restrict__ src, int C, int R, int Rl, int div,
                                                     const int64_t* __restrict__ ch, uint8_t* __restrict__ img,
                                                     int64_t row0, int lo, int hi) {
  const int cq = C >> 3, kb = C >> 6;                         // 16 B chunks per row, k-blocks per row tile
  constexpr int U = 8;
  const int e_lo = lo * cq, e_hi = hi * cq;
  for (int e0 = e_lo + threadIdx.x; e0 < e_hi; e0 += blockDim.x * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < e_hi) {
        const int i = e / cq, q = e - i * cq;
        int64_t pix = ch[i];
        pix = pix < 0 ? 0 : (pix >= (int64_t)R * R ? (int64_t)R * R - 1 : pix);
        const int64_t p = (int64_t)((int)pix / R / div) * Rl + ((int)pix % R) / div;   // intaghand_encoder.py:125-126
        v[u] = __ldg(reinterpret_cast<const uint4*>(src + p * C) + q);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < e_hi) {
        const int i = e / cq, q = e - i * cq;
        const int64_t row = row0 + i;
        uint8_t* blk = img + ((size_t)(row >> 7) * kb + (q >> 3)) * 16384;
        *reinterpret_cast<uint4*>(blk + sw128_off((uint32_t)(row & 127), (uint32_t)((q & 7) * 8))) = v[u];
      }
    }
  }
}

constexpr int PGB_PARTS = 4;                                  // CTAs per cloud and level

__global__ void __launch_bounds__(256)
pyramid_gather_bf16_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ choose, int clouds_per_frame,
                           int n_points, int n1, int n2, int R, const uint16_t* __restrict__ l0,
                           const uint16_t* __restrict__ l1, int C1, const uint16_t* __restrict__ l2, int C2,
                           const float* __restrict__ sft0, float* __restrict__ pts0, uint8_t* __restrict__ img1,
                           uint8_t* __restrict__ img2) {
  __shared__ float P[48];
  pdl_wait();
  pdl_trigger();
  const int64_t b = blockIdx.x;
  const int64_t f = b / clouds_per_frame;
  const int64_t* ch = choose + b * n_points;
  const int R2 = R / 2, R4 = R / 4;
  const int y = blockIdx.y;
  if (y == 0) {
    if (threadIdx.x < 48) P[threadIdx.x] = sft0[threadIdx.x];
    __syncthreads();
    const uint16_t* base = l0 + f * 3 * (int64_t)R * R;
    constexpr int U0 = 4;                                     // points per thread with every load issued up front
    for (int i0 = threadIdx.x; i0 < n_points; i0 += blockDim.x * U0) {
      int64_t pix[U0];
      float e[U0][3], p[U0][3];
#pragma unroll
      for (int u = 0; u < U0; ++u) {
        const int i = i0 + u * blockDim.x;
        pix[u] = i < n_points ? ch[i] : 0;
        pix[u] = pix[u] < 0 ? 0 : (pix[u] >= (int64_t)R * R ? (int64_t)R * R - 1 : pix[u]);
      }
#pragma unroll
      for (int u = 0; u < U0; ++u) {
        const int i = min(i0 + u * (int)blockDim.x, n_points - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          e[u][c] = bf16_to_f32(__ldg(base + pix[u] * 3 + c));
          p[u][c] = xyz[(b * n_points + i) * 3 + c];
        }
      }
#pragma unroll
      for (int u = 0; u < U0; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n_points) {
          sft0_apply_b(P, e[u], p[u]);
#pragma unroll
          for (int c = 0; c < 3; ++c) pts0[(b * n_points + i) * 3 + c] = p[u][c];
        }
      }
    }
  } else if (y <= PGB_PARTS) {
    const int part = y - 1, per = (n1 + PGB_PARTS - 1) / PGB_PARTS;
    gather_rows_to_image(l1 + f * C1 * (int64_t)R2 * R2, C1, R, R2, 2, ch, img1, b * n1, min(n1, part * per),
                         min(n1, (part + 1) * per));
  } else {
    const int part = y - 1 - PGB_PARTS, per = (n2 + PGB_PARTS - 1) / PGB_PARTS;
    gather_rows_to_image(l2 + f * C2 * (int64_t)R4 * R4, C2, R, R4, 4, ch, img2, b * n2, min(n2, part * per),
                         min(n2, (part + 1) * per));
  }
}

}  // namespace pdf

extern "C" int pdf_pyramid_gather_bf16(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                                       int n_points, int n1, int n2, int R, const void* l0, const void* l1, int C1,
                                       const void* l2, int C2, const float* sft0_params, float* pts0, void* cond1_img,
                                       void* cond2_img, void* stream) {
  PDF_REQUIRE(xyz && choose && l0 && l1 && l2 && sft0_params && pts0 && cond1_img && cond2_img, PDF_ERR_BAD_ARG,
              "pdf_pyramid_gather_bf16: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && clouds_per_frame > 0 && n_points > 0 && n1 >= 0 && n2 >= 0 && n1 <= n_points &&
                  n2 <= n_points && R >= 4,
              PDF_ERR_BAD_ARG, "pdf_pyramid_gather_bf16: bad size");
  PDF_REQUIRE(C1 > 0 && C2 > 0 && C1 % 64 == 0 && C2 % 64 == 0, PDF_ERR_UNSUPPORTED,
              "pdf_pyramid_gather_bf16: channel counts must be multiples of 64 (got %d, %d)", C1, C2);
  PDF_REQUIRE((n_clouds * n1) % 128 == 0 && (n_clouds * n2) % 128 == 0, PDF_ERR_UNSUPPORTED,
              "pdf_pyramid_gather_bf16: clouds * n1 and clouds * n2 must fill whole 128-row tiles");
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(n_clouds < (1ll << 31) && (int64_t)R * R < (1ll << 31), PDF_ERR_UNSUPPORTED,
              "pdf_pyramid_gather_bf16: too large");
  pdf::launch_pdl(pdf::pyramid_gather_bf16_kernel, dim3((unsigned)n_clouds, 1 + 2 * pdf::PGB_PARTS), dim3(256), 0,
                  (cudaStream_t)stream, xyz, choose, clouds_per_frame, n_points, n1, n2, R, (const uint16_t*)l0,
                  (const uint16_t*)l1, C1, (const uint16_t*)l2, C2, sft0_params, pts0, (uint8_t*)cond1_img,
                  (uint8_t*)cond2_img);
  return pdf::check_launch("pdf_pyramid_gather_bf16");
}
