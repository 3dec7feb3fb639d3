// Weight-gradient GEMM on tcgen05 straight from ROW tile images: dW[ca, cb] = sum_r A[r, ca] * B[r, cb]
// (A = dY, B = layer input X), the reduction running over the rows.
//
// A row image block is 128 rows x 64 channels, K-major / SWIZZLE_128B (umma.cuh).  Read as a TRANSPOSED
// operand (MN = channel, K = row) the very same bytes are the canonical MN-major SWIZZLE_128B layout
//   ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)),  T = 8 bf16
// with the 8-row groups SBO = 1024 B apart and the next 64 channels (the next k-block of the image)
// LBO = 16 KB apart.  So the operands of dW are the images the forward / data-gradient GEMMs already
// use: no transposed copy, no zero padding of 64-channel layers to 128 rows.  Setting a_major = b_major = 1
// in the instruction descriptor is all the tensor core needs.
//
// Split-bf16 operands ([hi|hi|lo] images from pdf_rows_to_image(split=1)): per row tile three stages
// (hi.hi, hi.lo, lo.hi), each 8 x tcgen05.mma 128x128x16.  Split-K over batches of row tiles; the caller
// sums the per-batch partials.  Persistent kernel, 320 threads: producer warp, MMA warp, 8 epilogue warps.
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {
using namespace umma;

constexpr int TN_STAGES = 3;
constexpr int TN_BLOCK = 16384;
constexpr int TN_STAGE_BYTES = 4 * TN_BLOCK;             // A: 2 channel blocks, B: 2 channel blocks
constexpr int TN_THREADS = 320;
constexpr int TN_SMEM = TN_STAGES * TN_STAGE_BYTES + 1024 + 256;

struct TnParams {
  const uint8_t* a_img; const uint8_t* b_img;
  int a_nkb, b_nkb;              // 64-channel blocks per part (image has 3 * nkb k-blocks per row tile)
  int a_tiles, b_tiles;          // 128-channel output tiles of A (rows of dW) and B (columns of dW)
  int row_tiles, tiles_per_batch, batches;
  int split;                     // 1: [hi|hi|lo] images (3 stages per row tile), 0: plain bf16 images
  float* out; int64_t ld_out, batch_stride;
  int ca, cb;                    // valid channels
};

// MN-major SWIZZLE_128B descriptor: LBO = 16 KB between 64-channel blocks, SBO = 1 KB between 8-row groups
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TN_BLOCK >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void tn_mbar_arrive(uint32_t saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr) : "memory");
}

__global__ void __launch_bounds__(TN_THREADS, 1) gemm_tn_bf16_kernel(const TnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TN_STAGES * TN_STAGE_BYTES);
  // bars: [0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * TN_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[TN_STAGES + s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars[2 * TN_STAGES + a]), 1); mbar_init(smem_u32(&bars[2 * TN_STAGES + 2 + a]), 8); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(s_tmem);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const int per_batch = P.a_tiles * P.b_tiles;
  const int n_work = per_batch * P.batches;
  const int parts = P.split ? 3 : 1;
  const int a_kbt = parts * P.a_nkb, b_kbt = parts * P.b_nkb;       // k-blocks per row tile in each image

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int bt = w / per_batch, wr = w - bt * per_batch;
        const int at = wr / P.b_tiles, bn = wr % P.b_tiles;
        // one or two valid 64-channel blocks per operand tile (the second half of the stage then keeps stale
        // bytes: they only reach accumulator rows / columns the epilogue never stores)
        const uint32_t a_bytes = (2 * at + 1 < P.a_nkb) ? 2 * TN_BLOCK : TN_BLOCK;
        const uint32_t b_bytes = (2 * bn + 1 < P.b_nkb) ? 2 * TN_BLOCK : TN_BLOCK;
        const int t0 = bt * P.tiles_per_batch;
        const int t1 = min(t0 + P.tiles_per_batch, P.row_tiles);
        for (int t = t0; t < t1; ++t) {
          for (int p = 0; p < parts; ++p) {
            const int pa = p, pb = p == 0 ? 0 : 3 - p;                 // (hi,hi), (hi,lo), (lo,hi)
            mbar_wait(smem_u32(&bars[TN_STAGES + stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bars[stage]);
            mbar_expect_tx(full, a_bytes + b_bytes);
            uint8_t* st = smem + stage * TN_STAGE_BYTES;
            bulk_g2s(smem_u32(st), P.a_img + ((size_t)t * a_kbt + (size_t)pa * P.a_nkb + 2 * at) * TN_BLOCK, a_bytes, full);
            bulk_g2s(smem_u32(st + 2 * TN_BLOCK), P.b_img + ((size_t)t * b_kbt + (size_t)pb * P.b_nkb + 2 * bn) * TN_BLOCK,
                     b_bytes, full);
            if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      const uint32_t idesc = idesc_bf16(128, 128) | (1u << 15) | (1u << 16);      // A and B MN-major
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int bt = w / per_batch;
        const int t0 = bt * P.tiles_per_batch;
        const int t1 = min(t0 + P.tiles_per_batch, P.row_tiles);
        mbar_wait(smem_u32(&bars[2 * TN_STAGES + 2 + as]), aphase ^ 1);
        fence_after_sync();
        const uint32_t acc = tmem_base + as * 128;
        bool first = true;
        for (int t = t0; t < t1; ++t) {
          for (int p = 0; p < parts; ++p) {
            mbar_wait(smem_u32(&bars[stage]), phase);
            fence_after_sync();
            const uint32_t sa = smem_u32(smem + stage * TN_STAGE_BYTES), sb = sa + 2 * TN_BLOCK;
#pragma unroll
            for (int k16 = 0; k16 < 8; ++k16) {                        // 16 rows = two 8-row groups per step
              mma_bf16(acc, desc_mn_sw128(sa + k16 * 2048), desc_mn_sw128(sb + k16 * 2048), idesc, !(first && k16 == 0));
            }
            first = false;
            commit(smem_u32(&bars[TN_STAGES + stage]));
            if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        commit(smem_u32(&bars[2 * TN_STAGES + as]));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int q4 = warp & 3;                               // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;                      // which 64 columns this warp drains
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = ((uint32_t)(q4 * 32)) << 16;
    int as = 0; uint32_t aphase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int bt = w / per_batch, wr = w - bt * per_batch;
      const int at = wr / P.b_tiles, bn = wr % P.b_tiles;
      mbar_wait(smem_u32(&bars[2 * TN_STAGES + as]), aphase);
      fence_after_sync();
      const uint32_t acc = tmem_base + as * 128 + lane_off;
      const int ca = at * 128 + row;                       // dW row (channel of A)
      const bool row_ok = ca < P.ca;
#pragma unroll 1
      for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(acc + c0, v);
        tmem_ld_wait();
        if (row_ok) {
          float* o = P.out + (int64_t)bt * P.batch_stride + (int64_t)ca * P.ld_out + bn * 128 + c0;
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (bn * 128 + c0 + q < P.cb) o[q] = __uint_as_float(v[q]);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) tn_mbar_arrive(smem_u32(&bars[2 * TN_STAGES + 2 + as]));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace pdf

extern "C" int pdf_gemm_tn_bf16(const void* a_img, int ca, const void* b_img, int cb, int64_t rows, int split,
                                int tiles_per_batch, float* out, int64_t ld_out, int64_t batch_stride, void* stream) {
  using namespace pdf;
  if (rows == 0) return PDF_OK;
  PDF_REQUIRE(a_img && b_img && out, PDF_ERR_BAD_ARG, "pdf_gemm_tn_bf16: null pointer");
  PDF_REQUIRE(rows > 0 && ca > 0 && cb > 0 && tiles_per_batch > 0 && ld_out >= cb && (split == 0 || split == 1),
              PDF_ERR_BAD_ARG, "pdf_gemm_tn_bf16: bad argument");
  TnParams P;
  memset(&P, 0, sizeof(P));
  P.a_img = (const uint8_t*)a_img; P.b_img = (const uint8_t*)b_img;
  P.a_nkb = (ca + 63) / 64; P.b_nkb = (cb + 63) / 64;
  P.a_tiles = (ca + 127) / 128; P.b_tiles = (cb + 127) / 128;
  P.row_tiles = (int)((rows + 127) / 128);
  P.tiles_per_batch = tiles_per_batch;
  P.batches = (P.row_tiles + tiles_per_batch - 1) / tiles_per_batch;
  P.split = split; P.out = out; P.ld_out = ld_out; P.batch_stride = batch_stride; P.ca = ca; P.cb = cb;
  PDF_REQUIRE((int64_t)P.a_tiles * P.b_tiles * P.batches < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_gemm_tn_bf16: too much work");
  static pdf::PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(gemm_tn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_work = (int64_t)P.a_tiles * P.b_tiles * P.batches;
  gemm_tn_bf16_kernel<<<(unsigned)(n_work < sms ? n_work : sms), TN_THREADS, TN_SMEM, (cudaStream_t)stream>>>(P);
  return check_launch("pdf_gemm_tn_bf16");
}
