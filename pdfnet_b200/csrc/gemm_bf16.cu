// Streaming bf16 GEMM on tcgen05 for the dense layers that follow the set-abstraction
// stages: SFT 1x1 convs, the global point-MLP (netR_3) and the fusion SFT(1024,1024).
//
// Both operands are "tile images": a matrix [rows, K] in bf16 stored as
// [row-tile (128 rows)][k-block (64 columns)] blocks of 16 KB, each block already in the
// K-major / 128-byte-swizzle layout the tensor core reads from shared memory.  A block is
// therefore ONE contiguous cp.async.bulk (TMA engine, no tensor map), and the epilogue of
// one layer writes the operand image of the next layer directly.
//
// Persistent warp-specialised kernel, one CTA per SM, 320 threads:
//   warp 0  : producer  - bulk copies of (M-operand block, N-operand block) into a 4-stage ring
//   warp 1  : MMA issue - 4 x tcgen05.mma (128x128x16) per k-block, commit frees the stage
//   warps 2-9: epilogue - drain one of two TMEM accumulator stages while the other fills
//              (two warps per TMEM lane quarter, 64 columns each; biases staged in smem)
// Epilogues (thread = TMEM lane):
//   ROW   : lane = activation row.  y = act(D + bias[n]) or, with two accumulators (k-blocks
//           below / above kb_split accumulate separately), the SFT modulation
//           y = F*(D0 + b0 + 1) + (D1 + b1)  (intaghand_encoder.py:217-219).  Output as fp32
//           rows and/or as the bf16 image of the next layer.
//   COLMAX: lane = output channel (weights are the M operand), columns = the 128 points of
//           one cloud: y[cloud, c] = relu(max_p D[c,p] + bias[c])   (netR_3 + MaxPool, :86-103).
#include <stdlib.h>
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {
using namespace umma;

constexpr int G_STAGES = 4;
constexpr int G_BLOCK = 16384;                 // one 128 x 64 bf16 block
constexpr int G_THREADS = 320;                // producer warp + MMA warp + 8 epilogue warps
constexpr int G_SMEM = G_STAGES * 2 * G_BLOCK + 1024 + 256 + 2 * 20 * 128 * 4 + 400 * 4;   // + staged biases, xyz weights
constexpr int G_MAX_NT = 20;                  // N tiles per GEMM in ROW mode (MANO blend shapes: 2334 columns = 19 tiles)

struct GemmParams {
  const uint8_t* m_img; const uint8_t* n_img;
  int m_kb, n_kb;                // k-blocks per row-tile in each image
  int m_tiles, n_tiles, KB, kb_split;
  const float* bias0; const float* bias1;
  int act;
  float* out_f32; int64_t ld_out; int64_t rows_valid;
  const float* F; int64_t ldf;
  uint8_t* out_img; int out_kb;
  int out_split;                 // out_img is a SPLIT image [hi | hi | lo] (3 x out_kb/3 k-blocks): the next GEMM's fp32-accurate operand
  uint16_t* out_bf16; int64_t ld_bf16;   // optional bf16 ROW-major output (same columns as out_f32, minus bf16_col_off)
  int bf16_col_off;
  float* out_max; int64_t ld_max;
  // XYZ mode (SFT1 hidden layer): the 128 columns are [64 scale-hidden | 64 shift-hidden]; besides the
  // bf16 image of lrelu(D+b) the epilogue applies the SFT modulation to the 3 xyz channels in fp32:
  // x[m,c] = x[m,c]*(w1s[c].h_s + b1s[c] + 1) + (w1h[c].h_h + b1h[c]).  xyz_w = [2][3][64] then [2][3].
  const float* xyz_w; float* xyz_x; int64_t xyz_ld;
  // batched mode (split-K weight gradients): work item = (batch, m-tile, n-tile); operands and the
  // fp32 output advance by these strides per batch (bytes / floats)
  int batches; int64_t m_batch_stride, n_batch_stride, out_batch_stride;
  // grouped mode (the two hands of the GCN decoder in ONE launch): row tiles [0, group_m_tiles) use the first weight
  // image / bias, row tiles from group_m_tiles on the second (n_group_stride bytes / bias_group_stride floats further);
  // rows_valid then counts inside each group of group_m_tiles * 128 rows.  0 = off.
  int group_m_tiles; int64_t n_group_stride, bias_group_stride;
  int tile_col[G_MAX_NT];        // ROW mode, per N-tile: first fp32 column (F and out_f32)
  int tile_nvalid[G_MAX_NT];     //   valid output columns of this N-tile (others are written as 0 / skipped)
  int tile_okb[G_MAX_NT];        //   first k-block of this N-tile in the output image
};

__device__ __forceinline__ void mbar_arrive(uint32_t saddr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr) : "memory");
}

// LIGHT: two operand stages and 256 TMEM columns instead of four and 512, so TWO CTAs fit on an SM.  The decoder's
// GEMMs are short (3 - 24 k-blocks, a few hundred work items) and come in left-hand / right-hand pairs on two
// streams; a full-size CTA owns its SM (150 KB of shared memory, all of tensor memory), which serialises the pair
// and exposes every launch's fixed cost (~10 us of prologue, pipeline fill and drain).  Single accumulator only.
template <bool COLMAX, bool LIGHT = false>
__global__ void __launch_bounds__(G_THREADS, LIGHT ? 2 : 1) gemm_bf16_kernel(const GemmParams P) {
  constexpr int G_STAGES = LIGHT ? 2 : pdf::G_STAGES;
  constexpr int TMEM_COLS = LIGHT ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * 2 * G_BLOCK);
  // bars: [0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 4);
  // warp index through a shuffle: provably warp-uniform, so the role branches are uniform branches and everything the
  // MMA warp derives from uniform values (stage, accumulator column, descriptors) stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const bool dual = LIGHT ? false : P.kb_split > 0;          // LIGHT: single accumulator, no xyz side channel (compile-time)
  const int acc_cols = dual ? 256 : 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[G_STAGES + s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars[2 * G_STAGES + a]), 1); mbar_init(smem_u32(&bars[2 * G_STAGES + 2 + a]), COLMAX ? 4 : 8); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(s_tmem);
  pdl_wait();                                      // everything above is on-chip set-up; global memory from here on
  pdl_trigger();
  float* s_bias = reinterpret_cast<float*>(smem + G_STAGES * 2 * G_BLOCK + 256);   // [2][n_tiles*128]
  float* s_xyz = s_bias + 2 * G_MAX_NT * 128;                                     // [390]
  const bool xyz = !LIGHT && !COLMAX && P.xyz_w != nullptr;
  if (!COLMAX) {
    for (int i = threadIdx.x; i < P.n_tiles * 128; i += G_THREADS) {
      s_bias[i] = P.bias0 ? P.bias0[i] : 0.f;
      s_bias[G_MAX_NT * 128 + i] = dual ? P.bias1[i] : (P.group_m_tiles ? P.bias0[P.bias_group_stride + i] : 0.f);
    }
    if (xyz) for (int i = threadIdx.x; i < 390; i += G_THREADS) s_xyz[i] = P.xyz_w[i];
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
  const int per_batch = P.m_tiles * P.n_tiles;
  const int n_work = per_batch * P.batches;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int bt = w / per_batch, wr = w - bt * per_batch;
        const int mt = wr / P.n_tiles, nt = wr % P.n_tiles;
        const uint8_t* mb = P.m_img + (size_t)bt * P.m_batch_stride + (size_t)mt * P.m_kb * G_BLOCK;
        const uint8_t* nb = P.n_img + (size_t)bt * P.n_batch_stride + (size_t)nt * P.n_kb * G_BLOCK +
                            ((P.group_m_tiles && mt >= P.group_m_tiles) ? (size_t)P.n_group_stride : 0);
        for (int kb = 0; kb < P.KB; ++kb) {
          mbar_wait(smem_u32(&bars[G_STAGES + stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bars[stage]);
          mbar_expect_tx(full, 2 * G_BLOCK);
          // the M image may hold fewer k-blocks than KB: they are reused cyclically (a bf16-exact activation
          // against a [hi | lo] split weight image reads the same activation block for both products)
          bulk_g2s(smem_u32(smem + stage * 2 * G_BLOCK), mb + (size_t)(kb % P.m_kb) * G_BLOCK, G_BLOCK, full);
          bulk_g2s(smem_u32(smem + stage * 2 * G_BLOCK + G_BLOCK), nb + (size_t)kb * G_BLOCK, G_BLOCK, full);
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the WHOLE warp runs the loop (uniform control flow, all operands in uniform registers); lane 0 alone issues
    // the MMAs and their commits
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    const uint32_t idesc = idesc_bf16(128, 128);
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      mbar_wait(smem_u32(&bars[2 * G_STAGES + 2 + as]), aphase ^ 1);
      fence_after_sync();
      const uint32_t acc = tmem_base + as * acc_cols;
      for (int kb = 0; kb < P.KB; ++kb) {
        mbar_wait(smem_u32(&bars[stage]), phase);
        fence_after_sync();
        const uint32_t sm = smem_u32(smem + stage * 2 * G_BLOCK), sn = sm + G_BLOCK;
        const bool second = dual && kb >= P.kb_split;
        const uint32_t d = acc + (second ? 128 : 0);
        const bool first_kb = second ? (kb == P.kb_split) : (kb == 0);
        if (lane == 0) {
          const uint64_t da = desc_sw128(sm), db = desc_sw128(sn);     // +32 B per K step = +2 in the address field
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16)
            mma_bf16(d, da + 2 * k16, db + 2 * k16, idesc, !(first_kb && k16 == 0));
          commit(smem_u32(&bars[G_STAGES + stage]));
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) commit(smem_u32(&bars[2 * G_STAGES + as]));
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (!COLMAX || warp < 6) {
    const int q4 = warp & 3;                               // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;                      // ROW mode: which 64 columns this warp drains
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = ((uint32_t)(q4 * 32)) << 16;
    int as = 0; uint32_t aphase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int bt = w / per_batch, wr = w - bt * per_batch;
      const int mt = wr / P.n_tiles, nt = wr % P.n_tiles;
      mbar_wait(smem_u32(&bars[2 * G_STAGES + as]), aphase);
      fence_after_sync();
      const uint32_t acc = tmem_base + as * acc_cols + lane_off;
      if (COLMAX) {
        float mx = -3.0e38f;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(acc + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; q += 2) mx = max3(mx, __uint_as_float(v[q]), __uint_as_float(v[q + 1]));
        }
        const int ch = mt * 128 + row;
        P.out_max[(int64_t)nt * P.ld_max + ch] = fmaxf(mx + P.bias0[ch], 0.f);
      } else {
        const int64_t m = (int64_t)mt * 128 + row;
        const bool grp1 = P.group_m_tiles && mt >= P.group_m_tiles;
        const bool row_ok = (grp1 ? m - (int64_t)P.group_m_tiles * 128 : m) < P.rows_valid;
        const int col0 = P.tile_col[nt], nvalid = P.tile_nvalid[nt], okb = P.tile_okb[nt];
        const uint32_t b0 = smem_u32(s_bias) + nt * 512 + (grp1 ? G_MAX_NT * 512 : 0);   // shared addresses (explicit LDS)
        const uint32_t b1 = b0 + G_MAX_NT * 512;
        const uint32_t sxyz = smem_u32(s_xyz);
        const bool full = row_ok && nvalid == 128;           // fast path: no per-element masking
        float xo[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        const int c_begin = xyz ? (half == 0 ? 0 : 128) : half * 64, c_end = xyz ? 128 : half * 64 + 64;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t v[32], u[32];
          float4 fin[8];                                   // all eight F loads of this chunk in flight at once
          if (dual) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              fin[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row_ok && c0 + 4 * q < nvalid)
                fin[q] = __ldg(reinterpret_cast<const float4*>(P.F + m * P.ldf + col0 + c0) + q);
            }
          }
          tmem_ld32(acc + c0, v);
          if (dual) tmem_ld32(acc + 128 + c0, u);
          tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 bb = ld_shared_f4(b0 + (c0 + q) * 4);
            float r[4] = {__uint_as_float(v[q]) + bb.x, __uint_as_float(v[q + 1]) + bb.y,
                          __uint_as_float(v[q + 2]) + bb.z, __uint_as_float(v[q + 3]) + bb.w};
            if (dual) {
              const float4 t = fin[q >> 2];
              const float4 b2 = ld_shared_f4(b1 + (c0 + q) * 4);
              r[0] = fmaf(t.x, r[0] + 1.f, __uint_as_float(u[q]) + b2.x);
              r[1] = fmaf(t.y, r[1] + 1.f, __uint_as_float(u[q + 1]) + b2.y);
              r[2] = fmaf(t.z, r[2] + 1.f, __uint_as_float(u[q + 2]) + b2.z);
              r[3] = fmaf(t.w, r[3] + 1.f, __uint_as_float(u[q + 3]) + b2.w);
            } else if (P.act == PDF_ACT_RELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) r[e] = fmaxf(r[e], 0.f);
            } else if (P.act == PDF_ACT_LEAKY01) {
#pragma unroll
              for (int e = 0; e < 4; ++e) r[e] = fmaxf(r[e], 0.1f * r[e]);
            }
            if (!full) {
#pragma unroll
              for (int e = 0; e < 4; ++e) r[e] = (row_ok && c0 + q + e < nvalid) ? r[e] : 0.f;
            }
            y[q] = r[0]; y[q + 1] = r[1]; y[q + 2] = r[2]; y[q + 3] = r[3];
          }
          if (xyz) {                                         // fp32 dots with the 3 xyz rows of the second conv
            const int br = c0 >> 6;
            const uint32_t w = sxyz + (br * 192 + (c0 & 63)) * 4;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float4 wv = ld_shared_f4(w + (c * 64 + q) * 4);
                xo[br][c] = fmaf(wv.x, y[q], fmaf(wv.y, y[q + 1], fmaf(wv.z, y[q + 2], fmaf(wv.w, y[q + 3], xo[br][c]))));
              }
            }
          }
          if (P.out_f32 != nullptr && row_ok) {
            float* o = P.out_f32 + (int64_t)bt * P.out_batch_stride + m * P.ld_out + col0 + c0;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              if (c0 + q + 3 < nvalid) *reinterpret_cast<float4*>(o + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
              else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c0 + q + e < nvalid) o[q + e] = y[q + e];
              }
            }
          }
          if (P.out_bf16 != nullptr && row_ok) {
            uint16_t* o = P.out_bf16 + m * P.ld_bf16 + (col0 - P.bf16_col_off) + c0;
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              if (c0 + q + 7 < nvalid) {
                uint4 wv;
                wv.x = pack_bf16(y[q], y[q + 1]); wv.y = pack_bf16(y[q + 2], y[q + 3]);
                wv.z = pack_bf16(y[q + 4], y[q + 5]); wv.w = pack_bf16(y[q + 6], y[q + 7]);
                *reinterpret_cast<uint4*>(o + q) = wv;
              }
            }
          }
          if (P.out_img != nullptr && c0 < ((nvalid + 63) & ~63)) {     // k-blocks past the valid columns do not exist
            uint8_t* blk = P.out_img + ((size_t)mt * P.out_kb + okb + (c0 >> 6)) * G_BLOCK;
            const size_t part = (size_t)(P.out_kb / 3) * G_BLOCK;       // split image: distance between the three parts
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              uint4 wv;
              wv.x = pack_bf16(y[q], y[q + 1]); wv.y = pack_bf16(y[q + 2], y[q + 3]);
              wv.z = pack_bf16(y[q + 4], y[q + 5]); wv.w = pack_bf16(y[q + 6], y[q + 7]);
              const uint32_t off = sw128_off(row, (c0 + q) & 63);
              *reinterpret_cast<uint4*>(blk + off) = wv;
              if (P.out_split) {                                          // [hi | hi | lo]
                uint4 wl;
                wl.x = pack_bf16(y[q] - __uint_as_float(wv.x << 16), y[q + 1] - __uint_as_float(wv.x & 0xffff0000u));
                wl.y = pack_bf16(y[q + 2] - __uint_as_float(wv.y << 16), y[q + 3] - __uint_as_float(wv.y & 0xffff0000u));
                wl.z = pack_bf16(y[q + 4] - __uint_as_float(wv.z << 16), y[q + 5] - __uint_as_float(wv.z & 0xffff0000u));
                wl.w = pack_bf16(y[q + 6] - __uint_as_float(wv.w << 16), y[q + 7] - __uint_as_float(wv.w & 0xffff0000u));
                *reinterpret_cast<uint4*>(blk + part + off) = wv;
                *reinterpret_cast<uint4*>(blk + 2 * part + off) = wl;
              }
            }
          }
        }
        if (xyz && half == 0 && row_ok) {
          float* xr = P.xyz_x + m * P.xyz_ld;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            xr[c] = __fadd_rn(__fmul_rn(xr[c], __fadd_rn(xo[0][c] + ld_shared_f1(sxyz + (384 + c) * 4), 1.f)),
                              xo[1][c] + ld_shared_f1(sxyz + (387 + c) * 4));
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[2 * G_STAGES + 2 + as]));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

constexpr int G_SMEM_LIGHT = 2 * 2 * G_BLOCK + 1024 + 256 + 2 * G_MAX_NT * 128 * 4 + 400 * 4;

// fp32 rows [M, ld] columns [col0, col0+K) -> bf16 image k-blocks [kb0, kb0 + ceil(K/64)) of every
// row-tile; rows >= M and columns >= K are written as zeros.  One thread per 16-byte chunk.
__global__ void rows_to_image_kernel(const float* __restrict__ X, int64_t ld, int64_t M, int col0, int K,
                                     uint8_t* __restrict__ img, int kb_total, int kb0, int nkb, int64_t n_chunks,
                                     int split) {
  pdl_wait();
  pdl_trigger();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_chunks;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(e & 7);                         // 16 B chunk inside a 64-column block row
    const int64_t r_all = e >> 3;
    const int kb = (int)(r_all % nkb);
    const int64_t m = r_all / nkb;
    const int k = kb * 64 + ch * 8;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (m < M && k + i < K) ? X[m * ld + col0 + k + i] : 0.f;
    uint4 w;
    w.x = pack_bf16(f[0], f[1]); w.y = pack_bf16(f[2], f[3]); w.z = pack_bf16(f[4], f[5]); w.w = pack_bf16(f[6], f[7]);
    const int64_t mt = m >> 7;
    const uint32_t off = sw128_off((uint32_t)(m & 127), ch * 8);
    uint8_t* tile = img + ((size_t)mt * kb_total + kb0) * G_BLOCK;
    *reinterpret_cast<uint4*>(tile + (size_t)kb * G_BLOCK + off) = w;
    if (split) {                                         // fp32-accurate products on bf16 tensor cores:
      float l[8];                                        // split 1 = [hi | hi | lo], split 2 = [hi | lo | hi]
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const uint32_t pk = i == 0 ? w.x : (i == 2 ? w.y : (i == 4 ? w.z : w.w));
        l[i] = f[i] - __uint_as_float(pk << 16);
        l[i + 1] = f[i + 1] - __uint_as_float(pk & 0xffff0000u);
      }
      uint4 wl;
      wl.x = pack_bf16(l[0], l[1]); wl.y = pack_bf16(l[2], l[3]); wl.z = pack_bf16(l[4], l[5]); wl.w = pack_bf16(l[6], l[7]);
      *reinterpret_cast<uint4*>(tile + (size_t)(nkb + kb) * G_BLOCK + off) = split == 2 ? wl : w;
      *reinterpret_cast<uint4*>(tile + (size_t)(2 * nkb + kb) * G_BLOCK + off) = split == 2 ? w : wl;
    }
  }
}

// Transposed image: fp32 rows X[M, ld], columns [col0, col0+C) -> bf16 image of X^T cut into batches of
// Mc rows of X (Mc % 64 == 0): image rows = channels (zero-padded to 128), k = row index inside the batch.
// Layout [batch][row-tile][k-block]; with split the k-blocks are tripled ([hi|hi|lo] or [hi|lo|hi]).
// One CTA = 64 rows of X x 128 channels, transposed through shared memory.
__global__ void __launch_bounds__(256)
rows_to_image_t_kernel(const float* __restrict__ X, int64_t ld, int64_t M, int col0, int C, uint8_t* __restrict__ img,
                       int64_t Mc, int split) {
  __shared__ float s[64][129];
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int rt = blockIdx.y, ch0 = rt * 128;
  for (int e = threadIdx.x; e < 64 * 128; e += 256) {
    const int r = e >> 7, c = e & 127;
    const int64_t m = m0 + r;
    s[r][c] = (m < M && ch0 + c < C) ? X[m * ld + col0 + ch0 + c] : 0.f;
  }
  __syncthreads();
  const int nkb = (int)(Mc / 64), parts = split ? 3 : 1, n_rt = (C + 127) / 128;
  const int64_t bt = m0 / Mc;
  const int kb = (int)((m0 - bt * Mc) / 64);
  uint8_t* base = img + (((size_t)bt * n_rt + rt) * (size_t)(nkb * parts)) * G_BLOCK;
  for (int e = threadIdx.x; e < 128 * 8; e += 256) {
    const int chunk = e & 7, c = e >> 3;                 // consecutive threads fill one 128 B image row
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = s[chunk * 8 + i][c];
    uint4 w;
    w.x = pack_bf16(f[0], f[1]); w.y = pack_bf16(f[2], f[3]); w.z = pack_bf16(f[4], f[5]); w.w = pack_bf16(f[6], f[7]);
    const uint32_t off = sw128_off((uint32_t)c, chunk * 8);
    *reinterpret_cast<uint4*>(base + (size_t)kb * G_BLOCK + off) = w;
    if (split) {
      float l[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const uint32_t pk = i == 0 ? w.x : (i == 2 ? w.y : (i == 4 ? w.z : w.w));
        l[i] = f[i] - __uint_as_float(pk << 16);
        l[i + 1] = f[i + 1] - __uint_as_float(pk & 0xffff0000u);
      }
      uint4 wl;
      wl.x = pack_bf16(l[0], l[1]); wl.y = pack_bf16(l[2], l[3]); wl.z = pack_bf16(l[4], l[5]); wl.w = pack_bf16(l[6], l[7]);
      *reinterpret_cast<uint4*>(base + (size_t)(nkb + kb) * G_BLOCK + off) = split == 2 ? wl : w;
      *reinterpret_cast<uint4*>(base + (size_t)(2 * nkb + kb) * G_BLOCK + off) = split == 2 ? w : wl;
    }
  }
}

// SFT on the three xyz channels in full fp32 (they feed the level-2 neighbour search):
// x[m,c] = x[m,c]*(scale_c+1)+shift_c, scale/shift = conv1(lrelu(conv0(cond))) restricted to c<3.
// Thread per row; conv0 weights (2 x [CC,CC]) broadcast from shared memory.
template <int CC>
__global__ void __launch_bounds__(128) sft_xyz_kernel(const float* __restrict__ cond, int64_t M,
                                                      const float* __restrict__ w0s, const float* __restrict__ b0s,
                                                      const float* __restrict__ w1s, const float* __restrict__ b1s,
                                                      const float* __restrict__ w0h, const float* __restrict__ b0h,
                                                      const float* __restrict__ w1h, const float* __restrict__ b1h,
                                                      float* __restrict__ x, int64_t ldx) {
  extern __shared__ float sw[];                            // [2][CC*CC] conv0, [2][CC] bias0, [2][3*CC] conv1 rows
  float* s_w0 = sw; float* s_b0 = s_w0 + 2 * CC * CC; float* s_w1 = s_b0 + 2 * CC;
  for (int i = threadIdx.x; i < CC * CC; i += blockDim.x) { s_w0[i] = w0s[i]; s_w0[CC * CC + i] = w0h[i]; }
  for (int i = threadIdx.x; i < CC; i += blockDim.x) { s_b0[i] = b0s[i]; s_b0[CC + i] = b0h[i]; }
  for (int i = threadIdx.x; i < 3 * CC; i += blockDim.x) { s_w1[i] = w1s[i]; s_w1[3 * CC + i] = w1h[i]; }
  __syncthreads();
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  float c[CC];
#pragma unroll
  for (int i = 0; i < CC; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(cond + m * CC + i);
    c[i] = t.x; c[i + 1] = t.y; c[i + 2] = t.z; c[i + 3] = t.w;
  }
  float out[2][3];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    const float* W0 = s_w0 + br * CC * CC;
    const float* W1 = s_w1 + br * 3 * CC;
#pragma unroll 4
    for (int j = 0; j < CC; ++j) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < CC; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(W0 + j * CC + i);
        a0 = fmaf(wv.x, c[i], a0); a1 = fmaf(wv.y, c[i + 1], a1);
        a0 = fmaf(wv.z, c[i + 2], a0); a1 = fmaf(wv.w, c[i + 3], a1);
      }
      float h = (a0 + a1) + s_b0[br * CC + j];
      h = h > 0.f ? h : 0.1f * h;
      o0 = fmaf(W1[j], h, o0); o1 = fmaf(W1[CC + j], h, o1); o2 = fmaf(W1[2 * CC + j], h, o2);
    }
    const float* b1 = br == 0 ? b1s : b1h;
    out[br][0] = o0 + b1[0]; out[br][1] = o1 + b1[1]; out[br][2] = o2 + b1[2];
  }
  float* xr = x + m * ldx;
#pragma unroll
  for (int k = 0; k < 3; ++k) xr[k] = __fadd_rn(__fmul_rn(xr[k], __fadd_rn(out[0][k], 1.f)), out[1][k]);
}

static inline uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace pdf

extern "C" int64_t pdf_image_bytes(int64_t rows, int cols) {
  if (rows < 0 || cols <= 0) return -1;
  return ((rows + 127) / 128) * (int64_t)((cols + 63) / 64) * pdf::G_BLOCK;
}

extern "C" int pdf_pack_image_host(const float* W, int64_t rows, int cols, int64_t ld, int split, void* out_host) {
  PDF_REQUIRE(W && out_host && rows > 0 && cols > 0 && ld >= cols, PDF_ERR_BAD_ARG, "pdf_pack_image_host: bad argument");
  const int nkb = (cols + 63) / 64, kbt = split ? 3 * nkb : nkb;
  uint8_t* out = (uint8_t*)out_host;
  memset(out, 0, (size_t)pdf_image_bytes(rows, cols) * (split ? 3 : 1));
  for (int64_t r = 0; r < rows; ++r)
    for (int k = 0; k < cols; ++k) {
      const float v = W[r * ld + k];
      const uint16_t h = pdf::f2bf_host(v);
      const uint32_t off = pdf::umma::sw128_off((uint32_t)(r & 127), k & 63);
      uint8_t* tile = out + (size_t)(r >> 7) * kbt * pdf::G_BLOCK;
      memcpy(tile + (size_t)(k >> 6) * pdf::G_BLOCK + off, &h, 2);
      if (split) {                                       // [hi | lo | hi], the mirror of the activation split
        uint32_t hb = (uint32_t)h << 16;
        float hf;
        memcpy(&hf, &hb, 4);
        const uint16_t l = pdf::f2bf_host(v - hf);
        memcpy(tile + (size_t)(nkb + (k >> 6)) * pdf::G_BLOCK + off, &l, 2);
        memcpy(tile + (size_t)(2 * nkb + (k >> 6)) * pdf::G_BLOCK + off, &h, 2);
      }
    }
  return PDF_OK;
}

extern "C" int pdf_rows_to_image(const float* X, int64_t ld, int64_t M, int col0, int K, void* img, int kb_total,
                                 int kb0, int split, void* stream) {
  if (M == 0) return PDF_OK;
  PDF_REQUIRE(X && img, PDF_ERR_BAD_ARG, "pdf_rows_to_image: null pointer");
  PDF_REQUIRE(M > 0 && K > 0 && col0 >= 0 && ld >= col0 + K && kb0 >= 0 &&
                  kb0 + (split ? 3 : 1) * ((K + 63) / 64) <= kb_total,
              PDF_ERR_BAD_ARG, "pdf_rows_to_image: bad size");
  const int nkb = (K + 63) / 64;
  const int64_t rows_pad = ((M + 127) / 128) * 128;
  const int64_t chunks = rows_pad * nkb * 8;
  int64_t grid = (chunks + 255) / 256;
  if (grid > 148 * 32) grid = 148 * 32;
  pdf::launch_pdl(pdf::rows_to_image_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, X, ld, M, col0, K,
                  (uint8_t*)img, kb_total, kb0, nkb, chunks, split);
  return pdf::check_launch("pdf_rows_to_image");
}

extern "C" int pdf_gemm_bf16(const void* m_img, int m_tiles, int m_kb, const void* n_img, int n_tiles, int n_kb,
                             int KB, int kb_split, int colmax, const float* bias0, const float* bias1, int act,
                             float* out_f32, int64_t ld_out, int64_t rows_valid, const float* F, int64_t ldf,
                             void* out_img, int out_kb, void* out_bf16, int64_t ld_bf16, int bf16_col_off,
                             const int32_t* tile_desc_host, float* out_max, int64_t ld_max, const float* xyz_w,
                             float* xyz_x, int64_t xyz_ld, void* stream) {
  using namespace pdf;
  if (m_tiles == 0 || n_tiles == 0) return PDF_OK;
  PDF_REQUIRE(m_img && n_img && bias0, PDF_ERR_BAD_ARG, "pdf_gemm_bf16: null pointer");
  PDF_REQUIRE(m_tiles > 0 && n_tiles > 0 && KB > 0 && m_kb > 0 && (KB <= m_kb || KB % m_kb == 0) && KB <= n_kb &&
                  kb_split >= 0 && kb_split < KB,
              PDF_ERR_BAD_ARG, "pdf_gemm_bf16: bad size");
  const int out_split = (act & PDF_GEMM_OUT_SPLIT) ? 1 : 0;     // flag bits on the activation argument
  // the half-footprint configuration only has two operand stages in flight: it pays for short K loops (<= 12 k-blocks, measured),
  // longer ones keep the four-stage pipeline (override: PDF_GEMM_LIGHT_MAX_KB)
  static const int light_max_kb = getenv("PDF_GEMM_LIGHT_MAX_KB") ? atoi(getenv("PDF_GEMM_LIGHT_MAX_KB")) : 12;
  const bool light = (act & PDF_GEMM_LIGHT) && !colmax && kb_split == 0 && KB <= light_max_kb;
  act &= ~(PDF_GEMM_OUT_SPLIT | PDF_GEMM_LIGHT);
  PDF_REQUIRE(act >= 0 && act <= 2, PDF_ERR_BAD_ARG, "pdf_gemm_bf16: bad activation");
  PDF_REQUIRE(!out_split || (out_img && out_kb % 3 == 0 && !colmax), PDF_ERR_BAD_ARG,
              "pdf_gemm_bf16: a split output image needs out_img with 3 x k-blocks");
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.out_split = out_split;
  P.m_img = (const uint8_t*)m_img; P.n_img = (const uint8_t*)n_img;
  P.m_kb = m_kb; P.n_kb = n_kb; P.m_tiles = m_tiles; P.n_tiles = n_tiles; P.KB = KB; P.kb_split = kb_split;
  P.bias0 = bias0; P.bias1 = bias1; P.act = act;
  P.out_f32 = out_f32; P.ld_out = ld_out; P.rows_valid = rows_valid; P.F = F; P.ldf = ldf;
  P.out_img = (uint8_t*)out_img; P.out_kb = out_kb; P.out_max = out_max; P.ld_max = ld_max;
  P.out_bf16 = (uint16_t*)out_bf16; P.ld_bf16 = ld_bf16; P.bf16_col_off = bf16_col_off;
  P.xyz_w = xyz_w; P.xyz_x = xyz_x; P.xyz_ld = xyz_ld;
  P.batches = 1;
  PDF_REQUIRE(!xyz_w || (!colmax && n_tiles == 1 && kb_split == 0 && xyz_x), PDF_ERR_BAD_ARG,
              "pdf_gemm_bf16: XYZ mode needs one N tile, one accumulator and xyz_x");
  if (colmax) {
    PDF_REQUIRE(out_max && kb_split == 0, PDF_ERR_BAD_ARG, "pdf_gemm_bf16: COLMAX needs out_max and one accumulator");
  } else {
    PDF_REQUIRE(n_tiles <= G_MAX_NT && tile_desc_host, PDF_ERR_BAD_ARG, "pdf_gemm_bf16: ROW mode needs <= %d N tiles and tile_desc", G_MAX_NT);
    PDF_REQUIRE(out_f32 || out_img || out_bf16, PDF_ERR_BAD_ARG, "pdf_gemm_bf16: no output");
    PDF_REQUIRE((!out_f32 || ld_out % 4 == 0) && (!F || ldf % 4 == 0), PDF_ERR_BAD_ARG,
                "pdf_gemm_bf16: fp32 row pitches must be multiples of 4 (16-byte vector access)");
    PDF_REQUIRE(!out_bf16 || ((ld_bf16 % 8) == 0 && (bf16_col_off % 4) == 0), PDF_ERR_BAD_ARG,
                "pdf_gemm_bf16: bf16 rows need 16-byte aligned pitch");
    PDF_REQUIRE(kb_split == 0 || (F && bias1), PDF_ERR_BAD_ARG, "pdf_gemm_bf16: SFT mode needs F and bias1");
    for (int i = 0; i < n_tiles; ++i) {
      P.tile_col[i] = tile_desc_host[3 * i]; P.tile_nvalid[i] = tile_desc_host[3 * i + 1]; P.tile_okb[i] = tile_desc_host[3 * i + 2];
      PDF_REQUIRE(P.tile_nvalid[i] >= 0 && P.tile_nvalid[i] <= 128 && (P.tile_col[i] % 4) == 0, PDF_ERR_BAD_ARG,
                  "pdf_gemm_bf16: bad tile descriptor %d", i);
      PDF_REQUIRE(!out_img || P.tile_okb[i] + (P.tile_nvalid[i] + 63) / 64 <= (out_split ? out_kb / 3 : out_kb),
                  PDF_ERR_BAD_ARG, "pdf_gemm_bf16: output image too narrow");
      PDF_REQUIRE(!out_bf16 || ((P.tile_col[i] - bf16_col_off) % 8 == 0 && P.tile_nvalid[i] % 8 == 0), PDF_ERR_BAD_ARG,
                  "pdf_gemm_bf16: bf16 row output needs 8-column aligned tiles");
    }
  }
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(gemm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    cudaFuncSetAttribute(gemm_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    cudaFuncSetAttribute(gemm_bf16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_LIGHT);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = m_tiles * n_tiles;
  if (light) {
    if (grid > 2 * sms) grid = 2 * sms;              // two resident CTAs per SM
    launch_pdl(gemm_bf16_kernel<false, true>, dim3(grid), dim3(G_THREADS), (size_t)G_SMEM_LIGHT, (cudaStream_t)stream, P);
    return check_launch("pdf_gemm_bf16");
  }
  if (grid > sms) grid = sms;
  if (colmax) launch_pdl(gemm_bf16_kernel<true, false>, dim3(grid), dim3(G_THREADS), (size_t)G_SMEM, (cudaStream_t)stream, P);
  else launch_pdl(gemm_bf16_kernel<false, false>, dim3(grid), dim3(G_THREADS), (size_t)G_SMEM, (cudaStream_t)stream, P);
  return check_launch("pdf_gemm_bf16");
}

extern "C" int pdf_sft_xyz_f32(const float* cond, int64_t M, int cc, const float* w0s, const float* b0s,
                               const float* w1s, const float* b1s, const float* w0h, const float* b0h,
                               const float* w1h, const float* b1h, float* x, int64_t ldx, void* stream) {
  if (M == 0) return PDF_OK;
  PDF_REQUIRE(cond && w0s && b0s && w1s && b1s && w0h && b0h && w1h && b1h && x, PDF_ERR_BAD_ARG,
              "pdf_sft_xyz_f32: null pointer");
  PDF_REQUIRE(cc == 64, PDF_ERR_UNSUPPORTED, "pdf_sft_xyz_f32: only 64 condition channels (level 1) are built");
  const size_t smem = (size_t)(2 * cc * cc + 2 * cc + 6 * cc) * sizeof(float);
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(pdf::sft_xyz_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  pdf::sft_xyz_kernel<64><<<(unsigned)((M + 127) / 128), 128, smem, (cudaStream_t)stream>>>(
      cond, M, w0s, b0s, w1s, b1s, w0h, b0h, w1h, b1h, x, ldx);
  return pdf::check_launch("pdf_sft_xyz_f32");
}


extern "C" int pdf_rows_to_image_t(const float* X, int64_t ld, int64_t M, int col0, int C, void* img, int64_t Mc,
                                   int split, void* stream) {
  if (M == 0) return PDF_OK;
  PDF_REQUIRE(X && img && M > 0 && C > 0 && col0 >= 0 && ld >= col0 + C && Mc > 0 && Mc % 64 == 0 && split >= 0 &&
                  split <= 2,
              PDF_ERR_BAD_ARG, "pdf_rows_to_image_t: bad argument");
  const int64_t batches = (M + Mc - 1) / Mc;
  const int64_t gx = batches * (Mc / 64);                 // padded: every k-block of every batch is written
  PDF_REQUIRE(gx < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_rows_to_image_t: too many rows");
  dim3 grid((unsigned)gx, (unsigned)((C + 127) / 128));
  pdf::rows_to_image_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, M, col0, C, (uint8_t*)img, Mc, split);
  return pdf::check_launch("pdf_rows_to_image_t");
}

extern "C" int pdf_gemm_bf16_batched(const void* m_img, int m_tiles, int m_kb, int64_t m_batch_stride, const void* n_img,
                                     int n_tiles, int n_kb, int64_t n_batch_stride, int KB, int batches,
                                     float* out_f32, int64_t ld_out, int64_t out_batch_stride, int64_t rows_valid,
                                     const int32_t* tile_desc_host, void* stream) {
  using namespace pdf;
  if (m_tiles == 0 || n_tiles == 0 || batches == 0) return PDF_OK;
  PDF_REQUIRE(m_img && n_img && out_f32 && tile_desc_host, PDF_ERR_BAD_ARG, "pdf_gemm_bf16_batched: null pointer");
  PDF_REQUIRE(m_tiles > 0 && n_tiles > 0 && n_tiles <= G_MAX_NT && batches > 0 && KB > 0 && KB <= m_kb && KB <= n_kb &&
                  ld_out % 4 == 0 && out_batch_stride % 4 == 0,
              PDF_ERR_BAD_ARG, "pdf_gemm_bf16_batched: bad size");
  PDF_REQUIRE((int64_t)m_tiles * n_tiles * batches < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_gemm_bf16_batched: too much work");
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.m_img = (const uint8_t*)m_img; P.n_img = (const uint8_t*)n_img;
  P.m_kb = m_kb; P.n_kb = n_kb; P.m_tiles = m_tiles; P.n_tiles = n_tiles; P.KB = KB;
  P.out_f32 = out_f32; P.ld_out = ld_out; P.rows_valid = rows_valid;
  P.batches = batches; P.m_batch_stride = m_batch_stride; P.n_batch_stride = n_batch_stride;
  P.out_batch_stride = out_batch_stride;
  for (int i = 0; i < n_tiles; ++i) {
    P.tile_col[i] = tile_desc_host[3 * i]; P.tile_nvalid[i] = tile_desc_host[3 * i + 1]; P.tile_okb[i] = tile_desc_host[3 * i + 2];
    PDF_REQUIRE(P.tile_nvalid[i] >= 0 && P.tile_nvalid[i] <= 128 && (P.tile_col[i] % 4) == 0, PDF_ERR_BAD_ARG,
                "pdf_gemm_bf16_batched: bad tile descriptor %d", i);
  }
  static pdf::PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(gemm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_work = (int64_t)m_tiles * n_tiles * batches;
  const int grid = (int)(n_work < sms ? n_work : sms);
  gemm_bf16_kernel<false><<<grid, G_THREADS, G_SMEM, (cudaStream_t)stream>>>(P);
  return check_launch("pdf_gemm_bf16_batched");
}

// The same ROW-mode GEMM for TWO groups of rows with their own weights and biases in one launch (the left and the
// right hand of the GCN decoder: identical shapes, different parameters): m_img holds 2 * group_m_tiles row tiles,
// n_img / bias the first group's weight image / padded bias with the second group's n_group_stride bytes /
// bias_group_stride floats further.  rows_valid counts inside each group.  Single accumulator, no SFT / xyz modes.
extern "C" int pdf_gemm_bf16_grouped(const void* m_img, int group_m_tiles, int m_kb, const void* n_img, int n_tiles,
                                     int n_kb, int64_t n_group_stride, int KB, const float* bias,
                                     int64_t bias_group_stride, int act, float* out_f32, int64_t ld_out,
                                     int64_t rows_valid, void* out_img, int out_kb, const int32_t* tile_desc_host,
                                     void* stream) {
  using namespace pdf;
  if (group_m_tiles == 0 || n_tiles == 0) return PDF_OK;
  PDF_REQUIRE(m_img && n_img && bias && tile_desc_host && (out_f32 || out_img), PDF_ERR_BAD_ARG,
              "pdf_gemm_bf16_grouped: null pointer");
  PDF_REQUIRE(group_m_tiles > 0 && n_tiles > 0 && n_tiles <= G_MAX_NT && KB > 0 && m_kb > 0 &&
                  (KB <= m_kb || KB % m_kb == 0) && KB <= n_kb && n_group_stride >= 0 && bias_group_stride >= 0 &&
                  rows_valid <= (int64_t)group_m_tiles * 128 && (!out_f32 || ld_out % 4 == 0),
              PDF_ERR_BAD_ARG, "pdf_gemm_bf16_grouped: bad size");
  const int out_split = (act & PDF_GEMM_OUT_SPLIT) ? 1 : 0;
  static const int light_max_kb = getenv("PDF_GEMM_LIGHT_MAX_KB") ? atoi(getenv("PDF_GEMM_LIGHT_MAX_KB")) : 12;
  const bool light = (act & PDF_GEMM_LIGHT) && KB <= light_max_kb;
  act &= ~(PDF_GEMM_OUT_SPLIT | PDF_GEMM_LIGHT);
  PDF_REQUIRE(act >= 0 && act <= 2, PDF_ERR_BAD_ARG, "pdf_gemm_bf16_grouped: bad activation");
  PDF_REQUIRE(!out_split || (out_img && out_kb % 3 == 0), PDF_ERR_BAD_ARG,
              "pdf_gemm_bf16_grouped: a split output image needs out_img with 3 x k-blocks");
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.out_split = out_split;
  P.m_img = (const uint8_t*)m_img; P.n_img = (const uint8_t*)n_img;
  P.m_kb = m_kb; P.n_kb = n_kb; P.m_tiles = 2 * group_m_tiles; P.n_tiles = n_tiles; P.KB = KB;
  P.bias0 = bias; P.act = act;
  P.out_f32 = out_f32; P.ld_out = ld_out; P.rows_valid = rows_valid;
  P.out_img = (uint8_t*)out_img; P.out_kb = out_kb;
  P.batches = 1;
  P.group_m_tiles = group_m_tiles; P.n_group_stride = n_group_stride; P.bias_group_stride = bias_group_stride;
  for (int i = 0; i < n_tiles; ++i) {
    P.tile_col[i] = tile_desc_host[3 * i]; P.tile_nvalid[i] = tile_desc_host[3 * i + 1]; P.tile_okb[i] = tile_desc_host[3 * i + 2];
    PDF_REQUIRE(P.tile_nvalid[i] >= 0 && P.tile_nvalid[i] <= 128 && (P.tile_col[i] % 4) == 0, PDF_ERR_BAD_ARG,
                "pdf_gemm_bf16_grouped: bad tile descriptor %d", i);
    PDF_REQUIRE(!out_img || P.tile_okb[i] + (P.tile_nvalid[i] + 63) / 64 <= (out_split ? out_kb / 3 : out_kb),
                PDF_ERR_BAD_ARG, "pdf_gemm_bf16_grouped: output image too narrow");
  }
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(gemm_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    cudaFuncSetAttribute(gemm_bf16_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_LIGHT);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = P.m_tiles * n_tiles;
  if (light) {
    if (grid > 2 * sms) grid = 2 * sms;
    launch_pdl(gemm_bf16_kernel<false, true>, dim3(grid), dim3(G_THREADS), (size_t)G_SMEM_LIGHT, (cudaStream_t)stream, P);
  } else {
    if (grid > sms) grid = sms;
    launch_pdl(gemm_bf16_kernel<false, false>, dim3(grid), dim3(G_THREADS), (size_t)G_SMEM, (cudaStream_t)stream, P);
  }
  return check_launch("pdf_gemm_bf16_grouped");
}
