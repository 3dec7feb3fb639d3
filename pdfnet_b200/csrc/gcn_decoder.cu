// GCN decoder kernels: the consumer of fuse_feat.
// Reference: lib/models/networks/intaghand_decoder.py:180-242 (decoder.forward),
// model_attn/gcn.py:34-110 (Chebyshev graph conv + GCN_ResBlock), self_attn.py:60-84,
// inter_attn.py:72-125.  The dense layers run on the GEMM kernels (linear_f32.cu / gemm_bf16.cu);
// everything between two GEMMs is ONE of the kernels below, activations are fp32 rows [B*V, C]:
//   row_combine   : t = a (+ b) (+ per-vertex row) with optional x2 vertex up-sampling, writes t and/or
//                   LayerNorm(t) (+ReLU)          -> residual adds, LayerNorms, position embeddings
//   graph_cheby_ln: t = U0 + L.U1 + bias (+ R + bias_r), LayerNorm(t) (+ReLU), L sparse (CSR)
//                   -> the second half of a K = 2 Chebyshev conv, fused with the norm that follows
//   mha           : softmax(q k^T / sqrt(d)) v for one (sample, head) per CTA, V <= 256 tokens
//   decoder_project: orthographic projection + GCN -> MANO vertex order (graph_upsample + GCN_to_vert)
#include "pdf_common.cuh"
#include "umma.cuh"

#include <type_traits>

namespace pdf {

constexpr int DEC_MAX_C = 1024;          // channels per row handled by one warp (32 per lane)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm (in place) of the lane-distributed row t[0..n) (element i of lane l is channel l + 32 i), two-pass
template <int NPL>
__device__ __forceinline__ void layer_norm_row(float (&t)[NPL], int C, int lane, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int relu) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) if (lane + 32 * i < C) s += t[i];
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) if (lane + 32 * i < C) { const float d = t[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      const float y = fmaf((t[i] - mean) * rstd, gamma[c], beta[c]);
      t[i] = relu ? fmaxf(y, 0.f) : y;
    }
  }
}

template <int NPL>
__device__ __forceinline__ void store_row(const float (&t)[NPL], int C, int lane, float* __restrict__ out) {
#pragma unroll
  for (int i = 0; i < NPL; ++i) if (lane + 32 * i < C) out[lane + 32 * i] = t[i];
}

// Row r of a split-bf16 tile image ([hi | hi | lo] k-blocks, the M operand pdf_gemm_bf16 expects): the producing
// kernel writes the next GEMM's operand directly instead of fp32 rows + a pdf_rows_to_image pass.
// C % 64 == 0; stage = this warp's C floats of shared memory.
template <int NPL>
__device__ __forceinline__ void store_split_image(const float (&t)[NPL], int C, int lane, float* stage,
                                                  uint8_t* __restrict__ img, int64_t r) {
#pragma unroll
  for (int i = 0; i < NPL; ++i) if (lane + 32 * i < C) stage[lane + 32 * i] = t[i];
  __syncwarp();
  const int nkb = C >> 6;
  uint8_t* base = img + (size_t)(r >> 7) * (size_t)(3 * nkb) * 16384;
  const uint32_t rr = (uint32_t)(r & 127);
  for (int q = lane; q < (C >> 3); q += 32) {
    const float4 a = *reinterpret_cast<const float4*>(stage + q * 8);
    const float4 b = *reinterpret_cast<const float4*>(stage + q * 8 + 4);
    uint4 w, wl;
    w.x = umma::pack_bf16(a.x, a.y); w.y = umma::pack_bf16(a.z, a.w);
    w.z = umma::pack_bf16(b.x, b.y); w.w = umma::pack_bf16(b.z, b.w);
    wl.x = umma::pack_bf16(a.x - __uint_as_float(w.x << 16), a.y - __uint_as_float(w.x & 0xffff0000u));
    wl.y = umma::pack_bf16(a.z - __uint_as_float(w.y << 16), a.w - __uint_as_float(w.y & 0xffff0000u));
    wl.z = umma::pack_bf16(b.x - __uint_as_float(w.z << 16), b.y - __uint_as_float(w.z & 0xffff0000u));
    wl.w = umma::pack_bf16(b.z - __uint_as_float(w.w << 16), b.w - __uint_as_float(w.w & 0xffff0000u));
    const int kb = q >> 3;
    const uint32_t off = umma::sw128_off(rr, (uint32_t)(q & 7) * 8);
    *reinterpret_cast<uint4*>(base + (size_t)kb * 16384 + off) = w;
    *reinterpret_cast<uint4*>(base + (size_t)(nkb + kb) * 16384 + off) = w;
    *reinterpret_cast<uint4*>(base + (size_t)(2 * nkb + kb) * 16384 + off) = wl;
  }
  __syncwarp();
}

// one warp per output row r' = (sample, v'); source row = sample * (V_out / up) + v' / up
template <int NPL>
__global__ void __launch_bounds__(256)
row_combine_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                   const float* __restrict__ rowvec, int64_t ldr, int V_out, int up, int C, int64_t rows_out,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                   float* __restrict__ sum_out, int64_t lds, float* __restrict__ ln_out, int64_t ldl,
                   uint8_t* __restrict__ sum_img, uint8_t* __restrict__ ln_img) {
  extern __shared__ __align__(16) float stage_all[];
  const int lane = threadIdx.x & 31;
  float* stage = stage_all + (threadIdx.x >> 5) * C;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows_out) return;
  const int64_t smp = r / V_out;
  const int v = (int)(r - smp * V_out);
  const int64_t src = smp * (V_out / up) + v / up;
  float t[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    t[i] = 0.f;
    if (c < C) {
      float x = a[src * lda + c];
      if (b) x += b[src * ldb + c];
      if (rowvec) x += rowvec[(int64_t)v * ldr + c];
      t[i] = x;
    }
  }
  if (sum_out) store_row<NPL>(t, C, lane, sum_out + r * lds);
  if (sum_img) store_split_image<NPL>(t, C, lane, stage, sum_img, r);
  if (ln_out || ln_img) {
    layer_norm_row<NPL>(t, C, lane, gamma, beta, eps, relu);
    if (ln_out) store_row<NPL>(t, C, lane, ln_out + r * ldl);
    if (ln_img) store_split_image<NPL>(t, C, lane, stage, ln_img, r);
  }
}

// t = U0[row] + bias + sum_u L[v,u] U1[sample, u] (+ R[row] + bias_r); out = LayerNorm(t) (+ReLU)
template <int NPL>
__global__ void __launch_bounds__(256)
graph_cheby_ln_kernel(const float* __restrict__ U0, const float* __restrict__ U1, int64_t ldu,
                      const float* __restrict__ bias, const float* __restrict__ R, int64_t ldr,
                      const float* __restrict__ bias_r, const int* __restrict__ rowptr,
                      const int* __restrict__ colidx, const float* __restrict__ vals, int V, int C, int64_t rows,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                      float* __restrict__ out, int64_t ldo, uint8_t* __restrict__ out_img) {
  extern __shared__ __align__(16) float stage_all[];
  const int lane = threadIdx.x & 31;
  float* stage = stage_all + (threadIdx.x >> 5) * C;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t smp = r / V;
  const int v = (int)(r - smp * V);
  float t[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    t[i] = 0.f;
    if (c < C) {
      float x = U0[r * ldu + c] + bias[c];
      if (R) x += R[r * ldr + c] + (bias_r ? bias_r[c] : 0.f);
      t[i] = x;
    }
  }
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int e = e0; e < e1; ++e) {
    const float w = vals[e];
    const float* u = U1 + (smp * V + colidx[e]) * ldu;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) t[i] = fmaf(w, u[c], t[i]);
    }
  }
  layer_norm_row<NPL>(t, C, lane, gamma, beta, eps, relu);
  if (out) store_row<NPL>(t, C, lane, out + r * ldo);
  if (out_img) store_split_image<NPL>(t, C, lane, stage, out_img, r);
}

// ---- vectorised variants for the decoder's own widths (C = 64 / 128 / 256 / 512) -------------------------------
// A row is handled by LPR lanes, each owning VEC float4 (lane l: channels 4 (l + LPR v) .. +3), so every load and
// store is a 16-byte access and a warp covers 32 / LPR rows: 4x fewer memory instructions than one float per lane,
// which is what bounded the generic kernels (they sit 4-5x under the HBM roofline, profiles/r02_*decoder*).
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
  // only the LPR lanes of this row's group take part (another group of the warp may already have left)
  const unsigned m = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x & 31) / LPR * LPR));
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
  return v;
}

template <int LPR, int VEC>
__device__ __forceinline__ void layer_norm_vec(float4 (&t)[VEC], int l, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, int relu) {
  constexpr int C = LPR * 4 * VEC;
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) s += (t[v].x + t[v].y) + (t[v].z + t[v].w);
  const float mean = group_sum<LPR>(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const float a = t[v].x - mean, b = t[v].y - mean, c = t[v].z - mean, d = t[v].w - mean;
    q = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, q))));
  }
  const float rstd = rsqrtf(group_sum<LPR>(q) / (float)C + eps);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = 4 * (l + LPR * v);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = fmaf((t[v].x - mean) * rstd, g.x, b.x); y.y = fmaf((t[v].y - mean) * rstd, g.y, b.y);
    y.z = fmaf((t[v].z - mean) * rstd, g.z, b.z); y.w = fmaf((t[v].w - mean) * rstd, g.w, b.w);
    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    t[v] = y;
  }
}

// the lane's 4 channels as 8 bytes of each part of the split image row ([hi | hi | lo])
template <int LPR, int VEC>
__device__ __forceinline__ void store_split_vec(const float4 (&t)[VEC], int l, uint8_t* __restrict__ img, int64_t r) {
  constexpr int C = LPR * 4 * VEC, NKB = C / 64;
  uint8_t* base = img + (size_t)(r >> 7) * (size_t)(3 * NKB) * 16384;
  const uint32_t rr = (uint32_t)(r & 127);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c = 4 * (l + LPR * v);
    uint2 w, wl;
    w.x = umma::pack_bf16(t[v].x, t[v].y); w.y = umma::pack_bf16(t[v].z, t[v].w);
    wl.x = umma::pack_bf16(t[v].x - __uint_as_float(w.x << 16), t[v].y - __uint_as_float(w.x & 0xffff0000u));
    wl.y = umma::pack_bf16(t[v].z - __uint_as_float(w.y << 16), t[v].w - __uint_as_float(w.y & 0xffff0000u));
    const int kb = c >> 6;
    const uint32_t off = umma::sw128_off(rr, (uint32_t)(c & 63));
    *reinterpret_cast<uint2*>(base + (size_t)kb * 16384 + off) = w;
    *reinterpret_cast<uint2*>(base + (size_t)(NKB + kb) * 16384 + off) = w;
    *reinterpret_cast<uint2*>(base + (size_t)(2 * NKB + kb) * 16384 + off) = wl;
  }
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma(float w, float4 u, float4 t) {
  return make_float4(fmaf(w, u.x, t.x), fmaf(w, u.y, t.y), fmaf(w, u.z, t.z), fmaf(w, u.w, t.w));
}

// Two groups of rows (the left and the right hand) in one launch: group g owns output rows
// [g * rows_per_group, g * rows_per_group + valid) - rows_per_group is a multiple of 128 so both groups start on a tile
// of the operand images - reads its sources from [g * src_per_group, ...) and its per-channel parameters (bias,
// LayerNorm weights) / per-vertex rows `pstride` / `rowvec_gstride` floats after the first group's.  rows_per_group = 0: off.
struct RowGroups {
  int64_t rows_per_group, valid, src_per_group, pstride, rowvec_gstride;
};

template <int LPR, int VEC>
__global__ void __launch_bounds__(256)
row_combine_vec_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                       const float* __restrict__ rowvec, int64_t ldr, int V_out, int up, int64_t rows_out,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                       float* __restrict__ sum_out, int64_t lds, float* __restrict__ ln_out, int64_t ldl,
                       uint8_t* __restrict__ sum_img, uint8_t* __restrict__ ln_img, const RowGroups G) {
  pdl_wait();
  pdl_trigger();
  const int l = threadIdx.x % LPR;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  if (r >= rows_out) return;                                   // whole LPR-lane groups leave together
  int64_t rl = r, src0 = 0;
  if (G.rows_per_group) {
    const int g = r >= G.rows_per_group;
    rl = r - g * G.rows_per_group;
    if (rl >= G.valid) return;                                 // padding rows between / after the groups
    src0 = g * G.src_per_group;
    if (gamma) { gamma += g * G.pstride; beta += g * G.pstride; }
    if (rowvec) rowvec += g * G.rowvec_gstride;
  }
  const int64_t smp = rl / V_out;
  const int v = (int)(rl - smp * V_out);
  const int64_t src = src0 + smp * (V_out / up) + v / up;
  float4 t[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = 4 * (l + LPR * i);
    float4 x = __ldg(reinterpret_cast<const float4*>(a + src * lda + c));
    if (b) x = f4_add(x, __ldg(reinterpret_cast<const float4*>(b + src * ldb + c)));
    if (rowvec) x = f4_add(x, __ldg(reinterpret_cast<const float4*>(rowvec + (int64_t)v * ldr + c)));
    t[i] = x;
  }
  if (sum_out) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) *reinterpret_cast<float4*>(sum_out + r * lds + 4 * (l + LPR * i)) = t[i];
  }
  if (sum_img) store_split_vec<LPR, VEC>(t, l, sum_img, r);
  if (ln_out || ln_img) {
    layer_norm_vec<LPR, VEC>(t, l, gamma, beta, eps, relu);
    if (ln_out) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) *reinterpret_cast<float4*>(ln_out + r * ldl + 4 * (l + LPR * i)) = t[i];
    }
    if (ln_img) store_split_vec<LPR, VEC>(t, l, ln_img, r);
  }
}

template <int LPR, int VEC>
__global__ void __launch_bounds__(256)
graph_cheby_ln_vec_kernel(const float* __restrict__ U0, const float* __restrict__ U1, int64_t ldu,
                          const float* __restrict__ bias, const float* __restrict__ R, int64_t ldr,
                          const float* __restrict__ bias_r, const int* __restrict__ rowptr,
                          const int* __restrict__ colidx, const float* __restrict__ vals, int V, int64_t rows,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                          float* __restrict__ out, int64_t ldo, uint8_t* __restrict__ out_img, const RowGroups G) {
  pdl_wait();
  pdl_trigger();
  const int l = threadIdx.x % LPR;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  if (r >= rows) return;
  int64_t rl = r, base = 0;
  if (G.rows_per_group) {
    const int g = r >= G.rows_per_group;
    base = g * G.rows_per_group;
    rl = r - base;
    if (rl >= G.valid) return;
    bias += g * G.pstride; gamma += g * G.pstride; beta += g * G.pstride;
    if (bias_r) bias_r += g * G.pstride;
  }
  const int64_t smp = rl / V;
  const int v = (int)(rl - smp * V);
  float4 t[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = 4 * (l + LPR * i);
    float4 x = f4_add(__ldg(reinterpret_cast<const float4*>(U0 + r * ldu + c)), __ldg(reinterpret_cast<const float4*>(bias + c)));
    if (R) {
      x = f4_add(x, __ldg(reinterpret_cast<const float4*>(R + r * ldr + c)));
      if (bias_r) x = f4_add(x, __ldg(reinterpret_cast<const float4*>(bias_r + c)));
    }
    t[i] = x;
  }
  const int e0 = rowptr[v], e1 = rowptr[v + 1];
  for (int e = e0; e < e1; ++e) {
    const float w = __ldg(vals + e);
    const float* u = U1 + (base + smp * V + __ldg(colidx + e)) * ldu;
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = f4_fma(w, __ldg(reinterpret_cast<const float4*>(u + 4 * (l + LPR * i))), t[i]);
  }
  layer_norm_vec<LPR, VEC>(t, l, gamma, beta, eps, relu);
  if (out) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) *reinterpret_cast<float4*>(out + r * ldo + 4 * (l + LPR * i)) = t[i];
  }
  if (out_img) store_split_vec<LPR, VEC>(t, l, out_img, r);
}

// C in {64, 128, 256, 512} with 16-byte aligned rows -> (LPR, VEC); returns false for anything else
template <typename F>
static bool dispatch_vec(int C, F&& f) {
  if (C == 64) { f(std::integral_constant<int, 16>(), std::integral_constant<int, 1>()); return true; }
  if (C == 128) { f(std::integral_constant<int, 32>(), std::integral_constant<int, 1>()); return true; }
  if (C == 256) { f(std::integral_constant<int, 32>(), std::integral_constant<int, 2>()); return true; }
  if (C == 512) { f(std::integral_constant<int, 32>(), std::integral_constant<int, 4>()); return true; }
  return false;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// one CTA per (sample, head): K and V of that head in shared memory; each warp takes FOUR query rows at a
// time: scores with lane = key (8 keys per lane, 32 FMAs per 9 shared loads), softmax by warp shuffles,
// probabilities through a per-warp shared tile, then P.V with lane = output channel.
constexpr int MHA_WARPS = 8, MHA_R = 4, MHA_MAXV = 256;

template <int D>
__global__ void __launch_bounds__(MHA_WARPS * 32)
mha_kernel(const float* __restrict__ Q, int64_t ldq, const float* __restrict__ K, int64_t ldk,
           const float* __restrict__ Vv, int64_t ldv, int V, int heads, float inv_norm,
           float* __restrict__ out, int64_t ldo) {
  extern __shared__ __align__(16) float sm[];
  constexpr int PITCH = D + 1;                   // odd pitch: lanes reading different keys hit different banks
  float* sP = sm;                                // [warps][MAXV][4]  (query fastest: one 16 B load = 4 rows)
  float* sQ = sP + MHA_WARPS * MHA_MAXV * MHA_R; // [warps][D][4]
  float* sV = sQ + MHA_WARPS * D * MHA_R;        // [V][D]
  float* sK = sV + (size_t)V * D;                // [V][D+1]
  const int smp = blockIdx.x / heads, h = blockIdx.x - smp * heads;
  const int64_t row0 = (int64_t)smp * V;
  for (int e = threadIdx.x; e < V * D; e += blockDim.x) {
    const int j = e / D, c = e - j * D;
    sK[j * PITCH + c] = K[(row0 + j) * ldk + h * D + c];
    sV[j * D + c] = Vv[(row0 + j) * ldv + h * D + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wP = sP + warp * MHA_MAXV * MHA_R;
  float* wQ = sQ + warp * D * MHA_R;
  for (int i0 = warp * MHA_R; i0 < V; i0 += MHA_WARPS * MHA_R) {
    // stage the four query rows, pre-scaled by 1/sqrt(d)
    for (int e = lane; e < D * MHA_R; e += 32) {
      const int r = e / D, c = e - r * D;
      wQ[c * MHA_R + r] = (i0 + r < V) ? Q[(row0 + i0 + r) * ldq + h * D + c] * inv_norm : 0.f;
    }
    __syncwarp();
    float s[MHA_R][8];
#pragma unroll
    for (int r = 0; r < MHA_R; ++r)
#pragma unroll
      for (int t = 0; t < 8; ++t) s[r][t] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float4 q = *reinterpret_cast<const float4*>(wQ + c * MHA_R);
      float kc[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) { const int j = lane + 32 * t; kc[t] = j < V ? sK[j * PITCH + c] : 0.f; }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        s[0][t] = fmaf(q.x, kc[t], s[0][t]); s[1][t] = fmaf(q.y, kc[t], s[1][t]);
        s[2][t] = fmaf(q.z, kc[t], s[2][t]); s[3][t] = fmaf(q.w, kc[t], s[3][t]);
      }
    }
#pragma unroll
    for (int r = 0; r < MHA_R; ++r) {
      float mx = -3.0e38f;
#pragma unroll
      for (int t = 0; t < 8; ++t) if (lane + 32 * t < V) mx = fmaxf(mx, s[r][t]);
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) { s[r][t] = (lane + 32 * t < V) ? expf(s[r][t] - mx) : 0.f; den += s[r][t]; }
      const float inv = 1.f / warp_sum(den);
#pragma unroll
      for (int t = 0; t < 8; ++t) if (lane + 32 * t < V) wP[(lane + 32 * t) * MHA_R + r] = s[r][t] * inv;
    }
    __syncwarp();
    // out[r][c] = sum_j P[j][r] V[j][c]
    if (D >= 32) {
      float o0[MHA_R] = {0.f, 0.f, 0.f, 0.f}, o1[MHA_R] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int j = 0; j < V; ++j) {
        const float4 pj = *reinterpret_cast<const float4*>(wP + j * MHA_R);
        const float v0 = sV[j * D + lane];
        o0[0] = fmaf(pj.x, v0, o0[0]); o0[1] = fmaf(pj.y, v0, o0[1]); o0[2] = fmaf(pj.z, v0, o0[2]); o0[3] = fmaf(pj.w, v0, o0[3]);
        if (D == 64) {
          const float v1 = sV[j * D + lane + 32];
          o1[0] = fmaf(pj.x, v1, o1[0]); o1[1] = fmaf(pj.y, v1, o1[1]); o1[2] = fmaf(pj.z, v1, o1[2]); o1[3] = fmaf(pj.w, v1, o1[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < MHA_R; ++r) {
        if (i0 + r < V) {
          float* o = out + (row0 + i0 + r) * ldo + h * D;
          o[lane] = o0[r];
          if (D == 64) o[lane + 32] = o1[r];
        }
      }
    } else {                                     // D == 16: two key groups per warp, combined by one shuffle
      const int c = lane & 15, g = lane >> 4;
      float o[MHA_R] = {0.f, 0.f, 0.f, 0.f};
      for (int j = g; j < V; j += 2) {
        const float4 pj = *reinterpret_cast<const float4*>(wP + j * MHA_R);
        const float v0 = sV[j * D + c];
        o[0] = fmaf(pj.x, v0, o[0]); o[1] = fmaf(pj.y, v0, o[1]); o[2] = fmaf(pj.z, v0, o[2]); o[3] = fmaf(pj.w, v0, o[3]);
      }
#pragma unroll
      for (int r = 0; r < MHA_R; ++r) {
        o[r] += __shfl_xor_sync(0xffffffffu, o[r], 16);
        if (g == 0 && i0 + r < V) out[(row0 + i0 + r) * ldo + h * D + c] = o[r];
      }
    }
    __syncwarp();
  }
}

// verts2d = scale*img * v[..., :2] + (trans2d*img/2 + img/2) for the coarse (Vc) and dense (Vd) meshes, and the
// MANO-order lists: mano[b, i] = coarse[b, rev[i] / rep]  (graph_upsample by rep = Vall / Vc, then GCN_to_vert)
__global__ void decoder_project_kernel(const float* __restrict__ vc, int Vc, const float* __restrict__ vd, int Vd,
                                       const float* __restrict__ params, int64_t ldp, float img,
                                       const int64_t* __restrict__ rev, int rep, int64_t B,
                                       float* __restrict__ vc2d, float* __restrict__ vd2d,
                                       float* __restrict__ mano3d, float* __restrict__ mano2d) {
  const int64_t total = B * (Vc + 2 * Vd);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / (Vc + 2 * Vd);
    const int k = (int)(e - b * (Vc + 2 * Vd));
    const float s = params[b * ldp] * img;
    const float tx = params[b * ldp + 1] * img / 2 + img / 2, ty = params[b * ldp + 2] * img / 2 + img / 2;
    if (k < Vc) {
      const float* p = vc + (b * Vc + k) * 3;
      vc2d[(b * Vc + k) * 2] = s * p[0] + tx;
      vc2d[(b * Vc + k) * 2 + 1] = s * p[1] + ty;
    } else if (k < Vc + Vd) {
      const int i = k - Vc;
      const float* p = vd + (b * Vd + i) * 3;
      vd2d[(b * Vd + i) * 2] = s * p[0] + tx;
      vd2d[(b * Vd + i) * 2 + 1] = s * p[1] + ty;
    } else {
      const int i = k - Vc - Vd;
      const float* p = vc + (b * Vc + (int)(rev[i] / rep)) * 3;
      float* m3 = mano3d + (b * Vd + i) * 3;
      m3[0] = p[0]; m3[1] = p[1]; m3[2] = p[2];
      mano2d[(b * Vd + i) * 2] = s * p[0] + tx;
      mano2d[(b * Vd + i) * 2 + 1] = s * p[1] + ty;
    }
  }
}

// Output heads of decoder.forward (intaghand_decoder.py:213-224) for one hand-sample per CTA: the [V, C] feature
// tile is staged once in shared memory (odd pitch) and feeds
//   temp[c]   = avg_head(f^T)        = sum_v aw[v] f[v,c] + ab          (Linear over the VERTEX axis)
//   params[o] = params_head(temp), root[o] = root_head(temp)            (o < 3)
//   v[v,o]    = coord_head(f[v,:])                                       (the coarse mesh, [V,3])
// instead of two transposed copies and four skinny FFMA GEMM launches per hand.
__global__ void __launch_bounds__(256)
decoder_heads_kernel(const float* __restrict__ f, int64_t ldf, int V, int C, const float* __restrict__ aw,
                     const float* __restrict__ ab, const float* __restrict__ Wp, const float* __restrict__ bp,
                     const float* __restrict__ Wr, const float* __restrict__ br, const float* __restrict__ Wc,
                     const float* __restrict__ bc, float* __restrict__ params, float* __restrict__ root,
                     float* __restrict__ verts) {
  extern __shared__ __align__(16) float hs[];
  pdl_wait();
  pdl_trigger();
  const int pitch = C + 1;
  float* sf = hs;                            // [V][C+1]
  float* st = sf + (size_t)V * pitch;        // temp [C]
  float* sw = st + C;                        // Wc [3][C], Wp [3][C], Wr [3][C]
  const int64_t n = blockIdx.x;
  const float* src = f + n * V * ldf;
  for (int e = threadIdx.x; e < V * C; e += blockDim.x) {
    const int v = e / C, c = e - v * C;
    sf[v * pitch + c] = src[(int64_t)v * ldf + c];
  }
  for (int e = threadIdx.x; e < 3 * C; e += blockDim.x) { sw[e] = Wc[e]; sw[3 * C + e] = Wp[e]; sw[6 * C + e] = Wr[e]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int v = 0; v < V; ++v) a = fmaf(__ldg(aw + v), sf[v * pitch + c], a);
    st[c] = a + ab[0];
  }
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = sf[v * pitch + c];
      a0 = fmaf(sw[c], x, a0); a1 = fmaf(sw[C + c], x, a1); a2 = fmaf(sw[2 * C + c], x, a2);
    }
    float* o = verts + (n * V + v) * 3;
    o[0] = a0 + bc[0]; o[1] = a1 + bc[1]; o[2] = a2 + bc[2];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int o = threadIdx.x % 3, which = threadIdx.x / 3;
    const float* w = sw + (which ? 6 : 3) * C + o * C;
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(w[c], st[c], a);
    if (which) root[n * 3 + o] = a + br[o];
    else params[n * 3 + o] = a + bp[o];
  }
}

template <typename F>
static int dispatch_npl(int C, F&& f) {
  if (C <= 128) return f(std::integral_constant<int, 4>());
  if (C <= 256) return f(std::integral_constant<int, 8>());
  if (C <= 512) return f(std::integral_constant<int, 16>());
  return f(std::integral_constant<int, 32>());
}

}  // namespace pdf

using namespace pdf;

extern "C" int pdf_row_combine(const float* a, int64_t lda, const float* b, int64_t ldb, const float* rowvec,
                               int64_t ldr, int V_out, int up, int C, int64_t rows_out, const float* gamma,
                               const float* beta, float eps, int relu, float* sum_out, int64_t lds, float* ln_out,
                               int64_t ldl, void* sum_img, void* ln_img, void* stream) {
  if (rows_out == 0) return PDF_OK;
  PDF_REQUIRE(a && (sum_out || ln_out || sum_img || ln_img), PDF_ERR_BAD_ARG, "pdf_row_combine: null pointer");
  PDF_REQUIRE(rows_out > 0 && C > 0 && C <= DEC_MAX_C && V_out > 0 && up >= 1 && V_out % up == 0 &&
                  rows_out % V_out == 0 && (!(ln_out || ln_img) || (gamma && beta)) &&
                  (!(sum_img || ln_img) || C % 64 == 0),
              PDF_ERR_BAD_ARG, "pdf_row_combine: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec_ok = lda % 4 == 0 && (!b || ldb % 4 == 0) && (!rowvec || ldr % 4 == 0) && (!sum_out || lds % 4 == 0) &&
                      (!ln_out || ldl % 4 == 0) && aligned16(a) && aligned16(b) && aligned16(rowvec) &&
                      aligned16(sum_out) && aligned16(ln_out) && aligned16(gamma) && aligned16(beta);
  if (vec_ok && dispatch_vec(C, [&](auto lpr, auto vec) {
        constexpr int LPR = decltype(lpr)::value, VEC = decltype(vec)::value;
        const unsigned g = (unsigned)((rows_out * LPR + 255) / 256);
        launch_pdl(row_combine_vec_kernel<LPR, VEC>, dim3(g), dim3(256), 0, s, a, lda, b, ldb, rowvec, ldr, V_out, up, rows_out,
                   gamma, beta, eps, relu, sum_out, lds, ln_out, ldl, (uint8_t*)sum_img, (uint8_t*)ln_img, RowGroups{0, 0, 0, 0, 0});
      }))
    return check_launch("pdf_row_combine");
  const unsigned grid = (unsigned)((rows_out * 32 + 255) / 256);
  const size_t smem = (sum_img || ln_img) ? (size_t)8 * C * sizeof(float) : 0;
  dispatch_npl(C, [&](auto npl) {
    row_combine_kernel<decltype(npl)::value><<<grid, 256, smem, s>>>(a, lda, b, ldb, rowvec, ldr, V_out, up, C, rows_out,
                                                                      gamma, beta, eps, relu, sum_out, lds, ln_out, ldl,
                                                                      (uint8_t*)sum_img, (uint8_t*)ln_img);
    return 0;
  });
  return check_launch("pdf_row_combine");
}

extern "C" int pdf_graph_cheby_ln(const float* U0, const float* U1, int64_t ldu, const float* bias, const float* R,
                                  int64_t ldr, const float* bias_r, const int32_t* rowptr, const int32_t* colidx,
                                  const float* vals, int V, int C, int64_t rows, const float* gamma,
                                  const float* beta, float eps, int relu, float* out, int64_t ldo, void* out_img,
                                  void* stream) {
  if (rows == 0) return PDF_OK;
  PDF_REQUIRE(U0 && U1 && bias && rowptr && colidx && vals && gamma && beta && (out || out_img), PDF_ERR_BAD_ARG,
              "pdf_graph_cheby_ln: null pointer");
  PDF_REQUIRE(rows > 0 && V > 0 && rows % V == 0 && C > 0 && C <= DEC_MAX_C && (!out_img || C % 64 == 0),
              PDF_ERR_BAD_ARG, "pdf_graph_cheby_ln: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec_ok = ldu % 4 == 0 && (!R || ldr % 4 == 0) && (!out || ldo % 4 == 0) && aligned16(U0) && aligned16(U1) &&
                      aligned16(bias) && aligned16(R) && aligned16(bias_r) && aligned16(out) && aligned16(gamma) &&
                      aligned16(beta);
  if (vec_ok && dispatch_vec(C, [&](auto lpr, auto vec) {
        constexpr int LPR = decltype(lpr)::value, VEC = decltype(vec)::value;
        const unsigned g = (unsigned)((rows * LPR + 255) / 256);
        launch_pdl(graph_cheby_ln_vec_kernel<LPR, VEC>, dim3(g), dim3(256), 0, s, U0, U1, ldu, bias, R, ldr, bias_r, rowptr,
                   colidx, vals, V, rows, gamma, beta, eps, relu, out, ldo, (uint8_t*)out_img, RowGroups{0, 0, 0, 0, 0});
      }))
    return check_launch("pdf_graph_cheby_ln");
  const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
  const size_t smem = out_img ? (size_t)8 * C * sizeof(float) : 0;
  dispatch_npl(C, [&](auto npl) {
    graph_cheby_ln_kernel<decltype(npl)::value><<<grid, 256, smem, s>>>(U0, U1, ldu, bias, R, ldr, bias_r, rowptr,
                                                                         colidx, vals, V, C, rows, gamma, beta, eps,
                                                                         relu, out, ldo, (uint8_t*)out_img);
    return 0;
  });
  return check_launch("pdf_graph_cheby_ln");
}

// Grouped forms (two hands in one launch, see RowGroups): buffers hold 2 * rows_per_group rows (rows_per_group % 128
// == 0, the first `valid` of each group are real); per-channel parameters are stacked [2, C] (pstride = C), the
// per-vertex rows [2, V_out, ldr] when rowvec_gstride != 0.  Vectorised widths only (C in {64,128,256,512}).
extern "C" int pdf_row_combine_grouped(const float* a, int64_t lda, const float* b, int64_t ldb, const float* rowvec,
                                       int64_t ldr, int64_t rowvec_gstride, int V_out, int up, int C,
                                       int64_t rows_per_group, int64_t valid, int64_t src_per_group, const float* gamma,
                                       const float* beta, float eps, int relu, float* sum_out, int64_t lds,
                                       float* ln_out, int64_t ldl, void* sum_img, void* ln_img, void* stream) {
  if (valid == 0) return PDF_OK;
  PDF_REQUIRE(a && (sum_out || ln_out || sum_img || ln_img), PDF_ERR_BAD_ARG, "pdf_row_combine_grouped: null pointer");
  PDF_REQUIRE(rows_per_group > 0 && rows_per_group % 128 == 0 && valid > 0 && valid <= rows_per_group && V_out > 0 &&
                  up >= 1 && V_out % up == 0 && valid % V_out == 0 && src_per_group >= valid / up &&
                  (!(ln_out || ln_img) || (gamma && beta)) && (!(sum_img || ln_img) || C % 64 == 0),
              PDF_ERR_BAD_ARG, "pdf_row_combine_grouped: bad argument");
  const bool vec_ok = lda % 4 == 0 && (!b || ldb % 4 == 0) && (!rowvec || (ldr % 4 == 0 && rowvec_gstride % 4 == 0)) &&
                      (!sum_out || lds % 4 == 0) && (!ln_out || ldl % 4 == 0) && aligned16(a) && aligned16(b) &&
                      aligned16(rowvec) && aligned16(sum_out) && aligned16(ln_out) && aligned16(gamma) && aligned16(beta);
  const int64_t rows_out = 2 * rows_per_group;
  const RowGroups G{rows_per_group, valid, src_per_group, (int64_t)C, rowvec_gstride};
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok && dispatch_vec(C, [&](auto lpr, auto vec) {
        constexpr int LPR = decltype(lpr)::value, VEC = decltype(vec)::value;
        const unsigned g = (unsigned)((rows_out * LPR + 255) / 256);
        launch_pdl(row_combine_vec_kernel<LPR, VEC>, dim3(g), dim3(256), 0, s, a, lda, b, ldb, rowvec, ldr, V_out, up, rows_out,
                   gamma, beta, eps, relu, sum_out, lds, ln_out, ldl, (uint8_t*)sum_img, (uint8_t*)ln_img, G);
      }))
    return check_launch("pdf_row_combine_grouped");
  set_error("pdf_row_combine_grouped: needs C in {64,128,256,512} and 16-byte aligned rows (got C=%d)", C);
  return PDF_ERR_UNSUPPORTED;
}

extern "C" int pdf_graph_cheby_ln_grouped(const float* U0, const float* U1, int64_t ldu, const float* bias, const float* R,
                                          int64_t ldr, const float* bias_r, const int32_t* rowptr, const int32_t* colidx,
                                          const float* vals, int V, int C, int64_t rows_per_group, int64_t valid,
                                          const float* gamma, const float* beta, float eps, int relu, float* out,
                                          int64_t ldo, void* out_img, void* stream) {
  if (valid == 0) return PDF_OK;
  PDF_REQUIRE(U0 && U1 && bias && rowptr && colidx && vals && gamma && beta && (out || out_img), PDF_ERR_BAD_ARG,
              "pdf_graph_cheby_ln_grouped: null pointer");
  PDF_REQUIRE(rows_per_group > 0 && rows_per_group % 128 == 0 && valid > 0 && valid <= rows_per_group && V > 0 &&
                  valid % V == 0 && (!out_img || C % 64 == 0),
              PDF_ERR_BAD_ARG, "pdf_graph_cheby_ln_grouped: bad argument");
  const bool vec_ok = ldu % 4 == 0 && (!R || ldr % 4 == 0) && (!out || ldo % 4 == 0) && aligned16(U0) && aligned16(U1) &&
                      aligned16(bias) && aligned16(R) && aligned16(bias_r) && aligned16(out) && aligned16(gamma) &&
                      aligned16(beta);
  const int64_t rows = 2 * rows_per_group;
  const RowGroups G{rows_per_group, valid, 0, (int64_t)C, 0};
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok && dispatch_vec(C, [&](auto lpr, auto vec) {
        constexpr int LPR = decltype(lpr)::value, VEC = decltype(vec)::value;
        const unsigned g = (unsigned)((rows * LPR + 255) / 256);
        launch_pdl(graph_cheby_ln_vec_kernel<LPR, VEC>, dim3(g), dim3(256), 0, s, U0, U1, ldu, bias, R, ldr, bias_r, rowptr,
                   colidx, vals, V, rows, gamma, beta, eps, relu, out, ldo, (uint8_t*)out_img, G);
      }))
    return check_launch("pdf_graph_cheby_ln_grouped");
  set_error("pdf_graph_cheby_ln_grouped: needs C in {64,128,256,512} and 16-byte aligned rows (got C=%d)", C);
  return PDF_ERR_UNSUPPORTED;
}

extern "C" int pdf_mha(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* Vv, int64_t ldv,
                       int64_t n_samples, int V, int heads, int d, float* out, int64_t ldo, void* stream) {
  if (n_samples == 0) return PDF_OK;
  PDF_REQUIRE(Q && K && Vv && out, PDF_ERR_BAD_ARG, "pdf_mha: null pointer");
  PDF_REQUIRE(n_samples > 0 && V > 0 && V <= MHA_MAXV && heads > 0 && (d == 16 || d == 32 || d == 64) &&
                  n_samples * heads < (1ll << 31),
              PDF_ERR_UNSUPPORTED, "pdf_mha: supports <= 256 tokens and head dim 16 / 32 / 64");
  const size_t smem = sizeof(float) * ((size_t)MHA_WARPS * MHA_MAXV * MHA_R + (size_t)MHA_WARPS * d * MHA_R +
                                       (size_t)V * d + (size_t)V * (d + 1));
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    const int mx = (int)(sizeof(float) * (MHA_WARPS * MHA_MAXV * MHA_R + MHA_WARPS * 64 * MHA_R + 256 * 64 + 256 * 65));
    cudaFuncSetAttribute(mha_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(mha_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(mha_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  }
  const unsigned grid = (unsigned)(n_samples * heads);
  const float inv_norm = 1.f / sqrtf((float)d);
  cudaStream_t s = (cudaStream_t)stream;
  if (d == 16) mha_kernel<16><<<grid, MHA_WARPS * 32, smem, s>>>(Q, ldq, K, ldk, Vv, ldv, V, heads, inv_norm, out, ldo);
  else if (d == 32) mha_kernel<32><<<grid, MHA_WARPS * 32, smem, s>>>(Q, ldq, K, ldk, Vv, ldv, V, heads, inv_norm, out, ldo);
  else mha_kernel<64><<<grid, MHA_WARPS * 32, smem, s>>>(Q, ldq, K, ldk, Vv, ldv, V, heads, inv_norm, out, ldo);
  return check_launch("pdf_mha");
}

extern "C" int pdf_decoder_project(const float* v_coarse, int Vc, const float* v_dense, int Vd, const float* params,
                                   int64_t ldp, float img_size, const int64_t* rev, int rep, int64_t B,
                                   float* coarse2d, float* dense2d, float* mano3d, float* mano2d, void* stream) {
  if (B == 0) return PDF_OK;
  PDF_REQUIRE(v_coarse && v_dense && params && rev && coarse2d && dense2d && mano3d && mano2d, PDF_ERR_BAD_ARG,
              "pdf_decoder_project: null pointer");
  PDF_REQUIRE(B > 0 && Vc > 0 && Vd > 0 && rep > 0 && ldp >= 3, PDF_ERR_BAD_ARG, "pdf_decoder_project: bad argument");
  const int64_t total = B * (Vc + 2 * Vd);
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  decoder_project_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(v_coarse, Vc, v_dense, Vd, params, ldp,
                                                                            img_size, rev, rep, B, coarse2d, dense2d,
                                                                            mano3d, mano2d);
  return check_launch("pdf_decoder_project");
}

extern "C" int pdf_decoder_heads(const float* f, int64_t ldf, int64_t n, int V, int C, const float* avg_w,
                                 const float* avg_b, const float* params_w, const float* params_b,
                                 const float* root_w, const float* root_b, const float* coord_w,
                                 const float* coord_b, float* params, float* root, float* verts, void* stream) {
  if (n == 0) return PDF_OK;
  PDF_REQUIRE(f && avg_w && avg_b && params_w && params_b && root_w && root_b && coord_w && coord_b && params && root &&
                  verts, PDF_ERR_BAD_ARG, "pdf_decoder_heads: null pointer");
  const size_t smem = sizeof(float) * ((size_t)V * (C + 1) + C + 9 * (size_t)C);
  PDF_REQUIRE(n > 0 && V > 0 && C > 0 && ldf >= C && smem <= 200 * 1024, PDF_ERR_UNSUPPORTED,
              "pdf_decoder_heads: bad size (the [V, C] tile must fit shared memory)");
  static pdf::PerDeviceOnce once;
  if (once.first()) cudaFuncSetAttribute(decoder_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  launch_pdl(decoder_heads_kernel, dim3((unsigned)n), dim3(256), smem, (cudaStream_t)stream, f, ldf, V, C, avg_w, avg_b,
             params_w, params_b, root_w, root_b, coord_w, coord_b, params, root, verts);
  return check_launch("pdf_decoder_heads");
}
