// Bandwidth-bound kernels: NCHW row gathers, the 3-level pyramid gather with the
// level-0 SFT fused, the grouping gather, back-projection and Split_coeff.
#include "pdf_common.cuh"

namespace pdf {

// A gather index outside the map never touches memory outside it (the reference's torch.gather raises a
// device assert instead; PDF_CHECK_INDICES=1 on the Python side reproduces that).
__device__ __forceinline__ int64_t clamp_index(int64_t i, int64_t n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// out[b,i,c] = feat[b/cpf, c, ind[b,i]] ; thread per output element, c fastest so
// stores are coalesced; loads hit one 32 B sector each (NCHW hand-off).
__global__ void gather_nchw_kernel(const float* __restrict__ feat, int clouds_per_frame, int C, int64_t HW,
                                   const int64_t* __restrict__ ind, int n, int64_t ind_stride,
                                   float* __restrict__ out, int64_t total) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t bi = e / C;
    const int i = (int)(bi % n);
    const int64_t b = bi / n;
    const int64_t pix = clamp_index(ind[b * ind_stride + i], HW);
    out[e] = __ldg(feat + ((b / clouds_per_frame) * C + c) * HW + pix);
  }
}

__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }

// sft0: 3 -> 3 -> 3 scale and shift branches (intaghand_encoder.py:205-219), fp32.
__device__ __forceinline__ void sft0_apply(const float* __restrict__ P, const float e[3], float xyz[3]) {
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
      h[o] = leaky01(fmaf(W0[o * 3 + 2], e[2], fmaf(W0[o * 3 + 1], e[1], fmaf(W0[o * 3], e[0], b0[o]))));
#pragma unroll
    for (int o = 0; o < 3; ++o)
      out[br][o] = fmaf(W1[o * 3 + 2], h[2], fmaf(W1[o * 3 + 1], h[1], fmaf(W1[o * 3], h[0], b1[o])));
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) xyz[o] = __fadd_rn(__fmul_rn(xyz[o], __fadd_rn(out[0][o], 1.f)), out[1][o]);
}

// grid = (cloud, 1 + 2*PG_SPLIT): blockIdx.y == 0 handles level 0 (3 channels + SFT0 for every
// point); the next PG_SPLIT blocks gather level 1, the last PG_SPLIT level 2.  Every thread keeps
// PG_UNROLL independent scattered loads in flight (each one is a separate 32 B sector of an NCHW
// plane), stores are coalesced ([cloud, point, channel], channel fastest).
constexpr int PG_SPLIT = 4, PG_UNROLL = 8;

__device__ __forceinline__ void gather_level(const float* __restrict__ plane0, int C, int64_t HW, int R, int Rl,
                                             int div, const int64_t* __restrict__ ch, int n,
                                             float* __restrict__ out, int part) {
  const int total = n * C;
  const int per = (total + PG_SPLIT - 1) / PG_SPLIT;
  const int lo = part * per, hi = min(total, lo + per);
  for (int e0 = lo + threadIdx.x; e0 < hi; e0 += blockDim.x * PG_UNROLL) {
    float v[PG_UNROLL];
#pragma unroll
    for (int u = 0; u < PG_UNROLL; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < hi) {
        const int i = e / C, c = e - i * C;
        const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
        const int pl = (pix / R / div) * Rl + (pix % R) / div;       // intaghand_encoder.py:125-126
        v[u] = __ldg(plane0 + (int64_t)c * HW + pl);
      }
    }
#pragma unroll
    for (int u = 0; u < PG_UNROLL; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < hi) out[e] = v[u];
    }
  }
}

// Window variant (levels 1 and 2).  HBM serves 64 B per access, so a point-wise NCHW gather moves
// 64 B per 4 B element.  The pixels of one hand are spatially compact, so a CTA instead loads the
// BOUNDING WINDOW of the cloud's pixels for CG channel planes with coalesced 16 B loads into shared
// memory and gathers from there; stores are CG contiguous floats per point.  Falls back to the
// point-wise gather when the window does not fit the shared-memory budget.
constexpr int PG_WIN_BYTES = 48 * 1024;

template <int CG>
__device__ __forceinline__ void gather_level_window(const float* __restrict__ plane0, int C, int64_t HW, int R,
                                                    int Rl, int div, const int64_t* __restrict__ ch, int n,
                                                    float* __restrict__ out, int c0, float* win, int* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int x0 = 1 << 30, x1 = -1, y0 = 1 << 30, y1 = -1;
  for (int i = tid; i < n; i += blockDim.x) {
    const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
    const int px = (pix % R) / div, py = pix / R / div;
    x0 = min(x0, px); x1 = max(x1, px); y0 = min(y0, py); y1 = max(y1, py);
  }
  x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
  y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
  if (lane == 0) { s_red[warp * 4] = x0; s_red[warp * 4 + 1] = x1; s_red[warp * 4 + 2] = y0; s_red[warp * 4 + 3] = y1; }
  __syncthreads();
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    x0 = min(x0, s_red[w * 4]); x1 = max(x1, s_red[w * 4 + 1]);
    y0 = min(y0, s_red[w * 4 + 2]); y1 = max(y1, s_red[w * 4 + 3]);
  }
  const int x0a = x0 & ~15;                                    // 64 B aligned start
  const int Wa = min(((x1 - x0a + 1) + 15) & ~15, ((Rl - x0a) + 3) & ~3);
  const int H = y1 - y0 + 1;
  const bool fits = (Rl % 4 == 0) && (x0a + Wa <= Rl) && ((int64_t)CG * H * Wa * 4 <= PG_WIN_BYTES);
  if (fits) {
    const int wq = Wa >> 2, per_plane = H * wq;                // float4 units
    constexpr int U = 8;                                       // independent 16 B loads in flight per thread
    const int total = CG * per_plane;
    for (int e0 = tid; e0 < total; e0 += blockDim.x * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < total) {
          const int c = e / per_plane, r = (e - c * per_plane) / wq, q = e - c * per_plane - r * wq;
          v[u] = __ldg(reinterpret_cast<const float4*>(plane0 + (int64_t)(c0 + c) * HW + (int64_t)(y0 + r) * Rl + x0a) + q);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < total) reinterpret_cast<float4*>(win)[e] = v[u];
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
      const int off = (pix / R / div - y0) * Wa + ((pix % R) / div - x0a);
      float v[CG];
#pragma unroll
      for (int c = 0; c < CG; ++c) v[c] = win[c * H * Wa + off];
      float4* o = reinterpret_cast<float4*>(out + (int64_t)i * C + c0);
#pragma unroll
      for (int c = 0; c < CG; c += 4) o[c >> 2] = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    }
  } else {
    for (int e = tid; e < n * CG; e += blockDim.x) {
      const int i = e / CG, c = e - i * CG;
      const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
      out[(int64_t)i * C + c0 + c] = __ldg(plane0 + (int64_t)(c0 + c) * HW + (pix / R / div) * Rl + (pix % R) / div);
    }
  }
}

constexpr int PG_CG1 = 4, PG_CG2 = 16;                         // channels per CTA at level 1 / level 2

__global__ void __launch_bounds__(256)
pyramid_gather_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ choose, int clouds_per_frame,
                      int n_points, int n1, int n2, int R,
                      const float* __restrict__ l0, const float* __restrict__ l1, int C1,
                      const float* __restrict__ l2, int C2, const float* __restrict__ sft0,
                      float* __restrict__ pts0, float* __restrict__ cond1, float* __restrict__ cond2,
                      int groups1, int groups2, int windowed) {
  extern __shared__ __align__(16) float win[];
  __shared__ int s_red[32];
  __shared__ float P[48];
  const int64_t b = blockIdx.x;
  const int64_t f = b / clouds_per_frame;
  const int64_t* ch = choose + b * n_points;
  const int R2 = R / 2, R4 = R / 4;
  const int64_t RR = (int64_t)R * R, HW2 = (int64_t)R2 * R2, HW4 = (int64_t)R4 * R4;
  const int y = blockIdx.y;
  if (y == 0) {
    if (threadIdx.x < 48) P[threadIdx.x] = sft0[threadIdx.x];
    __syncthreads();
    const float* base = l0 + f * 3 * RR;
    for (int i = threadIdx.x; i < n_points; i += blockDim.x) {
      const int64_t pix = clamp_index(ch[i], RR);
      float e[3], p[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e[c] = __ldg(base + c * RR + pix);
        p[c] = xyz[(b * n_points + i) * 3 + c];
      }
      sft0_apply(P, e, p);
#pragma unroll
      for (int c = 0; c < 3; ++c) pts0[(b * n_points + i) * 3 + c] = p[c];
    }
  } else if (y <= groups1) {
    if (windowed)
      gather_level_window<PG_CG1>(l1 + f * C1 * HW2, C1, HW2, R, R2, 2, ch, n1, cond1 + b * n1 * C1, (y - 1) * PG_CG1, win, s_red);
    else gather_level(l1 + f * C1 * HW2, C1, HW2, R, R2, 2, ch, n1, cond1 + b * n1 * C1, y - 1);
  } else {
    const int g = y - 1 - groups1;
    if (windowed)
      gather_level_window<PG_CG2>(l2 + f * C2 * HW4, C2, HW4, R, R4, 4, ch, n2, cond2 + b * n2 * C2, g * PG_CG2, win, s_red);
    else gather_level(l2 + f * C2 * HW4, C2, HW4, R, R4, 4, ch, n2, cond2 + b * n2 * C2, g);
  }
}

// ---- channels-last (NHWC) variants: the RGB neck hands over [F, H, W, C] maps, so a
// point's C features are one contiguous run and the gather reads exactly the algorithmic bytes. ----
// rows [lo, hi) x C of one level; thread = one 16 B chunk, 8 independent loads in flight.
__device__ __forceinline__ void gather_rows_nhwc(const float* __restrict__ src, int C, int R, int Rl, int div,
                                                 const int64_t* __restrict__ ch, float* __restrict__ out,
                                                 int lo, int hi) {
  if ((C & 3) == 0) {
    const int cq = C >> 2;
    constexpr int U = 8;
    const int e_lo = lo * cq, e_hi = hi * cq;
    for (int e0 = e_lo + threadIdx.x; e0 < e_hi; e0 += blockDim.x * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < e_hi) {
          const int i = e / cq, q = e - i * cq;
          const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
          const int64_t p = (int64_t)(pix / R / div) * Rl + (pix % R) / div;
          v[u] = __ldg(reinterpret_cast<const float4*>(src + p * C) + q);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < e_hi) reinterpret_cast<float4*>(out)[e] = v[u];
      }
    }
  } else {
    for (int e = lo * C + threadIdx.x; e < hi * C; e += blockDim.x) {
      const int i = e / C, c = e - i * C;
      const int pix = (int)clamp_index(ch[i], (int64_t)R * R);
      const int64_t p = (int64_t)(pix / R / div) * Rl + (pix % R) / div;
      out[e] = __ldg(src + p * C + c);
    }
  }
}

constexpr int PG_NHWC_PARTS = 4;                               // CTAs per cloud and level

__global__ void __launch_bounds__(256)
pyramid_gather_nhwc_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ choose, int clouds_per_frame,
                           int n_points, int n1, int n2, int R, const float* __restrict__ l0,
                           const float* __restrict__ l1, int C1, const float* __restrict__ l2, int C2,
                           const float* __restrict__ sft0, float* __restrict__ pts0, float* __restrict__ cond1,
                           float* __restrict__ cond2) {
  __shared__ float P[48];
  const int64_t b = blockIdx.x;
  const int64_t f = b / clouds_per_frame;
  const int64_t* ch = choose + b * n_points;
  const int R2 = R / 2, R4 = R / 4;
  const int y = blockIdx.y;
  if (y == 0) {
    if (threadIdx.x < 48) P[threadIdx.x] = sft0[threadIdx.x];
    __syncthreads();
    const float* base = l0 + f * 3 * (int64_t)R * R;
    for (int i = threadIdx.x; i < n_points; i += blockDim.x) {
      const int64_t pix = clamp_index(ch[i], (int64_t)R * R);
      float e[3], p[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        e[c] = __ldg(base + pix * 3 + c);
        p[c] = xyz[(b * n_points + i) * 3 + c];
      }
      sft0_apply(P, e, p);
#pragma unroll
      for (int c = 0; c < 3; ++c) pts0[(b * n_points + i) * 3 + c] = p[c];
    }
  } else if (y <= PG_NHWC_PARTS) {
    const int part = y - 1, per = (n1 + PG_NHWC_PARTS - 1) / PG_NHWC_PARTS;
    gather_rows_nhwc(l1 + f * C1 * (int64_t)R2 * R2, C1, R, R2, 2, ch, cond1 + b * n1 * C1, min(n1, part * per),
                     min(n1, (part + 1) * per));
  } else {
    const int part = y - 1 - PG_NHWC_PARTS, per = (n2 + PG_NHWC_PARTS - 1) / PG_NHWC_PARTS;
    gather_rows_nhwc(l2 + f * C2 * (int64_t)R4 * R4, C2, R, R4, 4, ch, cond2 + b * n2 * C2, min(n2, part * per),
                     min(n2, (part + 1) * per));
  }
}

// out[b,i,:] = feat[b / clouds_per_frame, ind[b,i], :]   (feat [F, HW, C] channels-last)
__global__ void gather_nhwc_kernel(const float* __restrict__ feat, int clouds_per_frame, int C, int64_t HW,
                                   const int64_t* __restrict__ ind, int n, int64_t ind_stride,
                                   float* __restrict__ out, int64_t total) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t bi = e / C;
    const int64_t b = bi / n;
    const int i = (int)(bi - b * n);
    out[e] = __ldg(feat + ((b / clouds_per_frame) * HW + clamp_index(ind[b * ind_stride + i], HW)) * C + c);
  }
}

// out[b,g,j,c] = pts[b,idx[b,g,j],c] - (c<3 ? pts[b,g,c] : 0).  One warp per output
// row; lanes stride over channels (coalesced for point-major sources).
__global__ void __launch_bounds__(256)
group_gather_kernel(const float* __restrict__ pts, int n_centroids, int k, int C,
                    int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                    const int32_t* __restrict__ idx, float* __restrict__ out, int64_t ld_out,
                    float* __restrict__ center, int64_t n_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int64_t bg = row / k;
    const int g = (int)(bg % n_centroids);
    const int64_t b = bg / n_centroids;
    const float* base = pts + b * stride_cloud;
    const float* src = base + (int64_t)idx[row] * stride_point;
    const float* cen = base + (int64_t)g * stride_point;
    float* dst = out + row * ld_out;
    for (int c = lane; c < C; c += 32) {
      float v = src[c * stride_ch];
      if (c < 3) v = __fsub_rn(v, cen[c * stride_ch]);
      dst[c] = v;
    }
    if (center != nullptr && (row % k) == 0 && lane < 3) center[bg * 3 + lane] = cen[lane * stride_ch];
  }
}

// xyz[b,r,v,u] = (Kinv[r,0]*u + Kinv[r,1]*v + Kinv[r,2]) * depth  (utils.py:251-262)
__global__ void backproject_kernel(const float* __restrict__ depth, const float* __restrict__ Kinv,
                                   int H, int W, float* __restrict__ xyz, int64_t total) {
  const int64_t HW = (int64_t)H * W;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / HW;
    const int64_t pix = e % HW;
    const float u = (float)(pix % W), v = (float)(pix / W);
    const float z = depth[e];
    const float* Ki = Kinv + b * 9;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float ray = fmaf(Ki[r * 3 + 1], v, fmaf(Ki[r * 3 + 0], u, Ki[r * 3 + 2]));
      xyz[(b * 3 + r) * HW + pix] = __fmul_rn(ray, z);
    }
  }
}

// Split_coeff, one thread per hand row (Mano_render.py:160-194).
__global__ void split_coeff_kernel(const float* __restrict__ theta, int64_t ld, int col0, int pair,
                                   const int64_t* __restrict__ index, const float* __restrict__ K, int64_t n,
                                   int input_res, int down_ratio, float* __restrict__ root,
                                   float* __restrict__ pose, float* __restrict__ shape, float* __restrict__ trans) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // pair mode: rows are (frame, hand); hand 0 reads the left slice [0,61), hand 1 the right slice [61,122)
  const float* t = theta + i * ld + (pair ? 61 * (int)(i & 1) : col0);
  for (int c = 0; c < 3; ++c) root[i * 3 + c] = t[c];
  for (int c = 0; c < 45; ++c) pose[i * 45 + c] = t[3 + c];
  for (int c = 0; c < 10; ++c) shape[i * 10 + c] = t[48 + c] * 0.f;
  const float* Ki = K + (pair ? (i >> 1) : i) * 9;
  const int g = input_res / down_ratio;
  const float cx = (float)((index[i] % g) * down_ratio), cy = (float)((index[i] / g) * down_ratio);
  const float tz = __fadd_rn(t[60], 0.6f);
  trans[i * 3 + 0] = __fdiv_rn(__fmul_rn(tz, __fsub_rn(__fadd_rn(t[58], cx), Ki[2])), Ki[0]);
  trans[i * 3 + 1] = __fdiv_rn(__fmul_rn(tz, __fsub_rn(__fadd_rn(t[59], cy), Ki[5])), Ki[4]);
  trans[i * 3 + 2] = tz;
}

// im2col of the 3x3 neighbourhood of outputs around each centre pixel, for evaluating
// center_feat_up1(center_feat_up0(x0)) ONLY at `ind` (intaghand_encoder.py:790-792):
// row (b, hand, pos) holds the 3x3x C input patch of conv0 output position pos (K order: tap-major,
// channel-minor); positions outside the map are all-zero rows (zero padding of the INTERMEDIATE).
__global__ void center_im2col_kernel(const float* __restrict__ x0, const int64_t* __restrict__ ind, int C, int H,
                                     int W, float* __restrict__ rows, int64_t total) {
  const int K = 9 * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int64_t row = e / K;
    const int pos = (int)(row % 9);
    const int64_t bh = row / 9;                        // b * 2 + hand
    const int64_t b = bh >> 1;
    const int centre = (int)clamp_index(ind[bh], (int64_t)H * W);
    const int py = centre / W + pos / 3 - 1, px = centre % W + pos % 3 - 1;
    const int tap = k / C, c = k - tap * C;
    const int y = py + tap / 3 - 1, x = px + tap % 3 - 1;
    float v = 0.f;
    if (py >= 0 && py < H && px >= 0 && px < W && y >= 0 && y < H && x >= 0 && x < W)
      v = __ldg(x0 + ((b * C + c) * H + y) * (int64_t)W + x);
    rows[e] = v;
  }
}

static inline unsigned grid_for(int64_t total, int block, int max_blocks = 148 * 16) {
  int64_t g = (total + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace pdf

extern "C" int pdf_gather_nchw(const float* feat, int64_t n_clouds, int clouds_per_frame, int C, int64_t HW,
                               const int64_t* ind, int n, int64_t ind_stride, float* out, void* stream) {
  if (n_clouds == 0 || n == 0) return PDF_OK;
  PDF_REQUIRE(feat && ind && out, PDF_ERR_BAD_ARG, "pdf_gather_nchw: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && clouds_per_frame > 0 && C > 0 && HW > 0 && n >= 0, PDF_ERR_BAD_ARG,
              "pdf_gather_nchw: bad size");
  const int64_t total = n_clouds * n * C;
  if (total == 0) return PDF_OK;
  pdf::gather_nchw_kernel<<<pdf::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      feat, clouds_per_frame, C, HW, ind, n, ind_stride, out, total);
  return pdf::check_launch("pdf_gather_nchw");
}

extern "C" int pdf_pyramid_gather(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                                  int n_points, int n1, int n2, int R, const float* l0, const float* l1, int C1,
                                  const float* l2, int C2, const float* sft0_params, float* pts0, float* cond1,
                                  float* cond2, void* stream) {
  PDF_REQUIRE(xyz && choose && l0 && l1 && l2 && sft0_params && pts0 && cond1 && cond2, PDF_ERR_BAD_ARG,
              "pdf_pyramid_gather: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && clouds_per_frame > 0 && n_points > 0 && n1 >= 0 && n2 >= 0 && n1 <= n_points &&
                  n2 <= n_points && R >= 4 && C1 > 0 && C2 > 0,
              PDF_ERR_BAD_ARG, "pdf_pyramid_gather: bad size");
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(n_clouds < (1ll << 31) && (int64_t)R * R < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_pyramid_gather: too large");
  // window variant needs whole channel groups and 16-byte aligned output rows
  const int windowed = (C1 % pdf::PG_CG1 == 0) && (C2 % pdf::PG_CG2 == 0) && (R % 16 == 0);
  const int groups1 = windowed ? C1 / pdf::PG_CG1 : pdf::PG_SPLIT, groups2 = windowed ? C2 / pdf::PG_CG2 : pdf::PG_SPLIT;
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(pdf::pyramid_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pdf::PG_WIN_BYTES);
  }
  pdf::pyramid_gather_kernel<<<dim3((unsigned)n_clouds, 1 + groups1 + groups2), 256,
                               windowed ? pdf::PG_WIN_BYTES : 0, (cudaStream_t)stream>>>(
      xyz, choose, clouds_per_frame, n_points, n1, n2, R, l0, l1, C1, l2, C2, sft0_params, pts0, cond1, cond2,
      groups1, groups2, windowed);
  return pdf::check_launch("pdf_pyramid_gather");
}

extern "C" int pdf_gather_nhwc(const float* feat, int64_t n_clouds, int clouds_per_frame, int C, int64_t HW,
                               const int64_t* ind, int n, int64_t ind_stride, float* out, void* stream) {
  if (n_clouds == 0 || n == 0) return PDF_OK;
  PDF_REQUIRE(feat && ind && out, PDF_ERR_BAD_ARG, "pdf_gather_nhwc: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && clouds_per_frame > 0 && C > 0 && HW > 0 && n >= 0, PDF_ERR_BAD_ARG,
              "pdf_gather_nhwc: bad size");
  const int64_t total = n_clouds * n * C;
  pdf::gather_nhwc_kernel<<<pdf::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      feat, clouds_per_frame, C, HW, ind, n, ind_stride, out, total);
  return pdf::check_launch("pdf_gather_nhwc");
}

extern "C" int pdf_pyramid_gather_nhwc(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                                       int n_points, int n1, int n2, int R, const float* l0, const float* l1, int C1,
                                       const float* l2, int C2, const float* sft0_params, float* pts0, float* cond1,
                                       float* cond2, void* stream) {
  PDF_REQUIRE(xyz && choose && l0 && l1 && l2 && sft0_params && pts0 && cond1 && cond2, PDF_ERR_BAD_ARG,
              "pdf_pyramid_gather_nhwc: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && clouds_per_frame > 0 && n_points > 0 && n1 >= 0 && n2 >= 0 && n1 <= n_points &&
                  n2 <= n_points && R >= 4 && C1 > 0 && C2 > 0,
              PDF_ERR_BAD_ARG, "pdf_pyramid_gather_nhwc: bad size");
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(n_clouds < (1ll << 31) && (int64_t)R * R < (1ll << 31), PDF_ERR_UNSUPPORTED,
              "pdf_pyramid_gather_nhwc: too large");
  pdf::pyramid_gather_nhwc_kernel<<<dim3((unsigned)n_clouds, 1 + 2 * pdf::PG_NHWC_PARTS), 256, 0,
                                    (cudaStream_t)stream>>>(xyz, choose, clouds_per_frame, n_points, n1, n2, R, l0, l1,
                                                            C1, l2, C2, sft0_params, pts0, cond1, cond2);
  return pdf::check_launch("pdf_pyramid_gather_nhwc");
}

extern "C" int pdf_group_gather(const float* pts, int64_t n_clouds, int n_centroids, int k, int C,
                                int64_t stride_cloud, int64_t stride_point, int64_t stride_ch, const int32_t* idx,
                                float* out, int64_t ld_out, float* center, void* stream) {
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(pts && idx && out, PDF_ERR_BAD_ARG, "pdf_group_gather: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && n_centroids > 0 && k > 0 && C >= 3 && ld_out >= C, PDF_ERR_BAD_ARG,
              "pdf_group_gather: bad size");
  const int64_t rows = n_clouds * n_centroids * k;
  if (rows == 0) return PDF_OK;
  pdf::group_gather_kernel<<<pdf::grid_for(rows * 32, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(
      pts, n_centroids, k, C, stride_cloud, stride_point, stride_ch, idx, out, ld_out, center, rows);
  return pdf::check_launch("pdf_group_gather");
}

extern "C" int pdf_backproject(const float* depth, const float* Kinv, int64_t B, int H, int W, float* xyz,
                               void* stream) {
  PDF_REQUIRE(depth && Kinv && xyz, PDF_ERR_BAD_ARG, "pdf_backproject: null pointer");
  PDF_REQUIRE(B >= 0 && H > 0 && W > 0, PDF_ERR_BAD_ARG, "pdf_backproject: bad size");
  const int64_t total = B * H * W;
  if (total == 0) return PDF_OK;
  pdf::backproject_kernel<<<pdf::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(depth, Kinv, H, W, xyz,
                                                                                       total);
  return pdf::check_launch("pdf_backproject");
}

extern "C" int pdf_split_coeff(const float* theta, int64_t ld_theta, int col0, int pair, const int64_t* index,
                               const float* K, int64_t n, int input_res, int down_ratio, float* root, float* pose,
                               float* shape, float* trans, void* stream) {
  PDF_REQUIRE(theta && index && K && root && pose && shape && trans, PDF_ERR_BAD_ARG, "pdf_split_coeff: null pointer");
  PDF_REQUIRE(n >= 0 && input_res > 0 && down_ratio > 0 && col0 >= 0 && ld_theta >= (pair ? 122 : col0 + 61),
              PDF_ERR_BAD_ARG, "pdf_split_coeff: bad size");
  if (n == 0) return PDF_OK;
  pdf::split_coeff_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      theta, ld_theta, col0, pair, index, K, n, input_res, down_ratio, root, pose, shape, trans);
  return pdf::check_launch("pdf_split_coeff");
}

extern "C" int pdf_center_im2col(const float* x0, const int64_t* ind, int64_t B, int C, int H, int W, float* rows,
                                 void* stream) {
  if (B == 0) return PDF_OK;
  PDF_REQUIRE(x0 && ind && rows, PDF_ERR_BAD_ARG, "pdf_center_im2col: null pointer");
  PDF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, PDF_ERR_BAD_ARG, "pdf_center_im2col: bad size");
  const int64_t total = B * 2 * 9 * 9 * C;
  pdf::center_im2col_kernel<<<pdf::grid_for(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(x0, ind, C, H, W,
                                                                                                  rows, total);
  return pdf::check_launch("pdf_center_im2col");
}
