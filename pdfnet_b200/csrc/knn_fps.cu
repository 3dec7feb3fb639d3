// Neighbour search (kNN-then-radius-mask) and farthest point sampling.
// Both are SM-latency/issue bound at N <= 1024 (12 KB per cloud): the cloud lives
// in shared memory, distances in registers, selection is done with warp votes.
#include "pdf_common.cuh"

namespace pdf {

// ---------------------------------------------------------------------------------
// kNN + ball mask.  One warp per centroid; lane l owns points j = 32 t + l.
// The k-th smallest distance is found by a bitwise radix descent over the fp32 bit
// pattern (distances are >= 0, so unsigned order == float order); one
// redux.sync per bit, early exit when a prefix splits off exactly k points.
// ---------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256, 2)
knn_ball_kernel(const float* __restrict__ xyz, int n_points, int n_centroids, int k, float r2,
                int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                int32_t* __restrict__ idx_out, int chunks_per_cloud, int centroids_per_cta) {
  constexpr int NP = T * 32;
  __shared__ float sx[NP], sy[NP], sz[NP];
  const int b = blockIdx.x / chunks_per_cloud;
  const int chunk = blockIdx.x % chunks_per_cloud;
  const float* base = xyz + (int64_t)b * stride_cloud;
  for (int j = threadIdx.x; j < NP; j += blockDim.x) {
    float x = __int_as_float(0x7f800000), y = x, z = x;   // padding: +inf, never selected
    if (j < n_points) {
      const float* p = base + (int64_t)j * stride_point;
      x = p[0]; y = p[stride_ch]; z = p[2 * stride_ch];
    }
    sx[j] = x; sy[j] = y; sz[j] = z;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int c_begin = chunk * centroids_per_cta;
  const int c_end = min(n_centroids, c_begin + centroids_per_cta);

  for (int i = c_begin + warp; i < c_end; i += nwarps) {
    const float cx = sx[i], cy = sy[i], cz = sz[i];
    uint32_t d[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int j = t * 32 + lane;
      d[t] = __float_as_uint(sqdist_rn(sx[j], sy[j], sz[j], cx, cy, cz));
    }
    // radix descent for the k-th smallest bit pattern.  Counting uses four independent
    // partial sums (setp + predicated add) so the per-bit latency is T/4 dependent adds.
    uint32_t prefix = 0, limit = 0;
    bool split = false;
    // bits on which every distance agrees need no vote: start below the common prefix
    uint32_t all_or = 0, all_and = 0xffffffffu;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (t * 32 < n_points) { all_or |= d[t]; all_and &= d[t]; }
    }
    all_or = __reduce_or_sync(0xffffffffu, all_or);
    all_and = __reduce_and_sync(0xffffffffu, all_and);
    int bit = 31 - __clz((all_or ^ all_and) | 1u);        // highest bit that differs (>= 0)
    prefix = (bit >= 30) ? 0u : (all_and & ~((2u << bit) - 1u));
    if (bit > 30) bit = 30;
#pragma unroll 1
    for (; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
      for (int t = 0; t < T; t += 4) {
        asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
            "setp.lt.u32 p0, %4, %8;\n\tsetp.lt.u32 p1, %5, %8;\n\t"
            "setp.lt.u32 p2, %6, %8;\n\tsetp.lt.u32 p3, %7, %8;\n\t"
            "@p0 add.s32 %0, %0, 1;\n\t@p1 add.s32 %1, %1, 1;\n\t"
            "@p2 add.s32 %2, %2, 1;\n\t@p3 add.s32 %3, %3, 1;\n\t}"
            : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3)
            : "r"(d[t]), "r"(d[t + 1]), "r"(d[t + 2]), "r"(d[t + 3]), "r"(cand));
      }
      const int cnt = __reduce_add_sync(0xffffffffu, (c0 + c1) + (c2 + c3));
      if (cnt == k) { limit = cand; split = true; break; }
      if (cnt < k) prefix = cand;
    }
    int need_eq = 0;
    if (!split) {
      limit = prefix;                       // == the k-th smallest value
      int cnt = 0;
#pragma unroll
      for (int t = 0; t < T; ++t) cnt += (d[t] < limit) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      need_eq = k - cnt;                    // ties at the k-th distance: lowest index first
    }
    int32_t* out = idx_out + ((int64_t)b * n_centroids + i) * k;
    int written = 0, eq_seen = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const bool lt = d[t] < limit;
      const bool eq = (need_eq > 0) && (d[t] == limit);
      const unsigned beq = __ballot_sync(0xffffffffu, eq);
      const bool sel = lt || (eq && (eq_seen + __popc(beq & lt_mask)) < need_eq);
      const unsigned bs = __ballot_sync(0xffffffffu, sel);
      if (sel) {
        const int pos = written + __popc(bs & lt_mask);
        out[pos] = (__uint_as_float(d[t]) > r2) ? i : (t * 32 + lane);
      }
      written += __popc(bs);
      eq_seen += __popc(beq);
    }
  }
}

// ---------------------------------------------------------------------------------
// FPS.  One CTA per cloud, 256 threads, thread owns points j = p*256 + tid with
// coordinates and running min-distance in registers; one __syncthreads per round.
// ---------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256)
fps_kernel(const float* __restrict__ xyz, int n_points, int n_sample, const int32_t* __restrict__ start_idx,
           int64_t stride_cloud, int64_t stride_point, int64_t stride_ch, int32_t* __restrict__ idx_out) {
  extern __shared__ float smem[];
  const int NP = P * 256;
  float* sx = smem; float* sy = sx + NP; float* sz = sy + NP;
  __shared__ uint32_t w_bits[2][8];
  __shared__ int w_idx[2][8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = xyz + (int64_t)b * stride_cloud;
  float px[P], py[P], pz[P], md[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int j = p * 256 + tid;
    float x = 0.f, y = 0.f, z = 0.f;
    if (j < n_points) {
      const float* q = base + (int64_t)j * stride_point;
      x = q[0]; y = q[stride_ch]; z = q[2 * stride_ch];
    }
    px[p] = x; py[p] = y; pz[p] = z;
    sx[j] = x; sy[j] = y; sz[j] = z;
  }
  __syncthreads();
  int cur = start_idx[b];
  cur = min(max(cur, 0), n_points - 1);
  int32_t* out = idx_out + (int64_t)b * n_sample;
  if (tid == 0) out[0] = cur;
  {
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
#pragma unroll
    for (int p = 0; p < P; ++p)
      md[p] = (p * 256 + tid < n_points) ? sqdist_rn(px[p], py[p], pz[p], cx, cy, cz) : 0.f;
  }
  for (int s = 1; s < n_sample; ++s) {
    // argmax of min_dist, first occurrence (interhand.py:169)
    uint32_t bb = __float_as_uint(md[0]);
    int bi = tid;
#pragma unroll
    for (int p = 1; p < P; ++p) {
      const uint32_t v = __float_as_uint(md[p]);
      if (v > bb) { bb = v; bi = p * 256 + tid; }
    }
    const uint32_t wm = __reduce_max_sync(0xffffffffu, bb);
    const int wi = __reduce_min_sync(0xffffffffu, (bb == wm) ? bi : 0x7fffffff);
    const int slot = s & 1;
    if (lane == 0) { w_bits[slot][warp] = wm; w_idx[slot][warp] = wi; }
    __syncthreads();
    uint32_t gm = w_bits[slot][0];
    int gi = w_idx[slot][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const uint32_t v = w_bits[slot][w];
      const int vi = w_idx[slot][w];
      if (v > gm || (v == gm && vi < gi)) { gm = v; gi = vi; }
    }
    cur = gi;
    if (tid == 0) out[s] = cur;
    // lower min_dist only where it is still > 1e-8 (interhand.py:172-175)
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float dnew = sqdist_rn(px[p], py[p], pz[p], cx, cy, cz);
      if (md[p] > 1e-8f) md[p] = fminf(md[p], dnew);
    }
  }
}

}  // namespace pdf

extern "C" int pdf_knn_ball(const float* xyz, int64_t n_clouds, int n_points, int n_centroids, int k, float r2,
                            int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                            int32_t* idx_out, void* stream) {
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(xyz && idx_out, PDF_ERR_BAD_ARG, "pdf_knn_ball: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && n_points > 0 && n_centroids > 0 && k > 0, PDF_ERR_BAD_ARG,
              "pdf_knn_ball: non-positive size");
  PDF_REQUIRE(n_points <= 1024 && k <= n_points && n_centroids <= n_points, PDF_ERR_UNSUPPORTED,
              "pdf_knn_ball: need k <= n_points <= 1024 and n_centroids <= n_points (got n=%d k=%d n1=%d)",
              n_points, k, n_centroids);
  if (n_clouds == 0) return PDF_OK;
  const int per_cta = 64;
  const int chunks = (n_centroids + per_cta - 1) / per_cta;
  PDF_REQUIRE(n_clouds * chunks < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_knn_ball: grid too large");
  dim3 grid((unsigned)(n_clouds * chunks));
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(T)                                                                                   \
  pdf::knn_ball_kernel<T><<<grid, 256, 0, s>>>(xyz, n_points, n_centroids, k, r2, stride_cloud,     \
                                               stride_point, stride_ch, idx_out, chunks, per_cta)
  if (n_points <= 128) LAUNCH(4);
  else if (n_points <= 256) LAUNCH(8);
  else if (n_points <= 512) LAUNCH(16);
  else LAUNCH(32);
#undef LAUNCH
  return pdf::check_launch("pdf_knn_ball");
}

extern "C" int pdf_fps(const float* xyz, int64_t n_clouds, int n_points, int n_sample, const int32_t* start_idx,
                       int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                       int32_t* idx_out, void* stream) {
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(xyz && idx_out && start_idx, PDF_ERR_BAD_ARG, "pdf_fps: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && n_points > 0 && n_sample > 0, PDF_ERR_BAD_ARG, "pdf_fps: non-positive size");
  PDF_REQUIRE(n_points <= 4096 && n_sample <= n_points, PDF_ERR_UNSUPPORTED,
              "pdf_fps: need n_sample <= n_points <= 4096 (got %d, %d)", n_sample, n_points);
  if (n_clouds == 0) return PDF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)n_clouds);
#define LAUNCH(P)                                                                                      \
  pdf::fps_kernel<P><<<grid, 256, 3 * (P) * 256 * sizeof(float), s>>>(xyz, n_points, n_sample, start_idx, \
                                                                     stride_cloud, stride_point, stride_ch, idx_out)
  if (n_points <= 512) LAUNCH(2);
  else if (n_points <= 1024) LAUNCH(4);
  else if (n_points <= 2048) LAUNCH(8);
  else {
    static bool attr = false;   // 48 KB of coordinates + static barriers exceeds the default dynamic limit
    if (!attr) {
      cudaFuncSetAttribute(pdf::fps_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * 256 * 4);
      attr = true;
    }
    LAUNCH(16);
  }
#undef LAUNCH
  return pdf::check_launch("pdf_fps");
}
