// Neighbour search (kNN-then-radius-mask) and farthest point sampling.
// Both are SM-latency/issue bound at N <= 1024 (12 KB per cloud): the cloud lives
// in shared memory, distances in registers, selection is done with warp votes.
#include "pdf_common.cuh"

namespace pdf {

// ---------------------------------------------------------------------------------
// kNN + ball mask.  One warp per centroid; lane l owns the 32 points j = 32 t + l.
// Selection of the k smallest distances is a BIT-SLICED radix descent: the 32 distance
// bit patterns of a lane (non-negative floats order as unsigned integers) are transposed
// in registers into 31 bit planes (plane[b] bit t = bit b of d[t]); per bit the lane's
// contribution to "how many still-active candidates have a 0 here" is one LOP3 + POPC,
// summed across the warp with one redux.sync.  ~8 instructions per bit instead of one
// compare per point per bit.  Ties at the k-th distance go to the lowest index.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
  // Hacker's Delight 7-3: afterwards a[r] bit c == old a[31-c] bit (31-r).
  // The 16- and 8-bit stages are pure byte permutes (2 PRMT per register pair instead of 5 ALU ops).
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t lo = a[k], hi = a[k + 16];
    a[k] = __byte_perm(lo, hi, 0x3276);        // (lo & 0xffff0000) | (hi >> 16)
    a[k + 16] = __byte_perm(lo, hi, 0x1054);   // (hi & 0x0000ffff) | (lo << 16)
  }
#pragma unroll
  for (int k = 0; k < 32; k = (k + 8 + 1) & ~8) {
    const uint32_t lo = a[k], hi = a[k + 8];
    a[k] = __byte_perm(lo, hi, 0x3715);        // bytes 0,2 of lo <- bytes 1,3 of hi
    a[k + 8] = __byte_perm(lo, hi, 0x2604);    // bytes 1,3 of hi <- bytes 0,2 of lo
  }
  uint32_t m = 0x0F0F0F0Fu;
#pragma unroll
  for (int j = 4; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
      const uint32_t t = (a[k] ^ (a[k + j] >> j)) & m;
      a[k] ^= t;
      a[k + j] ^= (t << j);
    }
  }
}

template <int T>   // T = rows of 32 points actually present (16 for n_points <= 512): the rest folds away at compile time
__global__ void __launch_bounds__(256, 2)
knn_ball_kernel(const float* __restrict__ xyz, int n_points, int n_centroids, int k, float r2,
                int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                int32_t* __restrict__ idx_out, int chunks_per_cloud, int centroids_per_cta) {
  constexpr int NP = T * 32;
  __shared__ float sx[NP], sy[NP], sz[NP];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x / chunks_per_cloud;
  const int chunk = blockIdx.x % chunks_per_cloud;
  const float* base = xyz + (int64_t)b * stride_cloud;
  for (int j = threadIdx.x; j < NP; j += blockDim.x) {
    float x = 0.f, y = 0.f, z = 0.f;
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
  const uint32_t r2b = __float_as_uint(fmaxf(r2, 0.f));
  // bit t set <=> point 32 t + lane exists
  uint32_t valid = 0;
#pragma unroll
  for (int t = 0; t < T; ++t) valid |= (t * 32 + lane < n_points) ? (1u << t) : 0u;

  for (int i = c_begin + warp; i < c_end; i += nwarps) {
    const float cx = sx[i], cy = sy[i], cz = sz[i];
    uint32_t a[32];                                   // a[31 - t] = bits of d(point 32 t + lane)
    uint32_t far = 0;                                 // bit t <=> d > r2 (radius mask, utils.py:149-151)
#pragma unroll
    for (int t = 31; t >= 0; --t) {
      if (t >= T) { a[31 - t] = 0; continue; }        // rows that do not exist: compile-time zeros
      const int j = t * 32 + lane;
      const uint32_t d = __float_as_uint(sqdist_rn(sx[j], sy[j], sz[j], cx, cy, cz));
      a[31 - t] = d;
      far = __funnelshift_l(r2b - d, far, 1);         // shifts in the sign of (r2 - d): 1 <=> d > r2
    }
    transpose32(a);                                   // now a[31 - B] bit t = bit B of d(point 32 t + lane)

    uint32_t active = valid, sel = 0;
    int kr = k;                                       // how many of the active set are still to be taken
#pragma unroll
    for (int B = 30; B >= 0; --B) {
      const uint32_t z = active & ~a[31 - B];         // active candidates with a 0 at this bit (the smaller ones)
      const int c = __reduce_add_sync(0xffffffffu, __popc(z));
      if (c >= kr) active = z;                        // the k-th smallest has a 0 here
      else { sel |= z; kr -= c; active &= a[31 - B]; }
    }
    // `active` now holds the candidates EQUAL to the k-th smallest value; kr >= 1 of them are needed
    const int n_eq = __reduce_add_sync(0xffffffffu, __popc(active));
    if (n_eq == kr) sel |= active;
    else {                                            // exact ties: lowest index first (index = 32 t + lane)
#pragma unroll 1
      for (int t = 0; t < 32 && kr > 0; ++t) {
        const bool mine = (active >> t) & 1u;
        const unsigned bt = __ballot_sync(0xffffffffu, mine);
        const int cnt = __popc(bt);
        if (mine && (cnt <= kr || __popc(bt & lt_mask) < kr)) sel |= 1u << t;
        kr -= min(cnt, kr);
      }
    }
    // each lane appends its selected points; order inside a group is irrelevant to every consumer
    int pos = __popc(sel);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pos, o);
      if (lane >= o) pos += v;
    }
    pos -= __popc(sel);
    int32_t* out = idx_out + ((int64_t)b * n_centroids + i) * k + pos;
    uint32_t m = sel;
    while (m) {
      const int t = __ffs(m) - 1;
      m &= m - 1;
      *out++ = ((far >> t) & 1u) ? i : (t * 32 + lane);
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
  if (n_points <= 512)
    pdf::launch_pdl(pdf::knn_ball_kernel<16>, grid, dim3(256), 0, s, xyz, n_points, n_centroids, k, r2, stride_cloud,
                    stride_point, stride_ch, idx_out, chunks, per_cta);
  else
    pdf::launch_pdl(pdf::knn_ball_kernel<32>, grid, dim3(256), 0, s, xyz, n_points, n_centroids, k, r2, stride_cloud,
                    stride_point, stride_ch, idx_out, chunks, per_cta);
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
    static pdf::PerDeviceOnce once;   // 48 KB of coordinates + static barriers exceeds the default dynamic limit
    if (once.first()) {
      cudaFuncSetAttribute(pdf::fps_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16 * 256 * 4);
    }
    LAUNCH(16);
  }
#undef LAUNCH
  return pdf::check_launch("pdf_fps");
}
