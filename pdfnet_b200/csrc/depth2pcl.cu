// Device-side, batched depth2pcl: per-hand cloud construction from raw depth.
// One CTA (1024 threads) per (frame, hand).  The frame is read from L2/HBM twice
// (z statistics, candidate flags); afterwards everything works on a candidate
// bitmask held in shared memory, so the algorithmic traffic is depth + 2 masks in,
// choose + cloud out.
#include "pdf_common.cuh"

namespace pdf {

constexpr int D2P_THREADS = 1024;

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  // 1024 threads = 32 warps; returns the exclusive prefix of v over the block.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                       // protect s_warp reuse
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;             // exclusive warp offsets
    if (lane == 31) s_warp[32] = winc;   // block total
  }
  __syncthreads();
  total = s_warp[32];
  return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(D2P_THREADS)
depth2pcl_kernel(const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ Kinv,
                 const float* __restrict__ valid, const int32_t* __restrict__ subset_keys,
                 const int32_t* __restrict__ perm, int H, int W, int n_points, int min_pixels,
                 int64_t* __restrict__ choose, float* __restrict__ cloud, int32_t* __restrict__ n_cand_out) {
  extern __shared__ uint32_t s_bits[];             // candidate bitmask, nwords
  __shared__ int s_sel[1024];
  __shared__ int s_warp[33];
  __shared__ double s_dsum[32];
  __shared__ int s_hist[256];
  __shared__ float s_lohi[2];
  __shared__ int s_misc[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f = blockIdx.x >> 1;
  const int hand = blockIdx.x & 1;                 // 0 = left, 1 = right
  const int mch = hand == 0 ? 1 : 0;               // intaghand_encoder.py:376-377
  const int npx = H * W;
  const int nwords = (npx + 31) >> 5;
  const float* dep = depth + f * npx;
  const float* msk = mask + (f * 2 + mch) * npx;
  const float* Ki = Kinv + f * 9;
  const float k20 = Ki[6], k21 = Ki[7], k22 = Ki[8];
  int64_t* ch_out = choose + (f * 2 + hand) * n_points;
  float* cl_out = cloud + (f * 2 + hand) * n_points * 3;

  if (valid[f * 2 + hand] != 1.f) {                // :401,:430-435 -> zeros
    for (int i = tid; i < n_points; i += D2P_THREADS) {
      ch_out[i] = 0;
      cl_out[i * 3] = 0.f; cl_out[i * 3 + 1] = 0.f; cl_out[i * 3 + 2] = 0.f;
    }
    if (tid == 0) n_cand_out[f * 2 + hand] = 0;
    return;
  }

  auto masked_depth = [&](int pix) -> float {      // depth * noise_mask * (mask > 0.5)   :392-395
    const float d = dep[pix];
    const bool keep = (0.2f < d) && (2.5f > d) && (msk[pix] > 0.5f);
    return keep ? d : 0.f;
  };
  auto z_uv = [&](float u, float v, float zm) -> float { return __fmul_rn(fmaf(k21, v, fmaf(k20, u, k22)), zm); };
  // pixel coordinates of lane `lane` of 32-pixel word w0 without a per-pixel division when W % 32 == 0
  const int wpr = (W & 31) == 0 ? (W >> 5) : 0;
  auto word_uv = [&](int w0, float& u, float& v) {
    if (wpr) { const int row = w0 / wpr; u = (float)(((w0 - row * wpr) << 5) + lane); v = (float)row; }
    else { const int pix = w0 * 32 + lane; u = (float)(pix % W); v = (float)(pix / W); }
  };

  // Every pass walks the frame one 32-pixel word per warp and keeps U words (U independent loads per
  // thread) in flight: the kernel is otherwise bound by the latency of one load per iteration.
  constexpr int U = 8;
  // pass A: mean z over non-zero pixels (:407)
  double sum = 0.0;
  int cnt = 0;
  for (int wb = warp * U; wb < nwords; wb += 32 * U) {        // U consecutive words per warp batch
    float d[U], m[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int pix = (wb + q) * 32 + lane;
      const bool ok = pix < npx;
      d[q] = ok ? __ldg(dep + pix) : 0.f;
      m[q] = ok ? __ldg(msk + pix) : 0.f;
    }
    int row = wpr ? wb / wpr : 0, wcol = wpr ? wb - row * wpr : 0;   // one division per batch
#pragma unroll
    for (int q = 0; q < U; ++q) {
      float u, v;
      if (wpr) { u = (float)((wcol << 5) + lane); v = (float)row; if (++wcol == wpr) { wcol = 0; ++row; } }
      else { const int pix = (wb + q) * 32 + lane; u = (float)(pix % W); v = (float)(pix / W); }
      const bool keep = (0.2f < d[q]) && (2.5f > d[q]) && (m[q] > 0.5f);
      const float z = z_uv(u, v, keep ? d[q] : 0.f);
      if (z != 0.f) { sum += (double)z; ++cnt; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { s_dsum[warp] = sum; s_warp[warp] = cnt; }
  __syncthreads();
  if (warp == 0) {
    double s = s_dsum[lane];
    int c = s_warp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      const float mean = c > 0 ? (float)(s / (double)c) : 0.f;
      s_lohi[0] = fmaxf(0.2f, __fsub_rn(mean, 0.08f));     // :408
      s_lohi[1] = fminf(2.5f, __fadd_rn(mean, 0.08f));
      s_misc[0] = c;
      s_misc[3] = 0;
    }
  }
  __syncthreads();
  const int n_nonzero = s_misc[0];
  const float lo = s_lohi[0], hi = s_lohi[1];
  uint32_t* s_wlist = s_bits + 3 * nwords;

  // pass B: candidate bitmask (:409)
  int my_cand = 0;
  for (int wb = warp * U; wb < nwords; wb += 32 * U) {
    float d[U], m[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int pix = (wb + q) * 32 + lane;
      const bool ok = pix < npx;
      d[q] = ok ? __ldg(dep + pix) : 0.f;
      m[q] = ok ? __ldg(msk + pix) : 0.f;
    }
    int row = wpr ? wb / wpr : 0, wcol = wpr ? wb - row * wpr : 0;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int w0 = wb + q;
      float u, v;
      if (wpr) { u = (float)((wcol << 5) + lane); v = (float)row; if (++wcol == wpr) { wcol = 0; ++row; } }
      else { const int pix = w0 * 32 + lane; u = (float)(pix % W); v = (float)(pix / W); }
      if (w0 < nwords) {                               // warp-uniform
        const bool keep = (0.2f < d[q]) && (2.5f > d[q]) && (m[q] > 0.5f);
        const float z = z_uv(u, v, keep ? d[q] : 0.f);
        const bool c = n_nonzero > 0 && (z > lo) && (z < hi);
        const unsigned b = __ballot_sync(0xffffffffu, c);
        if (lane == 0) {
          s_bits[w0] = b; my_cand += __popc(b);
          if (b) s_wlist[atomicAdd(&s_misc[3], 1)] = w0;   // non-empty words (unordered) for the key passes
        }
      }
    }
  }
  int n_cand;
  (void)block_exclusive_scan(my_cand, s_warp, n_cand);
  if (tid == 0) n_cand_out[f * 2 + hand] = n_cand;

  // words owned by this thread for the ordered enumerations below
  const int wpt = (nwords + D2P_THREADS - 1) / D2P_THREADS;
  const int wbeg = min(nwords, tid * wpt), wend = min(nwords, wbeg + wpt);

  int n_sel = 0;                                   // entries valid in s_sel
  if (n_cand >= min_pixels && n_cand <= n_points) {
    // keep every candidate in pixel order (:424 pads by wrapping)
    int c = 0;
    for (int w = wbeg; w < wend; ++w) c += __popc(s_bits[w]);
    int tot;
    int r = block_exclusive_scan(c, s_warp, tot);
    for (int w = wbeg; w < wend; ++w) {
      unsigned b = s_bits[w];
      while (b) { const int bit = __ffs(b) - 1; b &= b - 1; s_sel[r++] = w * 32 + bit; }
    }
    n_sel = n_cand;
  } else if (n_cand > n_points) {
    // random subset (:418-422): the n_points candidates with the smallest keys.  Every pass walks the
    // frame one 32-pixel word per warp iteration (lane = pixel): balanced, coalesced key loads.
    const int32_t* keys = subset_keys + (f * 2 + hand) * (int64_t)npx;
    uint32_t* s_lt = s_bits + nwords;              // selected (key < threshold) bitmask
    uint32_t* s_eq = s_bits + 2 * nwords;          // key == threshold bitmask
    const int n_words_used = s_misc[3];
    uint32_t prefix = 0;
    int remaining = n_points;                      // rank (1-based) of the threshold inside the prefix bucket
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const uint32_t hmask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int wb = warp; wb < n_words_used; wb += 32 * U) {
        uint32_t kx[U];
        bool on[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int li = wb + q * 32;
          const int w0 = li < n_words_used ? (int)s_wlist[li] : 0;
          on[q] = li < n_words_used && ((s_bits[w0] >> lane) & 1u);
          kx[q] = on[q] ? ((uint32_t)__ldg(keys + w0 * 32 + lane) ^ 0x80000000u) : 0u;   // signed order -> unsigned
        }
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (on[q] && (kx[q] & hmask) == prefix) atomicAdd(&s_hist[(kx[q] >> shift) & 255], 1);
      }
      __syncthreads();
      if (warp == 0) {                             // find the bucket holding the `remaining`-th key: warp scan over 256 bins
        int h[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { h[q] = s_hist[lane * 8 + q]; sum += h[q]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int acc = inc - sum;                       // keys in lower bins
        const bool here = acc < remaining && remaining <= inc;
        if (here) {
          int d = 0;
          for (; d < 8; ++d) { if (acc + h[d] >= remaining) break; acc += h[d]; }
          s_misc[1] = lane * 8 + d; s_misc[2] = remaining - acc;
        }
      }
      __syncthreads();
      prefix |= ((uint32_t)s_misc[1]) << shift;
      remaining = s_misc[2];
      __syncthreads();
    }
    const uint32_t thr = prefix;                   // n_points-th smallest key; `remaining` ties are taken
    for (int w = tid; w < nwords; w += D2P_THREADS) { s_lt[w] = 0; s_eq[w] = 0; }
    __syncthreads();
    for (int wb = warp; wb < n_words_used; wb += 32 * U) {
      uint32_t kx[U];
      bool on[U];
      int wq[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int li = wb + q * 32;
        wq[q] = li < n_words_used ? (int)s_wlist[li] : -1;
        on[q] = wq[q] >= 0 && ((s_bits[wq[q]] >> lane) & 1u);
        kx[q] = on[q] ? ((uint32_t)__ldg(keys + wq[q] * 32 + lane) ^ 0x80000000u) : 0u;
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        if (wq[q] >= 0) {                              // warp-uniform
          const unsigned blt = __ballot_sync(0xffffffffu, on[q] && kx[q] < thr);
          const unsigned beq = __ballot_sync(0xffffffffu, on[q] && kx[q] == thr);
          if (lane == 0) { s_lt[wq[q]] = blt; s_eq[wq[q]] = beq; }
        }
      }
    }
    __syncthreads();
    // ties at the threshold: the first `remaining` in pixel order join the selection
    int c_eq = 0;
    for (int w = wbeg; w < wend; ++w) c_eq += __popc(s_eq[w]);
    int tot;
    int er = block_exclusive_scan(c_eq, s_warp, tot);
    for (int w = wbeg; w < wend; ++w) {
      unsigned bq = s_eq[w], add = 0;
      while (bq && er < remaining) { add |= bq & (0u - bq); bq &= bq - 1; ++er; }
      s_lt[w] |= add;
    }
    int c_sel = 0;
    for (int w = wbeg; w < wend; ++w) c_sel += __popc(s_lt[w]);
    int r = block_exclusive_scan(c_sel, s_warp, tot);
    for (int w = wbeg; w < wend; ++w) {
      unsigned bsel = s_lt[w];
      while (bsel) { const int bit = __ffs(bsel) - 1; bsel &= bsel - 1; if (r < 1024) s_sel[r] = w * 32 + bit; ++r; }
    }
    n_sel = n_points;
  }
  __syncthreads();

  // final order + back-projection of the kept pixels (:427-428)
  const int32_t* pm = perm ? perm + (f * 2 + hand) * (int64_t)n_points : nullptr;
  for (int i = tid; i < n_points; i += D2P_THREADS) {
    const int src = pm ? pm[i] : i;
    const int pix = n_sel > 0 ? s_sel[src % n_sel] : 0;
    const float zm = masked_depth(pix);
    const float u = (float)(pix % W), v = (float)(pix / W);
    ch_out[i] = pix;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      cl_out[i * 3 + r] = __fmul_rn(fmaf(Ki[r * 3 + 1], v, fmaf(Ki[r * 3], u, Ki[r * 3 + 2])), zm);
  }
}

}  // namespace pdf

extern "C" int pdf_depth2pcl(const float* depth, const float* mask, const float* Kinv, const float* valid,
                             const int32_t* subset_keys, const int32_t* perm, int64_t B, int H, int W,
                             int n_points, int min_pixels, int64_t* choose, float* cloud, int32_t* n_cand,
                             void* stream) {
  PDF_REQUIRE(depth && mask && Kinv && valid && choose && cloud && n_cand, PDF_ERR_BAD_ARG,
              "pdf_depth2pcl: null pointer");
  PDF_REQUIRE(B >= 0 && H > 0 && W > 0 && min_pixels >= 1, PDF_ERR_BAD_ARG, "pdf_depth2pcl: bad size");
  PDF_REQUIRE(n_points == 1024, PDF_ERR_UNSUPPORTED, "pdf_depth2pcl: n_points must be 1024 (got %d)", n_points);
  PDF_REQUIRE((int64_t)H * W <= 640 * 640, PDF_ERR_UNSUPPORTED, "pdf_depth2pcl: frame larger than 640x640");
  PDF_REQUIRE(subset_keys != nullptr || (int64_t)H * W <= n_points, PDF_ERR_BAD_ARG,
              "pdf_depth2pcl: subset_keys required when a hand can exceed n_points pixels");
  if (B == 0) return PDF_OK;
  const size_t smem = (size_t)(((int64_t)H * W + 31) / 32) * 4 * 4;      // candidate / selected / tie bitmasks + word list
  static pdf::PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(pdf::depth2pcl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
  }
  pdf::depth2pcl_kernel<<<(unsigned)(B * 2), pdf::D2P_THREADS, smem, (cudaStream_t)stream>>>(
      depth, mask, Kinv, valid, subset_keys, perm, H, W, n_points, min_pixels, choose, cloud, n_cand);
  return pdf::check_launch("pdf_depth2pcl");
}
