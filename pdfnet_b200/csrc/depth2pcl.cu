// Device-side, batched depth2pcl: per-hand cloud construction from raw depth
// (intaghand_encoder.py:369-491; dataset twin interhand.py:758-908).
// One CTA (512 threads, two per SM) per (frame, hand).  The hand mask is scanned once (one 32-pixel word per lane;
// 32 B of a uint8 mask or 128 B of an fp32 mask); depth is only read where the mask is set: warps walk
// the ORDERED list of non-empty words with lane = pixel (coalesced 128 B), so the work after the mask
// scan scales with the hand, not with the frame.  Everything else lives in shared-memory bitmasks.
// Randomness (the two np.random.shuffle calls, :421,:427) is injected: explicit arrays
// (subset_keys / perm) or, when they are null, counter-based functions of (seed, cloud, pixel) that the
// host can reproduce bit for bit (pdfnet_b200.synth.hash_keys / feistel_perm).
#include "pdf_common.cuh"

namespace pdf {

constexpr int D2P_THREADS = 512;            // two CTAs per SM: one CTA's block-wide scans overlap the other's loads
constexpr int D2P_WARPS = D2P_THREADS / 32;

__host__ __device__ __forceinline__ uint32_t d2p_mix(uint32_t h) {     // murmur3 finaliser
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// subset key of pixel `pix` of cloud `cloud` (unsigned order; ties -> lower pixel)
__host__ __device__ __forceinline__ uint32_t d2p_key(uint32_t seed, uint32_t cloud, uint32_t pix) {
  return d2p_mix(pix * 0x9E3779B1u + d2p_mix(seed ^ (cloud * 0x27D4EB2Fu + 0x165667B1u)));
}
// pseudo-random bijection of [0, 1024): 4-round Feistel network on two 5-bit halves
__host__ __device__ __forceinline__ uint32_t d2p_perm1024(uint32_t seed, uint32_t cloud, uint32_t i) {
  const uint32_t k = d2p_mix(seed * 0x9E3779B1u + cloud + 0x7F4A7C15u);
  uint32_t l = i >> 5, r = i & 31u;
#pragma unroll
  for (uint32_t round = 0; round < 4; ++round) {
    const uint32_t f = d2p_mix(r + 32u * round + k) & 31u;
    const uint32_t t = l ^ f;
    l = r; r = t;
  }
  return (l << 5) | r;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  // D2P_WARPS warps; returns the exclusive prefix of v over the block.
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
    int w = lane < D2P_WARPS ? s_warp[lane] : 0;
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

template <bool MASK_U8>
__global__ void __launch_bounds__(D2P_THREADS)
depth2pcl_kernel(const float* __restrict__ depth, const void* __restrict__ mask_v, const float* __restrict__ Kinv,
                 const float* __restrict__ valid, const int32_t* __restrict__ subset_keys,
                 const int32_t* __restrict__ perm, uint32_t seed, int H, int W, int n_points, int min_pixels,
                 int64_t* __restrict__ choose, float* __restrict__ cloud, int32_t* __restrict__ n_cand_out) {
  extern __shared__ uint32_t s_bits[];             // [nwords] mask -> keep -> candidate bitmask
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
  const uint32_t cloud_id = (uint32_t)blockIdx.x;
  const int npx = H * W;
  const int nwords = (npx + 31) >> 5;
  const float* dep = depth + f * npx;
  const float* Ki = Kinv + f * 9;
  const float k20 = Ki[6], k21 = Ki[7], k22 = Ki[8];
  int64_t* ch_out = choose + (f * 2 + hand) * n_points;
  float* cl_out = cloud + (f * 2 + hand) * n_points * 3;
  uint32_t* s_lt = s_bits + nwords;                // selected (key < threshold) bitmask
  uint32_t* s_eq = s_bits + 2 * nwords;            // key == threshold bitmask
  uint32_t* s_wlist = s_bits + 3 * nwords;         // ordered list of non-empty words

  if (valid[f * 2 + hand] != 1.f) {                // :401,:430-435 -> zeros
    for (int i = tid; i < n_points; i += D2P_THREADS) {
      ch_out[i] = 0;
      cl_out[i * 3] = 0.f; cl_out[i * 3 + 1] = 0.f; cl_out[i * 3 + 2] = 0.f;
    }
    if (tid == 0) n_cand_out[f * 2 + hand] = 0;
    return;
  }

  // ---- pass 0: mask > 0.5 as one bit per pixel (:376-377,:395) ----
  if (MASK_U8) {
    // lane = word: 32 mask bytes = two 16 B loads.  (npx % 32 == 0 and 16 B alignment checked by the launcher.)
    const uint8_t* msk = reinterpret_cast<const uint8_t*>(mask_v) + (f * 2 + mch) * (int64_t)npx;
    for (int w = tid; w < nwords; w += D2P_THREADS) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(msk + (int64_t)w * 32));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(msk + (int64_t)w * 32) + 1);
      const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t bits = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        // byte != 0  <=>  uint8 value > 0.5; per-byte "non-zero" flags -> 4 bits
        const uint32_t x = v[q];
        const uint32_t nz = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | x) & 0x80808080u;     // bit 7 of each non-zero byte
        bits |= (((nz >> 7) & 1u) | ((nz >> 14) & 2u) | ((nz >> 21) & 4u) | ((nz >> 28) & 8u)) << (4 * q);
      }
      s_bits[w] = bits;
    }
  } else {
    // lane = pixel: one coalesced 128 B load per word, 8 words in flight per warp
    const float* msk = reinterpret_cast<const float*>(mask_v) + (f * 2 + mch) * (int64_t)npx;
    constexpr int U = 8;
    for (int wb = warp * U; wb < nwords; wb += D2P_WARPS * U) {
      float m[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int pix = (wb + q) * 32 + lane;
        m[q] = pix < npx ? __ldg(msk + pix) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const unsigned b = __ballot_sync(0xffffffffu, m[q] > 0.5f);
        if (lane == 0 && wb + q < nwords) s_bits[wb + q] = b;
      }
    }
  }
  __syncthreads();

  // ordered list of the non-empty words (each thread owns `wpt` consecutive words)
  const int wpt = (nwords + D2P_THREADS - 1) / D2P_THREADS;
  const int wbeg = min(nwords, tid * wpt), wend = min(nwords, wbeg + wpt);
  int n_list;
  {
    int c = 0;
    for (int w = wbeg; w < wend; ++w) c += s_bits[w] != 0u;
    int r = block_exclusive_scan(c, s_warp, n_list);
    for (int w = wbeg; w < wend; ++w)
      if (s_bits[w] != 0u) s_wlist[r++] = (uint32_t)w;
  }
  __syncthreads();

  // pixel coordinates of lane `lane` of word w without a per-pixel division when W % 32 == 0
  const int wpr = (W & 31) == 0 ? (W >> 5) : 0;
  auto word_uv = [&](int w0, float& u, float& v) {
    if (wpr) { const int row = w0 / wpr; u = (float)(((w0 - row * wpr) << 5) + lane); v = (float)row; }
    else { const int pix = w0 * 32 + lane; u = (float)(pix % W); v = (float)(pix / W); }
  };
  auto z_uv = [&](float u, float v, float zm) -> float { return __fmul_rn(fmaf(k21, v, fmaf(k20, u, k22)), zm); };

  // ---- pass A: keep = mask & (0.2 < d < 2.5) (:392-395); mean z over non-zero pixels (:407) ----
  constexpr int UA = 4;                            // words (independent depth loads) in flight per warp
  double sum = 0.0;
  int cnt = 0;
  for (int li = warp * UA; li < n_list; li += D2P_WARPS * UA) {
    int wq[UA];
    float d[UA];
#pragma unroll
    for (int q = 0; q < UA; ++q) {
      wq[q] = li + q < n_list ? (int)s_wlist[li + q] : -1;
      const int pix = wq[q] * 32 + lane;
      d[q] = (wq[q] >= 0 && pix < npx) ? __ldg(dep + pix) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < UA; ++q) {
      if (wq[q] >= 0) {                              // warp-uniform
        const bool keep = ((s_bits[wq[q]] >> lane) & 1u) && (0.2f < d[q]) && (2.5f > d[q]);
        float u, v;
        word_uv(wq[q], u, v);
        const float z = z_uv(u, v, keep ? d[q] : 0.f);
        if (z != 0.f) { sum += (double)z; ++cnt; }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (lane == 0) s_bits[wq[q]] = b;
      }
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
    double s = lane < D2P_WARPS ? s_dsum[lane] : 0.0;
    int c = lane < D2P_WARPS ? s_warp[lane] : 0;
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
    }
  }
  __syncthreads();
  const int n_nonzero = s_misc[0];
  const float lo = s_lohi[0], hi = s_lohi[1];

  // ---- pass B: candidate bitmask (:409): lo < z < hi ----
  int my_cand = 0;
  for (int li = warp * UA; li < n_list; li += D2P_WARPS * UA) {
    int wq[UA];
    float d[UA];
#pragma unroll
    for (int q = 0; q < UA; ++q) {
      wq[q] = li + q < n_list ? (int)s_wlist[li + q] : -1;
      const int pix = wq[q] * 32 + lane;
      d[q] = (wq[q] >= 0 && pix < npx) ? __ldg(dep + pix) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < UA; ++q) {
      if (wq[q] >= 0) {
        const bool keep = (s_bits[wq[q]] >> lane) & 1u;
        float u, v;
        word_uv(wq[q], u, v);
        const float z = z_uv(u, v, keep ? d[q] : 0.f);
        const bool c = n_nonzero > 0 && (z > lo) && (z < hi);
        const unsigned b = __ballot_sync(0xffffffffu, c);
        __syncwarp();
        if (lane == 0) { s_bits[wq[q]] = b; my_cand += __popc(b); }
      }
    }
  }
  int n_cand;
  (void)block_exclusive_scan(my_cand, s_warp, n_cand);
  if (tid == 0) n_cand_out[f * 2 + hand] = n_cand;

  // list entries owned by this thread for the ordered enumerations below
  const int lpt = (n_list + D2P_THREADS - 1) / D2P_THREADS;
  const int lbeg = min(n_list, tid * lpt), lend = min(n_list, lbeg + lpt);

  int n_sel = 0;                                   // entries valid in s_sel
  if (n_cand >= min_pixels && n_cand <= n_points) {
    // keep every candidate in pixel order (:424 pads by wrapping)
    int c = 0;
    for (int l = lbeg; l < lend; ++l) c += __popc(s_bits[s_wlist[l]]);
    int tot;
    int r = block_exclusive_scan(c, s_warp, tot);
    for (int l = lbeg; l < lend; ++l) {
      const int w = (int)s_wlist[l];
      unsigned b = s_bits[w];
      while (b) { const int bit = __ffs(b) - 1; b &= b - 1; s_sel[r++] = w * 32 + bit; }
    }
    n_sel = n_cand;
  } else if (n_cand > n_points) {
    // random subset (:418-422): the n_points candidates with the smallest keys (ties -> lower pixel).
    // Radix descent over the 32-bit keys, 8 bits per pass; one word per warp iteration (lane = pixel).
    const int32_t* keys = subset_keys ? subset_keys + (f * 2 + hand) * (int64_t)npx : nullptr;
    auto key_of = [&](int w0) -> uint32_t {
      const int pix = w0 * 32 + lane;
      return keys ? ((uint32_t)__ldg(keys + pix) ^ 0x80000000u)      // signed order -> unsigned
                  : d2p_key(seed, cloud_id, (uint32_t)pix);
    };
    constexpr int U = 4;
    uint32_t prefix = 0;
    int remaining = n_points;                      // rank (1-based) of the threshold inside the prefix bucket
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const uint32_t hmask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int li = warp * U; li < n_list; li += D2P_WARPS * U) {
        uint32_t kx[U];
        bool on[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int w0 = li + q < n_list ? (int)s_wlist[li + q] : -1;
          on[q] = w0 >= 0 && ((s_bits[w0] >> lane) & 1u);
          kx[q] = on[q] ? key_of(w0) : 0u;
        }
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (on[q] && (kx[q] & hmask) == prefix) atomicAdd(&s_hist[(kx[q] >> shift) & 255], 1);
      }
      __syncthreads();
      if (warp == 0) {                             // find the bucket holding the `remaining`-th key: warp scan over 256 bins
        int h[8], hs = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { h[q] = s_hist[lane * 8 + q]; hs += h[q]; }
        int inc = hs;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int acc = inc - hs;                        // keys in lower bins
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
    for (int li = warp * U; li < n_list; li += D2P_WARPS * U) {
      uint32_t kx[U];
      bool on[U];
      int wq[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        wq[q] = li + q < n_list ? (int)s_wlist[li + q] : -1;
        on[q] = wq[q] >= 0 && ((s_bits[wq[q]] >> lane) & 1u);
        kx[q] = on[q] ? key_of(wq[q]) : 0u;
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
    for (int l = lbeg; l < lend; ++l) c_eq += __popc(s_eq[s_wlist[l]]);
    int tot;
    int er = block_exclusive_scan(c_eq, s_warp, tot);
    for (int l = lbeg; l < lend; ++l) {
      const int w = (int)s_wlist[l];
      unsigned bq = s_eq[w], add = 0;
      while (bq && er < remaining) { add |= bq & (0u - bq); bq &= bq - 1; ++er; }
      s_lt[w] |= add;
    }
    int c_sel = 0;
    for (int l = lbeg; l < lend; ++l) c_sel += __popc(s_lt[s_wlist[l]]);
    int r = block_exclusive_scan(c_sel, s_warp, tot);
    for (int l = lbeg; l < lend; ++l) {
      const int w = (int)s_wlist[l];
      unsigned bsel = s_lt[w];
      while (bsel) { const int bit = __ffs(bsel) - 1; bsel &= bsel - 1; if (r < 1024) s_sel[r] = w * 32 + bit; ++r; }
    }
    n_sel = n_points;
  }
  __syncthreads();

  // final order + back-projection of the kept pixels (:427-428).  Every selected pixel passed
  // mask & depth window, so its masked depth is the depth itself; an empty selection yields pixel 0 / zeros.
  const int32_t* pm = perm ? perm + (f * 2 + hand) * (int64_t)n_points : nullptr;
  for (int i = tid; i < n_points; i += D2P_THREADS) {
    const int src = pm ? pm[i] : (int)d2p_perm1024(seed, cloud_id, (uint32_t)i);
    int pix = 0;
    float zm = 0.f;
    if (n_sel > 0) { pix = s_sel[src % n_sel]; zm = __ldg(dep + pix); }
    else {                                          // reference: choose = zeros -> xyz of pixel 0 of the masked map
      const float d0 = dep[0];
      bool m0;
      if (MASK_U8) m0 = reinterpret_cast<const uint8_t*>(mask_v)[(f * 2 + mch) * (int64_t)npx] != 0;
      else m0 = reinterpret_cast<const float*>(mask_v)[(f * 2 + mch) * (int64_t)npx] > 0.5f;
      zm = (m0 && 0.2f < d0 && 2.5f > d0) ? d0 : 0.f;
    }
    const float u = (float)(pix % W), v = (float)(pix / W);
    ch_out[i] = pix;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      cl_out[i * 3 + r] = __fmul_rn(fmaf(Ki[r * 3 + 1], v, fmaf(Ki[r * 3], u, Ki[r * 3 + 2])), zm);
  }
}

static int launch_depth2pcl(const float* depth, const void* mask, int mask_u8, const float* Kinv, const float* valid,
                            const int32_t* subset_keys, const int32_t* perm, uint32_t seed, int64_t B, int H, int W,
                            int n_points, int min_pixels, int64_t* choose, float* cloud, int32_t* n_cand,
                            void* stream, const char* what) {
  PDF_REQUIRE(depth && mask && Kinv && valid && choose && cloud && n_cand, PDF_ERR_BAD_ARG, "%s: null pointer", what);
  PDF_REQUIRE(B >= 0 && H > 0 && W > 0 && min_pixels >= 1, PDF_ERR_BAD_ARG, "%s: bad size", what);
  PDF_REQUIRE(n_points == 1024, PDF_ERR_UNSUPPORTED, "%s: n_points must be 1024 (got %d)", what, n_points);
  PDF_REQUIRE((int64_t)H * W <= 640 * 640, PDF_ERR_UNSUPPORTED, "%s: frame larger than 640x640", what);
  PDF_REQUIRE(!mask_u8 || (((int64_t)H * W) % 32 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0),
              PDF_ERR_UNSUPPORTED, "%s: uint8 masks need H*W %% 32 == 0 and a 16-byte aligned pointer", what);
  if (B == 0) return PDF_OK;
  const size_t smem = (size_t)(((int64_t)H * W + 31) / 32) * 4 * 4;      // candidate / selected / tie bitmasks + word list
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(depth2pcl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    cudaFuncSetAttribute(depth2pcl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
  }
  if (mask_u8)
    depth2pcl_kernel<true><<<(unsigned)(B * 2), D2P_THREADS, smem, (cudaStream_t)stream>>>(
        depth, mask, Kinv, valid, subset_keys, perm, seed, H, W, n_points, min_pixels, choose, cloud, n_cand);
  else
    depth2pcl_kernel<false><<<(unsigned)(B * 2), D2P_THREADS, smem, (cudaStream_t)stream>>>(
        depth, mask, Kinv, valid, subset_keys, perm, seed, H, W, n_points, min_pixels, choose, cloud, n_cand);
  return check_launch(what);
}

}  // namespace pdf

extern "C" int pdf_depth2pcl(const float* depth, const float* mask, const float* Kinv, const float* valid,
                             const int32_t* subset_keys, const int32_t* perm, int64_t B, int H, int W,
                             int n_points, int min_pixels, int64_t* choose, float* cloud, int32_t* n_cand,
                             void* stream) {
  PDF_REQUIRE(subset_keys != nullptr || (int64_t)H * W <= n_points, PDF_ERR_BAD_ARG,
              "pdf_depth2pcl: subset_keys required when a hand can exceed n_points pixels");
  PDF_REQUIRE(perm != nullptr, PDF_ERR_BAD_ARG, "pdf_depth2pcl: perm required (use pdf_depth2pcl_seeded for generated randomness)");
  return pdf::launch_depth2pcl(depth, mask, 0, Kinv, valid, subset_keys, perm, 0u, B, H, W, n_points, min_pixels, choose,
                               cloud, n_cand, stream, "pdf_depth2pcl");
}

extern "C" int pdf_depth2pcl_seeded(const float* depth, const void* mask, int mask_is_u8, const float* Kinv,
                                    const float* valid, const int32_t* subset_keys, const int32_t* perm,
                                    uint32_t seed, int64_t B, int H, int W, int n_points, int min_pixels,
                                    int64_t* choose, float* cloud, int32_t* n_cand, void* stream) {
  return pdf::launch_depth2pcl(depth, mask, mask_is_u8, Kinv, valid, subset_keys, perm, seed, B, H, W, n_points,
                               min_pixels, choose, cloud, n_cand, stream, "pdf_depth2pcl_seeded");
}

// Host mirrors of the counter-based randomness, so callers (and the parity tests) can materialise exactly
// the keys / permutation the kernel uses: keys[cloud][pix] as int32 in the SIGNED order pdf_depth2pcl
// expects, perm[cloud][i].
extern "C" int pdf_depth2pcl_host_randomness(uint32_t seed, int64_t n_clouds, int64_t npx, int32_t* keys_out,
                                             int32_t* perm_out) {
  PDF_REQUIRE(n_clouds >= 0 && npx >= 0, PDF_ERR_BAD_ARG, "pdf_depth2pcl_host_randomness: bad size");
  for (int64_t c = 0; c < n_clouds; ++c) {
    if (keys_out)
      for (int64_t p = 0; p < npx; ++p)
        keys_out[c * npx + p] = (int32_t)(pdf::d2p_key(seed, (uint32_t)c, (uint32_t)p) ^ 0x80000000u);
    if (perm_out)
      for (int i = 0; i < 1024; ++i) perm_out[c * 1024 + i] = (int32_t)pdf::d2p_perm1024(seed, (uint32_t)c, (uint32_t)i);
  }
  return PDF_OK;
}
