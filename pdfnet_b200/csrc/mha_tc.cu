// Multi-head attention of the GCN decoder on tensor cores (self_attn.py:60-72, inter_attn.py:84-108):
// out = softmax(q k^T / sqrt(d)) v per (sample, head), V <= 256 tokens, d in {16, 32, 64}.
//
// One CTA per (problem, sample, head).  K and V^T of the head are staged in shared memory as bf16 hi / lo
// pairs; a warp owns 16 query rows and walks the keys in chunks of 64 with an online softmax (running max
// and sum, rescaled accumulators; scores in the log2 domain, 2^x on the SFU), so the 16 x V score tile never
// exists in full.  Both contractions run on
// mma.sync m16n8k16 (bf16 operands, fp32 accumulate) with SPLIT operands: x = hi + lo, products hi.hi + hi.lo
// + lo.hi, i.e. ~2^-16 relative per product - the decoder holds a 1e-4 parity bound over ~40 chained layers,
// which plain bf16 operands do not (measured in round 1).  The probability tile goes from the accumulator
// layout of Q.K^T straight into the A-operand layout of P.V (two adjacent 8-key tiles = one 16-key step), no
// shared-memory round trip.  The tiles are far too small (V <= 256, d <= 64) for a tcgen05 / TMEM pipeline to
// pay off: one CTA's whole job is ~3 k MMAs.
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {

constexpr int MT_MAXV = 256;
constexpr int MT_PROBLEMS = 2;

struct MhaProblem { const float* q; const float* k; const float* v; float* out; uint8_t* out_img; int64_t img_row0; };
struct MhaParams {
  MhaProblem p[MT_PROBLEMS];
  int64_t ldq, ldk, ldv, ldo;
  int V, heads, n_samples;
  float inv_norm;
  int img_kb;                    // k-blocks per part of the split output image (= heads * d / 64)
};

__device__ __forceinline__ uint32_t mt_pack(float lo, float hi) {            // bf16x2, element 0 in the low half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// split two floats into packed bf16 hi parts and packed bf16 residuals
__device__ __forceinline__ void mt_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = mt_pack(a, b);
  lo = mt_pack(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// 2^x on the SFU (ex2.approx, max relative error 2^-22; -inf / very negative -> 0)
__device__ __forceinline__ float mt_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mt_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int D>
__global__ void __launch_bounds__(256)
mha_tc_kernel(const MhaParams P) {
  extern __shared__ __align__(16) uint8_t mt_smem[];
  constexpr int KP = D + 8;                       // K row pitch (bf16): 16 B of padding -> conflict-free fragment loads
  const int V = P.V, Vp = (V + 15) & ~15;         // keys padded to whole 16-key steps
  const int VP = Vp + 8;                          // V^T row pitch (bf16)
  uint16_t* sKh = reinterpret_cast<uint16_t*>(mt_smem);      // [Vp][KP]
  uint16_t* sKl = sKh + (size_t)Vp * KP;
  uint16_t* sVh = sKl + (size_t)Vp * KP;                     // [D][VP]  (V transposed: key index contiguous)
  uint16_t* sVl = sVh + (size_t)D * VP;

  pdl_wait();
  pdl_trigger();
  const int prob = blockIdx.y;
  const int smp = blockIdx.x / P.heads, h = blockIdx.x - smp * P.heads;
  const MhaProblem pb = P.p[prob];
  const int64_t row0 = (int64_t)smp * V;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  // ---- stage K (row-major) and V^T as split bf16; padded keys are zero ----
  for (int e = tid; e < Vp * (D / 2); e += blockDim.x) {
    const int j = e / (D / 2), c = (e - j * (D / 2)) * 2;
    float2 kv = make_float2(0.f, 0.f);
    if (j < V) kv = *reinterpret_cast<const float2*>(pb.k + (row0 + j) * P.ldk + h * D + c);
    uint32_t hi, lo;
    mt_split(kv.x, kv.y, hi, lo);
    *reinterpret_cast<uint32_t*>(sKh + (size_t)j * KP + c) = hi;
    *reinterpret_cast<uint32_t*>(sKl + (size_t)j * KP + c) = lo;
  }
  for (int e = tid; e < (Vp / 2) * D; e += blockDim.x) {
    const int c = e % D, j = (e / D) * 2;          // consecutive threads = consecutive channels (coalesced)
    const float v0 = j < V ? pb.v[(row0 + j) * P.ldv + h * D + c] : 0.f;
    const float v1 = j + 1 < V ? pb.v[(row0 + j + 1) * P.ldv + h * D + c] : 0.f;
    uint32_t hi, lo;
    mt_split(v0, v1, hi, lo);
    *reinterpret_cast<uint32_t*>(sVh + (size_t)c * VP + j) = hi;
    *reinterpret_cast<uint32_t*>(sVl + (size_t)c * VP + j) = lo;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;          // mma fragment coordinates
  const float qscale = P.inv_norm * 1.4426950408889634f;
  for (int qt = warp; qt * 16 < V; qt += nwarps) {
    const int r0 = qt * 16 + g, r1 = r0 + 8;      // this thread's two query rows
    // ---- Q fragments (pre-scaled by 1/sqrt(d)), split ----
    uint32_t qh[D / 16][4], ql[D / 16][4];
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = ks * 16 + half * 8 + 2 * t;
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        if (r0 < V) a = *reinterpret_cast<const float2*>(pb.q + (row0 + r0) * P.ldq + h * D + c);
        if (r1 < V) b = *reinterpret_cast<const float2*>(pb.q + (row0 + r1) * P.ldq + h * D + c);
        // scores are kept in the log2 domain: q * (1/sqrt(d)) * log2(e), softmax via 2^x
        mt_split(a.x * qscale, a.y * qscale, qh[ks][half * 2], ql[ks][half * 2]);
        mt_split(b.x * qscale, b.y * qscale, qh[ks][half * 2 + 1], ql[ks][half * 2 + 1]);
      }
    }
    float o[D / 8][4];
#pragma unroll
    for (int n = 0; n < D / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m0 = -3.0e38f, m1 = -3.0e38f, l0 = 0.f, l1 = 0.f;

    for (int k0 = 0; k0 < Vp; k0 += 64) {
      const int nt = min(8, (Vp - k0) >> 3);      // 8-key tiles in this chunk (Vp % 16 == 0 -> even)
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        if (n < nt) {
          const int key = k0 + n * 8 + g;          // B fragment: column n = key, k index = channel pair
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(sKh + (size_t)key * KP + ks * 16 + 2 * t);
            const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(sKh + (size_t)key * KP + ks * 16 + 8 + 2 * t);
            const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(sKl + (size_t)key * KP + ks * 16 + 2 * t);
            const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(sKl + (size_t)key * KP + ks * 16 + 8 + 2 * t);
            mt_mma(s[n], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], bh0, bh1);
            mt_mma(s[n], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bl0, bl1);
            mt_mma(s[n], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bh0, bh1);
          }
        }
      }
      // ---- online softmax: rows r0 (s[.][0..1]) and r1 (s[.][2..3]); columns 2t, 2t+1 of every tile ----
      float cm0 = -3.0e38f, cm1 = -3.0e38f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (n < nt) {
          const int key = k0 + n * 8 + 2 * t;
          if (key >= V) s[n][0] = s[n][2] = -3.0e38f;
          if (key + 1 >= V) s[n][1] = s[n][3] = -3.0e38f;
          cm0 = fmaxf(cm0, fmaxf(s[n][0], s[n][1]));
          cm1 = fmaxf(cm1, fmaxf(s[n][2], s[n][3]));
        }
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);
      const float sc0 = mt_exp2(m0 - nm0), sc1 = mt_exp2(m1 - nm1);
      m0 = nm0; m1 = nm1;
      l0 *= sc0; l1 *= sc1;
#pragma unroll
      for (int n = 0; n < D / 8; ++n) { o[n][0] *= sc0; o[n][1] *= sc0; o[n][2] *= sc1; o[n][3] *= sc1; }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (n < nt) {
          s[n][0] = mt_exp2(s[n][0] - m0); s[n][1] = mt_exp2(s[n][1] - m0);
          s[n][2] = mt_exp2(s[n][2] - m1); s[n][3] = mt_exp2(s[n][3] - m1);
          l0 += s[n][0] + s[n][1];
          l1 += s[n][2] + s[n][3];
        }
      }
      // ---- O += P . V : two adjacent 8-key tiles form one 16-key step of the A operand ----
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (2 * kk < nt) {
          uint32_t ph[4], pl[4];
          mt_split(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
          mt_split(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
          mt_split(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
          mt_split(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
          const int key = k0 + kk * 16 + 2 * t;    // B fragment: k index = key pair, column n = channel
#pragma unroll
          for (int n = 0; n < D / 8; ++n) {
            const int c = n * 8 + g;
            const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(sVh + (size_t)c * VP + key);
            const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(sVh + (size_t)c * VP + key + 8);
            const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(sVl + (size_t)c * VP + key);
            const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(sVl + (size_t)c * VP + key + 8);
            mt_mma(o[n], pl[0], pl[1], pl[2], pl[3], vh0, vh1);
            mt_mma(o[n], ph[0], ph[1], ph[2], ph[3], vl0, vl1);
            mt_mma(o[n], ph[0], ph[1], ph[2], ph[3], vh0, vh1);
          }
        }
      }
    }
    // row sums live in the 4 threads of a quad
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int n = 0; n < D / 8; ++n) {
      const int c = h * D + n * 8 + 2 * t;
      const float a0 = o[n][0] * i0, a1 = o[n][1] * i0, b0 = o[n][2] * i1, b1 = o[n][3] * i1;
      if (pb.out != nullptr) {
        if (r0 < V) *reinterpret_cast<float2*>(pb.out + (row0 + r0) * P.ldo + c) = make_float2(a0, a1);
        if (r1 < V) *reinterpret_cast<float2*>(pb.out + (row0 + r1) * P.ldo + c) = make_float2(b0, b1);
      }
      if (pb.out_img != nullptr) {
        // the output only feeds the `fc` GEMM: written as that GEMM's split-bf16 operand image [hi | hi | lo]
        const size_t part = (size_t)P.img_kb * 16384;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int64_t row = pb.img_row0 + row0 + (rr ? r1 : r0);
          if ((rr ? r1 : r0) < V) {
            uint32_t hi, lo;
            mt_split(rr ? b0 : a0, rr ? b1 : a1, hi, lo);
            uint8_t* blk = pb.out_img + ((size_t)(row >> 7) * 3 * P.img_kb + (c >> 6)) * 16384 +
                           umma::sw128_off((uint32_t)(row & 127), (uint32_t)(c & 63));
            *reinterpret_cast<uint32_t*>(blk) = hi;
            *reinterpret_cast<uint32_t*>(blk + part) = hi;
            *reinterpret_cast<uint32_t*>(blk + 2 * part) = lo;
          }
        }
      }
    }
  }
}

template <int D>
static void mha_tc_launch(const MhaParams& P, int n_problems, cudaStream_t s) {
  const int Vp = (P.V + 15) & ~15;
  const size_t smem = (size_t)2 * Vp * (D + 8) * 2 + (size_t)2 * D * (Vp + 8) * 2;
  static PerDeviceOnce once;
  if (once.first()) {
    const int mx = 2 * MT_MAXV * (64 + 8) * 2 + 2 * 64 * (MT_MAXV + 8) * 2;
    cudaFuncSetAttribute(mha_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
  }
  const int tiles = (P.V + 15) / 16;
  const int threads = 32 * (tiles < 8 ? (tiles < 4 ? 4 : tiles) : 8);
  launch_pdl(mha_tc_kernel<D>, dim3((unsigned)(P.n_samples * P.heads), (unsigned)n_problems), dim3(threads), smem, s, P);
}

}  // namespace pdf

// Up to two attention problems of identical shape in ONE launch (the two self-attentions, or the two cross
// directions R2L / L2R of inter_attn.py:84-108): problem i reads q[i], k[i], v[i] and writes out[i].
extern "C" int pdf_mha_tc(const float* const* q, const float* const* k, const float* const* v, float* const* out,
                          void* const* out_img, const int64_t* img_row0, int n_problems, int64_t ldq, int64_t ldk,
                          int64_t ldv, int64_t ldo, int64_t n_samples, int V, int heads, int d, void* stream) {
  using namespace pdf;
  if (n_samples == 0 || n_problems == 0) return PDF_OK;
  PDF_REQUIRE(q && k && v && (out || out_img) && n_problems >= 1 && n_problems <= MT_PROBLEMS, PDF_ERR_BAD_ARG,
              "pdf_mha_tc: null pointer / 1..%d problems", MT_PROBLEMS);
  PDF_REQUIRE(!out_img || (heads * d) % 64 == 0, PDF_ERR_BAD_ARG, "pdf_mha_tc: image output needs heads*d %% 64 == 0");
  PDF_REQUIRE(n_samples > 0 && V > 0 && V <= MT_MAXV && heads > 0 && (d == 16 || d == 32 || d == 64) &&
                  n_samples * heads < (1ll << 31),
              PDF_ERR_UNSUPPORTED, "pdf_mha_tc: supports <= 256 tokens and head dim 16 / 32 / 64");
  PDF_REQUIRE(ldq % 2 == 0 && ldk % 2 == 0 && ldo % 2 == 0, PDF_ERR_BAD_ARG, "pdf_mha_tc: row pitches must be even (8-byte vector access)");
  MhaParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < n_problems; ++i) {
    float* o = out ? out[i] : nullptr;
    uint8_t* oi = out_img ? (uint8_t*)out_img[i] : nullptr;
    PDF_REQUIRE(q[i] && k[i] && v[i] && (o || oi), PDF_ERR_BAD_ARG, "pdf_mha_tc: null problem pointer");
    PDF_REQUIRE(((reinterpret_cast<uintptr_t>(q[i]) | reinterpret_cast<uintptr_t>(k[i]) | reinterpret_cast<uintptr_t>(o)) & 7) == 0,
                PDF_ERR_BAD_ARG, "pdf_mha_tc: q / k / out must be 8-byte aligned");
    P.p[i] = MhaProblem{q[i], k[i], v[i], o, oi, img_row0 ? img_row0[i] : 0};
  }
  P.ldq = ldq; P.ldk = ldk; P.ldv = ldv; P.ldo = ldo;
  P.V = V; P.heads = heads; P.n_samples = (int)n_samples;
  P.inv_norm = 1.f / sqrtf((float)d);
  P.img_kb = heads * d / 64;
  cudaStream_t s = (cudaStream_t)stream;
  if (d == 16) mha_tc_launch<16>(P, n_problems, s);
  else if (d == 32) mha_tc_launch<32>(P, n_problems, s);
  else mha_tc_launch<64>(P, n_problems, s);
  return check_launch("pdf_mha_tc");
}
