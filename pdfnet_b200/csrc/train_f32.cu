// Training-mode (cfg5) kernels of the fusion path, fp32: train-mode BatchNorm over rows,
// the weight-gradient GEMM (reduction over rows), max-pool / grouping / pixel-gather
// backward scatters and the SFT modulation backward.  The data-gradient GEMM reuses
// pdf_linear_f32 with the transposed weight.  Everything is row-major [rows, channels] with a
// free row pitch, the layout the inference path already uses.
//
// Reference semantics: nn.BatchNorm2d(train) inside netR_1/2/3 (intaghand_encoder.py:52-99:
// biased variance for the normalisation, unbiased for running_var, momentum 0.1),
// nn.MaxPool2d backward (first maximum of the window receives the gradient), and autograd
// of group_points / group_points_2 (utils.py:134-191), _tranpose_and_gather_feat
// (models/utils.py:22-26) and SFTLayer.forward (intaghand_encoder.py:213-219).
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {

// ---------------------------------------------------------------------------------------------
// column reductions over rows: two sums per channel, accumulated in double with atomics.
// block (32 channels x 8 row lanes); each thread walks rows with a grid stride.
enum { RED_STATS = 0, RED_COLSUM = 1, RED_BN_BWD = 2 };

template <int V>
struct Vec {
  float v[V];
};
template <int V>
__device__ __forceinline__ Vec<V> vload(const float* p) {
  Vec<V> r;
  if (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = p[i];
  }
  return r;
}
template <int V>
__device__ __forceinline__ void vstore(float* p, const Vec<V>& r) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  else {
#pragma unroll
    for (int i = 0; i < V; ++i) p[i] = r.v[i];
  }
}

// 256 threads = LX channel-vectors (V channels each) x 256/LX row lanes; V = 4 needs C % 4 == 0,
// row pitches % 4 == 0 and 16-byte aligned bases (checked on the host).
template <int MODE, int V>
__global__ void __launch_bounds__(256)
col_reduce_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Yp, int64_t ldy,
                  const float* __restrict__ X, int64_t ldx, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  int relu, int64_t M, int C, int LX, double* __restrict__ sums) {
  const int cx = threadIdx.x % LX, ry = threadIdx.x / LX, RY = 256 / LX;
  const int c = (blockIdx.x * LX + cx) * V;
  const bool ok = c < C;
  float s0[V], s1[V], mu[V], rs[V], ga[V], be[V];
  double d0[V], d1[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { s0[e] = s1[e] = 0.f; d0[e] = d1[e] = 0.0; mu[e] = rs[e] = ga[e] = be[e] = 0.f; }
  if (ok) {
    // statistics are accumulated about the channel's first row (shifted sums): the variance
    // sum d^2 - (sum d)^2 / M then has no catastrophic cancellation when |mean| >> std
    if (MODE == RED_STATS && M > 0) { const Vec<V> t = vload<V>(A + c);
#pragma unroll
      for (int e = 0; e < V; ++e) mu[e] = t.v[e]; }
    if (MODE == RED_BN_BWD) {
      const Vec<V> t = vload<V>(mean + c), u = vload<V>(rstd + c);
#pragma unroll
      for (int e = 0; e < V; ++e) { mu[e] = t.v[e]; rs[e] = u.v[e]; }
      if (!Yp && relu) {                             // ReLU mask recomputed from x instead of read from Y
        const Vec<V> g2 = vload<V>(gamma + c), b2 = vload<V>(beta + c);
#pragma unroll
        for (int e = 0; e < V; ++e) { ga[e] = g2.v[e]; be[e] = b2.v[e]; }
      }
    }
  }
  // two-level accumulation keeps fp32 partial sums short (<= 64 terms) before going to double
  int n = 0;
  const int64_t stride = (int64_t)gridDim.y * RY;
  if (ok) {
    for (int64_t r = (int64_t)blockIdx.y * RY + ry; r < M; r += stride) {
      const Vec<V> a = vload<V>(A + r * lda + c);
      if (MODE == RED_STATS) {
#pragma unroll
        for (int e = 0; e < V; ++e) { const float d = a.v[e] - mu[e]; s0[e] += d; s1[e] = fmaf(d, d, s1[e]); }
      } else if (MODE == RED_COLSUM) {
#pragma unroll
        for (int e = 0; e < V; ++e) s0[e] += a.v[e];
      } else {
        const Vec<V> x = vload<V>(X + r * ldx + c);
        Vec<V> y;
        if (Yp) y = vload<V>(Yp + r * ldy + c);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float xhat = (x.v[e] - mu[e]) * rs[e];
          // same expression as the forward pass (pdf_bn_act_fwd), so the recomputed mask is bit-identical
          const bool on = !relu || (Yp ? y.v[e] > 0.f : fmaf(xhat, ga[e], be[e]) > 0.f);
          const float g = on ? a.v[e] : 0.f;
          s0[e] += g;
          s1[e] = fmaf(g, xhat, s1[e]);
        }
      }
      if (++n == 64) {
#pragma unroll
        for (int e = 0; e < V; ++e) { d0[e] += s0[e]; d1[e] += s1[e]; s0[e] = s1[e] = 0.f; }
        n = 0;
      }
    }
  }
  __shared__ double sh[2][256][V];
#pragma unroll
  for (int e = 0; e < V; ++e) { sh[0][threadIdx.x][e] = d0[e] + s0[e]; sh[1][threadIdx.x][e] = d1[e] + s1[e]; }
  __syncthreads();
  if (ry == 0 && ok) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (c + e >= C) break;
      double t0 = 0.0, t1 = 0.0;
      for (int i = 0; i < RY; ++i) { t0 += sh[0][i * LX + cx][e]; t1 += sh[1][i * LX + cx][e]; }
      atomicAdd(&sums[c + e], t0);
      if (MODE != RED_COLSUM) atomicAdd(&sums[C + c + e], t1);
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ X, int64_t M, int C,
                                   float eps, float momentum, float* running_mean, float* running_var,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double md = sums[c] / (double)M;                 // mean of (x - x[0,c])
  const double m = md + (double)X[c];
  double var = sums[C + c] / (double)M - md * md;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------------------------------------
// element-wise passes over [M, C] row-major matrices (grid-stride, channel fastest)
enum { EW_BN_FWD = 0, EW_BN_BWD = 1, EW_ACT_BWD = 2, EW_SFT_FWD = 3, EW_SFT_BWD = 4 };

struct EwArgs {
  const float* a; int64_t lda;      // BN_FWD: X      BN_BWD: dY    ACT_BWD: dY    SFT_FWD: fea    SFT_BWD: dout
  const float* b; int64_t ldb;      // BN_BWD: Y      ACT_BWD: Y    SFT_FWD: scale SFT_BWD: fea
  const float* c; int64_t ldc;      // BN_BWD: X      SFT_FWD: shift SFT_BWD: scale
  const float* v0; const float* v1; const float* v2; const float* v3;   // per-channel vectors
  const double* sums;               // BN_BWD: [sum g | sum g*xhat]
  float* o0; int64_t ldo0;          // primary output
  float* o1; int64_t ldo1;          // SFT_BWD: dscale
  int64_t M; int C; int flag;       // flag: relu (BN) / activation enum (ACT_BWD)
};

// V = 4: C and every row pitch are multiples of 4 and all pointers 16-byte aligned (checked on the host)
template <int MODE, int V>
__global__ void __launch_bounds__(256) elementwise_kernel(EwArgs p) {
  const int cv = p.C / V;
  const int64_t total = p.M * (int64_t)cv;
  const float invM = 1.f / (float)p.M;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv;
    const int c = (int)(i - r * cv) * V;
    Vec<V> o;
    if (MODE == EW_BN_FWD) {            // v0 mean, v1 rstd, v2 gamma, v3 beta
      const Vec<V> x = vload<V>(p.a + r * p.lda + c), mu = vload<V>(p.v0 + c), rs = vload<V>(p.v1 + c),
                   ga = vload<V>(p.v2 + c), be = vload<V>(p.v3 + c);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float y = fmaf((x.v[e] - mu.v[e]) * rs.v[e], ga.v[e], be.v[e]);
        o.v[e] = p.flag ? fmaxf(y, 0.f) : y;
      }
    } else if (MODE == EW_BN_BWD) {     // v0 mean, v1 rstd, v2 gamma
      const Vec<V> dy = vload<V>(p.a + r * p.lda + c), x = vload<V>(p.c + r * p.ldc + c), mu = vload<V>(p.v0 + c),
                   rs = vload<V>(p.v1 + c), ga = vload<V>(p.v2 + c);
      Vec<V> y, be;
      if (p.b) y = vload<V>(p.b + r * p.ldb + c);
      else if (p.flag) be = vload<V>(p.v3 + c);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float xhat = (x.v[e] - mu.v[e]) * rs.v[e];
        const bool on = !p.flag || (p.b ? y.v[e] > 0.f : fmaf(xhat, ga.v[e], be.v[e]) > 0.f);
        const float g = on ? dy.v[e] : 0.f;
        const float dbeta = (float)p.sums[c + e], dgamma = (float)p.sums[p.C + c + e];
        o.v[e] = ga.v[e] * rs.v[e] * (g - dbeta * invM - xhat * dgamma * invM);
      }
    } else if (MODE == EW_ACT_BWD) {
      const Vec<V> d = vload<V>(p.a + r * p.lda + c), y = vload<V>(p.b + r * p.ldb + c);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float s = p.flag == PDF_ACT_RELU ? (y.v[e] > 0.f ? 1.f : 0.f)
                                               : (p.flag == PDF_ACT_LEAKY01 ? (y.v[e] > 0.f ? 1.f : 0.1f) : 1.f);
        o.v[e] = d.v[e] * s;
      }
    } else if (MODE == EW_SFT_FWD) {    // fea * (scale + 1) + shift, rounded like the reference (mul, then add)
      const Vec<V> f = vload<V>(p.a + r * p.lda + c), sc = vload<V>(p.b + r * p.ldb + c),
                   sh = vload<V>(p.c + r * p.ldc + c);
#pragma unroll
      for (int e = 0; e < V; ++e) o.v[e] = __fadd_rn(__fmul_rn(f.v[e], __fadd_rn(sc.v[e], 1.f)), sh.v[e]);
    } else {                            // SFT_BWD: dfea = dout*(scale+1), dscale = dout*fea  (dshift = dout)
      const Vec<V> d = vload<V>(p.a + r * p.lda + c), f = vload<V>(p.b + r * p.ldb + c),
                   sc = vload<V>(p.c + r * p.ldc + c);
      Vec<V> o1;
#pragma unroll
      for (int e = 0; e < V; ++e) { o1.v[e] = d.v[e] * f.v[e]; o.v[e] = d.v[e] * (sc.v[e] + 1.f); }
      vstore<V>(p.o1 + r * p.ldo1 + c, o1);
    }
    vstore<V>(p.o0 + r * p.ldo0 + c, o);
  }
}


// BatchNorm(+ReLU) backward apply writing dX directly as the split-bf16 tile image ([hi|hi|lo]) that the
// data- and weight-gradient GEMMs read (pdf_gemm_bf16 / pdf_gemm_tn_bf16): one thread = 8 channels of a
// row = one 16-byte image chunk per part.  Rows >= M of the last row tile are written as zeros (the
// weight gradient reduces over them).  C % 64 == 0, 16-byte aligned rows.
__global__ void __launch_bounds__(256)
bn_bwd_image_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ X, int64_t ldx,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const double* __restrict__ sums, int relu, int64_t M, int C,
                    uint8_t* __restrict__ img, const float* __restrict__ dOut, int64_t lddo,
                    const uint8_t* __restrict__ arg, int G, int plain) {
  // thread = one 8-channel chunk (fixed for the thread's lifetime, its coefficients live in registers) of a run
  // of rows: 256 threads = (C/8 chunks) x (256 / (C/8) rows per pass)
  const int cq = C >> 3, nkb = C >> 6;
  const int q = threadIdx.x % cq, rl = threadIdx.x / cq, rows_cta = 256 / cq;
  if (rl >= rows_cta) return;
  const int c = q * 8;
  const float invM = 1.f / (float)M;
  float mu[8], rs[8], ga[8], be[8], k0[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = mean[c + e]; rs[e] = rstd[c + e]; ga[e] = gamma[c + e]; be[e] = beta[c + e];
    k0[e] = ga[e] * rs[e];
    k1[e] = (float)sums[c + e] * invM;               // dbeta / M
    k2[e] = (float)sums[C + c + e] * invM;           // dgamma / M
  }
  const int64_t rows_pad = ((M + 127) >> 7) << 7;
  const int kb = q >> 3;
  const uint32_t kcol = (uint32_t)(q & 7) * 8;
  for (int64_t r = (int64_t)blockIdx.x * rows_cta + rl; r < rows_pad; r += (int64_t)gridDim.x * rows_cta) {
    float d[8];
    if (r < M) {
      float dy[8];
      if (arg) {                                     // gradient of the max-pool: only the argmax row of a group is non-zero
        const int64_t g = r / G;
        const int rg = (int)(r - g * G);
        const uint2 aw = *reinterpret_cast<const uint2*>(arg + g * C + c);
        const float4 o0 = *reinterpret_cast<const float4*>(dOut + g * lddo + c), o1 = *reinterpret_cast<const float4*>(dOut + g * lddo + c + 4);
        const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int a = (int)(((e < 4 ? aw.x : aw.y) >> (8 * (e & 3))) & 0xffu);
          dy[e] = a == rg ? ov[e] : 0.f;
        }
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(dY + r * lddy + c), a1 = *reinterpret_cast<const float4*>(dY + r * lddy + c + 4);
        dy[0] = a0.x; dy[1] = a0.y; dy[2] = a0.z; dy[3] = a0.w; dy[4] = a1.x; dy[5] = a1.y; dy[6] = a1.z; dy[7] = a1.w;
      }
      const float4 x0 = *reinterpret_cast<const float4*>(X + r * ldx + c), x1 = *reinterpret_cast<const float4*>(X + r * ldx + c + 4);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xhat = (xv[e] - mu[e]) * rs[e];
        const bool on = !relu || fmaf(xhat, ga[e], be[e]) > 0.f;
        const float g = on ? dy[e] : 0.f;
        d[e] = k0[e] * (g - k1[e] - xhat * k2[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = 0.f;
    }
    uint4 w, wl;
    w.x = umma::pack_bf16(d[0], d[1]); w.y = umma::pack_bf16(d[2], d[3]);
    w.z = umma::pack_bf16(d[4], d[5]); w.w = umma::pack_bf16(d[6], d[7]);
    wl.x = umma::pack_bf16(d[0] - __uint_as_float(w.x << 16), d[1] - __uint_as_float(w.x & 0xffff0000u));
    wl.y = umma::pack_bf16(d[2] - __uint_as_float(w.y << 16), d[3] - __uint_as_float(w.y & 0xffff0000u));
    wl.z = umma::pack_bf16(d[4] - __uint_as_float(w.z << 16), d[5] - __uint_as_float(w.z & 0xffff0000u));
    wl.w = umma::pack_bf16(d[6] - __uint_as_float(w.w << 16), d[7] - __uint_as_float(w.w & 0xffff0000u));
    uint8_t* base = img + (size_t)(r >> 7) * (size_t)((plain ? 1 : 3) * nkb) * 16384;
    const uint32_t off = umma::sw128_off((uint32_t)(r & 127), kcol);
    *reinterpret_cast<uint4*>(base + (size_t)kb * 16384 + off) = w;
    if (!plain) {                                      // plain: one bf16 image (bf16 training mode), else [hi | hi | lo]
      *reinterpret_cast<uint4*>(base + (size_t)(nkb + kb) * 16384 + off) = w;
      *reinterpret_cast<uint4*>(base + (size_t)(2 * nkb + kb) * 16384 + off) = wl;
    }
  }
}

// BatchNorm(+ReLU) forward apply writing Y directly as the split-bf16 tile image of the NEXT layer's GEMM (and,
// optionally, as fp32 rows): thread = one 8-channel chunk of a run of rows, zero pad rows up to the tile edge.
__global__ void __launch_bounds__(256)
bn_fwd_image_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int relu, int64_t M, int C, float* __restrict__ Y, int64_t ldy, uint8_t* __restrict__ img, int plain) {
  const int cq = C >> 3, nkb = C >> 6;
  const int q = threadIdx.x % cq, rl = threadIdx.x / cq, rows_cta = 256 / cq;
  if (rl >= rows_cta) return;
  const int c = q * 8;
  float mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { mu[e] = mean[c + e]; rs[e] = rstd[c + e]; ga[e] = gamma[c + e]; be[e] = beta[c + e]; }
  const int64_t rows_pad = ((M + 127) >> 7) << 7;
  const int kb = q >> 3;
  const uint32_t kcol = (uint32_t)(q & 7) * 8;
  for (int64_t r = (int64_t)blockIdx.x * rows_cta + rl; r < rows_pad; r += (int64_t)gridDim.x * rows_cta) {
    float d[8];
    if (r < M) {
      const float4 x0 = *reinterpret_cast<const float4*>(X + r * ldx + c), x1 = *reinterpret_cast<const float4*>(X + r * ldx + c + 4);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = fmaf((xv[e] - mu[e]) * rs[e], ga[e], be[e]);       // same expression as elementwise BN_FWD
        d[e] = relu ? fmaxf(y, 0.f) : y;
      }
      if (Y) {
        *reinterpret_cast<float4*>(Y + r * ldy + c) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<float4*>(Y + r * ldy + c + 4) = make_float4(d[4], d[5], d[6], d[7]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = 0.f;
    }
    uint4 w, wl;
    w.x = umma::pack_bf16(d[0], d[1]); w.y = umma::pack_bf16(d[2], d[3]);
    w.z = umma::pack_bf16(d[4], d[5]); w.w = umma::pack_bf16(d[6], d[7]);
    wl.x = umma::pack_bf16(d[0] - __uint_as_float(w.x << 16), d[1] - __uint_as_float(w.x & 0xffff0000u));
    wl.y = umma::pack_bf16(d[2] - __uint_as_float(w.y << 16), d[3] - __uint_as_float(w.y & 0xffff0000u));
    wl.z = umma::pack_bf16(d[4] - __uint_as_float(w.z << 16), d[5] - __uint_as_float(w.z & 0xffff0000u));
    wl.w = umma::pack_bf16(d[6] - __uint_as_float(w.w << 16), d[7] - __uint_as_float(w.w & 0xffff0000u));
    uint8_t* base = img + (size_t)(r >> 7) * (size_t)((plain ? 1 : 3) * nkb) * 16384;
    const uint32_t off = umma::sw128_off((uint32_t)(r & 127), kcol);
    *reinterpret_cast<uint4*>(base + (size_t)kb * 16384 + off) = w;
    if (!plain) {
      *reinterpret_cast<uint4*>(base + (size_t)(nkb + kb) * 16384 + off) = w;
      *reinterpret_cast<uint4*>(base + (size_t)(2 * nkb + kb) * 16384 + off) = wl;
    }
  }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

template <int MODE>
static int launch_ew(const EwArgs& p, cudaStream_t s, const char* what) {
  if (p.M * (int64_t)p.C == 0) return PDF_OK;
  const bool vec = p.C % 4 == 0 && p.lda % 4 == 0 && p.ldb % 4 == 0 && p.ldc % 4 == 0 && p.ldo0 % 4 == 0 &&
                   p.ldo1 % 4 == 0 && aligned16(p.a) && aligned16(p.b) && aligned16(p.c) && aligned16(p.o0) &&
                   aligned16(p.o1) && aligned16(p.v0) && aligned16(p.v1) && aligned16(p.v2) && aligned16(p.v3);
  const int64_t total = p.M * (int64_t)(vec ? p.C / 4 : p.C);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec) elementwise_kernel<MODE, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
  else elementwise_kernel<MODE, 1><<<(unsigned)blocks, 256, 0, s>>>(p);
  return check_launch(what);
}

// ---------------------------------------------------------------------------------------------
// weight gradient: Cout[N, K] += sum_r A[r, n] * B[r, k]   (A = dY, B = layer input)
// 64 x 64 output tile per CTA over a chunk of rows, 4x4 micro-tile per thread, fp32 atomics.
constexpr int TN_ROWS = 16;

__global__ void __launch_bounds__(256)
linear_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t M, int N,
                 int K, int64_t rows_per_cta, float* __restrict__ Cout, int64_t ldc) {
  __shared__ float As[TN_ROWS][64 + 4];
  __shared__ float Bs[TN_ROWS][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < M ? r_begin + rows_per_cta : M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 4, lc = (tid & 15) * 4;        // loader: 16 rows x 64 columns, 4 columns per thread
  for (int64_t r0 = r_begin; r0 < r_end; r0 += TN_ROWS) {
    const int64_t r = r0 + lr;
    float a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a[q] = (r < r_end && n0 + lc + q < N) ? A[r * lda + n0 + lc + q] : 0.f;
      b[q] = (r < r_end && k0 + lc + q < K) ? B[r * ldb + k0 + lc + q] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[lr][lc + q] = a[q]; Bs[lr][lc + q] = b[q]; }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < TN_ROWS; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(Cout + (int64_t)n * ldc + k, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// max over groups of G consecutive rows, and its backward (gradient to the FIRST maximum)
__global__ void group_max_kernel(const float* __restrict__ Y, int64_t ldy, int G, int64_t groups, int C,
                                 float* __restrict__ out, int64_t ldo, uint8_t* __restrict__ arg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * C) return;
  const int64_t g = i / C;
  const int c = (int)(i - g * C);
  const float* y = Y + g * G * ldy + c;
  float m = y[0];
  int a = 0;
  for (int r = 1; r < G; ++r) {
    const float v = y[(int64_t)r * ldy];
    if (v > m) { m = v; a = r; }                      // first maximum, as nn.MaxPool2d
  }
  out[g * ldo + c] = m;
  if (arg) arg[g * C + c] = (uint8_t)a;
}

// BatchNorm(+ReLU) backward for a layer whose output goes straight into the max-pool: the incoming gradient is
// dY[r, c] = dOut[g, c] if r is the argmax row of (group g = r / G, channel c), else 0.  The two sums then
// need only the `groups` argmax rows per channel (G times less data than a dense reduction).
__global__ void __launch_bounds__(256)
bn_bwd_sparse_reduce_kernel(const float* __restrict__ dOut, int64_t lddo, const uint8_t* __restrict__ arg,
                            const float* __restrict__ X, int64_t ldx, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                            const float* __restrict__ beta, int relu, int G, int64_t groups, int C,
                            double* __restrict__ sums) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), gl = threadIdx.x >> 6;
  double s0 = 0.0, s1 = 0.0;
  if (c < C) {
    const float mu = mean[c], rs = rstd[c], ga = gamma[c], be = beta[c];
    for (int64_t g = (int64_t)blockIdx.y * 4 + gl; g < groups; g += (int64_t)gridDim.y * 4) {
      const int64_t r = g * G + arg[g * C + c];
      const float xhat = (X[r * ldx + c] - mu) * rs;
      const bool on = !relu || fmaf(xhat, ga, be) > 0.f;
      const float d = on ? dOut[g * lddo + c] : 0.f;
      s0 += d;
      s1 += (double)d * xhat;
    }
  }
  __shared__ double red[2][4][64];
  red[0][gl][threadIdx.x & 63] = s0;
  red[1][gl][threadIdx.x & 63] = s1;
  __syncthreads();
  if (gl == 0 && c < C) {
    const int t = threadIdx.x;
    atomicAdd(&sums[c], red[0][0][t] + red[0][1][t] + red[0][2][t] + red[0][3][t]);
    atomicAdd(&sums[C + c], red[1][0][t] + red[1][1][t] + red[1][2][t] + red[1][3][t]);
  }
}

__global__ void group_max_bwd_kernel(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ dOut,
                                     int64_t lddo, int G, int64_t groups, int C, float* __restrict__ dY,
                                     int64_t lddy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * C) return;
  const int64_t g = i / C;
  const int c = (int)(i - g * C);
  const float* y = Y + g * G * ldy + c;
  float m = y[0];
  int arg = 0;
  for (int r = 1; r < G; ++r) {
    const float v = y[(int64_t)r * ldy];
    if (v > m) { m = v; arg = r; }
  }
  const float d = dOut[g * lddo + c];
  float* dy = dY + g * G * lddy + c;
  for (int r = 0; r < G; ++r) dy[(int64_t)r * lddy] = r == arg ? d : 0.f;
}

// ---------------------------------------------------------------------------------------------
// backward of the grouping gather: dG [B, N1, K, C] -> dPts [B, N, C] (+= neighbour rows,
// -= the xyz columns at the centroid, which is source point g)
__global__ void group_scatter_add_kernel(const float* __restrict__ dG, const int* __restrict__ idx, int64_t B, int N,
                                         int N1, int K, int C, float* __restrict__ dPts, int64_t ldp) {
  const int64_t total = B * N1 * (int64_t)K * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t row = i / C;                 // (b, g, k)
    const int64_t bg = row / K;
    const int64_t b = bg / N1;
    const int g = (int)(bg - b * N1);
    const float d = dG[i];
    if (d == 0.f) continue;
    int j = idx[row];
    j = j < 0 ? 0 : (j >= N ? N - 1 : j);        // never scatter outside the cloud
    atomicAdd(dPts + (b * N + j) * ldp + c, d);
    if (c < 3) atomicAdd(dPts + (b * N + g) * ldp + c, -d);
  }
}

// backward of the pixel->point gather: dOut [B, n, C] -> dFeat [B, C, HW] (+=)
__global__ void gather_nchw_bwd_kernel(const float* __restrict__ dOut, const int64_t* __restrict__ ind, int64_t B,
                                       int C, int64_t HW, int n, float* __restrict__ dFeat) {
  const int64_t total = B * n * (int64_t)C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bp = i / C;
    const int64_t b = bp / n;
    int64_t pix = ind[bp];
    pix = pix < 0 ? 0 : (pix >= HW ? HW - 1 : pix);   // never scatter outside the map (see gather.cu: clamp_index)
    atomicAdd(dFeat + (b * C + c) * HW + pix, dOut[i]);
  }
}

static unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <int MODE>
static int launch_reduce(const float* A, int64_t lda, const float* Y, int64_t ldy, const float* X, int64_t ldx,
                         const float* mean, const float* rstd, int relu, int64_t M, int C, double* sums,
                         cudaStream_t s, const char* what, const float* gamma = nullptr, const float* beta = nullptr) {
  const int nsum = MODE == RED_COLSUM ? C : 2 * C;
  cudaMemsetAsync(sums, 0, sizeof(double) * nsum, s);
  if (M == 0) return PDF_OK;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = C % 4 == 0 && lda % 4 == 0 && ldy % 4 == 0 && ldx % 4 == 0 && al(A) && al(Y) && al(X) &&
                   al(mean) && al(rstd) && al(gamma) && al(beta);
  const int cv = vec ? C / 4 : C;                       // channel-vectors per row
  int LX = 1;
  while (LX < cv && LX < 32) LX <<= 1;
  const int RY = 256 / LX;
  const int gx = (cv + LX - 1) / LX;
  int64_t gy = (M + (int64_t)RY * 16 - 1) / ((int64_t)RY * 16);
  const int64_t cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (vec) col_reduce_kernel<MODE, 4><<<grid, 256, 0, s>>>(A, lda, Y, ldy, X, ldx, mean, rstd, gamma, beta, relu, M, C, LX, sums);
  else col_reduce_kernel<MODE, 1><<<grid, 256, 0, s>>>(A, lda, Y, ldy, X, ldx, mean, rstd, gamma, beta, relu, M, C, LX, sums);
  return check_launch(what);
}

}  // namespace pdf

using namespace pdf;

extern "C" int pdf_bn_stats(const float* X, int64_t ldx, int64_t M, int C, double* sums, void* stream) {
  PDF_REQUIRE(X && sums && M >= 0 && C > 0 && ldx >= C, PDF_ERR_BAD_ARG, "pdf_bn_stats: bad argument");
  return launch_reduce<RED_STATS>(X, ldx, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, M, C, sums,
                                  (cudaStream_t)stream, "pdf_bn_stats");
}

extern "C" int pdf_col_sum(const float* A, int64_t lda, int64_t M, int C, double* sums, void* stream) {
  PDF_REQUIRE(A && sums && M >= 0 && C > 0 && lda >= C, PDF_ERR_BAD_ARG, "pdf_col_sum: bad argument");
  return launch_reduce<RED_COLSUM>(A, lda, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, M, C, sums,
                                   (cudaStream_t)stream, "pdf_col_sum");
}

extern "C" int pdf_bn_finalize(const double* sums, const float* X, int64_t M, int C, float eps, float momentum,
                               float* running_mean, float* running_var, float* mean, float* rstd, void* stream) {
  PDF_REQUIRE(sums && X && mean && rstd && M > 0 && C > 0, PDF_ERR_BAD_ARG, "pdf_bn_finalize: bad argument");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, X, M, C, eps, momentum, running_mean,
                                                                         running_var, mean, rstd);
  return check_launch("pdf_bn_finalize");
}

extern "C" int pdf_bn_act_fwd(const float* X, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, int relu, int64_t M, int C, float* Y, int64_t ldy, void* Y_img,
                              void* stream) {
  PDF_REQUIRE(X && mean && rstd && gamma && beta && (Y || Y_img) && M >= 0 && C > 0, PDF_ERR_BAD_ARG,
              "pdf_bn_act_fwd: bad argument");
  const int plain = (relu & PDF_BN_PLAIN_IMAGE) ? 1 : 0;      // flag bit on `relu`: image written as one plain bf16 image
  relu &= 1;
  if (Y_img) {
    PDF_REQUIRE(C % 64 == 0 && C <= 2048 && ldx % 4 == 0 && aligned16(X) && (!Y || (ldy % 4 == 0 && aligned16(Y))),
                PDF_ERR_BAD_ARG, "pdf_bn_act_fwd: the image output needs C %% 64 == 0 and 16-byte aligned rows");
    if (M == 0) return PDF_OK;
    const int rows_cta = 256 / (C >> 3);
    int64_t gx = ((((M + 127) >> 7) << 7) + rows_cta - 1) / rows_cta;
    if (gx > 148 * 16) gx = 148 * 16;
    bn_fwd_image_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(X, ldx, mean, rstd, gamma, beta, relu, M, C, Y,
                                                                         ldy, (uint8_t*)Y_img, plain);
    return check_launch("pdf_bn_act_fwd");
  }
  EwArgs p = {};
  p.a = X; p.lda = ldx; p.v0 = mean; p.v1 = rstd; p.v2 = gamma; p.v3 = beta;
  p.o0 = Y; p.ldo0 = ldy; p.M = M; p.C = C; p.flag = relu;
  return launch_ew<EW_BN_FWD>(p, (cudaStream_t)stream, "pdf_bn_act_fwd");
}

extern "C" int pdf_bn_act_bwd(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* X, int64_t ldx,
                              const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                              int64_t M, int C, double* sums, float* dX, int64_t lddx, void* dX_img, void* stream) {
  const int plain = (relu & PDF_BN_PLAIN_IMAGE) ? 1 : 0;
  relu &= 1;
  PDF_REQUIRE(dY && X && mean && rstd && gamma && sums && (dX || dX_img) && M >= 0 && C > 0 && (Y || beta || !relu),
              PDF_ERR_BAD_ARG, "pdf_bn_act_bwd: bad argument");
  PDF_REQUIRE(!dX_img || (C % 64 == 0 && beta && !Y && lddy % 4 == 0 && ldx % 4 == 0 && aligned16(dY) && aligned16(X)),
              PDF_ERR_BAD_ARG, "pdf_bn_act_bwd: the image output needs C %% 64 == 0, beta (mask from x) and aligned rows");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = launch_reduce<RED_BN_BWD>(dY, lddy, Y, ldy, X, ldx, mean, rstd, relu, M, C, sums, s, "pdf_bn_act_bwd", gamma,
                                     beta);
  if (rc != PDF_OK) return rc;
  EwArgs p = {};
  p.a = dY; p.lda = lddy; p.b = Y; p.ldb = ldy; p.c = X; p.ldc = ldx;
  p.v0 = mean; p.v1 = rstd; p.v2 = gamma; p.v3 = beta; p.sums = sums;
  p.o0 = dX; p.ldo0 = lddx; p.M = M; p.C = C; p.flag = relu;
  if (dX_img && M > 0) {
    PDF_REQUIRE(C <= 2048, PDF_ERR_UNSUPPORTED, "pdf_bn_act_bwd: the image output supports C <= 2048");
    const int rows_cta = 256 / (C >> 3);
    int64_t gx = ((((M + 127) >> 7) << 7) + rows_cta - 1) / rows_cta;
    if (gx > 148 * 16) gx = 148 * 16;
    bn_bwd_image_kernel<<<(unsigned)gx, 256, 0, s>>>(dY, lddy, X, ldx, mean, rstd, gamma, beta, sums, relu, M, C,
                                                     (uint8_t*)dX_img, nullptr, 0, nullptr, 1, plain);
    rc = check_launch("pdf_bn_act_bwd");
    if (rc != PDF_OK || !dX) return rc;
  }
  return launch_ew<EW_BN_BWD>(p, s, "pdf_bn_act_bwd");
}

extern "C" int pdf_act_bwd(const float* dY, int64_t lddy, const float* Y, int64_t ldy, int act, int64_t M, int C,
                           float* dX, int64_t lddx, void* stream) {
  PDF_REQUIRE(dY && Y && dX && M >= 0 && C > 0 && act >= 0 && act <= 2, PDF_ERR_BAD_ARG, "pdf_act_bwd: bad argument");
  EwArgs p = {};
  p.a = dY; p.lda = lddy; p.b = Y; p.ldb = ldy; p.o0 = dX; p.ldo0 = lddx; p.M = M; p.C = C; p.flag = act;
  return launch_ew<EW_ACT_BWD>(p, (cudaStream_t)stream, "pdf_act_bwd");
}

extern "C" int pdf_sft_modulate(const float* fea, int64_t ldf, const float* scale, int64_t lds, const float* shift,
                                int64_t ldh, int64_t M, int C, float* out, int64_t ldo, void* stream) {
  PDF_REQUIRE(fea && scale && shift && out && M >= 0 && C > 0, PDF_ERR_BAD_ARG, "pdf_sft_modulate: bad argument");
  EwArgs p = {};
  p.a = fea; p.lda = ldf; p.b = scale; p.ldb = lds; p.c = shift; p.ldc = ldh;
  p.o0 = out; p.ldo0 = ldo; p.M = M; p.C = C;
  return launch_ew<EW_SFT_FWD>(p, (cudaStream_t)stream, "pdf_sft_modulate");
}

extern "C" int pdf_sft_modulate_bwd(const float* dout, int64_t ldd, const float* fea, int64_t ldf, const float* scale,
                                    int64_t lds, int64_t M, int C, float* dfea, int64_t lddf, float* dscale,
                                    int64_t ldds, void* stream) {
  PDF_REQUIRE(dout && fea && scale && dfea && dscale && M >= 0 && C > 0, PDF_ERR_BAD_ARG,
              "pdf_sft_modulate_bwd: bad argument");
  EwArgs p = {};
  p.a = dout; p.lda = ldd; p.b = fea; p.ldb = ldf; p.c = scale; p.ldc = lds;
  p.o0 = dfea; p.ldo0 = lddf; p.o1 = dscale; p.ldo1 = ldds; p.M = M; p.C = C;
  return launch_ew<EW_SFT_BWD>(p, (cudaStream_t)stream, "pdf_sft_modulate_bwd");
}

extern "C" int pdf_linear_tn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int N, int K,
                                 float* C, int64_t ldc, void* stream) {
  PDF_REQUIRE(A && B && C && M >= 0 && N > 0 && K > 0 && lda >= N && ldb >= K && ldc >= K, PDF_ERR_BAD_ARG,
              "pdf_linear_tn_f32: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * K, N, s);
  if (M == 0) return PDF_OK;
  const int gx = (N + 63) / 64, gy = (K + 63) / 64;
  // enough row chunks to fill the machine a few times over, each at least 256 rows
  int64_t chunks = (148 * 8 + gx * gy - 1) / (gx * gy);
  int64_t rows = (M + chunks - 1) / chunks;
  if (rows < 256) rows = 256;
  rows = (rows + TN_ROWS - 1) / TN_ROWS * TN_ROWS;
  chunks = (M + rows - 1) / rows;
  PDF_REQUIRE(chunks <= 65535, PDF_ERR_UNSUPPORTED, "pdf_linear_tn_f32: too many row chunks");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)chunks);
  linear_tn_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, M, N, K, rows, C, ldc);
  return check_launch("pdf_linear_tn_f32");
}

extern "C" int pdf_group_max(const float* Y, int64_t ldy, int G, int64_t groups, int C, float* out, int64_t ldo,
                             uint8_t* arg_out, void* stream) {
  PDF_REQUIRE(Y && out && G > 0 && groups >= 0 && C > 0 && (!arg_out || G <= 256), PDF_ERR_BAD_ARG,
              "pdf_group_max: bad argument");
  if (groups == 0) return PDF_OK;
  const int64_t total = groups * C;
  group_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Y, ldy, G, groups, C, out, ldo,
                                                                                       arg_out);
  return check_launch("pdf_group_max");
}

extern "C" int pdf_bn_maxpool_bwd(const float* dOut, int64_t lddo, const uint8_t* arg, int G, const float* X, int64_t ldx,
                                  const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                                  int64_t M, int C, double* sums, void* dX_img, void* stream) {
  const int plain = (relu & PDF_BN_PLAIN_IMAGE) ? 1 : 0;
  relu &= 1;
  PDF_REQUIRE(dOut && arg && X && mean && rstd && gamma && beta && sums && dX_img && M >= 0 && C > 0 && G > 0 &&
                  G <= 256 && M % G == 0,
              PDF_ERR_BAD_ARG, "pdf_bn_maxpool_bwd: bad argument");
  PDF_REQUIRE(C % 64 == 0 && C <= 2048 && lddo % 4 == 0 && ldx % 4 == 0 && aligned16(dOut) && aligned16(X) &&
                  (reinterpret_cast<uintptr_t>(arg) & 7) == 0,
              PDF_ERR_BAD_ARG, "pdf_bn_maxpool_bwd: needs C %% 64 == 0 and aligned rows");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
  if (M == 0) return PDF_OK;
  const int64_t groups = M / G;
  int64_t gy = (groups + 4 * 32 - 1) / (4 * 32);
  const int gx = (C + 63) / 64;
  const int64_t cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  bn_bwd_sparse_reduce_kernel<<<dim3((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy)), 256, 0, s>>>(
      dOut, lddo, arg, X, ldx, mean, rstd, gamma, beta, relu, G, groups, C, sums);
  int rc = check_launch("pdf_bn_maxpool_bwd");
  if (rc != PDF_OK) return rc;
  const int rows_cta = 256 / (C >> 3);
  int64_t gxi = ((((M + 127) >> 7) << 7) + rows_cta - 1) / rows_cta;
  if (gxi > 148 * 16) gxi = 148 * 16;
  bn_bwd_image_kernel<<<(unsigned)gxi, 256, 0, s>>>(nullptr, 0, X, ldx, mean, rstd, gamma, beta, sums, relu, M, C,
                                                    (uint8_t*)dX_img, dOut, lddo, arg, G, plain);
  return check_launch("pdf_bn_maxpool_bwd");
}

extern "C" int pdf_group_max_bwd(const float* Y, int64_t ldy, const float* dOut, int64_t lddo, int G, int64_t groups,
                                 int C, float* dY, int64_t lddy, void* stream) {
  PDF_REQUIRE(Y && dOut && dY && G > 0 && groups >= 0 && C > 0, PDF_ERR_BAD_ARG, "pdf_group_max_bwd: bad argument");
  if (groups == 0) return PDF_OK;
  const int64_t total = groups * C;
  group_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Y, ldy, dOut, lddo, G,
                                                                                           groups, C, dY, lddy);
  return check_launch("pdf_group_max_bwd");
}

extern "C" int pdf_group_scatter_add(const float* dG, const int* idx, int64_t B, int N, int N1, int K, int C,
                                     float* dPts, int64_t ldp, void* stream) {
  PDF_REQUIRE(dG && idx && dPts && B >= 0 && N > 0 && N1 > 0 && N1 <= N && K > 0 && C >= 3 && ldp >= C,
              PDF_ERR_BAD_ARG, "pdf_group_scatter_add: bad argument");
  if (B == 0) return PDF_OK;
  const int64_t total = B * N1 * (int64_t)K * C;
  group_scatter_add_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dG, idx, B, N, N1, K, C, dPts, ldp);
  return check_launch("pdf_group_scatter_add");
}

extern "C" int pdf_gather_nchw_bwd(const float* dOut, const int64_t* ind, int64_t B, int C, int64_t HW, int n,
                                   float* dFeat, void* stream) {
  PDF_REQUIRE(dOut && ind && dFeat && B >= 0 && C > 0 && HW > 0 && n > 0, PDF_ERR_BAD_ARG,
              "pdf_gather_nchw_bwd: bad argument");
  if (B == 0) return PDF_OK;
  const int64_t total = B * n * (int64_t)C;
  gather_nchw_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dOut, ind, B, C, HW, n, dFeat);
  return check_launch("pdf_gather_nchw_bwd");
}

// ---------------------------------------------------------------------------------------------
// The K = 3 first layer of the level-1 point-MLP (netR_1[0], intaghand_encoder.py:50) and its two
// gradients are pure streaming: 12 B in / 4*N B out per row.  The 64x64-tile FFMA kernels waste 95 % of a
// tile on them, so they get their own bandwidth-shaped kernels (K <= 4).
namespace pdf {

// Y[m, n] = sum_k X[m,k] W[n,k] + b[n]; one thread = 4 consecutive channels of 4 rows per iteration
// (N % 4 == 0; blockDim.x = 256 = (N/4 channel groups) x (rows per CTA pass))
__global__ void __launch_bounds__(256)
linear_smallk_fwd_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw,
                         const float* __restrict__ bias, int64_t M, int N, int K, float* __restrict__ Y, int64_t ldy) {
  extern __shared__ float sw[];                      // [N][4] weights (zero padded) then [N] bias
  for (int i = threadIdx.x; i < N * 4; i += blockDim.x) sw[i] = (i & 3) < K ? W[(int64_t)(i >> 2) * ldw + (i & 3)] : 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sw[N * 4 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int nq = N >> 2;                             // channel groups per row
  const int rows_cta = 256 / nq > 0 ? 256 / nq : 1;  // rows covered by one CTA pass (nq <= 256 by the host check)
  const int q = threadIdx.x % nq, rl = threadIdx.x / nq;
  if (rl >= rows_cta) return;
  const int n0 = q * 4;
  float4 w[4];
  float bb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { w[j] = *reinterpret_cast<const float4*>(sw + (n0 + j) * 4); bb[j] = sw[N * 4 + n0 + j]; }
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * rows_cta;
  for (int64_t m0 = (int64_t)blockIdx.x * rows_cta + rl; m0 < M; m0 += U * stride) {
    float x[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t m = m0 + u * stride;
#pragma unroll
      for (int k = 0; k < 4; ++k) x[u][k] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t m = m0 + u * stride;
      if (m >= M) break;
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        y[j] = fmaf(w[j].w, x[u][3], fmaf(w[j].z, x[u][2], fmaf(w[j].y, x[u][1], fmaf(w[j].x, x[u][0], bb[j]))));
      *reinterpret_cast<float4*>(Y + m * ldy + n0) = make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// dX[m, k] = sum_n dY[m,n] W[n,k] (k < K <= 4): 16 lanes x float4 per row, two row groups per warp, four rows
// in flight per group
__global__ void __launch_bounds__(256)
linear_smallk_dx_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ W, int64_t ldw,
                        int64_t M, int N, int K, float* __restrict__ dX, int64_t lddx) {
  extern __shared__ float sw[];                      // [N][4]
  for (int i = threadIdx.x; i < N * 4; i += blockDim.x) sw[i] = (i & 3) < K ? W[(int64_t)(i >> 2) * ldw + (i & 3)] : 0.f;
  __syncthreads();
  const int lane16 = threadIdx.x & 15;
  const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);   // the two 16-lane row groups of a warp may leave the loop apart
  const int64_t row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
  const int64_t rstep = ((int64_t)gridDim.x * blockDim.x) >> 4;
  constexpr int U = 4;
  for (int64_t m0 = row0; m0 < M; m0 += U * rstep) {   // the 16 lanes of a row group run the same trip count
    float a[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u][0] = a[u][1] = a[u][2] = a[u][3] = 0.f;
    for (int n0 = lane16 * 4; n0 < N; n0 += 64) {
      float4 d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t m = m0 + u * rstep;
        d[u] = m < M ? *reinterpret_cast<const float4*>(dY + m * lddy + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + (n0 + j) * 4);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float dv = j == 0 ? d[u].x : (j == 1 ? d[u].y : (j == 2 ? d[u].z : d[u].w));
          a[u][0] = fmaf(dv, w.x, a[u][0]); a[u][1] = fmaf(dv, w.y, a[u][1]);
          a[u][2] = fmaf(dv, w.z, a[u][2]); a[u][3] = fmaf(dv, w.w, a[u][3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[u][k] += __shfl_xor_sync(hmask, a[u][k], o, 16);
      const int64_t m = m0 + u * rstep;
      if (lane16 == 0 && m < M)
        for (int k = 0; k < K; ++k) dX[m * lddx + k] = a[u][k];
    }
  }
}

// dW[n, k] = sum_m dY[m,n] X[m,k]: block = 64 channels x 4 row lanes, eight rows in flight per thread,
// fp32 partials, atomics at the end
__global__ void __launch_bounds__(256)
linear_smallk_dw_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ X, int64_t ldx, int64_t M,
                        int N, int K, float* __restrict__ dW, int64_t lddw) {
  const int c = blockIdx.y * 64 + (threadIdx.x & 63), rl = threadIdx.x >> 6;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int U = 8;
  const int64_t stride = (int64_t)gridDim.x * 4;
  if (c < N) {
    for (int64_t m0 = (int64_t)blockIdx.x * 4 + rl; m0 < M; m0 += U * stride) {
      float d[U], x[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t m = m0 + u * stride;
        d[u] = m < M ? dY[m * lddy + c] : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) x[u][k] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = fmaf(d[u], x[u][k], a[k]);
    }
  }
  __shared__ float red[4][64][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) red[rl][threadIdx.x & 63][k] = a[k];
  __syncthreads();
  if (rl == 0 && c < N)
    for (int k = 0; k < K; ++k)
      atomicAdd(dW + (int64_t)c * lddw + k, red[0][threadIdx.x][k] + red[1][threadIdx.x][k] + red[2][threadIdx.x][k] +
                                                red[3][threadIdx.x][k]);
}

}  // namespace pdf

extern "C" int pdf_linear_smallk_f32(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias,
                                     int64_t M, int N, int K, float* out, int64_t ldo, void* stream) {
  if (M == 0 && mode != 2) return PDF_OK;
  PDF_REQUIRE(A && B && out && M >= 0 && N > 0 && K > 0 && K <= 4 && mode >= 0 && mode <= 2, PDF_ERR_BAD_ARG,
              "pdf_linear_smallk_f32: bad argument");
  PDF_REQUIRE(N % 4 == 0 && N <= 2048, PDF_ERR_UNSUPPORTED, "pdf_linear_smallk_f32: N must be a multiple of 4, <= 2048");
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) {                                     // forward: A = X [M,K], B = W [N,K] -> out [M,N]
    PDF_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, PDF_ERR_BAD_ARG,
                "pdf_linear_smallk_f32: output rows must be 16-byte aligned");
    PDF_REQUIRE(N <= 1024, PDF_ERR_UNSUPPORTED, "pdf_linear_smallk_f32: forward supports N <= 1024");
    const int rows_cta = 256 / (N / 4);
    int64_t gx = (M + (int64_t)rows_cta * 4 - 1) / ((int64_t)rows_cta * 4);
    if (gx > 148 * 16) gx = 148 * 16;
    linear_smallk_fwd_kernel<<<(unsigned)(gx < 1 ? 1 : gx), 256, (size_t)N * 5 * sizeof(float), s>>>(A, lda, B, ldb, bias,
                                                                                                     M, N, K, out, ldo);
  } else if (mode == 1) {                              // data gradient: A = dY [M,N], B = W [N,K] -> out [M,K]
    PDF_REQUIRE(lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, PDF_ERR_BAD_ARG,
                "pdf_linear_smallk_f32: dY rows must be 16-byte aligned");
    linear_smallk_dx_kernel<<<grid_for(M * 4), 256, (size_t)N * 4 * sizeof(float), s>>>(A, lda, B, ldb, M, N, K, out, ldo);
  } else {                                             // weight gradient: A = dY [M,N], B = X [M,K] -> out [N,K]
    cudaMemset2DAsync(out, sizeof(float) * ldo, 0, sizeof(float) * K, N, s);
    if (M == 0) return PDF_OK;
    int64_t gx = (M + 4 * 64 - 1) / (4 * 64);
    if (gx > 148 * 8) gx = 148 * 8;
    linear_smallk_dw_kernel<<<dim3((unsigned)gx, (unsigned)((N + 63) / 64)), 256, 0, s>>>(A, lda, B, ldb, M, N, K, out, ldo);
  }
  return check_launch("pdf_linear_smallk_f32");
}
