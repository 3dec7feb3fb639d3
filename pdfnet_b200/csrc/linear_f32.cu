// fp32 linear layer Y = epilogue(X * W^T + b) on the FFMA pipe.
// This is the exact-precision ("fp32 parity", 1e-4) path of the shared point-MLP,
// the SFT 1x1 convs and mano_head; the bf16 tcgen05 kernel in sa_mlp_bf16.cu is the
// throughput path for the two set-abstraction stages.
//
// 64x64 output tile per CTA, BK = 16, 256 threads, 4x4 micro-tile per thread,
// operands staged through shared memory (K-major, padded against bank conflicts).
#include "pdf_common.cuh"

namespace pdf {

constexpr int BM = 64, BN = 64, BK = 16;

template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == PDF_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == PDF_ACT_LEAKY01) return v > 0.f ? v : 0.1f * v;
  return v;
}

template <int ACT, int EPI>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ X, int64_t lda, const float* __restrict__ W, int64_t ldw,
                  const float* __restrict__ bias, int64_t M, int N, int K, int group,
                  const float* __restrict__ F, int64_t ldf, float* __restrict__ Y, int64_t ldy) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int gmax[BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 256 threads cover a 64 x 16 tile as 4 k-columns x 64 rows
  const int lr = tid >> 2;            // 0..63 row inside tile
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += BK) {
    float a[4], b[4];
    const int64_t am = m0 + lr;
    const int bn = n0 + lr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk = k0 + lk + q;
      a[q] = (am < M && kk < K) ? X[am * lda + kk] : 0.f;
      b[q] = (bn < N && kk < K) ? __ldg(W + (int64_t)bn * ldw + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) { As[lk + q][lr] = a[q]; Bs[lk + q][lr] = b[q]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  if (EPI == PDF_EPI_GROUP_MAX) {
    // ReLU outputs are >= 0, so max commutes with the int reinterpretation.
    if (tid < BN) gmax[tid] = 0;
    __syncthreads();
    const bool one_group = (group % BM) == 0;   // the whole tile belongs to one group
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float bj = bias ? bias[n] : 0.f;
      if (one_group) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (m0 + ty * 4 + i < M) v = fmaxf(v, activate<ACT>(acc[i][j] + bj));
        atomicMax(&gmax[tx * 4 + j], __float_as_int(v));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t m = m0 + ty * 4 + i;
          if (m < M) atomicMax(reinterpret_cast<int*>(Y + (m / group) * ldy + n),
                               __float_as_int(activate<ACT>(acc[i][j] + bj)));
        }
      }
    }
    if (one_group) {
      __syncthreads();
      if (tid < BN && n0 + tid < N) {
        int* dst = reinterpret_cast<int*>(Y + (m0 / group) * ldy + n0 + tid);
        if (group == BM) *dst = gmax[tid];
        else atomicMax(dst, gmax[tid]);
      }
    }
    return;
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float s = __fadd_rn(acc[i][j], bias ? bias[n] : 0.f);
      float* y = Y + m * ldy + n;
      if (EPI == PDF_EPI_STORE) *y = activate<ACT>(s);
      else if (EPI == PDF_EPI_SFT_SCALE) *y = __fmul_rn(F[m * ldf + n], __fadd_rn(s, 1.f));
      else if (EPI == PDF_EPI_ACCUM) *y = __fadd_rn(*y, s);
    }
  }
}

}  // namespace pdf

extern "C" int pdf_linear_f32(const float* X, int64_t lda, const float* W, int64_t ldw, const float* bias,
                              int64_t M, int N, int K, int act, int epilogue, int group, const float* F,
                              int64_t ldf, float* Y, int64_t ldy, void* stream) {
  using namespace pdf;
  if (M == 0) return PDF_OK;
  PDF_REQUIRE(X && W && Y, PDF_ERR_BAD_ARG, "pdf_linear_f32: null pointer");
  PDF_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K, PDF_ERR_BAD_ARG, "pdf_linear_f32: bad size");
  PDF_REQUIRE(act >= 0 && act <= 2 && epilogue >= 0 && epilogue <= 3, PDF_ERR_BAD_ARG, "pdf_linear_f32: bad enum");
  PDF_REQUIRE(epilogue != PDF_EPI_SFT_SCALE || F != nullptr, PDF_ERR_BAD_ARG, "pdf_linear_f32: SFT_SCALE needs F");
  PDF_REQUIRE(epilogue != PDF_EPI_GROUP_MAX || (act == PDF_ACT_RELU && group > 0 && M % group == 0),
              PDF_ERR_BAD_ARG, "pdf_linear_f32: GROUP_MAX needs RELU and M %% group == 0");
  PDF_REQUIRE(epilogue == PDF_EPI_STORE || epilogue == PDF_EPI_GROUP_MAX || act == PDF_ACT_NONE, PDF_ERR_BAD_ARG,
              "pdf_linear_f32: SFT epilogues take no activation");
  if (M == 0) return PDF_OK;
  const int64_t gm = (M + BM - 1) / BM;
  PDF_REQUIRE(gm < (1ll << 31), PDF_ERR_UNSUPPORTED, "pdf_linear_f32: M too large");
  dim3 grid((unsigned)gm, (unsigned)((N + BN - 1) / BN));
  cudaStream_t s = (cudaStream_t)stream;
#define GO(A, E) linear_f32_kernel<A, E><<<grid, 256, 0, s>>>(X, lda, W, ldw, bias, M, N, K, group, F, ldf, Y, ldy)
  if (epilogue == PDF_EPI_STORE) {
    if (act == PDF_ACT_NONE) GO(PDF_ACT_NONE, PDF_EPI_STORE);
    else if (act == PDF_ACT_RELU) GO(PDF_ACT_RELU, PDF_EPI_STORE);
    else GO(PDF_ACT_LEAKY01, PDF_EPI_STORE);
  } else if (epilogue == PDF_EPI_SFT_SCALE) GO(PDF_ACT_NONE, PDF_EPI_SFT_SCALE);
  else if (epilogue == PDF_EPI_ACCUM) GO(PDF_ACT_NONE, PDF_EPI_ACCUM);
  else GO(PDF_ACT_RELU, PDF_EPI_GROUP_MAX);
#undef GO
  return check_launch("pdf_linear_f32");
}
