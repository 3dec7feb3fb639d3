// MANO linear blend skinning, four hands per CTA, fp32 end to end (the 1e-5 m
// tolerance rules out bf16 bases, SURVEY.md section 7 "LBS precision").
// The whole op is 1.17 MFLOP per hand and a 16-step dependent chain: it is
// latency bound, so the kernel keeps everything of one hand in shared memory and
// reads the (L2-resident, 1.4 MB) basis tables in a coalesced, transposed layout.
#include "pdf_common.cuh"

namespace pdf {

constexpr int NV = 778, NE = NV * 3;
constexpr int VT_PITCH = PDF_MANO_VT_PITCH;   // row pitch of the precomputed blend-shape rows (16-byte aligned rows)

__constant__ int c_parent[16] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};  // kintree_table[0]
__constant__ int c_new_order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

struct Tips { int v[5]; };
struct ManoTables {                      // one side's constants (see include/pdfnet_b200.h)
  const float *v_template, *shapedirs_t, *posedirs_t, *j_template, *j_shapedirs, *weights_t;
};

// HPC hands per CTA (1: the kernel is latency bound; the blend-shape contraction can be taken out of
// it entirely, see v_tpose_in / pdf_mano_pose_feature).
constexpr int HPC = 1;

__global__ void __launch_bounds__(256)
mano_lbs_kernel(ManoTables T0, ManoTables T1, int pair, int root_is_mat,
                const float* __restrict__ root, const float* __restrict__ pose, const float* __restrict__ shape,
                const float* __restrict__ trans, const float* __restrict__ scale, int64_t n_hands, Tips tips0,
                Tips tips1, int center_idx, int new_skel, const float* __restrict__ v_tpose_in,
                float* __restrict__ v_out, float* __restrict__ j_out) {
  // pair mode: hands are laid out (frame, side); even hands use table set 0 (left), odd hands set 1 (right)
  const bool alt = pair && (blockIdx.x & 1);
  const ManoTables& T = alt ? T1 : T0;
  const Tips tips = alt ? tips1 : tips0;
  const float* __restrict__ v_template = T.v_template;
  const float* __restrict__ shapedirs_t = T.shapedirs_t;
  const float* __restrict__ posedirs_t = T.posedirs_t;
  const float* __restrict__ j_template = T.j_template;
  const float* __restrict__ j_shapedirs = T.j_shapedirs;
  const float* __restrict__ weights_t = T.weights_t;
  __shared__ float s_aa[HPC][48];          // axis-angle: root + 15 joints
  __shared__ float s_beta[HPC][10];
  __shared__ float s_R[HPC][16][9];
  __shared__ float s_pf[HPC][135];         // pose feature R - I
  __shared__ float s_jt[HPC][16][3];       // rest joints
  __shared__ float s_G[HPC][16][12];       // global 3x4 transforms
  __shared__ float s_v[HPC][NE];           // v_tpose, then posed vertices
  __shared__ float s_j[HPC][21][3];
  const int64_t h0 = (int64_t)blockIdx.x * HPC;
  const int nh = (int)min((int64_t)HPC, n_hands - h0);
  const int tid = threadIdx.x;

  for (int i = tid; i < HPC * 58; i += 256) {
    const int hh = i / 58, k = i % 58;
    float val = 0.f;
    if (hh < nh) {
      const int64_t h = h0 + hh;
      val = k < 3 ? (root_is_mat ? 0.f : root[h * 3 + k]) : (k < 48 ? pose[h * 45 + k - 3] : shape[h * 10 + k - 48]);
    }
    if (k < 48) s_aa[hh][k] = val; else s_beta[hh][k - 48] = val;
  }
  __syncthreads();

  if (tid < HPC * 16) {                     // rodrigues_batch, manolayer.py:32-48
    const int hh = tid >> 4, jn = tid & 15;
    const float ax = s_aa[hh][jn * 3], ay = s_aa[hh][jn * 3 + 1], az = s_aa[hh][jn * 3 + 2];
    const float angle = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;
    const float x = ax / angle, y = ay / angle, z = az / angle;
    const float sn = sinf(angle), cs = cosf(angle);
    const float L[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
    const float oc = 1.f - cs;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float ll = L[r * 3] * L[c] + L[r * 3 + 1] * L[3 + c] + L[r * 3 + 2] * L[6 + c];
        float v = (r == c ? 1.f : 0.f) + sn * L[r * 3 + c] + oc * ll;
        // use_pca layers receive the root rotation as a 3x3 matrix and use it as is (manolayer.py:266-267,285)
        if (jn == 0 && root_is_mat) v = root[(h0 + hh) * 9 + r * 3 + c];
        s_R[hh][jn][r * 3 + c] = v;
        if (jn > 0) s_pf[hh][(jn - 1) * 9 + r * 3 + c] = v - (r == c ? 1.f : 0.f);
      }
  } else if (tid >= 64 && tid < 64 + HPC * 48) {   // rest joints = J_regressor (v_template + shapedirs beta)
    const int hh = (tid - 64) / 48, e = (tid - 64) % 48;
    float a = j_template[e];
#pragma unroll
    for (int k = 0; k < 10; ++k) a = fmaf(j_shapedirs[e * 10 + k], s_beta[hh][k], a);
    s_jt[hh][e / 3][e % 3] = a;
  }
  __syncthreads();

  // blend shapes: v_tpose = v_template + shapedirs beta + posedirs (R - I)   (:274-282)
  if (v_tpose_in != nullptr) {              // contraction already done as one [n,145]x[145,2334] GEMM
    for (int i = tid; i < nh * NE; i += 256) s_v[i / NE][i % NE] = v_tpose_in[(h0 + i / NE) * VT_PITCH + i % NE];
  } else
  for (int e = tid; e < NE; e += 256) {
    float a[HPC], p0[HPC], p1[HPC];
#pragma unroll
    for (int hh = 0; hh < HPC; ++hh) { a[hh] = 0.f; p0[hh] = 0.f; p1[hh] = 0.f; }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float w = __ldg(shapedirs_t + k * NE + e);
#pragma unroll
      for (int hh = 0; hh < HPC; ++hh) a[hh] = fmaf(w, s_beta[hh][k], a[hh]);
    }
    const float vt = __ldg(v_template + e);
#pragma unroll 5
    for (int k = 0; k < 134; k += 2) {
      const float w0 = __ldg(posedirs_t + k * NE + e), w1 = __ldg(posedirs_t + (k + 1) * NE + e);
#pragma unroll
      for (int hh = 0; hh < HPC; ++hh) {
        p0[hh] = fmaf(w0, s_pf[hh][k], p0[hh]);
        p1[hh] = fmaf(w1, s_pf[hh][k + 1], p1[hh]);
      }
    }
    const float wl = __ldg(posedirs_t + 134 * NE + e);
#pragma unroll
    for (int hh = 0; hh < HPC; ++hh) s_v[hh][e] = (a[hh] + vt) + (fmaf(wl, s_pf[hh][134], p0[hh]) + p1[hh]);
  }

  // kinematic chain (:284-293): G_i = G_parent * [R_i | (I - R_i) j_i]; warp hh handles hand hh
  if (tid < HPC * 32) {
    const int hh = tid >> 5, l = tid & 31;
    const int r = l / 4, c = l % 4;               // lanes 0..11 hold one 3x4 entry
    for (int i = 0; i < 16; ++i) {
      if (l < 12) {
        float lc[3];                              // column c of the local transform (rows 0..2)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (c < 3) lc[k] = s_R[hh][i][k * 3 + c];
          else {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 3; ++q) t = fmaf(((k == q) ? 1.f : 0.f) - s_R[hh][i][k * 3 + q], s_jt[hh][i][q], t);
            lc[k] = t;
          }
        }
        float g;
        if (i == 0) g = lc[r];
        else {
          const float* P = s_G[hh][c_parent[i]];
          g = P[r * 4] * lc[0] + P[r * 4 + 1] * lc[1] + P[r * 4 + 2] * lc[2] + (c == 3 ? P[r * 4 + 3] : 0.f);
        }
        s_G[hh][i][r * 4 + c] = g;
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // skinning (:301-304)
  for (int v = tid; v < NV; v += 256) {
    float w[16];
#pragma unroll
    for (int jn = 0; jn < 16; ++jn) w[jn] = __ldg(weights_t + jn * NV + v);
#pragma unroll 1
    for (int hh = 0; hh < HPC; ++hh) {
      float g[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) g[q] = 0.f;
#pragma unroll
      for (int jn = 0; jn < 16; ++jn)
#pragma unroll
        for (int q = 0; q < 12; ++q) g[q] = fmaf(w[jn], s_G[hh][jn][q], g[q]);
      const float x = s_v[hh][v * 3], y = s_v[hh][v * 3 + 1], z = s_v[hh][v * 3 + 2];
      s_v[hh][v * 3] = g[0] * x + g[1] * y + g[2] * z + g[3];
      s_v[hh][v * 3 + 1] = g[4] * x + g[5] * y + g[6] * z + g[7];
      s_v[hh][v * 3 + 2] = g[8] * x + g[9] * y + g[10] * z + g[11];
    }
  }
  __syncthreads();

  // joints: 16 posed joints + 5 tips, reordered (:295-311)
  for (int i = tid; i < HPC * 63; i += 256) {
    const int hh = i / 63, jo = (i % 63) / 3, c = i % 3;
    const int src = c_new_order[jo];
    float val;
    if (src == 0) val = s_jt[hh][0][c];
    else if (src < 16) {
      const float* P = s_G[hh][c_parent[src]];
      val = P[c * 4] * s_jt[hh][src][0] + P[c * 4 + 1] * s_jt[hh][src][1] + P[c * 4 + 2] * s_jt[hh][src][2] + P[c * 4 + 3];
    } else val = s_v[hh][tips.v[src - 16] * 3 + c];
    s_j[hh][jo][c] = val;
  }
  __syncthreads();

  for (int hh = 0; hh < nh; ++hh) {
    const int64_t h = h0 + hh;
    float cen[3] = {0.f, 0.f, 0.f};
    if (center_idx >= 0) { cen[0] = s_j[hh][center_idx][0]; cen[1] = s_j[hh][center_idx][1]; cen[2] = s_j[hh][center_idx][2]; }
    const float sc = scale ? scale[h] : 1.f;
    float tr[3] = {0.f, 0.f, 0.f};
    if (trans) { tr[0] = trans[h * 3]; tr[1] = trans[h * 3 + 1]; tr[2] = trans[h * 3 + 2]; }
    __syncthreads();
    for (int e = tid; e < NE; e += 256) {
      float val = s_v[hh][e];
      if (center_idx >= 0) val = val - cen[e % 3];
      if (scale) val = val * sc;
      if (trans) val = val + tr[e % 3];
      s_v[hh][e] = val;
      v_out[h * NE + e] = val;
    }
    if (tid < 63) {
      float val = s_j[hh][tid / 3][tid % 3];
      if (center_idx >= 0) val = val - cen[tid % 3];
      if (scale) val = val * sc;
      if (trans) val = val + tr[tid % 3];
      s_j[hh][tid / 3][tid % 3] = val;
    }
    __syncthreads();
    if (tid < 63) {
      const int jo = tid / 3, c = tid % 3;
      float val = s_j[hh][jo][c];
      if (new_skel) {                                 // :328-332
        if (jo == 5) val = (s_v[hh][63 * 3 + c] + s_v[hh][144 * 3 + c]) / 2.f;
        else if (jo == 9) val = (s_v[hh][271 * 3 + c] + s_v[hh][220 * 3 + c]) / 2.f;
        else if (jo == 13) val = (s_v[hh][148 * 3 + c] + s_v[hh][290 * 3 + c]) / 2.f;
        else if (jo == 17) val = (s_v[hh][770 * 3 + c] + s_v[hh][83 * 3 + c]) / 2.f;
      }
      j_out[h * 63 + tid] = val;
    }
  }
}

// Blend-shape coefficients of every hand: X[h] = [beta(10) | (R_1..R_15 - I)(135)]  (:274-281), so that
// v_tpose = X * [shapedirs | posedirs]^T + v_template is a single dense GEMM over all hands.
__global__ void mano_pose_feature_kernel(const float* __restrict__ pose, const float* __restrict__ shape,
                                         int64_t n, float* __restrict__ X) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  const int64_t h = i >> 4;
  const int jn = (int)(i & 15);
  float* x = X + h * 145;
  if (jn == 0) {
    for (int k = 0; k < 10; ++k) x[k] = shape[h * 10 + k];
    return;
  }
  const float ax = pose[h * 45 + (jn - 1) * 3], ay = pose[h * 45 + (jn - 1) * 3 + 1], az = pose[h * 45 + (jn - 1) * 3 + 2];
  const float angle = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;
  const float xx = ax / angle, yy = ay / angle, zz = az / angle;
  const float sn = sinf(angle), oc = 1.f - cosf(angle);
  const float L[9] = {0.f, -zz, yy, zz, 0.f, -xx, -yy, xx, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ll = L[r * 3] * L[c] + L[r * 3 + 1] * L[3 + c] + L[r * 3 + 2] * L[6 + c];
      x[10 + (jn - 1) * 9 + r * 3 + c] = ((r == c ? 1.f : 0.f) + sn * L[r * 3 + c] + oc * ll) - (r == c ? 1.f : 0.f);
    }
}

// rodrigues_batch (manolayer.py:32-48): axis-angle [n,3] -> rotation matrices [n,3,3]; one thread per rotation.
__global__ void rodrigues_kernel(const float* __restrict__ axis, int64_t n, float* __restrict__ Rm) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ax = axis[i * 3], ay = axis[i * 3 + 1], az = axis[i * 3 + 2];
  const float angle = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;
  const float x = ax / angle, y = ay / angle, z = az / angle;
  const float sn = sinf(angle), oc = 1.f - cosf(angle);
  const float L[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ll = L[r * 3] * L[c] + L[r * 3 + 1] * L[3 + c] + L[r * 3 + 2] * L[6 + c];
      Rm[i * 9 + r * 3 + c] = (r == c ? 1.f : 0.f) + sn * L[r * 3 + c] + oc * ll;
    }
}

// joints[h, j, c] = sum_v reg[j, v] * verts[h, v, c]   (full_regressor @ verts, Mano_model.py:309-323,
// demo.py:217-218).  One CTA per mesh: the 778x3 vertices are staged in shared memory, one warp per output
// (j, c), lanes stride over the vertices of the dense (L2-resident, 65 KB) regressor row.
__global__ void __launch_bounds__(256)
joint_regress_kernel(const float* __restrict__ reg, int n_joints, const float* __restrict__ verts,
                     float* __restrict__ joints) {
  __shared__ float s_v[NE];
  const int64_t h = blockIdx.x;
  for (int e = threadIdx.x; e < NE; e += 256) s_v[e] = verts[h * NE + e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < n_joints * 3; o += 8) {
    const int jn = o / 3, c = o - jn * 3;
    float a = 0.f;
    for (int v = lane; v < NV; v += 32) a = fmaf(__ldg(reg + jn * NV + v), s_v[v * 3 + c], a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) joints[(h * n_joints + jn) * 3 + c] = a;
  }
}

}  // namespace pdf

extern "C" int pdf_rodrigues(const float* axis, int64_t n, float* rot, void* stream) {
  if (n == 0) return PDF_OK;
  PDF_REQUIRE(axis && rot && n > 0, PDF_ERR_BAD_ARG, "pdf_rodrigues: null pointer / bad size");
  pdf::rodrigues_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(axis, n, rot);
  return pdf::check_launch("pdf_rodrigues");
}

extern "C" int pdf_joint_regress(const float* reg, int n_joints, const float* verts, int64_t n, float* joints,
                                 void* stream) {
  if (n == 0) return PDF_OK;
  PDF_REQUIRE(reg && verts && joints, PDF_ERR_BAD_ARG, "pdf_joint_regress: null pointer");
  PDF_REQUIRE(n > 0 && n_joints > 0 && n_joints <= 64, PDF_ERR_BAD_ARG, "pdf_joint_regress: bad size");
  pdf::joint_regress_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(reg, n_joints, verts, joints);
  return pdf::check_launch("pdf_joint_regress");
}

static int mano_lbs_launch(const pdf::ManoTables& T0, const pdf::ManoTables& T1, int pair, const float* root,
                           const float* pose, const float* shape, const float* trans, const float* scale, int64_t n,
                           const int32_t* tip0, const int32_t* tip1, int center_idx, int new_skel,
                           const float* v_tpose, float* v, float* j, void* stream, int root_is_mat = 0);

extern "C" int pdf_mano_lbs(const float* v_template, const float* shapedirs_t, const float* posedirs_t,
                            const float* j_template, const float* j_shapedirs, const float* weights_t,
                            const float* root, const float* pose, const float* shape, const float* trans,
                            const float* scale, int64_t n, const int32_t* tip_idx_host, int center_idx,
                            int new_skel, const float* v_tpose, float* v, float* j, void* stream) {
  PDF_REQUIRE(v_template && shapedirs_t && posedirs_t && j_template && j_shapedirs && weights_t, PDF_ERR_BAD_ARG,
              "pdf_mano_lbs: null table pointer");
  pdf::ManoTables T{v_template, shapedirs_t, posedirs_t, j_template, j_shapedirs, weights_t};
  return mano_lbs_launch(T, T, 0, root, pose, shape, trans, scale, n, tip_idx_host, tip_idx_host, center_idx, new_skel,
                         v_tpose, v, j, stream);
}

extern "C" int pdf_mano_lbs_rootmat(const float* v_template, const float* shapedirs_t, const float* posedirs_t,
                                    const float* j_template, const float* j_shapedirs, const float* weights_t,
                                    const float* root_mat, const float* pose, const float* shape, const float* trans,
                                    const float* scale, int64_t n, const int32_t* tip_idx_host, int center_idx,
                                    int new_skel, const float* v_tpose, float* v, float* j, void* stream) {
  PDF_REQUIRE(v_template && shapedirs_t && posedirs_t && j_template && j_shapedirs && weights_t, PDF_ERR_BAD_ARG,
              "pdf_mano_lbs_rootmat: null table pointer");
  pdf::ManoTables T{v_template, shapedirs_t, posedirs_t, j_template, j_shapedirs, weights_t};
  return mano_lbs_launch(T, T, 0, root_mat, pose, shape, trans, scale, n, tip_idx_host, tip_idx_host, center_idx,
                         new_skel, v_tpose, v, j, stream, 1);
}

extern "C" int pdf_mano_lbs_pair(const float* const* tables_left, const float* const* tables_right, const float* root,
                                 const float* pose, const float* shape, const float* trans, const float* scale,
                                 int64_t n, const int32_t* tips_left_host, const int32_t* tips_right_host,
                                 int center_idx, int new_skel, const float* v_tpose, float* v, float* j, void* stream) {
  PDF_REQUIRE(tables_left && tables_right, PDF_ERR_BAD_ARG, "pdf_mano_lbs_pair: null table array");
  for (int i = 0; i < 6; ++i)
    PDF_REQUIRE(tables_left[i] && tables_right[i], PDF_ERR_BAD_ARG, "pdf_mano_lbs_pair: null table pointer");
  PDF_REQUIRE((n & 1) == 0, PDF_ERR_BAD_ARG, "pdf_mano_lbs_pair: n must be even ((frame, side) layout)");
  pdf::ManoTables L{tables_left[0], tables_left[1], tables_left[2], tables_left[3], tables_left[4], tables_left[5]};
  pdf::ManoTables R{tables_right[0], tables_right[1], tables_right[2], tables_right[3], tables_right[4], tables_right[5]};
  return mano_lbs_launch(L, R, 1, root, pose, shape, trans, scale, n, tips_left_host, tips_right_host, center_idx,
                         new_skel, v_tpose, v, j, stream);
}

static int mano_lbs_launch(const pdf::ManoTables& T0, const pdf::ManoTables& T1, int pair, const float* root,
                           const float* pose, const float* shape, const float* trans, const float* scale, int64_t n,
                           const int32_t* tip_idx_host, const int32_t* tip1_host, int center_idx, int new_skel,
                           const float* v_tpose, float* v, float* j, void* stream, int root_is_mat) {
  if (n == 0) return PDF_OK;
  PDF_REQUIRE(root && pose && shape && v && j && tip_idx_host, PDF_ERR_BAD_ARG, "pdf_mano_lbs: null pointer");
  PDF_REQUIRE(n >= 0 && center_idx < 21, PDF_ERR_BAD_ARG, "pdf_mano_lbs: bad size");
  PDF_REQUIRE(tip1_host != nullptr, PDF_ERR_BAD_ARG, "pdf_mano_lbs: null pointer");
  pdf::Tips tips, tips1;
  for (int i = 0; i < 5; ++i) {
    PDF_REQUIRE(tip_idx_host[i] >= 0 && tip_idx_host[i] < pdf::NV && tip1_host[i] >= 0 && tip1_host[i] < pdf::NV,
                PDF_ERR_BAD_ARG, "pdf_mano_lbs: tip index");
    tips.v[i] = tip_idx_host[i];
    tips1.v[i] = tip1_host[i];
  }
  if (n == 0) return PDF_OK;
  pdf::mano_lbs_kernel<<<(unsigned)((n + pdf::HPC - 1) / pdf::HPC), 256, 0, (cudaStream_t)stream>>>(
      T0, T1, pair, root_is_mat, root, pose, shape, trans, scale, n, tips, tips1, center_idx, new_skel, v_tpose, v, j);
  return pdf::check_launch("pdf_mano_lbs");
}

extern "C" int pdf_mano_pose_feature(const float* pose, const float* shape, int64_t n, float* X, void* stream) {
  if (n == 0) return PDF_OK;
  PDF_REQUIRE(pose && shape && X, PDF_ERR_BAD_ARG, "pdf_mano_pose_feature: null pointer");
  pdf::mano_pose_feature_kernel<<<(unsigned)((n * 16 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pose, shape, n, X);
  return pdf::check_launch("pdf_mano_pose_feature");
}
