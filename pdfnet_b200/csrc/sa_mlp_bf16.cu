// Fused set-abstraction stage on 5th-gen tensor cores:
//   neighbour gather -> centroid subtraction -> 3-layer shared point-MLP (BN folded,
//   ReLU) -> max over the 64 neighbours, with NO intermediate in HBM.
//
// Design (one persistent CTA per SM, S independent "slots" of 128 threads each):
//   * a tile = 128 neighbour rows = 2 groups (centroids) x 64 neighbours;
//   * folded bf16 weights of all three layers stay resident in shared memory for the
//     lifetime of the CTA, in the exact UMMA operand layout, brought in with one
//     cp.async.bulk (TMA engine) per 32 KB chunk from a host-packed image;
//   * layers 1 and 2 run "points as M": D[p, c] (TMEM lane = point) so each thread
//     repacks ITS row to bf16 (cvt.rn.relu.bf16x2) straight into the next operand tile;
//   * layer 3 runs transposed, "channels as M": D[c, p] (TMEM lane = channel), so the
//     max over the 64 neighbours is a per-thread reduction over TMEM columns
//     (3-input FMNMX), no shuffles;
//   * biases never touch the CUDA cores: every operand tile carries a 16-column
//     auxiliary K block [x_hi y_hi z_hi 1 | x_lo y_lo z_lo 1 | 0...] against weight
//     columns [w_x w_y w_z b_hi | w_x w_y w_z b_lo | 0...] (layer 1) or
//     [0 0 0 b_hi | 0 0 0 b_lo | 0...] (layers 2,3): the relative coordinates enter
//     as a bf16 hi+lo pair (~fp32 accuracy) and the bias is added by the MMA itself;
//   * each slot issues its own MMAs (one elected thread) and waits on its own
//     mbarrier; slots overlap each other's CUDA-core phases with tensor-core phases.
#include <stdlib.h>
#include "pdf_common.cuh"
#include "umma.cuh"

namespace pdf {
using namespace umma;

template <int CF_, int C1_, int C2_, int C3_, int SLOTS_, bool COMPACT_ = false, bool TS_ = true, bool EARLY_ = true>
struct SaCfg {
  // EARLY: the inputs of tile t+1 are staged while tile t is still in its last MMA phase / its max epilogue: the
  // geometry block is double-buffered (written before waiting for layer 3 of tile t) and the level-2 feature rows
  // are copied (cp.async) into the operand tile the moment layer 3 has finished reading it, i.e. under the max
  // epilogue, instead of at the top of the next iteration where their latency sat on the slot's serial chain.
  static constexpr bool EARLY = EARLY_ && !COMPACT_;
  // TS: layer 2 reads its A operand (the layer-1 activations) from TENSOR MEMORY: the layer-1 epilogue packs
  // ReLU(D1) to bf16 in place over D1's own columns (tcgen05.st) together with the constant bias block, so the
  // activations never cross the shared-memory port (whose 128 B/clk - MMA operand reads plus the repack stores -
  // is what bounds these kernels), and the generic->async proxy fence of that stage goes.
  static constexpr bool TS = TS_;
  static constexpr int CF = CF_;        // feature channels gathered with each neighbour (0 or 128)
  static constexpr int C1 = C1_, C2 = C2_, C3 = C3_, SLOTS = SLOTS_;
  // COMPACT: every accumulator of a tile (D1, D2, and layer 3 one 64-point group at a time) reuses
  // the SAME 64 TMEM columns, so twice as many tiles are in flight per SM (TMEM, not issue slots or
  // tensor throughput, is what limits the number of concurrent tiles); costs one more MMA phase.
  static constexpr bool COMPACT = COMPACT_;
  static constexpr int KB1 = CF / 64, KB2 = C1 / 64, KB3 = C2 / 64;       // 64-wide SW128 K blocks per layer
  // weight image (bytes): [W1 feat blocks][W1 aux][W2 feat][W2 aux][W3 feat][W3 aux]
  static constexpr int W1_FEAT = KB1 * C1 * 128, W1_AUX = C1 * 32;
  static constexpr int W2_FEAT = KB2 * C2 * 128, W2_AUX = C2 * 32;
  static constexpr int W3_FEAT = KB3 * C3 * 128, W3_AUX = C3 * 32;
  static constexpr int OFF_W1A = W1_FEAT, OFF_W2 = OFF_W1A + W1_AUX, OFF_W2A = OFF_W2 + W2_FEAT;
  static constexpr int OFF_W3 = OFF_W2A + W2_AUX, OFF_W3A = OFF_W3 + W3_FEAT;
  static constexpr int W_BYTES = OFF_W3A + W3_AUX;                        // the part that lives in shared memory
  static constexpr int OFF_B3 = W_BYTES;                                  // fp32 bias of layer 3, read from global once per thread
  static constexpr int PACK_BYTES = W_BYTES + C3 * 4;                     //   (early kernel: added after the max)
  static constexpr int KBMAX = (KB1 > KB2 ? (KB1 > KB3 ? KB1 : KB3) : (KB2 > KB3 ? KB2 : KB3));
  static constexpr int SLOT_FEAT = KBMAX * 128 * 128;                     // activation tile, in place
  static constexpr int AUX_BUFS = EARLY ? 2 : 1;
  static constexpr int SLOT_BYTES = SLOT_FEAT + AUX_BUFS * 128 * 32;      // + geometry/bias aux block(s)
  static constexpr int ROWIDX_BYTES = (EARLY && CF > 0) ? SLOTS * 2 * 128 * 2 : 0;   // uint16 row indices of two tiles
  static constexpr int SMEM_BYTES = W_BYTES + SLOTS * SLOT_BYTES + 1024 /*align*/ + 256 /*barriers*/ + ROWIDX_BYTES;
  static constexpr int TMEM_PER_SLOT = COMPACT ? 64 : ((C3 > C1 + C2) ? C3 : ((C1 + C2) > 128 ? C1 + C2 : 128));
  static_assert(!COMPACT || (C1 <= 64 && C2 <= 64 && C3 == 128), "COMPACT is the level-1 plan");
  static constexpr int THREADS = SLOTS * 128;
  static_assert(SLOTS * TMEM_PER_SLOT <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// (the COMPACT 8-slot plan measured slower on B200: level 1 is bound by shared-memory operand
// bandwidth, not by the number of tiles in flight — see DESIGN.md section 3.1)
#ifndef PDF_SA_TS
#define PDF_SA_TS 1
#endif
#ifndef PDF_SA_EARLY
#define PDF_SA_EARLY 1
#endif
using Sa1Cfg = SaCfg<0, 64, 64, 128, 4, false, PDF_SA_TS != 0, PDF_SA_EARLY != 0>;
using Sa2Cfg = SaCfg<128, 128, 128, 256, 2, false, PDF_SA_TS != 0, PDF_SA_EARLY != 0>;
using Sa1CfgLate = SaCfg<0, 64, 64, 128, 4, false, PDF_SA_TS != 0, false>;        // PDF_SA_EARLY=0 at run time (A/B)
using Sa2CfgLate = SaCfg<128, 128, 128, 256, 2, false, PDF_SA_TS != 0, false>;

// One layer = (KB SW128 K-blocks x 4 K-steps) + 1 aux K-step, accumulating into d_tmem.
template <int KB, bool AUX = true>
__device__ __forceinline__ void issue_layer(uint32_t a_feat, uint32_t a_blk, uint32_t a_aux, uint32_t b_feat,
                                            uint32_t b_blk, uint32_t b_aux, uint32_t d_tmem, uint32_t idesc) {
  bool acc = false;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
    for (int k16 = 0; k16 < 4; ++k16) {
      mma_bf16(d_tmem, desc_sw128(a_feat + kb * a_blk + k16 * 32), desc_sw128(b_feat + kb * b_blk + k16 * 32), idesc,
               acc);
      acc = true;
    }
  }
  if (AUX) mma_bf16(d_tmem, desc_none(a_aux), desc_none(b_aux), idesc, acc);
}

// Layer with the A operand in tensor memory: KB*4 K-steps over the packed activations (8 columns per step)
// + 1 K-step over the constant bias block that follows them.
template <int KB>
__device__ __forceinline__ void issue_layer_ts(uint32_t a_tmem, uint32_t b_feat, uint32_t b_blk, uint32_t b_aux,
                                               uint32_t d_tmem, uint32_t idesc) {
  bool acc = false;
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
    for (int k16 = 0; k16 < 4; ++k16) {
      mma_bf16_ts(d_tmem, a_tmem + (kb * 4 + k16) * 8, desc_sw128(b_feat + kb * b_blk + k16 * 32), idesc, acc);
      acc = true;
    }
  }
  mma_bf16_ts(d_tmem, a_tmem + KB * 32, desc_none(b_aux), idesc, acc);
}

// Epilogue of layer 1 in TS mode: TMEM row (fp32, C columns) -> ReLU -> bf16 pairs written back IN PLACE over
// the first C/2 columns, followed by the 8-column bias block [0 0 0 1 | 0 0 0 1 | 0 x 8] (K = 16 step).
template <int C>
__device__ __forceinline__ void epilogue_repack_tmem(uint32_t tmem_row) {
  uint32_t w[C / 2];
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q)
      w[c0 / 2 + q] = pack_relu_bf16(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
  }
#pragma unroll
  for (int c0 = 0; c0 < C / 2; c0 += 16) {
    uint32_t t[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) t[q] = w[c0 + q];
    tmem_st16(tmem_row + c0, t);
  }
  const uint32_t one_hi = 0x3F800000u;                    // bf16 (0, 1): element k = 3 / k = 7 of the block
  const uint32_t aux[8] = {0u, one_hi, 0u, one_hi, 0u, 0u, 0u, 0u};
  tmem_st8(tmem_row + C / 2, aux);
  tmem_st_wait();
}

// Epilogue of layers 1/2: TMEM row (this thread's point) -> ReLU -> bf16 -> SW128 tile row.
template <int C>
__device__ __forceinline__ void epilogue_repack(uint32_t tmem_row, uint32_t feat, int p) {   // feat: shared address
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_row + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {                         // 4 chunks of 8 columns
      const int col = c0 + q * 8;
      uint4 w;
      w.x = pack_relu_bf16(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
      w.y = pack_relu_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
      w.z = pack_relu_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
      w.w = pack_relu_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
      st_shared_v4(feat + (col >> 6) * (128 * 128) + sw128_off(p, col & 63), w);
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
sa_mlp_max_kernel(const float* __restrict__ pts, int n_src, int64_t ld_pts, const uint16_t* __restrict__ feat_bf16,
                  const int32_t* __restrict__ idx, int n_centroids, const uint8_t* __restrict__ wpack,
                  float* __restrict__ out, int64_t ld_out, int out_col0, int64_t n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + Cfg::W_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_slots + Cfg::SLOTS * Cfg::SLOT_BYTES);   // [0]=weights, [1+s]=slot
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = tid >> 7, p = tid & 127;              // p: tile row (layers 1,2) / channel lane (layer 3)
  const int wslot = warp & 3;                            // TMEM lane quarter of this warp

  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    for (int s = 0; s < Cfg::SLOTS; ++s) mbar_init(smem_u32(&s_bar[1 + s]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(s_tmem);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  pdl_wait();                                            // on-chip set-up above; global memory from here on
  pdl_trigger();
  if (tid == 0) {                                        // resident weights: bulk copies on the TMA engine
    mbar_expect_tx(smem_u32(&s_bar[0]), Cfg::W_BYTES);
    for (int off = 0; off < Cfg::W_BYTES; off += 32768) {
      const int n = (Cfg::W_BYTES - off) < 32768 ? (Cfg::W_BYTES - off) : 32768;
      bulk_g2s(smem_u32(s_w + off), wpack + off, n, smem_u32(&s_bar[0]));
    }
  }

  uint8_t* my_feat = s_slots + slot * Cfg::SLOT_BYTES;
  uint8_t* my_aux = my_feat + Cfg::SLOT_FEAT;
  const uint32_t sa_feat = smem_u32(my_feat), sa_aux = smem_u32(my_aux);
  const uint32_t sw = smem_u32(s_w);
  const uint32_t bar = smem_u32(&s_bar[1 + slot]);
  const uint32_t d_base = tmem_base + slot * Cfg::TMEM_PER_SLOT;
  const uint32_t lane_off = ((uint32_t)(wslot * 32)) << 16;
  const uint32_t d1 = d_base, d2 = Cfg::COMPACT ? d_base : d_base + Cfg::C1, d3 = d_base;
  uint32_t phase = 0;
  const int tiles_per_cloud = n_centroids >> 1;
  // tile -> (cloud, first centroid): a shift when the tile count per cloud is a power of two (it is for the
  // reference's 512 / 128 centroids) instead of ~20-instruction integer divisions on the per-tile path
  const bool pow2 = (tiles_per_cloud & (tiles_per_cloud - 1)) == 0;
  const int tshift = 31 - __clz(tiles_per_cloud);
  auto tile_cloud = [&](int64_t t) -> int64_t {
    return pow2 ? (int64_t)((uint32_t)t >> tshift) : (int64_t)((uint32_t)t / (uint32_t)tiles_per_cloud);
  };
  auto tile_g0 = [&](int64_t t) -> int {
    return (pow2 ? (int)((uint32_t)t & (uint32_t)(tiles_per_cloud - 1)) : (int)((uint32_t)t % (uint32_t)tiles_per_cloud)) * 2;
  };

  mbar_wait(smem_u32(&s_bar[0]), 0);                     // weights resident

  // Two-deep software prefetch of the geometry (warps issue in order, so a load only overlaps
  // with other work if its first USE is an iteration away): the neighbour index of tile t+2 and the
  // raw coordinates of tile t+1 are in flight while tile t runs its three MMA phases.
  auto load_idx = [&](int64_t t) -> int {
    const int64_t b = tile_cloud(t);                                  // n_tiles < 2^31 (checked by the launcher)
    const int g0 = tile_g0(t);
    return __ldg(idx + (b * n_centroids + g0) * 64 + p);
  };
  auto load_raw = [&](int64_t t, int j, float (&sp)[3], float (&cp)[3]) {
    const int64_t b = tile_cloud(t);
    const int g0 = tile_g0(t);
    const float* cloud = pts + b * n_src * ld_pts;
    const float* src = cloud + (int64_t)j * ld_pts;
    const float* cen = cloud + (int64_t)(g0 + (p >> 6)) * ld_pts;
#pragma unroll
    for (int c = 0; c < 3; ++c) { sp[c] = __ldg(src + c); cp[c] = __ldg(cen + c); }
  };
  const int64_t t_first = (int64_t)blockIdx.x * Cfg::SLOTS + slot, t_step = (int64_t)gridDim.x * Cfg::SLOTS;
  float sp[3] = {0.f, 0.f, 0.f}, cp[3] = {0.f, 0.f, 0.f};
  int j_next = 0;
  if (t_first < n_tiles) load_raw(t_first, load_idx(t_first), sp, cp);
  if (t_first + t_step < n_tiles) j_next = load_idx(t_first + t_step);

  for (int64_t t = t_first; t < n_tiles; t += t_step) {
    const int64_t b = tile_cloud(t);
    const int g0 = tile_g0(t);
    const float* cloud = pts + b * n_src * ld_pts;
    const int32_t* tidx = idx + (b * n_centroids + g0) * 64;
    // p_j - c_i in fp32, the reference's operand order (utils.py:142-143)
    const float rx = __fsub_rn(sp[0], cp[0]), ry = __fsub_rn(sp[1], cp[1]), rz = __fsub_rn(sp[2], cp[2]);
    if ((p & 63) == 0 && out_col0 > 0) {                   // centroid xyz (+ zero pad) into the leading columns,
      float* o = out + (b * n_centroids + g0 + (p >> 6)) * ld_out;   // straight from the prefetched registers
      // constant indices only: a runtime index would put cp[] in local memory and make every prefetch of the
      // centroid coordinates wait for its loads at the spill store
      if (out_col0 > 0) o[0] = cp[0];
      if (out_col0 > 1) o[1] = cp[1];
      if (out_col0 > 2) o[2] = cp[2];
      if (out_col0 > 3) o[3] = 0.f;
    }

    // ---- gather: geometry/bias block (thread = row, prefetched one tile ahead) ----
    {
      const float hx = __uint_as_float(pack_bf16(rx, 0.f) << 16), hy = __uint_as_float(pack_bf16(ry, 0.f) << 16),
                  hz = __uint_as_float(pack_bf16(rz, 0.f) << 16);
      uint4 w;
      w.x = pack_bf16(hx, hy);
      w.y = pack_bf16(hz, 1.f);
      w.z = pack_bf16(rx - hx, ry - hy);
      w.w = pack_bf16(rz - hz, 1.f);
      st_shared_v4(sa_aux + aux_off(p, 0), w);
      st_shared_v4(sa_aux + aux_off(p, 8), make_uint4(0, 0, 0, 0));
    }
    if (Cfg::CF > 0) {
      if (feat_bf16 != nullptr) {
        // bf16 feature rows (256 B each): 16 consecutive threads copy one row with cp.async
        // (LDGSTS, no register staging), 8 rows per step, straight into the SW128 tile.
        const uint16_t* fcloud = feat_bf16 + b * n_src * (int64_t)Cfg::CF;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int row = i * 8 + (p >> 4), ch = p & 15;           // 16-byte chunk ch of row
          const int j = __ldg(tidx + row);
          const uint32_t dst = sa_feat + (ch >> 3) * (128 * 128) + sw128_off(row, (ch & 7) * 8);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(fcloud + (int64_t)j * Cfg::CF + ch * 8)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        // fp32 source rows [xyz, pad, 128 features]: one warp per row, lane l loads float4 #l
        // (coalesced 512 B) and stores 4 bf16 into the SW128 tile.
        const int w4 = warp & 3;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const int row = w4 * 32 + r;
          const int j = tidx[row];
          const float4 f = __ldg(reinterpret_cast<const float4*>(cloud + (int64_t)j * ld_pts + 4) + lane);
          uint2 w;
          w.x = pack_bf16(f.x, f.y);
          w.y = pack_bf16(f.z, f.w);
          const int col = lane * 4;
          st_shared_v2(sa_feat + (col >> 6) * (128 * 128) + sw128_off(row, col & 63), w);
        }
      }
    }
    fence_async_smem();
    fence_before_sync();                                 // previous tile's TMEM reads are complete
    named_bar_sync(1 + slot, 128);

    // ---- layer 1: D1[p, c1] = X[p, :] . W1[c1, :] ----
    if (p == 0) {
      fence_after_sync();
      issue_layer<Cfg::KB1>(sa_feat, 128 * 128, sa_aux, sw, Cfg::C1 * 128, sw + Cfg::OFF_W1A, d1,
                            idesc_bf16(128, Cfg::C1));
      commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    if (Cfg::TS) {
      epilogue_repack_tmem<Cfg::C1>(d1 + lane_off);      // H1 stays in tensor memory (in place over D1)
    } else {
      epilogue_repack<Cfg::C1>(d1 + lane_off, sa_feat, p);
      fence_async_smem();
    }
    fence_before_sync();
    named_bar_sync(1 + slot, 128);

    // ---- layer 2: D2[p, c2] = H1[p, :] . W2[c2, :] ----
    if (p == 0) {
      fence_after_sync();
      if (Cfg::TS)
        issue_layer_ts<Cfg::KB2>(d1, sw + Cfg::OFF_W2, Cfg::C2 * 128, sw + Cfg::OFF_W2A, d2, idesc_bf16(128, Cfg::C2));
      else
        issue_layer<Cfg::KB2>(sa_feat, 128 * 128, sa_aux, sw + Cfg::OFF_W2, Cfg::C2 * 128, sw + Cfg::OFF_W2A, d2,
                              idesc_bf16(128, Cfg::C2));
      commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    epilogue_repack<Cfg::C2>(d2 + lane_off, sa_feat, p);
    fence_async_smem();
    fence_before_sync();
    named_bar_sync(1 + slot, 128);

    // ---- layer 3 (transposed): D3[c3, p] = W3[c3, :] . H2[p, :] ; max over each 64-point group ----
    if (Cfg::COMPACT) {
      // one group (64 points = 8 row-groups of the H2 tile) per MMA phase, same 64 TMEM columns
#pragma unroll 1
      for (int grp = 0; grp < 2; ++grp) {
        if (p == 0) {
          fence_after_sync();
          issue_layer<Cfg::KB3>(sw + Cfg::OFF_W3, Cfg::C3 * 128, sw + Cfg::OFF_W3A, sa_feat + grp * (64 * 128),
                                128 * 128, sa_aux + grp * (64 * 32), d3, idesc_bf16(128, 64));
          commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after_sync();
        float mm = 0.f;                                  // ReLU floor
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(d3 + lane_off + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; q += 2) mm = max3(mm, __uint_as_float(v[q]), __uint_as_float(v[q + 1]));
        }
        out[(b * n_centroids + g0 + grp) * ld_out + out_col0 + p] = mm;
        if (grp == 0) {                                  // D3 of group 1 overwrites the columns just read
          fence_before_sync();
          named_bar_sync(1 + slot, 128);
        }
      }
    } else {
    if (p == 0) {
      fence_after_sync();
#pragma unroll
      for (int h = 0; h < Cfg::C3 / 128; ++h)
        issue_layer<Cfg::KB3>(sw + Cfg::OFF_W3 + h * (128 * 128), Cfg::C3 * 128, sw + Cfg::OFF_W3A + h * (128 * 32),
                              sa_feat, 128 * 128, sa_aux, d3 + h * 128, idesc_bf16(128, 128));
      commit(bar);
    }
    // prefetch for the next tiles, issued AFTER the last proxy fence of this tile (the fence is a
    // MEMBAR and would otherwise wait for these loads): first use is the top of the next iteration
    if (t + t_step < n_tiles) load_raw(t + t_step, j_next, sp, cp);
    if (t + 2 * t_step < n_tiles) j_next = load_idx(t + 2 * t_step);
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
#pragma unroll
    for (int h = 0; h < Cfg::C3 / 128; ++h) {
      float m[2] = {0.f, 0.f};                           // ReLU floor
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(d3 + lane_off + h * 128 + c0, v);
        tmem_ld_wait();
        float mm = m[c0 >> 6];
#pragma unroll
        for (int q = 0; q < 32; q += 2) mm = max3(mm, __uint_as_float(v[q]), __uint_as_float(v[q + 1]));
        m[c0 >> 6] = mm;
      }
      const int ch = h * 128 + p;
      float* o = out + (b * n_centroids + g0) * ld_out + out_col0 + ch;
      o[0] = m[0];
      o[ld_out] = m[1];
    }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// The same stage with the inputs of tile t+1 staged under the tail of tile t (Cfg::EARLY):
//   * tile t covers neighbour rows [128 t, 128 t + 128) of the index array and output rows 2t, 2t+1 (two centroids
//     per tile, n_centroids even), so the per-tile addresses need no division besides cloud = t / tiles_per_cloud;
//   * geometry block double-buffered: the relative coordinates of tile t+1 are written right after layer 3 of tile
//     t has been ISSUED (their raw coordinates were loaded an iteration earlier), not after its epilogue;
//   * level 2: the bf16 feature rows of tile t+1 are copied (cp.async) into the operand tile the moment layer 3 of
//     tile t has COMPLETED (nothing reads the tile any more), so the copies fly under the max epilogue; their
//     row indices travel through a small shared array (thread = row writes, 16 threads per row read).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
sa_mlp_max_early_kernel(const float* __restrict__ pts, int n_src, int64_t ld_pts, const uint16_t* __restrict__ feat_bf16,
                        const int32_t* __restrict__ idx, int n_centroids, const uint8_t* __restrict__ wpack,
                        float* __restrict__ out, int64_t ld_out, int out_col0, int64_t n_tiles, int early_feat) {
  static_assert(Cfg::EARLY && !Cfg::COMPACT, "early staging plan");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_slots = smem + Cfg::W_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_slots + Cfg::SLOTS * Cfg::SLOT_BYTES);   // [0]=weights, [1+s]=slot
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);
  uint16_t* s_rowidx = reinterpret_cast<uint16_t*>(s_bar + 32);                           // [SLOTS][2][128] (level 2)

  const int tid = threadIdx.x;
  // warp and slot indices through a shuffle: the compiler then KNOWS they are warp-uniform and keeps every address
  // derived from them (operand tiles, barriers, tensor-memory columns) in uniform registers - the tcgen05 operands
  // need no per-instruction R2UR / elect waterfall, which kept the K = 16 steps of layer 2 issue-bound
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int slot = warp >> 2, p = tid & 127;             // p: tile row (layers 1,2) / channel lane (layer 3)
  const int wslot = warp & 3;                            // TMEM lane quarter of this warp

  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    for (int s = 0; s < Cfg::SLOTS; ++s) mbar_init(smem_u32(&s_bar[1 + s]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(s_tmem);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  pdl_wait();                                            // on-chip set-up above; global memory from here on
  pdl_trigger();
  if (tid == 0) {                                        // resident weights: bulk copies on the TMA engine
    mbar_expect_tx(smem_u32(&s_bar[0]), Cfg::W_BYTES);
    for (int off = 0; off < Cfg::W_BYTES; off += 32768) {
      const int n = (Cfg::W_BYTES - off) < 32768 ? (Cfg::W_BYTES - off) : 32768;
      bulk_g2s(smem_u32(s_w + off), wpack + off, n, smem_u32(&s_bar[0]));
    }
  }

  const uint32_t sa_feat = smem_u32(s_slots + slot * Cfg::SLOT_BYTES);
  const uint32_t sa_aux0 = sa_feat + Cfg::SLOT_FEAT;     // two geometry blocks of 4 KB
  const uint32_t sw = smem_u32(s_w);
  const uint32_t bar = smem_u32(&s_bar[1 + slot]);
  const uint32_t d_base = tmem_base + slot * Cfg::TMEM_PER_SLOT;
  const uint32_t lane_off = ((uint32_t)(wslot * 32)) << 16;
  const uint32_t d1 = d_base, d2 = d_base + Cfg::C1, d3 = d_base;
  uint16_t* my_rowidx = s_rowidx + slot * 256;
  uint32_t phase = 0;
  const uint32_t tiles_per_cloud = (uint32_t)(n_centroids >> 1);
  const bool pow2 = (tiles_per_cloud & (tiles_per_cloud - 1)) == 0;
  const int tshift = 31 - __clz((int)tiles_per_cloud);
  const int64_t cloud_pitch = (int64_t)n_src * ld_pts;
  const bool feat16 = Cfg::CF > 0 && feat_bf16 != nullptr;
  // cloud of tile t and its first centroid (n_tiles < 2^31, checked by the launcher)
  auto cloud_of = [&](uint32_t t) -> uint32_t { return pow2 ? (t >> tshift) : (t / tiles_per_cloud); };
  // raw coordinates of this thread's neighbour row and of its centroid, tile t
  auto load_raw = [&](uint32_t t, int j, float& s0, float& s1, float& s2, float& c0, float& c1, float& c2) {
    const uint32_t b = cloud_of(t);
    const uint32_t g = ((t - b * tiles_per_cloud) << 1) + (uint32_t)(p >> 6);
    const float* cloud = pts + (int64_t)b * cloud_pitch;
    const float* src = cloud + (int64_t)j * ld_pts;
    const float* cen = cloud + (int64_t)g * ld_pts;
    s0 = __ldg(src); s1 = __ldg(src + 1); s2 = __ldg(src + 2);
    c0 = __ldg(cen); c1 = __ldg(cen + 1); c2 = __ldg(cen + 2);
  };
  // geometry / bias block of one tile (thread = row) + the centroid's leading output columns
  auto write_aux = [&](uint32_t aux, uint32_t t, float s0, float s1, float s2, float c0, float c1, float c2) {
    // p_j - c_i in fp32, the reference's operand order (utils.py:142-143)
    const float rx = __fsub_rn(s0, c0), ry = __fsub_rn(s1, c1), rz = __fsub_rn(s2, c2);
    if ((p & 63) == 0 && out_col0 > 0) {
      float* o = out + ((int64_t)t * 2 + (p >> 6)) * ld_out;
      o[0] = c0;
      if (out_col0 > 1) o[1] = c1;
      if (out_col0 > 2) o[2] = c2;
      if (out_col0 > 3) o[3] = 0.f;
    }
    const float hx = __uint_as_float(pack_bf16(rx, 0.f) << 16), hy = __uint_as_float(pack_bf16(ry, 0.f) << 16),
                hz = __uint_as_float(pack_bf16(rz, 0.f) << 16);
    uint4 w;
    w.x = pack_bf16(hx, hy);
    w.y = pack_bf16(hz, 1.f);
    w.z = pack_bf16(rx - hx, ry - hy);
    w.w = pack_bf16(rz - hz, 1.f);
    st_shared_v4(aux + aux_off(p, 0), w);                // columns 8..15 of the block stay zero (written once below)
    fence_async_smem();                                  // visible to the tensor core's (async-proxy) operand reads
  };
  // bf16 feature rows (256 B each) of tile t: 16 consecutive threads copy one row with cp.async (LDGSTS, no
  // register staging), 8 rows per step, straight into the SW128 operand tile; row indices from shared memory
  auto issue_features = [&](uint32_t t, const uint16_t* rowidx) {
    const uint16_t* fcloud = feat_bf16 + (int64_t)cloud_of(t) * n_src * (int64_t)Cfg::CF;
    const int ch = p & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int row = i * 8 + (p >> 4);
      const int j = rowidx[row];
      const uint32_t dst = sa_feat + (ch >> 3) * (128 * 128) + sw128_off(row, (ch & 7) * 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(fcloud + (int64_t)j * Cfg::CF + ch * 8)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  st_shared_v4(sa_aux0 + aux_off(p, 8), make_uint4(0, 0, 0, 0));             // zero half of both geometry blocks
  st_shared_v4(sa_aux0 + 128 * 32 + aux_off(p, 8), make_uint4(0, 0, 0, 0));
  const uint32_t t_first = blockIdx.x * Cfg::SLOTS + slot, t_step = gridDim.x * Cfg::SLOTS;
  const uint32_t nt = (uint32_t)n_tiles;
  // invariants at the top of iteration t: geometry block [buf] = tile t (and, level 2, its feature rows are in
  // the operand tile); (s*, c*) = raw coordinates of tile t+1; j_next = this row's neighbour index of tile t+2
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  int j_next = 0;
  if (t_first < nt) {
    const int j0 = __ldg(idx + (int64_t)t_first * 128 + p);
    load_raw(t_first, j0, s0, s1, s2, c0, c1, c2);
    write_aux(sa_aux0, t_first, s0, s1, s2, c0, c1, c2);
    if (feat16) {
      my_rowidx[p] = (uint16_t)j0;
      named_bar_sync(1 + slot, 128);
      issue_features(t_first, my_rowidx);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (t_first + t_step < nt) {
      j_next = __ldg(idx + (int64_t)(t_first + t_step) * 128 + p);
      if (feat16) my_rowidx[128 + p] = (uint16_t)j_next;
      load_raw(t_first + t_step, j_next, s0, s1, s2, c0, c1, c2);
    }
    if (t_first + 2 * t_step < nt) j_next = __ldg(idx + (int64_t)(t_first + 2 * t_step) * 128 + p);
  }
  mbar_wait(smem_u32(&s_bar[0]), 0);                     // weights resident (their copy ran under the loads above)

  float b3[Cfg::C3 / 128];                               // layer-3 bias of this thread's channel(s)
#pragma unroll
  for (int h = 0; h < Cfg::C3 / 128; ++h) b3[h] = __ldg(reinterpret_cast<const float*>(wpack + Cfg::OFF_B3) + h * 128 + p);

  uint32_t buf = 0;
  for (uint32_t t = t_first; t < nt; t += t_step) {
    const uint32_t sa_aux = sa_aux0 + buf * (128 * 32);
    const bool has_next = t + t_step < nt;
    if (Cfg::CF > 0 && !feat16) {
      // fp32 source rows [xyz, pad, 128 features]: one warp per row, lane l loads float4 #l (coalesced 512 B) and
      // stores 4 bf16 into the SW128 tile.  (Legacy input form: staged here, not early.)
      const int lane = tid & 31;
      const float* cloud = pts + (int64_t)cloud_of(t) * cloud_pitch;
      const int32_t* tidx = idx + (int64_t)t * 128;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int row = wslot * 32 + r;
        const int j = tidx[row];
        const float4 f = __ldg(reinterpret_cast<const float4*>(cloud + (int64_t)j * ld_pts + 4) + lane);
        uint2 w;
        w.x = pack_bf16(f.x, f.y);
        w.y = pack_bf16(f.z, f.w);
        const int col = lane * 4;
        st_shared_v2(sa_feat + (col >> 6) * (128 * 128) + sw128_off(row, col & 63), w);
      }
    }
    if (feat16 && !early_feat && t != t_first) {         // late plan: the copies sit at the top of the iteration
      issue_features(t, my_rowidx + buf * 128);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ---- layer 1: D1[p, c1] = X[p, :] . W1[c1, :] ----
    // Level 1 (no feature rows): layer 1 of tile t+1 only needs its geometry block and D3's FIRST 64 columns, so it
    // was issued from inside tile t's max epilogue (below) and its round trip ran under the second half of that
    // epilogue; only a slot's first tile issues it here.
    constexpr bool EARLY_L1 = Cfg::CF == 0;
    if (!EARLY_L1 || t == t_first) {
      if (Cfg::CF > 0) fence_async_smem();               // feature rows (the geometry block was fenced when written)
      fence_before_sync();                               // previous tile's TMEM reads are complete
      named_bar_sync(1 + slot, 128);
      if (p == 0) {
        fence_after_sync();
        issue_layer<Cfg::KB1>(sa_feat, 128 * 128, sa_aux, sw, Cfg::C1 * 128, sw + Cfg::OFF_W1A, d1,
                              idesc_bf16(128, Cfg::C1));
        commit(bar);
      }
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    if (Cfg::TS) {
      epilogue_repack_tmem<Cfg::C1>(d1 + lane_off);      // H1 stays in tensor memory (in place over D1)
    } else {
      epilogue_repack<Cfg::C1>(d1 + lane_off, sa_feat, p);
      fence_async_smem();
    }
    fence_before_sync();
    named_bar_sync(1 + slot, 128);

    // ---- layer 2: D2[p, c2] = H1[p, :] . W2[c2, :] ----
    if (p == 0) {
      fence_after_sync();
      if (Cfg::TS)
        issue_layer_ts<Cfg::KB2>(d1, sw + Cfg::OFF_W2, Cfg::C2 * 128, sw + Cfg::OFF_W2A, d2, idesc_bf16(128, Cfg::C2));
      else
        issue_layer<Cfg::KB2>(sa_feat, 128 * 128, sa_aux, sw + Cfg::OFF_W2, Cfg::C2 * 128, sw + Cfg::OFF_W2A, d2,
                              idesc_bf16(128, Cfg::C2));
      commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    epilogue_repack<Cfg::C2>(d2 + lane_off, sa_feat, p);
    fence_async_smem();
    fence_before_sync();
    named_bar_sync(1 + slot, 128);

    // ---- layer 3 (transposed): D3[c3, p] = W3[c3, :] . H2[p, :] ; max over each 64-point group ----
    if (p == 0) {
      fence_after_sync();
      // no bias K-step here: max_p(x_p + b) = max_p(x_p) + b, the fp32 bias is added once per channel after the max
#pragma unroll
      for (int h = 0; h < Cfg::C3 / 128; ++h)
        issue_layer<Cfg::KB3, false>(sw + Cfg::OFF_W3 + h * (128 * 128), Cfg::C3 * 128, 0, sa_feat, 128 * 128, 0,
                                     d3 + h * 128, idesc_bf16(128, 128));
      commit(bar);
    }
    // ---- under layer 3: geometry of tile t+1 into the other block; raw coordinates of t+2, index of t+3 ----
    if (has_next) write_aux(sa_aux0 + (buf ^ 1) * (128 * 32), t + t_step, s0, s1, s2, c0, c1, c2);
    if (t + 2 * t_step < nt) {
      if (feat16) my_rowidx[buf * 128 + p] = (uint16_t)j_next;      // rows of tile t+2 (block [buf] is free: tile t's were read an iteration ago)
      load_raw(t + 2 * t_step, j_next, s0, s1, s2, c0, c1, c2);
    }
    if (t + 3 * t_step < nt) j_next = __ldg(idx + (int64_t)(t + 3 * t_step) * 128 + p);
    mbar_wait(bar, phase); phase ^= 1;
    fence_after_sync();
    // layer 3 no longer reads the operand tile: the feature rows of tile t+1 land under the max epilogue
    if (feat16 && early_feat && has_next) issue_features(t + t_step, my_rowidx + (buf ^ 1) * 128);
#pragma unroll
    for (int h = 0; h < Cfg::C3 / 128; ++h) {
      float m[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {                      // one 64-point group; independent max chains (ReLU floor = 0)
        if constexpr (Cfg::THREADS <= 256) {             // register budget allows both 32-column loads in flight
          uint32_t v[32], u[32];
          tmem_ld32(d3 + lane_off + h * 128 + g * 64, v);
          tmem_ld32(d3 + lane_off + h * 128 + g * 64 + 32, u);
          tmem_ld_wait();
          float a0 = -3.0e38f, a1 = -3.0e38f, a2 = -3.0e38f, a3 = -3.0e38f;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            a0 = max3(a0, __uint_as_float(v[q]), __uint_as_float(v[q + 1]));
            a1 = max3(a1, __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
            a2 = max3(a2, __uint_as_float(u[q]), __uint_as_float(u[q + 1]));
            a3 = max3(a3, __uint_as_float(u[q + 2]), __uint_as_float(u[q + 3]));
          }
          m[g] = fmaxf(fmaxf(max3(a0, a1, a2), a3) + b3[h], 0.f);     // relu(max + bias)
        } else {
          float a0 = -3.0e38f, a1 = -3.0e38f;
#pragma unroll
          for (int cc = 0; cc < 64; cc += 32) {
            uint32_t v[32];
            tmem_ld32(d3 + lane_off + h * 128 + g * 64 + cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              a0 = max3(a0, __uint_as_float(v[q]), __uint_as_float(v[q + 1]));
              a1 = max3(a1, __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
            }
          }
          m[g] = fmaxf(fmaxf(a0, a1) + b3[h], 0.f);                   // relu(max + bias)
        }
        if (EARLY_L1 && g == 0 && has_next) {
          // every thread has read columns [0, 64) of D3 (group 0): D1 of tile t+1 may overwrite them now, while the
          // second group's columns are still being reduced
          fence_before_sync();
          named_bar_sync(1 + slot, 128);
          if (p == 0) {
            fence_after_sync();
            issue_layer<Cfg::KB1>(sa_feat, 128 * 128, sa_aux0 + (buf ^ 1) * (128 * 32), sw, Cfg::C1 * 128, sw + Cfg::OFF_W1A,
                                  d1, idesc_bf16(128, Cfg::C1));
            commit(bar);
          }
        }
      }
      float* o = out + (int64_t)t * 2 * ld_out + out_col0 + h * 128 + p;
      o[0] = m[0];
      o[ld_out] = m[1];
    }
    if (feat16 && early_feat) asm volatile("cp.async.wait_group 0;" ::: "memory");
    buf ^= 1;
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ---- host-side weight packer -------------------------------------------------------------------
static inline uint16_t f2bf(float f) {                   // round-to-nearest-even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

template <class Cfg>
static void pack_weights(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                         const float* b3, int c_in, uint8_t* out) {
  memset(out, 0, Cfg::PACK_BYTES);
  auto put = [&](int off, uint32_t byte, float v) {
    const uint16_t h = f2bf(v);
    memcpy(out + off + byte, &h, 2);
  };
  auto aux_bias = [&](int off, int r, float b) {
    const uint16_t hi = f2bf(b);
    put(off, aux_off(r, 3), b);
    put(off, aux_off(r, 7), b - bf2f(hi));
  };
  // layer 1: xyz columns 0..2 -> aux (hi and lo copies), features (source columns 3..) -> SW128 blocks
  const int xyz_cols = 3, feat0 = c_in - Cfg::CF;        // feature k <-> W1 column feat0 + k
  for (int r = 0; r < Cfg::C1; ++r) {
    for (int c = 0; c < xyz_cols; ++c) {
      put(Cfg::OFF_W1A, aux_off(r, c), W1[r * c_in + c]);
      put(Cfg::OFF_W1A, aux_off(r, 4 + c), W1[r * c_in + c]);
    }
    aux_bias(Cfg::OFF_W1A, r, b1[r]);
    for (int k = 0; k < Cfg::CF; ++k) put((k >> 6) * Cfg::C1 * 128, sw128_off(r, k & 63), W1[r * c_in + feat0 + k]);
  }
  for (int r = 0; r < Cfg::C2; ++r) {
    aux_bias(Cfg::OFF_W2A, r, b2[r]);
    for (int k = 0; k < Cfg::C1; ++k) put(Cfg::OFF_W2 + (k >> 6) * Cfg::C2 * 128, sw128_off(r, k & 63), W2[r * Cfg::C1 + k]);
  }
  memcpy(out + Cfg::OFF_B3, b3, (size_t)Cfg::C3 * sizeof(float));
  for (int r = 0; r < Cfg::C3; ++r) {
    aux_bias(Cfg::OFF_W3A, r, b3[r]);
    for (int k = 0; k < Cfg::C2; ++k) put(Cfg::OFF_W3 + (k >> 6) * Cfg::C3 * 128, sw128_off(r, k & 63), W3[r * Cfg::C2 + k]);
  }
}

template <class Cfg>
static int launch_sa(const float* pts, int64_t n_clouds, int n_src, int64_t ld_pts, const void* feat_bf16,
                     const int32_t* idx, int n_centroids, const void* wpack, float* out, int64_t ld_out, int out_col0,
                     cudaStream_t stream) {
  static PerDeviceOnce once;            // one flag array per template instantiation
  if (once.first()) {
    cudaError_t e;
    if constexpr (Cfg::EARLY)
      e = cudaFuncSetAttribute(sa_mlp_max_early_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    else
      e = cudaFuncSetAttribute(sa_mlp_max_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("pdf_sa_mlp_max_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PDF_ERR_CUDA;
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = n_clouds * (n_centroids / 2);
  if (n_tiles >= (1ll << 31)) {
    set_error("pdf_sa_mlp_max_bf16: too many tiles (%lld)", (long long)n_tiles);
    return PDF_ERR_UNSUPPORTED;
  }
  int64_t grid = (n_tiles + Cfg::SLOTS - 1) / Cfg::SLOTS;
  if (grid > sms) grid = sms;
  if constexpr (Cfg::EARLY) {
    // level-2 feature copies under the max epilogue (1) or at the top of the iteration (0, measured faster with the two
    // slots of level 2: the early copies change how the slots' tensor phases interleave)
    static const int early_feat = getenv("PDF_SA_EARLY_FEAT") ? atoi(getenv("PDF_SA_EARLY_FEAT")) : 0;
    launch_pdl(sa_mlp_max_early_kernel<Cfg>, dim3((unsigned)grid), dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, stream, pts,
               n_src, ld_pts, reinterpret_cast<const uint16_t*>(feat_bf16), idx, n_centroids,
               reinterpret_cast<const uint8_t*>(wpack), out, ld_out, out_col0, n_tiles, early_feat);
  } else {
    launch_pdl(sa_mlp_max_kernel<Cfg>, dim3((unsigned)grid), dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, stream, pts, n_src,
               ld_pts, reinterpret_cast<const uint16_t*>(feat_bf16), idx, n_centroids,
               reinterpret_cast<const uint8_t*>(wpack), out, ld_out, out_col0, n_tiles);
  }
  return check_launch("pdf_sa_mlp_max_bf16");
}

}  // namespace pdf

extern "C" int64_t pdf_sa_pack_size(int c_in, int c1, int c2, int c3) {
  if (c_in == 3 && c1 == 64 && c2 == 64 && c3 == 128) return pdf::Sa1Cfg::PACK_BYTES;
  if (c_in == 131 && c1 == 128 && c2 == 128 && c3 == 256) return pdf::Sa2Cfg::PACK_BYTES;
  return -1;
}

extern "C" int pdf_sa_pack_weights_host(const float* W1, const float* b1, const float* W2, const float* b2,
                                        const float* W3, const float* b3, int c_in, int c1, int c2, int c3,
                                        void* out_host) {
  PDF_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3 && out_host, PDF_ERR_BAD_ARG, "pdf_sa_pack_weights_host: null pointer");
  if (c_in == 3 && c1 == 64 && c2 == 64 && c3 == 128)
    pdf::pack_weights<pdf::Sa1Cfg>(W1, b1, W2, b2, W3, b3, c_in, (uint8_t*)out_host);
  else if (c_in == 131 && c1 == 128 && c2 == 128 && c3 == 256)
    pdf::pack_weights<pdf::Sa2Cfg>(W1, b1, W2, b2, W3, b3, c_in, (uint8_t*)out_host);
  else {
    pdf::set_error("pdf_sa_pack_weights_host: unsupported channel plan (%d,%d,%d,%d)", c_in, c1, c2, c3);
    return PDF_ERR_UNSUPPORTED;
  }
  return PDF_OK;
}

extern "C" int pdf_sa_mlp_max_bf16(const float* pts, int64_t n_clouds, int n_src, int64_t ld_pts, int c_in,
                                   const void* feat_bf16, const int32_t* idx, int n_centroids, int k,
                                   const void* wpack, int c1, int c2, int c3, float* out, int64_t ld_out,
                                   int out_col0, void* stream) {
  if (n_clouds == 0) return PDF_OK;
  PDF_REQUIRE(pts && idx && wpack && out, PDF_ERR_BAD_ARG, "pdf_sa_mlp_max_bf16: null pointer");
  PDF_REQUIRE(n_clouds >= 0 && n_src > 0 && n_centroids > 0 && out_col0 >= 0 && out_col0 <= 4, PDF_ERR_BAD_ARG,
              "pdf_sa_mlp_max_bf16: bad size");
  PDF_REQUIRE(k == 64 && (n_centroids % 2) == 0, PDF_ERR_UNSUPPORTED,
              "pdf_sa_mlp_max_bf16: needs k == 64 and an even number of centroids (got k=%d, n=%d)", k, n_centroids);
  PDF_REQUIRE(ld_out >= out_col0 + c3, PDF_ERR_BAD_ARG, "pdf_sa_mlp_max_bf16: ld_out too small");
  if (n_clouds == 0) return PDF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  // early staging of the next tile (default); PDF_SA_EARLY=0 keeps the round-1 schedule for A/B timing.  The early
  // level-2 plan passes row indices through 16-bit shared cells.
  static const bool early_env = !(getenv("PDF_SA_EARLY") && atoi(getenv("PDF_SA_EARLY")) == 0);
  const bool early = early_env && n_src <= 65536;
  if (c_in == 3 && c1 == 64 && c2 == 64 && c3 == 128) {
    PDF_REQUIRE(ld_pts >= 3, PDF_ERR_BAD_ARG, "pdf_sa_mlp_max_bf16: ld_pts too small");
    if (!early)
      return pdf::launch_sa<pdf::Sa1CfgLate>(pts, n_clouds, n_src, ld_pts, nullptr, idx, n_centroids, wpack, out, ld_out,
                                             out_col0, s);
    return pdf::launch_sa<pdf::Sa1Cfg>(pts, n_clouds, n_src, ld_pts, nullptr, idx, n_centroids, wpack, out, ld_out,
                                       out_col0, s);
  }
  if (c_in == 131 && c1 == 128 && c2 == 128 && c3 == 256) {
    PDF_REQUIRE(feat_bf16 != nullptr ? ld_pts >= 3 : (ld_pts >= 132 && (ld_pts % 4) == 0), PDF_ERR_UNSUPPORTED,
                "pdf_sa_mlp_max_bf16: level-2 source rows must be [xyz,pad,128 features] with pitch %% 4 == 0 "
                "unless the features are given as bf16 rows");
    if (!early)
      return pdf::launch_sa<pdf::Sa2CfgLate>(pts, n_clouds, n_src, ld_pts, feat_bf16, idx, n_centroids, wpack, out, ld_out,
                                             out_col0, s);
    return pdf::launch_sa<pdf::Sa2Cfg>(pts, n_clouds, n_src, ld_pts, feat_bf16, idx, n_centroids, wpack, out, ld_out,
                                       out_col0, s);
  }
  pdf::set_error("pdf_sa_mlp_max_bf16: unsupported channel plan (%d,%d,%d,%d)", c_in, c1, c2, c3);
  return PDF_ERR_UNSUPPORTED;
}
