/*
 * pdfnet_b200 — C ABI of the B200-native depth-branch / fusion / MANO hot path.
 *
 * The reference (zijinxuxu/PDFNet) has no FFI layer: its boundary is Python call
 * level.  Every entry point below replaces the body of one
 * reference function (cited as path:line relative to the reference root); the
 * Python mirror in pdfnet_b200/*.py binds them with ctypes and keeps the
 * reference's names, argument meaning and error behaviour.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *    name ends in _host.  The caller owns every buffer; the library never
 *    allocates or frees persistent device memory.
 *  - every call enqueues asynchronously on `stream` (a cudaStream_t passed as
 *    void*) and never synchronises.
 *  - return 0 on success, a negative pdf_status on failure; the message is
 *    available from pdf_last_error() (thread-local).  Nothing ever exit()s.
 *  - no CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef PDFNET_B200_H
#define PDFNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PDF_OK = 0,
  PDF_ERR_BAD_ARG = -1,       /* null pointer, non-positive size, bad enum        */
  PDF_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels are built for     */
  PDF_ERR_CUDA = -3           /* launch / runtime error (message has the detail)  */
} pdf_status;

/* activation codes for pdf_linear_f32 */
enum { PDF_ACT_NONE = 0, PDF_ACT_RELU = 1, PDF_ACT_LEAKY01 = 2 };
/* OR-ed onto the `act` argument of pdf_gemm_bf16: out_img is written as a SPLIT image [hi | hi | lo]
 * (out_kb = 3 x k-blocks of the output matrix), i.e. directly as the fp32-accurate M operand of the next GEMM. */
enum { PDF_GEMM_OUT_SPLIT = 256,
       /* also OR-ed onto `act`: run the half-footprint configuration (2 operand stages, 256 TMEM columns, two CTAs
        * per SM) - for short GEMMs that come in concurrent pairs on two streams; single-accumulator ROW mode only */
       PDF_GEMM_LIGHT = 512 };
/* OR-ed onto the `relu` argument of pdf_bn_act_fwd / pdf_bn_act_bwd / pdf_bn_maxpool_bwd: the image output is ONE plain
 * bf16 tile image (C/64 k-blocks per row tile) instead of the split image [hi | hi | lo] - the bf16 training mode */
enum { PDF_BN_PLAIN_IMAGE = 2 };
#define PDF_MANO_VT_PITCH 2336
/* epilogue modes for pdf_linear_f32 */
enum {
  PDF_EPI_STORE = 0,     /* Y = act(acc + bias)                                   */
  PDF_EPI_SFT_SCALE = 1, /* Y = F * ((acc + bias) + 1)      SFTLayer scale branch */
  PDF_EPI_ACCUM = 2,     /* Y = Y + (acc + bias)            SFTLayer shift branch */
  PDF_EPI_GROUP_MAX = 3  /* Y[m/G] = max over G consecutive rows of act(acc+bias);
                            act must be RELU, Y must be zero-filled by the caller */
};

int pdf_version(void);
const char* pdf_last_error(void);
/* number of kernel launches enqueued by this library since load (bench bookkeeping) */
int64_t pdf_launch_count(void);

/* kNN-then-radius-mask neighbour search.
 * Replaces the index half of group_points (lib/utils/utils.py:142-151) and
 * group_points_2 (:169-179): d2 = (dx*dx + dy*dy) + dz*dz in fp32 without FMA,
 * diff = p_j - c_i; the k smallest per centroid; neighbours with d2 > r2 are
 * replaced by the centroid's own index.  Centroids are points 0..n_centroids-1.
 * xyz element (cloud b, point j, channel c) is xyz[b*stride_cloud + j*stride_point
 * + c*stride_ch] so both [B,N,C] and channel-major [B,C,N] inputs work.
 * idx_out: int32 [n_clouds, n_centroids, k]; the order inside a group is unspecified (as with the
 * reference's topk(sorted=False)) but deterministic; exact distance ties at the k-th neighbour
 * are resolved towards the lower index.  Supported: k <= n_points <= 1024, r2 >= 0. */
int pdf_knn_ball(const float* xyz, int64_t n_clouds, int n_points, int n_centroids, int k, float r2,
                 int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                 int32_t* idx_out, void* stream);

/* Farthest point sampling in selection order.
 * Replaces the loop of InterHandDataset.farthest_point_sampling_fast
 * (lib/datasets/interhand.py:159-175): fp32 (dx2+dy2)+dz2, first-occurrence
 * argmax, min-distance lowered only where it is > 1e-8.  start_idx[b] is the
 * injected first sample (:159 draws it from np.random).  idx_out int32
 * [n_clouds, n_sample].  Supported: n_points <= 4096, 1 <= n_sample <= n_points. */
int pdf_fps(const float* xyz, int64_t n_clouds, int n_points, int n_sample, const int32_t* start_idx,
            int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
            int32_t* idx_out, void* stream);

/* Row gather from an NCHW map: out[b,i,c] = feat[b / clouds_per_frame, c, ind[b,i]].
 * Replaces _tranpose_and_gather_feat (lib/models/utils.py:22-26) without the
 * full-map permute+copy.  feat fp32 [n_frames,C,HW]; ind int64 [n_clouds,n] with
 * row pitch ind_stride; out fp32 [n_clouds,n,C].  Indices outside [0,HW) fail
 * the call's contract (checked on the host side of the Python mirror). */
int pdf_gather_nchw(const float* feat, int64_t n_clouds, int clouds_per_frame, int C, int64_t HW,
                    const int64_t* ind, int n, int64_t ind_stride, float* out, void* stream);

/* Three-level pixel->point pyramid gather with the level-0 SFT fused.
 * Replaces intaghand_encoder.py:120-128 (+ SFTLayer sft0, :205-219):
 *   e0 = l0[:, :, choose]                      (3 channels, all n_points)
 *   pts0 = xyz * (scale(e0) + 1) + shift(e0)   (fp32; feeds the neighbour search)
 *   cond1 = l1[:, :, (choose//R//2)*(R//2) + (choose%R)//2]  first n1 points
 *   cond2 = l2[:, :, (choose//R//4)*(R//4) + (choose%R)//4]  first n2 points
 * sft0_params: 48 floats = scale_conv0 W[3x3],b[3], scale_conv1 W,b, shift_conv0
 * W,b, shift_conv1 W,b (row-major [out][in]).  l0 [F,3,R,R], l1 [F,C1,R/2,R/2],
 * l2 [F,C2,R/4,R/4] fp32.  pts0 [n_clouds,n_points,3], cond1 [n_clouds,n1,C1],
 * cond2 [n_clouds,n2,C2] fp32. */
int pdf_pyramid_gather(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                       int n_points, int n1, int n2, int R,
                       const float* l0, const float* l1, int C1, const float* l2, int C2,
                       const float* sft0_params, float* pts0, float* cond1, float* cond2, void* stream);

/* Channels-last (NHWC) forms of the two gathers above (hand-off from the RGB neck,
 * intaghand_encoder.py:711,715,741-744, in torch.channels_last memory format): feat / l0 / l1 / l2 are
 * [F, H, W, C] so each point reads C contiguous floats.  Results are bit-identical to the NCHW forms. */
int pdf_gather_nhwc(const float* feat, int64_t n_clouds, int clouds_per_frame, int C, int64_t HW,
                    const int64_t* ind, int n, int64_t ind_stride, float* out, void* stream);
int pdf_pyramid_gather_nhwc(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                            int n_points, int n1, int n2, int R,
                            const float* l0, const float* l1, int C1, const float* l2, int C2,
                            const float* sft0_params, float* pts0, float* cond1, float* cond2, void* stream);
/* bf16 channels-last pyramid (what an autocast / channels_last RGB neck emits; SURVEY 8d "bf16 features",
 * 8f row f4): l0 [F,R,R,3], l1 [F,R/2,R/2,C1], l2 [F,R/4,R/4,C2] bf16.  pts0 as pdf_pyramid_gather (SFT0 in
 * fp32 on the widened bf16 values).  The condition rows are written DIRECTLY as the bf16 tile images the SFT
 * GEMMs read: cond1_img = image of [n_clouds*n1, C1], cond2_img = image of [n_clouds*n2, C2]
 * (pdf_image_bytes each; C1, C2 multiples of 64; row counts multiples of 128). */
int pdf_pyramid_gather_bf16(const float* xyz, const int64_t* choose, int64_t n_clouds, int clouds_per_frame,
                            int n_points, int n1, int n2, int R, const void* l0, const void* l1, int C1,
                            const void* l2, int C2, const float* sft0_params, float* pts0, void* cond1_img,
                            void* cond2_img, void* stream);
/* Zero-copy hand-off of a HOST-resident pyramid: l0 / l1 / l2 of pdf_pyramid_gather_bf16 may each be the device
 * alias of a page-locked, mapped host buffer (this call; cudaHostGetDevicePointer).  The gather then reads only the
 * pixels `choose` selects (2 x (1024 x 6 + 512 x 128 + 128 x 512) B of a 4.6 MB frame) over the PCIe link, in
 * place of a dense host->device copy of the maps (the reference moves whole maps: `.cuda()` of the batch,
 * lib/trains/base_trainer.py:66-68, then _tranpose_and_gather_feat copies them once more, lib/models/utils.py:12-26).
 * Returns PDF_ERR_BAD_ARG when `host` is not mapped page-locked memory. */
int pdf_host_device_pointer(const void* host, void** device_out);


/* Grouping gather: out[b,g,j,c] = pts[b, idx[b,g,j], c] - (c < 3 ? pts[b,g,c] : 0).
 * Replaces lib/utils/utils.py:153-158 and :181-186.  pts addressed with
 * (stride_cloud, stride_point, stride_ch) as in pdf_knn_ball; out fp32
 * [n_clouds, n_centroids, k, C] with row pitch ld_out >= C;
 * center (optional, may be null) fp32 [n_clouds, n_centroids, 3]. */
int pdf_group_gather(const float* pts, int64_t n_clouds, int n_centroids, int k, int C,
                     int64_t stride_cloud, int64_t stride_point, int64_t stride_ch,
                     const int32_t* idx, float* out, int64_t ld_out, float* center, void* stream);

/* Y = epilogue(X[M,K] * W[N,K]^T + bias[N]) in fp32 (FFMA, fp32 accumulate).
 * The shared point-MLP layers (1x1 conv + folded BN + ReLU, intaghand_encoder.py:
 * 48-103), the SFT 1x1 convs (:205-219) and mano_head (:630-643) are all this.
 * lda/ldf/ldy are row pitches in elements.  See the PDF_EPI_* modes; `group` is
 * the row-group size for PDF_EPI_GROUP_MAX. */
int pdf_linear_f32(const float* X, int64_t lda, const float* W, int64_t ldw, const float* bias,
                   int64_t M, int N, int K, int act, int epilogue, int group,
                   const float* F, int64_t ldf, float* Y, int64_t ldy, void* stream);

/* Fused set-abstraction stage on tensor cores (tcgen05 + TMEM, bf16 operands,
 * fp32 accumulate): neighbour gather + centroid subtraction + 3-layer shared
 * point-MLP (folded BN, ReLU) + max over the k neighbours, one pass, no
 * intermediate in HBM.  Replaces utils.py:153-158/181-186 + netR_1 / netR_2
 * (intaghand_encoder.py:48-84,132,143).
 * pts fp32 [n_clouds, n_src, ld_pts] (xyz = columns 0..2; for c_in = 131 the 128 feature
 * columns start at column 4, or, when feat_bf16 is non-null, are read from it instead:
 * bf16 rows [n_clouds, n_src, 128] copied with cp.async straight into the operand tile);
 * idx int32 [n_clouds, n_centroids, 64]; wpack = pdf_sa_pack_weights image;
 * out fp32 [n_clouds, n_centroids, ld_out] columns out_col0 .. out_col0+c3-1.
 * Supported (c_in,c1,c2,c3): (3,64,64,128) and (131,128,128,256); k = 64. */
int pdf_sa_mlp_max_bf16(const float* pts, int64_t n_clouds, int n_src, int64_t ld_pts, int c_in,
                        const void* feat_bf16, const int32_t* idx, int n_centroids, int k,
                        const void* wpack, int c1, int c2, int c3,
                        float* out, int64_t ld_out, int out_col0, void* stream);
/* Size in bytes of, and host-side packer for, the weight image above: folded
 * fp32 weights W1[c1,c_in],b1, W2[c2,c1],b2, W3[c3,c2],b3 -> bf16 tiles in the
 * UMMA shared-memory layout + fp32 biases.  Pure host code (no CUDA call). */
int64_t pdf_sa_pack_size(int c_in, int c1, int c2, int c3);
int pdf_sa_pack_weights_host(const float* W1, const float* b1, const float* W2, const float* b2,
                             const float* W3, const float* b3, int c_in, int c1, int c2, int c3,
                             void* out_host);

/* ---- streaming bf16 GEMM on tcgen05 over "tile images" -------------------------------------
 * A tile image stores a matrix [rows, K] in bf16 as [row-tile of 128 rows][k-block of 64
 * columns] blocks of 16 KB, each block in the K-major 128-byte-swizzle layout the tensor core
 * reads from shared memory (rows / K zero-padded to multiples of 128 / 64).
 * Replaces the dense layers behind the set-abstraction stages: SFT 1x1 convs
 * (intaghand_encoder.py:205-219), netR_3 + MaxPool (:86-103,152-154) and the fusion
 * SFT(1024,1024) (:809). */
int64_t pdf_image_bytes(int64_t rows, int cols);
/* host-side packer: fp32 W[rows, cols] (row pitch ld) -> bf16 tile image (pure host code).
 * split != 0 packs [hi | lo | hi] (3x the k-blocks, 3x pdf_image_bytes): against an activation
 * image written with split != 0 ([hi | hi | lo]) the GEMM then accumulates a_hi*w_hi + a_hi*w_lo +
 * a_lo*w_hi, i.e. fp32-accurate products (error ~2^-16 relative) on the bf16 tensor cores. */
int pdf_pack_image_host(const float* W, int64_t rows, int cols, int64_t ld, int split, void* out_host);
/* device: fp32 rows X[M, ld], columns [col0, col0+K) -> k-blocks [kb0, kb0+ceil(K/64)) of an image
 * that has kb_total k-blocks per row-tile; padding rows/columns are written as zeros */
int pdf_rows_to_image(const float* X, int64_t ld, int64_t M, int col0, int K, void* img, int kb_total, int kb0,
                      int split, void* stream);
/* D[m,n] = sum_k Mop[m,k] * Nop[n,k] over KB k-blocks, 128x128 tiles, fp32 accumulate in TMEM.
 * colmax = 0 (ROW epilogue, thread = M row): y = act(D + bias0[n]); with kb_split > 0 the
 *   k-blocks below / from kb_split accumulate separately and y = F*(D0+bias0+1) + (D1+bias1)
 *   (SFT modulation).  Output: fp32 rows out_f32[m, tile_col + n] (m < rows_valid) and/or a bf16
 *   tile image (out_kb k-blocks per row-tile) and/or bf16 rows out_bf16[m, tile_col -
 *   bf16_col_off + n] (row pitch ld_bf16 elements).  tile_desc_host: int32 [n_tiles][3] =
 *   {first fp32 column of this N-tile in F/out_f32, valid columns (<=128), first output k-block}.
 * colmax = 1: M operand = weights (rows = channels), N operand = activations, one N-tile = the
 *   128 points of one cloud: out_max[n_tile, m] = relu(max_n D[m,n] + bias0[m]).
 * xyz_w != NULL (XYZ mode, one N-tile whose 128 columns are [64 scale-hidden | 64 shift-hidden] of
 *   an SFTLayer): besides the regular outputs the epilogue applies the SFT modulation to the 3 xyz
 *   channels in fp32, x[m,c] = x[m,c]*(w1s[c].h_s + b1s[c] + 1) + (w1h[c].h_h + b1h[c]) with
 *   xyz_w = {w1s[3][64], w1h[3][64], b1s[3], b1h[3]} and x = xyz_x[m*xyz_ld + c]. */
int pdf_gemm_bf16(const void* m_img, int m_tiles, int m_kb, const void* n_img, int n_tiles, int n_kb, int KB,
                  int kb_split, int colmax, const float* bias0, const float* bias1, int act, float* out_f32,
                  int64_t ld_out, int64_t rows_valid, const float* F, int64_t ldf, void* out_img, int out_kb,
                  void* out_bf16, int64_t ld_bf16, int bf16_col_off, const int32_t* tile_desc_host, float* out_max,
                  int64_t ld_max, const float* xyz_w, float* xyz_x, int64_t xyz_ld, void* stream);
/* Transposed tile image for reductions over rows (weight gradients): X[M, ld] columns
 * [col0, col0+C) -> bf16 image of X^T cut into batches of Mc rows of X (Mc % 64 == 0): layout
 * [batch][row-tile of 128 channels][k-block of 64 rows], zero padded; split as in pdf_rows_to_image
 * (1 = [hi|hi|lo], 2 = [hi|lo|hi], tripling the k-blocks).  pdf_rows_to_image accepts split = 2 too. */
int pdf_rows_to_image_t(const float* X, int64_t ld, int64_t M, int col0, int C, void* img, int64_t Mc, int split,
                        void* stream);
/* Batched ROW-mode GEMM without bias/activation: for every batch b, out[b][m, n] = sum_k Mop_b[m,k] *
 * Nop_b[n,k]; operands advance by *_batch_stride bytes, the fp32 output by out_batch_stride floats.
 * With pdf_rows_to_image_t images this is the split-K weight gradient dW = dY^T X (autograd of the
 * 1x1 convs, intaghand_encoder.py:48-103,205-219); the caller sums the per-batch partials. */
int pdf_gemm_bf16_batched(const void* m_img, int m_tiles, int m_kb, int64_t m_batch_stride, const void* n_img,
                          int n_tiles, int n_kb, int64_t n_batch_stride, int KB, int batches, float* out_f32,
                          int64_t ld_out, int64_t out_batch_stride, int64_t rows_valid,
                          const int32_t* tile_desc_host, void* stream);
/* ROW-mode GEMM for TWO groups of rows with their own weights and biases in ONE launch - the left and the right hand
 * of the GCN decoder (intaghand_decoder.py:180-242: graph_left / graph_right, L_self_attn_layer / R_self_attn_layer,
 * ffL / ffR have identical shapes and different parameters).  m_img holds 2 * group_m_tiles row tiles (every group
 * padded to whole 128-row tiles), n_img / bias are the FIRST group's weight image / bias (padded to n_tiles * 128)
 * and the second group's lie n_group_stride bytes / bias_group_stride floats further.  rows_valid counts inside each
 * group; out_f32 / out_img cover all 2 * group_m_tiles * 128 rows.  act takes the PDF_GEMM_OUT_SPLIT / PDF_GEMM_LIGHT
 * flags like pdf_gemm_bf16.  Single accumulator (no SFT / xyz / COLMAX modes). */
int pdf_gemm_bf16_grouped(const void* m_img, int group_m_tiles, int m_kb, const void* n_img, int n_tiles, int n_kb,
                          int64_t n_group_stride, int KB, const float* bias, int64_t bias_group_stride, int act,
                          float* out_f32, int64_t ld_out, int64_t rows_valid, void* out_img, int out_kb,
                          const int32_t* tile_desc_host, void* stream);
/* Weight gradient straight from ROW tile images (no transposed copy): out[b][ca, cb] = sum over the rows of
 * batch b of A[r, ca] * B[r, cb], A / B = images of dY / X as written by pdf_rows_to_image (split = 1:
 * [hi|hi|lo], three products per fp32 product; split = 0: plain bf16).  The tensor core reads the K-major
 * blocks as MN-major operands (instruction-descriptor a_major = b_major = 1).  Batches are runs of
 * tiles_per_batch row tiles (split-K); the caller sums out over b (batch_stride floats apart).
 * Rows >= `rows` inside the last tile must be zero in both images (pdf_rows_to_image guarantees it). */
int pdf_gemm_tn_bf16(const void* a_img, int ca, const void* b_img, int cb, int64_t rows, int split,
                     int tiles_per_batch, float* out, int64_t ld_out, int64_t batch_stride, void* stream);
/* SFT on the three xyz channels of level 1 in full fp32 (they feed the level-2 neighbour
 * search): x[m,c] = x[m,c]*(scale_c+1)+shift_c for c < 3; cond fp32 [M,cc]; conv weights as in
 * SFTLayer ([out,in] row-major; only rows 0..2 of the second convs are read). cc must be 64. */
int pdf_sft_xyz_f32(const float* cond, int64_t M, int cc, const float* w0s, const float* b0s, const float* w1s,
                    const float* b1s, const float* w0h, const float* b0h, const float* w1h, const float* b1h,
                    float* x, int64_t ldx, void* stream);

/* Centre features evaluated only where they are used.  The reference runs two 3x3
 * convolutions over the whole (R/4)^2 map, center_feat_up1(center_feat_up0(x0)), and then keeps the
 * 2 pixels at `ind` (intaghand_encoder.py:627-628,790-792).  This builds, for every (frame, hand),
 * the im2col rows of the 3x3 conv0 outputs that conv1 needs at the centre pixel:
 * rows fp32 [B*2*9, 9*C], row (b,hand,pos) = 3x3xC input patch around output position pos (K order:
 * tap-major, channel-minor; zero padding; all-zero rows for positions outside the map).  The two
 * convolutions are then two GEMMs (pdf_linear_f32 / pdf_gemm_bf16) with re-laid weights.
 * x0 fp32 [B,C,H,W]; ind int64 [B,2] flat index on the HxW map. */
int pdf_center_im2col(const float* x0, const int64_t* ind, int64_t B, int C, int H, int W, float* rows, void* stream);

/* Depth back-projection xyz[b,:,v,u] = (Kinv[b] * [u,v,1]) * depth[b,v,u].
 * Replaces get_points_coordinate (lib/utils/utils.py:251-262).  depth fp32
 * [B,H,W], Kinv fp32 [B,3,3] (inverse intrinsics, computed by the caller as the
 * reference does with np.linalg.inv, :269), xyz fp32 [B,3,H,W]. */
int pdf_backproject(const float* depth, const float* Kinv, int64_t B, int H, int W, float* xyz, void* stream);

/* Batched per-hand cloud construction from raw depth (device-side depth2pcl).
 * Replaces depth2pcl (intaghand_encoder.py:369-491) / the dataset twin
 * (lib/datasets/interhand.py:758-797) for every frame of a batch at once:
 * noise gate 0.2<z<2.5, hand mask > 0.5, mean z of the non-zero pixels, window
 * mean +- 0.08 clipped to [0.2,2.5], candidate pixels; < min_pixels -> zeros,
 * > n_points -> the n_points candidates with the smallest subset_keys (pixel
 * order kept), else wrap-pad; final order choose[i] = sel[perm[i]]; cloud =
 * back-projected masked depth at choose.
 * depth fp32 [B,H,W]; mask fp32 [B,2,H,W] (channel 0 = right hand, 1 = left,
 * :376-377); Kinv fp32 [B,3,3]; valid fp32 [B,2] (0 = left, 1 = right);
 * subset_keys int32 [B,2,H*W] (distinct per row; may be null when no hand can
 * exceed n_points); perm int32 [B,2,n_points] (null = identity).
 * choose int64 [B,2,n_points] (row 0 = left), cloud fp32 [B,2,n_points,3],
 * n_cand int32 [B,2] (number of candidate pixels, for diagnostics).
 * Supported: n_points == 1024, H*W <= 640*640. */
int pdf_depth2pcl(const float* depth, const float* mask, const float* Kinv, const float* valid,
                  const int32_t* subset_keys, const int32_t* perm, int64_t B, int H, int W,
                  int n_points, int min_pixels, int64_t* choose, float* cloud, int32_t* n_cand, void* stream);
/* Same cloud builder with (a) the hand masks as fp32 (mask_is_u8 = 0) or uint8 / bool (mask_is_u8 = 1: any
 * non-zero byte is "mask > 0.5", 4x fewer bytes across PCIe and HBM) and (b) generated randomness: when
 * subset_keys / perm are null they are replaced by counter-based functions of (seed, cloud = 2*frame + hand,
 * pixel / slot): key = murmur-style hash (unsigned order, ties -> lower pixel), perm = a 4-round Feistel
 * bijection of [0,1024).  pdf_depth2pcl_host_randomness materialises exactly these on the host (keys as int32
 * in the signed order pdf_depth2pcl expects), so a caller - and the parity tests - can hand the same
 * randomness to the reference's depth2pcl (intaghand_encoder.py:418-427, np.random.shuffle x2). */
int pdf_depth2pcl_seeded(const float* depth, const void* mask, int mask_is_u8, const float* Kinv, const float* valid,
                         const int32_t* subset_keys, const int32_t* perm, uint32_t seed, int64_t B, int H, int W,
                         int n_points, int min_pixels, int64_t* choose, float* cloud, int32_t* n_cand, void* stream);
int pdf_depth2pcl_host_randomness(uint32_t seed, int64_t n_clouds, int64_t npx, int32_t* keys_out_host,
                                  int32_t* perm_out_host);


/* MANO linear blend skinning, one hand per CTA, fp32.
 * Replaces ManoLayer.forward with use_pca=False (lib/models/networks/manolayer.py:
 * 257-334) including rodrigues_batch (:32-48).  Tables (device, fp32):
 *   v_template [778*3]; shapedirs_t [10, 778*3] and posedirs_t [135, 778*3]
 *   (basis index outermost so a CTA reads them coalesced); j_template [16,3] =
 *   J_regressor*v_template and j_shapedirs [16,3,10] = J_regressor*shapedirs
 *   (precomputed on the host in fp64); weights_t [16,778].
 * Inputs: root [n,3] and pose [n,45] axis-angle, shape [n,10], trans [n,3] or
 * null, scale [n] or null.  tip_idx_host: the 5 finger-tip vertex ids (:305-308).
 * center_idx < 0 disables centring (:313-316).  new_skel as :328-332.
 * v_tpose (optional, may be null): blend-shaped rest vertices, rows of 778*3 floats with row pitch
 * PDF_MANO_VT_PITCH (= 2336: 16-byte aligned rows, so a GEMM can write them with vector stores), computed beforehand as ONE
 * dense GEMM over all hands (pdf_mano_pose_feature + pdf_linear_f32); when null the kernel
 * evaluates the blend shapes itself.
 * Outputs v [n,778,3], j [n,21,3] (joints in the reference's new_order, :110-115). */
int pdf_mano_lbs(const float* v_template, const float* shapedirs_t, const float* posedirs_t,
                 const float* j_template, const float* j_shapedirs, const float* weights_t,
                 const float* root, const float* pose, const float* shape, const float* trans,
                 const float* scale, int64_t n, const int32_t* tip_idx_host, int center_idx, int new_skel,
                 const float* v_tpose, float* v, float* j, void* stream);
/* Same, for layers built with use_pca=True (manolayer.py:266-267, the dataset layers of
 * interhand.py:192,220-223): the root rotation arrives as a 3x3 MATRIX root_mat [n,9] and is used as is
 * (:285); pose is the axis-angle vector after pca2axis (:159-162, a [n,ncomps]x[ncomps,45] product done by
 * the caller). */
int pdf_mano_lbs_rootmat(const float* v_template, const float* shapedirs_t, const float* posedirs_t,
                         const float* j_template, const float* j_shapedirs, const float* weights_t,
                         const float* root_mat, const float* pose, const float* shape, const float* trans,
                         const float* scale, int64_t n, const int32_t* tip_idx_host, int center_idx, int new_skel,
                         const float* v_tpose, float* v, float* j, void* stream);
/* rodrigues_batch (manolayer.py:32-48): axis [n,3] -> rot [n,3,3], theta = |a| + 1e-8. */
int pdf_rodrigues(const float* axis, int64_t n, float* rot, void* stream);
/* joints = full_regressor @ verts (Mano_model.py:246-247,309-323; demo.py:217-218, simplified.py:431-434):
 * reg [n_joints,778] dense fp32 (J_regressor + one-hot tip rows, reordered), verts [n,778,3] ->
 * joints [n,n_joints,3]. */
int pdf_joint_regress(const float* reg, int n_joints, const float* verts, int64_t n, float* joints, void* stream);
/* Both hands of every frame in ONE launch: hands laid out (frame, side), even = left, odd = right.
 * tables_left / tables_right: host arrays of the 6 device table pointers in the order of
 * pdf_mano_lbs (v_template, shapedirs_t, posedirs_t, j_template, j_shapedirs, weights_t).
 * Replaces the two per-side ManoLayer calls of CtdetLoss.origforward (lib/trains/simplified.py:733-736). */
int pdf_mano_lbs_pair(const float* const* tables_left, const float* const* tables_right, const float* root,
                      const float* pose, const float* shape, const float* trans, const float* scale, int64_t n,
                      const int32_t* tips_left_host, const int32_t* tips_right_host, int center_idx, int new_skel,
                      const float* v_tpose, float* v, float* j, void* stream);
/* X[h] = [shape(10) | (rodrigues(pose_j) - I) for the 15 joints (135)], fp32 [n,145] (manolayer.py:274-281) */
int pdf_mano_pose_feature(const float* pose, const float* shape, int64_t n, float* X, void* stream);

/* Split_coeff (lib/models/hand3d/Mano_render.py:160-194, non-PCA) for one hand:
 * theta [n,ld_theta] (61 used columns starting at col0), index int64 [n], K [n,3,3];
 * writes root [n,3], pose [n,45], shape [n,10] (zeros: betas*0), trans [n,3].
 * pair != 0: rows are (frame, side) with one 122-vector per row (point2mano_left / point2mano_right,
 * simplified.py:722-732): even rows read the left slice [0,61), odd rows the right slice [61,122),
 * and K is indexed per frame (K [n/2,3,3]). */
int pdf_split_coeff(const float* theta, int64_t ld_theta, int col0, int pair, const int64_t* index, const float* K,
                    int64_t n, int input_res, int down_ratio,
                    float* root, float* pose, float* shape, float* trans, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-mode (BASELINE cfg5) entry points: fp32, rows [M, C] with a free row pitch.
 * Together with pdf_linear_f32 (data gradient: dX = dY * W, call it with W^T) these are the
 * backward of PointNet_Plus.forward / SFTLayer.forward (intaghand_encoder.py:118-159, 205-219)
 * that the reference gets from torch.autograd.  `sums` buffers are caller-owned fp64 scratch.
 * ------------------------------------------------------------------------------------------- */

/* Per-channel shifted sums over the M rows, d = x - x[0,c]: sums[0:C] = sum d, sums[C:2C] = sum d^2
 * (nn.BatchNorm2d batch statistics, intaghand_encoder.py:52,57,62 ...; the shift keeps the
 * variance free of cancellation). */
int pdf_bn_stats(const float* X, int64_t ldx, int64_t M, int C, double* sums, void* stream);
/* mean / rstd from pdf_bn_stats of the same X (biased variance, eps inside the sqrt) and the
 * running-stat update running = (1-momentum)*running + momentum*batch (unbiased variance), as
 * torch.nn.functional.batch_norm(training=True).  running_* may be null. */
int pdf_bn_finalize(const double* sums, const float* X, int64_t M, int C, float eps, float momentum,
                    float* running_mean, float* running_var, float* mean, float* rstd, void* stream);
/* Y = [relu]((X - mean) * rstd * gamma + beta).  Y_img (optional, C % 64 == 0): Y written as the split-bf16 tile
 * image of the next layer's GEMM (zero rows up to the next multiple of 128); Y (fp32 rows) may then be null. */
int pdf_bn_act_fwd(const float* X, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, int relu, int64_t M, int C, float* Y, int64_t ldy, void* Y_img, void* stream);
/* Backward of pdf_bn_act_fwd: g = dY * [Y > 0]; sums[0:C] = dbeta = sum g, sums[C:2C] = dgamma =
 * sum g*xhat; dX = gamma*rstd*(g - dbeta/M - xhat*dgamma/M).  dX may alias dY.  Y may be null when
 * beta is given: the ReLU mask is then recomputed from X with the forward's own expression (bit-identical),
 * which saves reading Y twice.  dX_img (optional; C % 64 == 0, Y null, beta given): dX written as the
 * split-bf16 tile image the gradient GEMMs read (zero rows up to the next multiple of 128); dX (fp32) may
 * then be null. */
int pdf_bn_act_bwd(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* X, int64_t ldx,
                   const float* mean, const float* rstd, const float* gamma, const float* beta, int relu, int64_t M,
                   int C, double* sums, float* dX, int64_t lddx, void* dX_img, void* stream);
/* sums[0:C] = column sums of A (bias gradients) */
int pdf_col_sum(const float* A, int64_t lda, int64_t M, int C, double* sums, void* stream);
/* dX = dY * act'(Y) for the PDF_ACT_* enum (Y is the activation OUTPUT); dX may alias dY */
int pdf_act_bwd(const float* dY, int64_t lddy, const float* Y, int64_t ldy, int act, int64_t M, int C, float* dX,
                int64_t lddx, void* stream);
/* out = fea * (scale + 1) + shift (SFTLayer.forward :219) and its backward:
 * dfea = dout*(scale+1), dscale = dout*fea (dshift = dout). */
int pdf_sft_modulate(const float* fea, int64_t ldf, const float* scale, int64_t lds, const float* shift, int64_t ldh,
                     int64_t M, int C, float* out, int64_t ldo, void* stream);
int pdf_sft_modulate_bwd(const float* dout, int64_t ldd, const float* fea, int64_t ldf, const float* scale,
                         int64_t lds, int64_t M, int C, float* dfea, int64_t lddf, float* dscale, int64_t ldds,
                         void* stream);
/* Weight gradient C[N,K] = A[M,N]^T * B[M,K] (A = dY, B = layer input); C is overwritten. */
int pdf_linear_tn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int N, int K, float* C,
                      int64_t ldc, void* stream);
/* Streaming forms of a linear layer with K <= 4 input channels (netR_1[0]: 3 -> 64, intaghand_encoder.py:50)
 * and of its two gradients (12 B in / 4N B out per row; the 64x64-tile kernels waste 95 % of a tile there):
 *   mode 0: out[M,N] = A[M,K] * B[N,K]^T + bias      (forward; N % 4 == 0, 16-byte aligned output rows)
 *   mode 1: out[M,K] = A[M,N] * B[N,K]               (data gradient, A = dY, B = W)
 *   mode 2: out[N,K] = A[M,N]^T * B[M,K]             (weight gradient, A = dY, B = X; out is overwritten) */
int pdf_linear_smallk_f32(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias,
                          int64_t M, int N, int K, float* out, int64_t ldo, void* stream);
/* nn.MaxPool2d over groups of G consecutive rows (intaghand_encoder.py:63,81,99) and its
 * backward: the FIRST maximum of each (group, channel) receives dOut, all other rows 0.  arg_out (optional,
 * uint8 [groups, C], G <= 256): row index of that maximum inside its group. */
int pdf_group_max(const float* Y, int64_t ldy, int G, int64_t groups, int C, float* out, int64_t ldo,
                  uint8_t* arg_out, void* stream);
/* BatchNorm(+ReLU) backward of a layer whose output feeds the max-pool directly (netR_x[6..8] -> MaxPool,
 * intaghand_encoder.py:60-63,78-81,96-99): the incoming gradient is dOut at the argmax row of each (group,
 * channel) and zero elsewhere, so it is never materialised; sums (dbeta | dgamma) are reduced over the argmax
 * rows only and dX is written as the split-bf16 tile image.  arg = pdf_group_max's arg_out (G <= 256). */
int pdf_bn_maxpool_bwd(const float* dOut, int64_t lddo, const uint8_t* arg, int G, const float* X, int64_t ldx,
                       const float* mean, const float* rstd, const float* gamma, const float* beta, int relu, int64_t M,
                       int C, double* sums, void* dX_img, void* stream);
int pdf_group_max_bwd(const float* Y, int64_t ldy, const float* dOut, int64_t lddo, int G, int64_t groups, int C,
                      float* dY, int64_t lddy, void* stream);
/* Backward of pdf_group_gather: dPts[b, idx[b,g,j], c] += dG[b,g,j,c]; dPts[b,g,c<3] -= dG[b,g,j,c].
 * dG fp32 [n_clouds, n_centroids, k, C] contiguous; dPts [n_clouds, n_points, ld] pre-zeroed by the caller. */
int pdf_group_scatter_add(const float* dG, const int32_t* idx, int64_t n_clouds, int n_points, int n_centroids, int k,
                          int C, float* dPts, int64_t ldp, void* stream);
/* Backward of pdf_gather_nchw (one cloud per frame): dFeat[b, c, ind[b,i]] += dOut[b,i,c];
 * dFeat [n_clouds, C, HW] pre-zeroed by the caller. */
int pdf_gather_nchw_bwd(const float* dOut, const int64_t* ind, int64_t n_clouds, int C, int64_t HW, int n,
                        float* dFeat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GCN decoder (lib/models/networks/intaghand_decoder.py:180-242): everything
 * between two dense layers of decoder.forward.  Activations are fp32 rows [n_samples * V, C].
 * ------------------------------------------------------------------------------------------- */

/* t = a[src] (+ b[src]) (+ rowvec[v]) for output row (sample, v), src = sample*(V_out/up) + v/up
 * (up = 2 is graph_upsample, DualGraph.py:11-18; rowvec is the position embedding, :76-80);
 * writes t to sum_out and/or LayerNorm(t)*gamma+beta (+ReLU) to ln_out (nn.LayerNorm, eps inside the
 * sqrt; gcn.py:92-98, self_attn.py:20,55).  C <= 1024.  sum_img / ln_img (optional, C % 64 == 0): the
 * same rows written as a split-bf16 tile image ([hi|hi|lo], 3*C/64 k-blocks per row-tile, as
 * pdf_rows_to_image(split=1) would produce), i.e. directly as the next GEMM's operand. */
int pdf_row_combine(const float* a, int64_t lda, const float* b, int64_t ldb, const float* rowvec, int64_t ldr,
                    int V_out, int up, int C, int64_t rows_out, const float* gamma, const float* beta, float eps,
                    int relu, float* sum_out, int64_t lds, float* ln_out, int64_t ldl, void* sum_img, void* ln_img,
                    void* stream);
/* Second half of a K = 2 Chebyshev graph convolution fused with the LayerNorm that follows
 * (graph_conv_cheby gcn.py:34-69, GCN_ResBlock.forward :100-110): with U = x [W0;W1]^T already
 * computed by a GEMM (W0 = fc.weight[:, 0::2], W1 = fc.weight[:, 1::2]),
 *   t = U0 + bias + L.U1 (+ R + bias_r),  out = LayerNorm(t) (+ReLU);
 * L [V,V] in CSR (rowptr int32 [V+1], colidx, vals); R is the shortcut branch (:108).  out (fp32 rows)
 * and/or out_img (split-bf16 tile image, C % 64 == 0) receive the result. */
int pdf_graph_cheby_ln(const float* U0, const float* U1, int64_t ldu, const float* bias, const float* R, int64_t ldr,
                       const float* bias_r, const int32_t* rowptr, const int32_t* colidx, const float* vals, int V,
                       int C, int64_t rows, const float* gamma, const float* beta, float eps, int relu, float* out,
                       int64_t ldo, void* out_img, void* stream);
/* The two kernels above for TWO groups of rows (left / right hand) in one launch: every buffer holds
 * 2 * rows_per_group rows (rows_per_group % 128 == 0 so that both groups start on a tile of the operand images; the
 * first `valid` rows of each group are real, the rest is padding that is neither read nor written); per-channel
 * parameters (bias, bias_r, gamma, beta) are stacked [2, C]; rowvec is shared (rowvec_gstride = 0) or stacked
 * rowvec_gstride floats apart; group g reads its source rows from a / b rows [g * src_per_group, ...).
 * C in {64, 128, 256, 512}, 16-byte aligned rows. */
int pdf_row_combine_grouped(const float* a, int64_t lda, const float* b, int64_t ldb, const float* rowvec, int64_t ldr,
                            int64_t rowvec_gstride, int V_out, int up, int C, int64_t rows_per_group, int64_t valid,
                            int64_t src_per_group, const float* gamma, const float* beta, float eps, int relu,
                            float* sum_out, int64_t lds, float* ln_out, int64_t ldl, void* sum_img, void* ln_img,
                            void* stream);
int pdf_graph_cheby_ln_grouped(const float* U0, const float* U1, int64_t ldu, const float* bias, const float* R,
                               int64_t ldr, const float* bias_r, const int32_t* rowptr, const int32_t* colidx,
                               const float* vals, int V, int C, int64_t rows_per_group, int64_t valid,
                               const float* gamma, const float* beta, float eps, int relu, float* out, int64_t ldo,
                               void* out_img, void* stream);
/* softmax(q k^T / sqrt(d)) v per (sample, head) (self_attn.py:60-72, inter_attn.py:84-108); q/k/v/out
 * rows [n_samples*V, heads*d] with free pitches (q and k/v may come from different hands).
 * V <= 256, d in {16, 32, 64}. */
int pdf_mha(const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* Vv, int64_t ldv,
            int64_t n_samples, int V, int heads, int d, float* out, int64_t ldo, void* stream);
/* The same attention on tensor cores (mma.sync m16n8k16, bf16 hi/lo SPLIT operands = fp32-accurate products,
 * fp32 accumulate, online softmax), up to two problems of identical shape per launch: problem i reads
 * q[i] / k[i] / v[i] (host arrays of device pointers; q and k may come from different hands: the R2L / L2R
 * directions of inter_attn.py:84-108 are one launch) and writes out[i] (fp32 rows, may be null) and / or
 * out_img[i] (may be null): the split-bf16 tile image [hi | hi | lo] of the [n_samples*V, heads*d] result, i.e. the
 * operand of the `fc` GEMM that follows; problem i's rows start at image row img_row0[i] (null = 0), so two
 * problems can fill the two halves of ONE image.  Row pitches in floats, even. */
int pdf_mha_tc(const float* const* q, const float* const* k, const float* const* v, float* const* out,
               void* const* out_img, const int64_t* img_row0, int n_problems, int64_t ldq, int64_t ldk, int64_t ldv,
               int64_t ldo, int64_t n_samples, int V, int heads, int d, void* stream);

/* projection_batch (lib/utils/utils.py:231-249) of the coarse [B,Vc,3] and dense [B,Vd,3] meshes with
 * params [B, >=3] = (scale, tx, ty), and the MANO-order lists of intaghand_decoder.py:231-240:
 * mano[b,i] = coarse[b, rev[i] / rep] (graph_upsample by rep, then GCN_to_vert). */
int pdf_decoder_project(const float* v_coarse, int Vc, const float* v_dense, int Vd, const float* params, int64_t ldp,
                        float img_size, const int64_t* rev, int rep, int64_t B, float* coarse2d, float* dense2d,
                        float* mano3d, float* mano2d, void* stream);
/* Output heads of decoder.forward (intaghand_decoder.py:213-224) for n hand-samples (both hands stacked: the
 * heads are shared modules): f [n*V, C] fp32 rows (pitch ldf) -> params [n,3] = params_head(avg_head(f^T)),
 * root [n,3] = root_head(avg_head(f^T)), verts [n,V,3] = coord_head(f).  avg_w [V], avg_b [1]; the three
 * 3-output heads as [3,C] weights + [3] biases. */
int pdf_decoder_heads(const float* f, int64_t ldf, int64_t n, int V, int C, const float* avg_w, const float* avg_b,
                      const float* params_w, const float* params_b, const float* root_w, const float* root_b,
                      const float* coord_w, const float* coord_b, float* params, float* root, float* verts,
                      void* stream);


#ifdef __cplusplus
}
#endif
#endif /* PDFNET_B200_H */
