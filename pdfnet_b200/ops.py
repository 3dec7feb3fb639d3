"""Tensor-level wrappers over the C-ABI kernels (one function per entry point).

Everything here takes and returns CUDA tensors, allocates outputs with torch
(the library never allocates), and enqueues on torch's current stream.
"""
import ctypes
import os

import torch

from . import _lib as L


def _strides_point_major(t):
    """(cloud, point, channel) element strides of a [B,N,C] tensor."""
    return t.stride(0), t.stride(1), t.stride(2)


def knn_ball(xyz, n_centroids, k, r2, channel_major=False):
    """Neighbour indices int32 [B,n_centroids,k] (see pdf_knn_ball).
    xyz: fp32 [B,N,C>=3] or, with channel_major=True, [B,C>=3,N]; any strides."""
    L.require_cuda(xyz)
    if xyz.dtype != torch.float32:
        xyz = xyz.float()
    if channel_major:
        B, _, N = xyz.shape
        sc, sch, sp = xyz.stride(0), xyz.stride(1), xyz.stride(2)
    else:
        B, N, _ = xyz.shape
        sc, sp, sch = _strides_point_major(xyz)
    out = torch.empty((B, n_centroids, k), dtype=torch.int32, device=xyz.device)
    L.call("pdf_knn_ball", L.ptr(xyz), B, N, n_centroids, k, float(r2), sc, sp, sch, L.ptr(out), L.stream())
    return out


def fps(xyz, n_sample, start_idx):
    """FPS selection order int32 [B,n_sample]; xyz fp32 [B,N,C>=3], start_idx int [B]."""
    L.require_cuda(xyz, start_idx)
    if xyz.dtype != torch.float32:
        xyz = xyz.float()
    B, N, _ = xyz.shape
    start = start_idx.to(torch.int32).contiguous()
    out = torch.empty((B, n_sample), dtype=torch.int32, device=xyz.device)
    sc, sp, sch = _strides_point_major(xyz)
    L.call("pdf_fps", L.ptr(xyz), B, N, n_sample, L.ptr(start), sc, sp, sch, L.ptr(out), L.stream())
    return out


CHECK_INDICES = bool(int(os.environ.get("PDF_CHECK_INDICES", "0")))


def _check_index(ind, limit, what):
    """The kernels clamp gather / scatter indices into [0, limit) so a bad index can never touch memory outside
    the map.  torch.gather (what the reference calls, models/utils.py:15) raises a device-side assert instead;
    PDF_CHECK_INDICES=1 reproduces that: an asynchronous device assert, no host synchronisation."""
    if CHECK_INDICES and ind.numel():
        torch._assert_async(((ind >= 0) & (ind < limit)).all(), "pdfnet_b200: %s index out of range" % what)


def _is_nhwc(t):
    """[F,C,H,W] tensor stored channels-last (and not also plain-contiguous, as C == 1 or H*W == 1 would be)."""
    return (t.dim() == 4 and t.dtype == torch.float32 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last))


def gather_nchw(feat, ind, clouds_per_frame=1):
    """out[b,i,:] = feat[b // clouds_per_frame, :, ind[b,i]] ; feat [F,C,H,W] fp32, ind [B,n] int64.
    A channels-last ``feat`` is gathered in place (pdf_gather_nhwc), no layout conversion."""
    L.require_cuda(feat, ind)
    nhwc = _is_nhwc(feat)
    if not nhwc:
        feat = L.f32c(feat)
    ind = ind.long()
    if ind.stride(-1) != 1:
        ind = ind.contiguous()
    Fr, C = feat.shape[0], feat.shape[1]
    HW = feat[0, 0].numel()
    B, n = ind.shape
    if B != Fr * clouds_per_frame:
        raise RuntimeError("gather_nchw: %d index rows for %d maps x %d clouds per frame" % (B, Fr, clouds_per_frame))
    _check_index(ind, HW, "gather_nchw")
    out = torch.empty((B, n, C), dtype=torch.float32, device=feat.device)
    L.call("pdf_gather_nhwc" if nhwc else "pdf_gather_nchw", L.ptr(feat), B, clouds_per_frame, C, HW, L.ptr(ind), n,
           ind.stride(0), L.ptr(out), L.stream())
    return out


def pyramid_gather(xyz, choose, emb, sft0_params, n1, n2, R, clouds_per_frame=1):
    """Fused 3-level gather + SFT0 (see pdf_pyramid_gather).
    Returns pts0 [B,N,3], cond1 [B,n1,C1], cond2 [B,n2,C2] (fp32)."""
    L.require_cuda(xyz, choose, emb[0], emb[1], emb[2], sft0_params)
    xyz = L.f32c(xyz)
    choose = choose.long().contiguous()
    nhwc = all(_is_nhwc(e) for e in emb)                 # channels-last hand-off from the RGB neck (f4)
    l0, l1, l2 = emb if nhwc else (L.f32c(e) for e in emb)
    B, N, _ = xyz.shape
    C1, C2 = l1.shape[1], l2.shape[1]
    # the index arithmetic (intaghand_encoder.py:125-126) assumes maps of R, R/2 and R/4 pixels a side
    for lvl, (e, r) in enumerate(((l0, R), (l1, R // 2), (l2, R // 4))):
        if tuple(e.shape[-2:]) != (r, r):
            raise RuntimeError("pyramid_gather: level %d map is %s, expected %dx%d for default_resolution=%d"
                               % (lvl, tuple(e.shape[-2:]), r, r, R))
    if l0.shape[1] != 3 or l0.shape[0] * clouds_per_frame < B:
        raise RuntimeError("pyramid_gather: level 0 must be [frames,3,R,R] with frames * clouds_per_frame >= clouds")
    _check_index(choose, R * R, "pyramid_gather (choose)")
    dev = xyz.device
    pts0 = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    cond1 = torch.empty((B, n1, C1), dtype=torch.float32, device=dev)
    cond2 = torch.empty((B, n2, C2), dtype=torch.float32, device=dev)
    L.call("pdf_pyramid_gather_nhwc" if nhwc else "pdf_pyramid_gather", L.ptr(xyz), L.ptr(choose), B, clouds_per_frame,
           N, n1, n2, R, L.ptr(l0), L.ptr(l1), C1, L.ptr(l2), C2, L.ptr(sft0_params), L.ptr(pts0), L.ptr(cond1),
           L.ptr(cond2), L.stream())
    return pts0, cond1, cond2


def _is_bf16_nhwc(t):
    return (t.dim() == 4 and t.dtype == torch.bfloat16
            and (t.is_contiguous(memory_format=torch.channels_last) or t.shape[1] == 1))


def pyramid_gather_bf16(xyz, choose, emb, sft0_params, n1, n2, R, clouds_per_frame=1):
    """3-level gather from a bf16 channels-last pyramid + SFT0 (pdf_pyramid_gather_bf16).
    Returns pts0 [B,N,3] fp32 and the two condition operands as bf16 TILE IMAGES (uint8 tensors):
    image of [B*n1, C1] and of [B*n2, C2] - what the SFT GEMMs read, no intermediate rows."""
    L.require_cuda(xyz, choose, sft0_params)       # the maps may be page-locked HOST tensors (read in place, below)
    xyz = L.f32c(xyz)
    choose = choose.long().contiguous()
    l0, l1, l2 = emb
    if not all(_is_bf16_nhwc(e) for e in emb):
        raise RuntimeError("pyramid_gather_bf16: the three maps must be bf16 in torch.channels_last memory format")
    B, N, _ = xyz.shape
    C1, C2 = l1.shape[1], l2.shape[1]
    for lvl, (e, r) in enumerate(((l0, R), (l1, R // 2), (l2, R // 4))):
        if tuple(e.shape[-2:]) != (r, r):
            raise RuntimeError("pyramid_gather_bf16: level %d map is %s, expected %dx%d for default_resolution=%d"
                               % (lvl, tuple(e.shape[-2:]), r, r, R))
    if l0.shape[1] != 3 or l0.shape[0] * clouds_per_frame < B:
        raise RuntimeError("pyramid_gather_bf16: level 0 must be [frames,3,R,R] with frames * clouds_per_frame >= clouds")
    _check_index(choose, R * R, "pyramid_gather_bf16 (choose)")
    dev = xyz.device
    pts0 = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    img1 = torch.empty((image_bytes(B * n1, C1),), dtype=torch.uint8, device=dev)
    img2 = torch.empty((image_bytes(B * n2, C2),), dtype=torch.uint8, device=dev)
    # a map living in page-locked host memory is gathered in place over the PCIe link (zero-copy): only the pixels
    # `choose` selects cross the link instead of the whole map (L.host_ptr raises for pageable memory)
    L.call("pdf_pyramid_gather_bf16", L.ptr(xyz), L.ptr(choose), B, clouds_per_frame, N, n1, n2, R, L.host_ptr(l0),
           L.host_ptr(l1), C1, L.host_ptr(l2), C2, L.ptr(sft0_params), L.ptr(pts0), L.ptr(img1), L.ptr(img2), L.stream())
    return pts0, img1, img2


def group_gather(pts, idx, channel_major=False, out=None, want_center=True):
    """Grouped rows [B,N1,k,C] with centroid-relative xyz, and centres [B,N1,3]."""
    L.require_cuda(pts, idx)
    if pts.dtype != torch.float32:
        pts = pts.float()
    if channel_major:
        B, C, _ = pts.shape
        sc, sch, sp = pts.stride(0), pts.stride(1), pts.stride(2)
    else:
        B, _, C = pts.shape
        sc, sp, sch = _strides_point_major(pts)
    idx = idx.to(torch.int32).contiguous()
    _, N1, k = idx.shape
    if out is None:
        out = torch.empty((B, N1, k, C), dtype=torch.float32, device=pts.device)
    center = torch.empty((B, N1, 3), dtype=torch.float32, device=pts.device) if want_center else None
    L.call("pdf_group_gather", L.ptr(pts), B, N1, k, C, sc, sp, sch, L.ptr(idx), L.ptr(out), out.stride(2),
           L.ptr(center), L.stream())
    return out, center


def linear(x, w, bias=None, act=L.ACT_NONE, epilogue=L.EPI_STORE, group=0, f=None, out=None):
    """Y = epilogue(x @ w.T + bias); x [M,K] (row pitch free), w [N,K]; see pdf_linear_f32."""
    L.require_cuda(x, w, bias, f, out)
    assert x.dim() == 2 and w.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        if epilogue == L.EPI_GROUP_MAX:
            out = torch.zeros((M // group, N), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    assert out.stride(1) == 1
    ldf = f.stride(0) if f is not None else 0
    L.call("pdf_linear_f32", L.ptr(x), x.stride(0), L.ptr(w), w.stride(0), L.ptr(bias), M, N, K, act, epilogue, group,
           L.ptr(f), ldf, L.ptr(out), out.stride(0), L.stream())
    return out


def sa_pack_weights(w1, b1, w2, b2, w3, b3):
    """Host-side packing of folded fp32 weights into the tcgen05 kernel's bf16 image."""
    c_in, c1, c2, c3 = w1.shape[1], w1.shape[0], w2.shape[0], w3.shape[0]
    lib = L.load()
    size = int(lib.pdf_sa_pack_size(c_in, c1, c2, c3))
    if size <= 0:
        raise RuntimeError("pdf_sa_pack_size: unsupported channel plan %s" % ((c_in, c1, c2, c3),))
    buf = torch.empty((size,), dtype=torch.uint8)
    host = [t.detach().cpu().float().contiguous() for t in (w1, b1, w2, b2, w3, b3)]
    args = [ctypes.c_void_p(t.data_ptr()) for t in host]
    L.call("pdf_sa_pack_weights_host", *args, c_in, c1, c2, c3, ctypes.c_void_p(buf.data_ptr()))
    return buf


def sa_mlp_max_bf16(pts, idx, wpack, c_in, c1, c2, c3, out, out_col0=0, feat_bf16=None):
    """Fused gather + 3-layer point-MLP + max over k on tcgen05 (see pdf_sa_mlp_max_bf16).
    pts fp32 [B,n_src,ld] contiguous, idx int32 [B,N1,k]; writes out[:, :, out_col0:out_col0+c3]."""
    L.require_cuda(pts, idx, wpack, out, feat_bf16)
    assert pts.dtype == torch.float32 and pts.is_contiguous() and idx.dtype == torch.int32 and idx.is_contiguous()
    assert feat_bf16 is None or (feat_bf16.dtype == torch.bfloat16 and feat_bf16.is_contiguous()
                                 and feat_bf16.shape[-1] == 128)
    assert out.dtype == torch.float32 and out.stride(2) == 1
    B, n_src, ld = pts.shape
    _, N1, k = idx.shape
    L.call("pdf_sa_mlp_max_bf16", L.ptr(pts), B, n_src, ld, c_in, L.ptr(feat_bf16), L.ptr(idx), N1, k, L.ptr(wpack), c1, c2, c3,
           L.ptr(out), out.stride(1), out_col0, L.stream())
    return out


def image_bytes(rows, cols):
    return int(L.load().pdf_image_bytes(rows, cols))


def pack_image(w, split=False):
    """Host-side packing of an fp32 matrix [rows, cols] into a bf16 tile image (uint8 CPU tensor).
    split=True packs [hi | lo | hi] (3x k-blocks) for fp32-accurate GEMMs against a split activation image."""
    w = w.detach().cpu().float().contiguous()
    rows, cols = w.shape
    buf = torch.empty((image_bytes(rows, cols) * (3 if split else 1),), dtype=torch.uint8)
    L.call("pdf_pack_image_host", ctypes.c_void_p(w.data_ptr()), rows, cols, w.stride(0), 1 if split else 0,
           ctypes.c_void_p(buf.data_ptr()))
    return buf


def rows_to_image(x, col0, K, img=None, kb_total=None, kb0=0, split=False):
    """fp32 rows x[M, ld] columns [col0, col0+K) -> bf16 tile image (uint8 device tensor).
    split=True/1 writes [hi | hi | lo], split=2 writes [hi | lo | hi] (3x k-blocks)."""
    L.require_cuda(x, img)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32
    M = x.shape[0]
    nkb = (K + 63) // 64 * (3 if split else 1)
    if kb_total is None:
        kb_total = nkb
    if img is None:
        img = torch.empty((((M + 127) // 128) * kb_total * 16384,), dtype=torch.uint8, device=x.device)
    L.call("pdf_rows_to_image", L.ptr(x), x.stride(0), M, col0, K, L.ptr(img), kb_total, kb0, int(split),
           L.stream())
    return img


def gemm_bf16(m_img, m_tiles, m_kb, n_img, n_tiles, n_kb, KB, bias0, kb_split=0, bias1=None, act=L.ACT_NONE,
              out_f32=None, rows_valid=0, F=None, out_img=None, out_kb=0, tile_desc=None, out_max=None,
              out_bf16=None, bf16_col_off=0, xyz_w=None, xyz_x=None):
    """Streaming tcgen05 GEMM over tile images (see pdf_gemm_bf16)."""
    L.require_cuda(m_img, n_img, bias0, bias1, out_f32, F, out_img, out_max, out_bf16, xyz_w, xyz_x)
    colmax = out_max is not None
    desc = None
    if tile_desc is not None:
        flat = [int(v) for t in tile_desc for v in t]
        desc = (ctypes.c_int32 * len(flat))(*flat)
    L.call("pdf_gemm_bf16", L.ptr(m_img), m_tiles, m_kb, L.ptr(n_img), n_tiles, n_kb, KB, kb_split, 1 if colmax else 0,
           L.ptr(bias0), L.ptr(bias1), act, L.ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0, rows_valid,
           L.ptr(F), F.stride(0) if F is not None else 0, L.ptr(out_img), out_kb, L.ptr(out_bf16),
           out_bf16.stride(0) if out_bf16 is not None else 0, bf16_col_off,
           ctypes.cast(desc, ctypes.c_void_p) if desc is not None else None, L.ptr(out_max),
           out_max.stride(0) if colmax else 0, L.ptr(xyz_w), L.ptr(xyz_x),
           xyz_x.stride(0) if xyz_x is not None else 0, L.stream())


def sft_xyz(cond_rows, weights, x_rows):
    """fp32 SFT on columns 0..2 of x_rows (in place); weights = SFTLayer.weights() tuple."""
    L.require_cuda(cond_rows, x_rows)
    ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = weights
    assert cond_rows.is_contiguous() and x_rows.stride(1) == 1
    L.call("pdf_sft_xyz_f32", L.ptr(cond_rows), cond_rows.shape[0], cond_rows.shape[1], L.ptr(ws0), L.ptr(bs0),
           L.ptr(ws1), L.ptr(bs1), L.ptr(wh0), L.ptr(bh0), L.ptr(wh1), L.ptr(bh1), L.ptr(x_rows), x_rows.stride(0),
           L.stream())


def center_im2col(x0, ind):
    """im2col rows [B*2*9, 9*C] of the conv0 outputs that conv1 needs at the centre pixels (pdf_center_im2col)."""
    L.require_cuda(x0, ind)
    x0 = L.f32c(x0)
    ind = ind.long().contiguous()
    B, C, H, W = x0.shape
    rows = torch.empty((B * 2 * 9, 9 * C), dtype=torch.float32, device=x0.device)
    L.call("pdf_center_im2col", L.ptr(x0), L.ptr(ind), B, C, H, W, L.ptr(rows), L.stream())
    return rows


def backproject(depth, Kinv):
    """xyz [B,3,H,W] = (Kinv @ [u,v,1]) * depth ; depth [B,H,W] fp32, Kinv [B,3,3] fp32."""
    L.require_cuda(depth, Kinv)
    depth, Kinv = L.f32c(depth), L.f32c(Kinv)
    B, H, W = depth.shape
    xyz = torch.empty((B, 3, H, W), dtype=torch.float32, device=depth.device)
    L.call("pdf_backproject", L.ptr(depth), L.ptr(Kinv), B, H, W, L.ptr(xyz), L.stream())
    return xyz


def depth2pcl(depth, mask, Kinv, valid, subset_keys=None, perm=None, n_points=1024, min_pixels=10, seed=None):
    """Batched device-side cloud builder (see pdf_depth2pcl / pdf_depth2pcl_seeded).
    mask: fp32 (the reference's dtype) or uint8 / bool (non-zero = hand).  Randomness: explicit
    ``subset_keys`` int32 [B,2,H*W] / ``perm`` int32 [B,2,n]; whatever is None is generated inside the
    kernel from ``seed`` (counter-based, reproducible on the host with d2p_host_randomness).
    Returns choose int64 [B,2,n], cloud fp32 [B,2,n,3], n_cand int32 [B,2]."""
    L.require_cuda(depth, mask, Kinv, valid, subset_keys, perm)
    depth, Kinv, valid = L.f32c(depth), L.f32c(Kinv), L.f32c(valid)
    u8 = mask.dtype in (torch.uint8, torch.bool)
    mask = mask.contiguous() if u8 else L.f32c(mask)
    B, H, W = depth.shape
    dev = depth.device
    if (subset_keys is None or perm is None) and seed is None:
        if perm is None:                                # historic behaviour of the unseeded entry: identity order
            perm = torch.arange(n_points, dtype=torch.int32, device=dev).expand(B, 2, n_points)
        if subset_keys is None and H * W > n_points:
            raise RuntimeError("depth2pcl: pass subset_keys or a seed (a hand can exceed n_points pixels)")
    if subset_keys is not None:
        subset_keys = subset_keys.to(torch.int32).contiguous()
    if perm is not None:
        perm = perm.to(torch.int32).contiguous()
    choose = torch.empty((B, 2, n_points), dtype=torch.int64, device=dev)
    cloud = torch.empty((B, 2, n_points, 3), dtype=torch.float32, device=dev)
    n_cand = torch.empty((B, 2), dtype=torch.int32, device=dev)
    if u8 or seed is not None:
        L.call("pdf_depth2pcl_seeded", L.ptr(depth), L.ptr(mask), 1 if u8 else 0, L.ptr(Kinv), L.ptr(valid),
               L.ptr(subset_keys), L.ptr(perm), int(seed or 0) & 0xFFFFFFFF, B, H, W, n_points, min_pixels,
               L.ptr(choose), L.ptr(cloud), L.ptr(n_cand), L.stream())
    else:
        L.call("pdf_depth2pcl", L.ptr(depth), L.ptr(mask), L.ptr(Kinv), L.ptr(valid), L.ptr(subset_keys), L.ptr(perm),
               B, H, W, n_points, min_pixels, L.ptr(choose), L.ptr(cloud), L.ptr(n_cand), L.stream())
    return choose, cloud, n_cand


def d2p_host_randomness(seed, n_clouds, npx, want_keys=True, want_perm=True):
    """The keys int32 [n_clouds, npx] / perm int32 [n_clouds, 1024] that pdf_depth2pcl_seeded generates for
    ``seed`` (host tensors; cloud = 2*frame + hand)."""
    keys = torch.empty((n_clouds, npx), dtype=torch.int32) if want_keys else None
    perm = torch.empty((n_clouds, 1024), dtype=torch.int32) if want_perm else None
    L.call("pdf_depth2pcl_host_randomness", int(seed) & 0xFFFFFFFF, n_clouds, npx,
           ctypes.c_void_p(keys.data_ptr()) if want_keys else None, ctypes.c_void_p(perm.data_ptr()) if want_perm else None)
    return keys, perm


def rodrigues(axis):
    """rodrigues_batch (manolayer.py:32-48): axis-angle [n,3] -> rotation matrices [n,3,3]."""
    L.require_cuda(axis)
    axis = L.f32c(axis).reshape(-1, 3)
    out = torch.empty((axis.shape[0], 3, 3), dtype=torch.float32, device=axis.device)
    L.call("pdf_rodrigues", L.ptr(axis), axis.shape[0], L.ptr(out), L.stream())
    return out


def joint_regress(reg, verts):
    """full_regressor @ verts: reg [J,778] fp32, verts [n,778,3] -> joints [n,J,3] (pdf_joint_regress)."""
    L.require_cuda(reg, verts)
    reg, verts = L.f32c(reg), L.f32c(verts)
    assert reg.dim() == 2 and reg.shape[1] == 778 and verts.shape[1:] == (778, 3)
    out = torch.empty((verts.shape[0], reg.shape[0], 3), dtype=torch.float32, device=verts.device)
    L.call("pdf_joint_regress", L.ptr(reg), reg.shape[0], L.ptr(verts), verts.shape[0], L.ptr(out), L.stream())
    return out


MANO_VT_PITCH = 2336         # PDF_MANO_VT_PITCH: row pitch of the precomputed blend-shape rows
MANO_TC_MIN_HANDS = 128      # from here on the blend-shape contraction runs on the tcgen05 GEMM


def _blend_shapes(tables, X, out):
    """v_tpose rows = X [n,145] @ blend_w^T [145,2334] + v_template into ``out`` [n, >=2334] (row pitch free).
    From MANO_TC_MIN_HANDS hands on: the streaming tcgen05 GEMM with split-bf16 operands ([hi|hi|lo] x [hi|lo|hi],
    ~2^-16 relative per product: the 1e-5 m vertex bound holds with two orders of margin); the weight image is
    packed once per table set.  Below: the FFMA kernel."""
    n = X.shape[0]
    if n < MANO_TC_MIN_HANDS:
        linear(X, tables["blend_w"], tables["v_template"], out=out[:, :2334])
        return
    packed = tables.get("_blend_tc")
    if packed is None:
        packed = tables["_blend_tc"] = pack_linear_tc(tables["blend_w"], tables["v_template"], split=True)
    w_img, b, N, K, _ = packed
    kb = 3 * ((K + 63) // 64)
    x_img = rows_to_image(X, 0, K, split=1)
    gemm_bf16(x_img, (n + 127) // 128, kb, w_img, (N + 127) // 128, kb, kb, b, out_f32=out, rows_valid=n,
              tile_desc=_tile_desc(N))


def mano_lbs(tables, root, pose, shape, trans, scale, tips, center_idx, new_skel, root_is_matrix=False):
    """tables: dict of device tensors in the kernel layout (see manolayer.ManoTables).
    root_is_matrix: ``root`` is [n,3,3] rotation matrices (use_pca layers, manolayer.py:266-267)."""
    L.require_cuda(root, pose, shape, trans, scale)
    root, pose, shape = L.f32c(root), L.f32c(pose), L.f32c(shape)
    trans = L.f32c(trans) if trans is not None else None
    scale = L.f32c(scale) if scale is not None else None
    n = root.shape[0]
    v = torch.empty((n, 778, 3), dtype=torch.float32, device=root.device)
    j = torch.empty((n, 21, 3), dtype=torch.float32, device=root.device)
    tip_arr = (ctypes.c_int32 * 5)(*[int(t) for t in tips])
    v_tpose = None
    if n >= 16 and "blend_w" in tables:
        # blend shapes of all hands as one dense GEMM [n,145] x [145,2334] (+ v_template as bias)
        X = torch.empty((n, 145), dtype=torch.float32, device=root.device)
        L.call("pdf_mano_pose_feature", L.ptr(pose), L.ptr(shape), n, L.ptr(X), L.stream())
        v_tpose = torch.empty((n, MANO_VT_PITCH), dtype=torch.float32, device=root.device)
        _blend_shapes(tables, X, v_tpose)
    L.call("pdf_mano_lbs_rootmat" if root_is_matrix else "pdf_mano_lbs", L.ptr(tables["v_template"]),
           L.ptr(tables["shapedirs_t"]), L.ptr(tables["posedirs_t"]),
           L.ptr(tables["j_template"]), L.ptr(tables["j_shapedirs"]), L.ptr(tables["weights_t"]), L.ptr(root),
           L.ptr(pose), L.ptr(shape), L.ptr(trans), L.ptr(scale), n, ctypes.cast(tip_arr, ctypes.c_void_p),
           -1 if center_idx is None else int(center_idx), 1 if new_skel else 0, L.ptr(v_tpose), L.ptr(v), L.ptr(j),
           L.stream())
    return v, j


def mano_lbs_pair(tables_l, tables_r, root, pose, shape, trans, scale, tips_l, tips_r, center_idx, new_skel):
    """Both hands of every frame in one launch; inputs [B,2,...] ((frame, side), side 0 = left)."""
    L.require_cuda(root, pose, shape, trans, scale)
    B = root.shape[0]
    n = 2 * B
    root, pose, shape = L.f32c(root).view(n, 3), L.f32c(pose).view(n, 45), L.f32c(shape).view(n, 10)
    trans = L.f32c(trans).view(n, 3) if trans is not None else None
    scale = L.f32c(scale).view(n) if scale is not None else None
    dev = root.device
    v = torch.empty((B, 2, 778, 3), dtype=torch.float32, device=dev)
    j = torch.empty((B, 2, 21, 3), dtype=torch.float32, device=dev)
    names = ("v_template", "shapedirs_t", "posedirs_t", "j_template", "j_shapedirs", "weights_t")
    tl = (ctypes.c_void_p * 6)(*[tables_l[k].data_ptr() for k in names])
    tr = (ctypes.c_void_p * 6)(*[tables_r[k].data_ptr() for k in names])
    tipl = (ctypes.c_int32 * 5)(*[int(t) for t in tips_l])
    tipr = (ctypes.c_int32 * 5)(*[int(t) for t in tips_r])
    v_tpose = None
    if n >= 16:
        # blend shapes: one [n,145] coefficient matrix, then one fp32 GEMM per side on strided rows
        X = torch.empty((n, 145), dtype=torch.float32, device=dev)
        L.call("pdf_mano_pose_feature", L.ptr(pose), L.ptr(shape), n, L.ptr(X), L.stream())
        v_tpose = torch.empty((n, MANO_VT_PITCH), dtype=torch.float32, device=dev)
        _blend_shapes(tables_l, X[0::2], v_tpose[0::2])
        _blend_shapes(tables_r, X[1::2], v_tpose[1::2])
    L.call("pdf_mano_lbs_pair", ctypes.cast(tl, ctypes.c_void_p), ctypes.cast(tr, ctypes.c_void_p), L.ptr(root),
           L.ptr(pose), L.ptr(shape), L.ptr(trans), L.ptr(scale), n, ctypes.cast(tipl, ctypes.c_void_p),
           ctypes.cast(tipr, ctypes.c_void_p), -1 if center_idx is None else int(center_idx), 1 if new_skel else 0,
           L.ptr(v_tpose), L.ptr(v), L.ptr(j), L.stream())
    return v, j


def split_coeff_pair(theta, index, K, input_res, down_ratio):
    """theta [B,2,122] (row (b,0) = point2mano_left, (b,1) = point2mano_right), index [B,2], K [B,3,3] ->
    root [B,2,3], pose [B,2,45], shape [B,2,10], trans [B,2,3] in one launch."""
    L.require_cuda(theta, index, K)
    theta, K = L.f32c(theta), L.f32c(K)
    index = index.long().contiguous()
    B = theta.shape[0]
    dev = theta.device
    root = torch.empty((B, 2, 3), dtype=torch.float32, device=dev)
    pose = torch.empty((B, 2, 45), dtype=torch.float32, device=dev)
    shape = torch.empty((B, 2, 10), dtype=torch.float32, device=dev)
    trans = torch.empty((B, 2, 3), dtype=torch.float32, device=dev)
    L.call("pdf_split_coeff", L.ptr(theta), 122, 0, 1, L.ptr(index), L.ptr(K), 2 * B, input_res, down_ratio,
           L.ptr(root), L.ptr(pose), L.ptr(shape), L.ptr(trans), L.stream())
    return root, pose, shape, trans


def split_coeff(theta, col0, index, K, input_res, down_ratio):
    """One hand's slice of Split_coeff; returns root [n,3], pose [n,45], shape [n,10], trans [n,3]."""
    L.require_cuda(theta, index, K)
    theta, K = L.f32c(theta), L.f32c(K)
    index = index.long().contiguous()
    n = theta.shape[0]
    dev = theta.device
    root = torch.empty((n, 3), dtype=torch.float32, device=dev)
    pose = torch.empty((n, 45), dtype=torch.float32, device=dev)
    shape = torch.empty((n, 10), dtype=torch.float32, device=dev)
    trans = torch.empty((n, 3), dtype=torch.float32, device=dev)
    L.call("pdf_split_coeff", L.ptr(theta), theta.stride(0), col0, 0, L.ptr(index), L.ptr(K), n, input_res, down_ratio,
           L.ptr(root), L.ptr(pose), L.ptr(shape), L.ptr(trans), L.stream())
    return root, pose, shape, trans


# ------------------------------------------------------------------------------------------------
# training-mode (cfg5) primitives: thin wrappers over the pdf_* entry points in train_f32.cu
def _rows(t):
    assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.float32, "fp32 rows with unit column stride"
    return t


def bn_batch_stats(x, eps, momentum, running_mean=None, running_var=None):
    """Train-mode BatchNorm statistics of rows x [M,C]: (mean, rstd); updates the running buffers."""
    L.require_cuda(x, running_mean, running_var)
    M, C = _rows(x).shape
    sums = torch.empty((2 * C,), dtype=torch.float64, device=x.device)
    L.call("pdf_bn_stats", L.ptr(x), x.stride(0), M, C, L.ptr(sums), L.stream())
    mean = torch.empty((C,), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    L.call("pdf_bn_finalize", L.ptr(sums), L.ptr(x), M, C, float(eps), float(momentum), L.ptr(running_mean),
           L.ptr(running_var), L.ptr(mean), L.ptr(rstd), L.stream())
    return mean, rstd


def plain_image_empty(rows, C, device):
    """Uninitialised plain bf16 tile image for a [rows, C] matrix, C % 64 == 0."""
    assert C % 64 == 0
    return torch.empty((((rows + 127) // 128) * (C // 64) * 16384,), dtype=torch.uint8, device=device)


def bn_act_fwd(x, mean, rstd, gamma, beta, relu=True, rows=True, image=False, plain=False):
    """[relu](BatchNorm(x)) as fp32 rows and/or (image=True, C % 64 == 0) as the tile image of the next layer's
    GEMM (split-bf16, or one plain bf16 image with plain=True).  Returns rows, or (rows or None, image)."""
    M, C = _rows(x).shape
    y = torch.empty((M, C), dtype=torch.float32, device=x.device) if rows else None
    img = (plain_image_empty if plain else split_image_empty)(M, C, x.device) if image else None
    relu = int(relu) | (L.BN_PLAIN_IMAGE if plain else 0)
    L.call("pdf_bn_act_fwd", L.ptr(x), x.stride(0), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(beta), int(relu), M,
           C, L.ptr(y), y.stride(0) if y is not None else 0, L.ptr(img), L.stream())
    return (y, img) if image else y


def bn_act_bwd(dy, y, x, mean, rstd, gamma, relu=True, beta=None, image=False, plain=False):
    """-> (dx, dgamma, dbeta).  y=None with beta given: the ReLU mask is recomputed from x (no read of y).
    image=True (needs y=None, beta, C % 64 == 0): dx is returned as the split-bf16 tile image the gradient
    GEMMs read instead of fp32 rows."""
    M, C = _rows(x).shape
    dy = _rows(dy if dy.stride(1) == 1 else dy.contiguous())
    sums = torch.empty((2 * C,), dtype=torch.float64, device=x.device)
    dx = None if image else torch.empty((M, C), dtype=torch.float32, device=x.device)
    img = (plain_image_empty if plain else split_image_empty)(M, C, x.device) if image else None
    relu = int(relu) | (L.BN_PLAIN_IMAGE if plain else 0)
    L.call("pdf_bn_act_bwd", L.ptr(dy), dy.stride(0), L.ptr(y), y.stride(0) if y is not None else 0, L.ptr(x),
           x.stride(0), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(beta), int(relu), M, C, L.ptr(sums), L.ptr(dx),
           dx.stride(0) if dx is not None else 0, L.ptr(img), L.stream())
    return (img if image else dx), sums[C:].float(), sums[:C].float()


def col_sum(a):
    M, C = _rows(a).shape
    sums = torch.empty((C,), dtype=torch.float64, device=a.device)
    L.call("pdf_col_sum", L.ptr(a), a.stride(0), M, C, L.ptr(sums), L.stream())
    return sums.float()


def act_bwd(dy, y, act):
    M, C = _rows(y).shape
    dy = _rows(dy if dy.stride(1) == 1 else dy.contiguous())
    dx = torch.empty((M, C), dtype=torch.float32, device=y.device)
    L.call("pdf_act_bwd", L.ptr(dy), dy.stride(0), L.ptr(y), y.stride(0), int(act), M, C, L.ptr(dx), dx.stride(0),
           L.stream())
    return dx


def sft_modulate(fea, scale, shift):
    M, C = _rows(fea).shape
    out = torch.empty((M, C), dtype=torch.float32, device=fea.device)
    L.call("pdf_sft_modulate", L.ptr(fea), fea.stride(0), L.ptr(_rows(scale)), scale.stride(0), L.ptr(_rows(shift)),
           shift.stride(0), M, C, L.ptr(out), out.stride(0), L.stream())
    return out


def sft_modulate_bwd(dout, fea, scale):
    """-> (dfea, dscale); dshift is dout itself."""
    M, C = _rows(fea).shape
    dout = _rows(dout if dout.stride(1) == 1 else dout.contiguous())
    dfea = torch.empty((M, C), dtype=torch.float32, device=fea.device)
    dscale = torch.empty_like(dfea)
    L.call("pdf_sft_modulate_bwd", L.ptr(dout), dout.stride(0), L.ptr(fea), fea.stride(0), L.ptr(scale),
           scale.stride(0), M, C, L.ptr(dfea), dfea.stride(0), L.ptr(dscale), dscale.stride(0), L.stream())
    return dfea, dscale


def linear_tn(a, b):
    """a [M,N], b [M,K] -> a^T b [N,K]  (weight gradient: a = dY, b = layer input)."""
    L.require_cuda(a, b)
    M, N = _rows(a).shape
    K = _rows(b).shape[1]
    assert b.shape[0] == M
    c = torch.empty((N, K), dtype=torch.float32, device=a.device)
    L.call("pdf_linear_tn_f32", L.ptr(a), a.stride(0), L.ptr(b), b.stride(0), M, N, K, L.ptr(c), c.stride(0),
           L.stream())
    return c


def linear_smallk(mode, a, b, bias=None):
    """K <= 4 streaming linear layer: mode 0 forward (a = x [M,K], b = w [N,K]), 1 data gradient
    (a = dy [M,N], b = w [N,K] -> [M,K]), 2 weight gradient (a = dy [M,N], b = x [M,K] -> [N,K])."""
    L.require_cuda(a, b, bias)
    a, b = _rows(a), _rows(b)
    if mode == 0:
        M, K = a.shape
        N = b.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    elif mode == 1:
        M, N = a.shape
        K = b.shape[1]
        out = torch.empty((M, K), dtype=torch.float32, device=a.device)
    else:
        M, N = a.shape
        K = b.shape[1]
        out = torch.empty((N, K), dtype=torch.float32, device=a.device)
    L.call("pdf_linear_smallk_f32", mode, L.ptr(a), a.stride(0), L.ptr(b), b.stride(0), L.ptr(bias), M, N, K, L.ptr(out),
           out.stride(0), L.stream())
    return out


def group_max(y, G, want_arg=False):
    """max over groups of G consecutive rows; want_arg: also the uint8 row index of the (first) maximum."""
    M, C = _rows(y).shape
    assert M % G == 0
    out = torch.empty((M // G, C), dtype=torch.float32, device=y.device)
    arg = torch.empty((M // G, C), dtype=torch.uint8, device=y.device) if want_arg else None
    L.call("pdf_group_max", L.ptr(y), y.stride(0), G, M // G, C, L.ptr(out), out.stride(0), L.ptr(arg), L.stream())
    return (out, arg) if want_arg else out


def bn_maxpool_bwd(dout, arg, G, x, mean, rstd, gamma, beta, relu=True, plain=False):
    """BatchNorm(+ReLU) backward fed by the max-pool gradient (never materialised): -> (dx split tile image,
    dgamma, dbeta); see pdf_bn_maxpool_bwd."""
    M, C = _rows(x).shape
    dout = _rows(dout if dout.stride(1) == 1 else dout.contiguous())
    sums = torch.empty((2 * C,), dtype=torch.float64, device=x.device)
    img = (plain_image_empty if plain else split_image_empty)(M, C, x.device)
    relu = int(relu) | (L.BN_PLAIN_IMAGE if plain else 0)
    L.call("pdf_bn_maxpool_bwd", L.ptr(dout), dout.stride(0), L.ptr(arg), G, L.ptr(x), x.stride(0), L.ptr(mean),
           L.ptr(rstd), L.ptr(gamma), L.ptr(beta), int(relu), M, C, L.ptr(sums), L.ptr(img), L.stream())
    return img, sums[C:].float(), sums[:C].float()


def group_max_bwd(y, dout, G):
    M, C = _rows(y).shape
    dout = _rows(dout if dout.stride(1) == 1 else dout.contiguous())
    dy = torch.empty((M, C), dtype=torch.float32, device=y.device)
    L.call("pdf_group_max_bwd", L.ptr(y), y.stride(0), L.ptr(dout), dout.stride(0), G, M // G, C, L.ptr(dy),
           dy.stride(0), L.stream())
    return dy


def group_scatter_add(dg, idx, n_points):
    """dg [B,N1,K,C] contiguous, idx int32 [B,N1,K] -> dpts [B,n_points,C]."""
    L.require_cuda(dg, idx)
    dg = L.f32c(dg)
    B, N1, K, C = dg.shape
    dpts = torch.zeros((B, n_points, C), dtype=torch.float32, device=dg.device)
    L.call("pdf_group_scatter_add", L.ptr(dg), L.ptr(idx), B, n_points, N1, K, C, L.ptr(dpts), C, L.stream())
    return dpts


def gather_nchw_bwd(dout, ind, shape):
    """dout [B,n,C], ind int64 [B,n] -> dfeat of ``shape`` [B,C,H,W]."""
    L.require_cuda(dout, ind)
    dout = L.f32c(dout)
    ind = ind.long().contiguous()
    B, n, C = dout.shape
    _check_index(ind, int(shape[2]) * int(shape[3]), "gather_nchw_bwd")
    dfeat = torch.zeros(tuple(shape), dtype=torch.float32, device=dout.device)
    L.call("pdf_gather_nchw_bwd", L.ptr(dout), L.ptr(ind), B, C, dfeat[0, 0].numel(), n, L.ptr(dfeat), L.stream())
    return dfeat


def _pad4(n):
    return (n + 3) // 4 * 4


def _tile_desc(n):
    return [(128 * i, min(128, n - 128 * i), 0) for i in range((n + 127) // 128)]


def pack_linear_tc(w, bias=None, split=True):
    """Pre-pack a weight [N,K] (device, fp32) for linear_tc: (bf16 tile image, bias padded to whole
    N-tiles, N, K, split).  Inference code packs once and reuses it; training re-packs every call."""
    N, K = _rows(w).shape
    nt = (N + 127) // 128
    w_img = rows_to_image(w, 0, K, split=2 if split else 0)
    b = torch.zeros((nt * 128,), dtype=torch.float32, device=w.device)
    if bias is not None:
        b[:N] = bias
    return (w_img, b, N, K, split)


def linear_tc(x, w, bias=None, act=L.ACT_NONE, split=True, packed=None, x_img=None, M=None, out_image=False,
              light=False):
    """act(x @ w.T + bias) on the tcgen05 GEMM.  split=True: split-bf16 operands (three bf16 products
    per fp32 product: fp32-accurate, ~2^-16 relative); split=False: plain bf16 operands, fp32
    accumulate.  x [M,K] fp32 rows (or ``x_img`` = its tile image and ``M``), w [N,K] fp32 (device) or
    ``packed`` = pack_linear_tc(w, bias, split).
    Returns a [M,N] view of a buffer whose row pitch is padded to a multiple of 4."""
    L.require_cuda(x, w, bias)
    if packed is None:
        packed = pack_linear_tc(w, bias, split)
    w_img, b, N, K, split = packed
    kb = (3 if split else 1) * ((K + 63) // 64)
    nt = (N + 127) // 128
    if x_img is None:                                   # otherwise the producer already wrote the operand image
        M = _rows(x).shape[0]
        assert x.shape[1] == K
        x_img = rows_to_image(x, 0, K, split=1 if split else 0)
    if light:
        act = act | L.GEMM_LIGHT
    if out_image:
        # the result only feeds another GEMM: written ONLY as that GEMM's split operand image (N % 64 == 0)
        assert split and N % 64 == 0
        img = split_image_empty(M, N, x_img.device)
        gemm_bf16(x_img, (M + 127) // 128, kb, w_img, nt, kb, kb, b, act=act | L.GEMM_OUT_SPLIT, out_img=img,
                  out_kb=3 * (N // 64), rows_valid=M, tile_desc=[(128 * i, min(128, N - 128 * i), 2 * i) for i in range(nt)])
        return img
    out = torch.empty((M, _pad4(N)), dtype=torch.float32, device=x_img.device)
    gemm_bf16(x_img, (M + 127) // 128, kb, w_img, nt, kb, kb, b, act=act, out_f32=out, rows_valid=M,
              tile_desc=_tile_desc(N))
    return out[:, :N]


def pack_linear_tc_grouped(ws, biases):
    """Two weights [N,K] of identical shape (the left / right hand's copy of one layer) packed for
    linear_tc_grouped: (stacked split weight images, stacked padded biases, N, K, image bytes, bias floats)."""
    packs = [pack_linear_tc(w, b, True) for w, b in zip(ws, biases)]
    assert len(packs) == 2 and packs[0][2:4] == packs[1][2:4] and packs[0][0].numel() == packs[1][0].numel()
    return (torch.cat([packs[0][0], packs[1][0]]), torch.cat([packs[0][1], packs[1][1]]), packs[0][2], packs[0][3],
            packs[0][0].numel(), packs[0][1].numel())


def linear_tc_grouped(packed2, x_img, Mp, M, act=L.ACT_NONE, out_image=False, light=True):
    """act(x @ w_g.T + bias_g) for the two row groups of a stacked split image (pdf_gemm_bf16_grouped): x_img holds
    2 * Mp rows (Mp % 128 == 0, the first M of each group real).  Returns fp32 rows [2*Mp, N] (a view of a buffer
    with a pitch padded to 4) or, with out_image, only the split operand image of the result (N % 64 == 0)."""
    w_img, b, N, K, wstride, bstride = packed2
    L.require_cuda(x_img, w_img)
    assert Mp % 128 == 0 and 0 < M <= Mp
    kb = 3 * ((K + 63) // 64)
    nt = (N + 127) // 128
    act = act | (L.GEMM_LIGHT if light else 0)
    if out_image:
        assert N % 64 == 0
        img = split_image_empty(2 * Mp, N, x_img.device)
        flat = [int(v) for i in range(nt) for v in (128 * i, min(128, N - 128 * i), 2 * i)]
        desc = (ctypes.c_int32 * len(flat))(*flat)
        L.call("pdf_gemm_bf16_grouped", L.ptr(x_img), Mp // 128, kb, L.ptr(w_img), nt, kb, wstride, kb, L.ptr(b), bstride,
               act | L.GEMM_OUT_SPLIT, None, 0, M, L.ptr(img), 3 * (N // 64), ctypes.cast(desc, ctypes.c_void_p), L.stream())
        return img
    out = torch.empty((2 * Mp, _pad4(N)), dtype=torch.float32, device=x_img.device)
    flat = [int(v) for t in _tile_desc(N) for v in t]
    desc = (ctypes.c_int32 * len(flat))(*flat)
    L.call("pdf_gemm_bf16_grouped", L.ptr(x_img), Mp // 128, kb, L.ptr(w_img), nt, kb, wstride, kb, L.ptr(b), bstride, act,
           L.ptr(out), out.stride(0), M, None, 0, ctypes.cast(desc, ctypes.c_void_p), L.stream())
    return out[:, :N]


def linear_tn_mn(a_img, ca, b_img, cb, M, split=True):
    """a^T b [ca, cb] from the ROW tile images of a [M,ca] and b [M,cb] (pdf_gemm_tn_bf16: the tensor core
    reads the K-major blocks as MN-major operands; no transposed copies).  Images as written by
    rows_to_image(split=1) (or split=0 with split=False)."""
    L.require_cuda(a_img, b_img)
    row_tiles = (M + 127) // 128
    at, bt = (ca + 127) // 128, (cb + 127) // 128
    want = max(1, (2 * 148 + at * bt - 1) // (at * bt))                  # about two waves of work items
    tpb = max(4, (row_tiles + want - 1) // want)
    batches = (row_tiles + tpb - 1) // tpb
    ld = _pad4(cb)
    part = torch.zeros((batches, ca, ld), dtype=torch.float32, device=a_img.device)
    L.call("pdf_gemm_tn_bf16", L.ptr(a_img), ca, L.ptr(b_img), cb, M, 1 if split else 0, tpb, L.ptr(part), ld, ca * ld,
           L.stream())
    if batches == 1:
        return part[0, :, :cb]
    return col_sum(part.view(batches, ca * ld)).view(ca, ld)[:, :cb]


def linear_tn_tc(a, b, split=True):
    """a [M,N], b [M,K] -> a^T b [N,K] on the tcgen05 GEMM: split-K over batches of rows
    (pdf_rows_to_image_t + pdf_gemm_bf16_batched; split-bf16 or plain bf16 operands), partials
    summed in fp64."""
    L.require_cuda(a, b)
    M, N = _rows(a).shape
    K = _rows(b).shape[1]
    mt, nt = (N + 127) // 128, (K + 127) // 128
    want = max(1, (2 * 148 + mt * nt - 1) // (mt * nt))                 # batches: about two waves of work items
    Mc = max(512, ((M + want - 1) // want + 63) // 64 * 64)
    batches = (M + Mc - 1) // Mc
    kb = (3 if split else 1) * (Mc // 64)
    dev = a.device
    a_img = torch.empty((batches * mt * kb * 16384,), dtype=torch.uint8, device=dev)
    b_img = torch.empty((batches * nt * kb * 16384,), dtype=torch.uint8, device=dev)
    L.call("pdf_rows_to_image_t", L.ptr(a), a.stride(0), M, 0, N, L.ptr(a_img), Mc, 1 if split else 0, L.stream())
    L.call("pdf_rows_to_image_t", L.ptr(b), b.stride(0), M, 0, K, L.ptr(b_img), Mc, 2 if split else 0, L.stream())
    ldk = _pad4(K)
    part = torch.zeros((batches, N, ldk), dtype=torch.float32, device=dev)
    flat = [int(v) for t in _tile_desc(K) for v in t]
    desc = (ctypes.c_int32 * len(flat))(*flat)
    L.call("pdf_gemm_bf16_batched", L.ptr(a_img), mt, kb, mt * kb * 16384, L.ptr(b_img), nt, kb, nt * kb * 16384, kb,
           batches, L.ptr(part), ldk, N * ldk, N, ctypes.cast(desc, ctypes.c_void_p), L.stream())
    if batches == 1:
        return part[0, :, :K]
    return col_sum(part.view(batches, N * ldk)).view(N, ldk)[:, :K]


# ------------------------------------------------------------------------------------------------
# GCN decoder primitives (gcn_decoder.cu)
def split_image_empty(rows, C, device):
    """Uninitialised split-bf16 tile image ([hi|hi|lo]) for a [rows, C] matrix, C % 64 == 0."""
    assert C % 64 == 0
    return torch.empty((((rows + 127) // 128) * 3 * (C // 64) * 16384,), dtype=torch.uint8, device=device)


def row_combine(a, b=None, rowvec=None, V_out=None, up=1, ln=None, relu=False, want_sum=False, eps=1e-6,
                ln_out=None, sum_img=False, ln_img=False, ln_rows=True, sum_out=None):
    """t = a[src] (+ b[src]) (+ rowvec[v]); returns (t or None, LayerNorm(t) or None); see pdf_row_combine.
    ``ln`` = (gamma, beta).  Output rows = a.shape[0] * up.  sum_img / ln_img: also (or, with
    want_sum=False / ln_rows=False, only) write the result as a split-bf16 tile image; an image argument
    may be a preallocated uint8 view; the return value then is (t, ln, t_image, ln_image)."""
    L.require_cuda(a, b, rowvec, ln_out)
    Ma, C = _rows(a).shape
    rows = Ma * up
    if V_out is None:
        V_out = up
    dev = a.device
    s_out = sum_out if sum_out is not None else (torch.empty((rows, C), dtype=torch.float32, device=dev) if want_sum else None)
    assert s_out is None or (s_out.shape == (rows, C) and s_out.is_contiguous())
    if ln is not None and ln_out is None and ln_rows:
        ln_out = torch.empty((rows, C), dtype=torch.float32, device=dev)
    s_im = split_image_empty(rows, C, dev) if sum_img is True else (sum_img if sum_img is not False else None)
    l_im = split_image_empty(rows, C, dev) if ln_img is True else (ln_img if ln_img is not False else None)
    gamma, beta = ln if ln is not None else (None, None)
    L.call("pdf_row_combine", L.ptr(a), a.stride(0), L.ptr(b), b.stride(0) if b is not None else 0, L.ptr(rowvec),
           rowvec.stride(0) if rowvec is not None else 0, V_out, up, C, rows, L.ptr(gamma), L.ptr(beta), float(eps),
           int(relu), L.ptr(s_out), C, L.ptr(ln_out), ln_out.stride(0) if ln_out is not None else 0, L.ptr(s_im),
           L.ptr(l_im), L.stream())
    if s_im is None and l_im is None:
        return s_out, ln_out
    return s_out, ln_out, s_im, l_im


def row_combine_grouped(a, b, Mp, M, src_per_group, rowvec=None, rowvec_gstride=0, V_out=None, up=1, ln=None,
                        relu=False, want_sum=False, sum_img=False, ln_img=False, eps=1e-6):
    """pdf_row_combine_grouped: the two row groups (hands) of stacked buffers in one launch.  a / b hold the
    sources ([2 * src_per_group, C]); outputs hold 2 * Mp rows (the first M of each group are written).  ``ln`` =
    (gamma [2,C], beta [2,C]).  Returns (sum rows or None, sum image or None, LayerNorm image or None)."""
    L.require_cuda(a, b, rowvec)
    C = _rows(a).shape[1]
    if V_out is None:
        V_out = up
    dev = a.device
    s_out = torch.empty((2 * Mp, C), dtype=torch.float32, device=dev) if want_sum else None
    s_im = split_image_empty(2 * Mp, C, dev) if sum_img else None
    l_im = split_image_empty(2 * Mp, C, dev) if ln_img else None
    gamma, beta = ln if ln is not None else (None, None)
    L.call("pdf_row_combine_grouped", L.ptr(a), a.stride(0), L.ptr(b), b.stride(0) if b is not None else 0, L.ptr(rowvec),
           rowvec.stride(-2) if rowvec is not None else 0, rowvec_gstride, V_out, up, C, Mp, M, src_per_group,
           L.ptr(gamma), L.ptr(beta), float(eps), int(relu), L.ptr(s_out), C, None, 0, L.ptr(s_im), L.ptr(l_im), L.stream())
    return s_out, s_im, l_im


def graph_cheby_ln_grouped(U0, U1, bias2, csr, V, Mp, M, ln2, relu, R=None, bias_r2=None, eps=1e-6, want_rows=True,
                           want_img=False):
    """pdf_graph_cheby_ln_grouped: U0 / U1 / R are column slices of a [2*Mp, .] GEMM output, bias2 / bias_r2 /
    ln2 = (gamma, beta) are stacked [2, C].  Returns (rows or None, image or None)."""
    L.require_cuda(U0, U1, bias2, R)
    C = _rows(U0).shape[1]
    assert U1.stride(0) == U0.stride(0) and U0.shape[0] == 2 * Mp
    rowptr, colidx, vals = csr
    out = torch.empty((2 * Mp, C), dtype=torch.float32, device=U0.device) if want_rows else None
    img = split_image_empty(2 * Mp, C, U0.device) if want_img else None
    L.call("pdf_graph_cheby_ln_grouped", L.ptr(U0), L.ptr(U1), U0.stride(0), L.ptr(bias2), L.ptr(R),
           R.stride(0) if R is not None else 0, L.ptr(bias_r2), L.ptr(rowptr), L.ptr(colidx), L.ptr(vals), V, C, Mp, M,
           L.ptr(ln2[0]), L.ptr(ln2[1]), float(eps), int(relu), L.ptr(out), C, L.ptr(img), L.stream())
    return out, img


def graph_cheby_ln(U0, U1, bias, csr, V, ln, relu, R=None, bias_r=None, eps=1e-6, want_rows=True, want_img=False):
    """LayerNorm(U0 + bias + L.U1 (+ R + bias_r)) (+ReLU); U0/U1 (and R) are column slices of GEMM outputs.
    Returns fp32 rows, or (rows or None, split-bf16 tile image) with want_img."""
    L.require_cuda(U0, U1, bias, R)
    rows, C = _rows(U0).shape
    assert U1.stride(0) == U0.stride(0) and U1.shape == U0.shape
    rowptr, colidx, vals = csr
    out = torch.empty((rows, C), dtype=torch.float32, device=U0.device) if want_rows else None
    img = split_image_empty(rows, C, U0.device) if want_img else None
    L.call("pdf_graph_cheby_ln", L.ptr(U0), L.ptr(U1), U0.stride(0), L.ptr(bias), L.ptr(R),
           R.stride(0) if R is not None else 0, L.ptr(bias_r), L.ptr(rowptr), L.ptr(colidx), L.ptr(vals), V, C, rows,
           L.ptr(ln[0]), L.ptr(ln[1]), float(eps), int(relu), L.ptr(out), C, L.ptr(img), L.stream())
    return (out, img) if want_img else out


def mha(q, k, v, n_samples, V, heads, out=None):
    """softmax(q k^T / sqrt(d)) v per (sample, head); q/k/v [n_samples*V, heads*d] column slices."""
    L.require_cuda(q, k, v, out)
    M, f = _rows(q).shape
    assert M == n_samples * V and f % heads == 0
    if out is None:
        out = torch.empty((M, f), dtype=torch.float32, device=q.device)
    L.call("pdf_mha", L.ptr(q), q.stride(0), L.ptr(_rows(k)), k.stride(0), L.ptr(_rows(v)), v.stride(0), n_samples, V,
           heads, f // heads, L.ptr(out), out.stride(0), L.stream())
    return out


def mha_tc(problems, n_samples, V, heads, rows=True, image=None, image_rows=None):
    """Tensor-core attention (pdf_mha_tc) for one or two problems of identical shape in ONE launch.
    problems: [(q, k, v, out-or-None), ...] with q/k/v [n_samples*V, heads*d] column slices of fp32 rows that
    share their row pitches.  rows=True: fp32 row outputs (returned as a list).  image: True (allocate) or a
    uint8 tensor: the results are ALSO / ONLY written as the split-bf16 tile image of the stacked
    [len(problems) * n_samples * V, heads*d] matrix (problem i fills rows [i*M, (i+1)*M), or the rows starting at
    image_rows[i]) - the operand of the ``fc`` GEMM that follows.  Returns outs, or (outs, image) when an image is
    requested."""
    outs, ptrs = [], ([], [], [], [])
    q0, k0, v0, _ = problems[0]
    M, f = _rows(q0).shape
    assert M == n_samples * V and f % heads == 0 and 1 <= len(problems) <= 2
    for q, k, v, out in problems:
        L.require_cuda(q, k, v, out)
        assert _rows(q).shape == (M, f) and _rows(k).shape == (M, f) and _rows(v).shape == (M, f)
        assert q.stride(0) == q0.stride(0) and k.stride(0) == k0.stride(0) and v.stride(0) == v0.stride(0)
        if out is None and rows:
            out = torch.empty((M, f), dtype=torch.float32, device=q.device)
        if out is not None:
            assert out.stride(0) == (outs[0].stride(0) if outs else out.stride(0)) and out.stride(1) == 1
            outs.append(out)
        for lst, t in zip(ptrs, (q, k, v, out)):
            if t is not None:
                L.ptr(t)                                    # registers the device for the launch guard
            lst.append(t.data_ptr() if t is not None else None)
    n = len(problems)
    arr = [(ctypes.c_void_p * n)(*lst) for lst in ptrs]
    img_arr = row0 = None
    if image is not None and image is not False:
        if image is True:
            image = split_image_empty(n * M, f, q0.device)
        L.ptr(image)
        img_arr = ctypes.cast((ctypes.c_void_p * n)(*([image.data_ptr()] * n)), ctypes.c_void_p)
        starts = list(image_rows) if image_rows is not None else [i * M for i in range(n)]
        assert len(starts) == n
        row0 = ctypes.cast((ctypes.c_int64 * n)(*starts), ctypes.c_void_p)
    L.call("pdf_mha_tc", ctypes.cast(arr[0], ctypes.c_void_p), ctypes.cast(arr[1], ctypes.c_void_p),
           ctypes.cast(arr[2], ctypes.c_void_p), ctypes.cast(arr[3], ctypes.c_void_p) if outs else None, img_arr, row0, n,
           q0.stride(0), k0.stride(0), v0.stride(0), outs[0].stride(0) if outs else 0, n_samples, V, heads, f // heads,
           L.stream())
    return (outs, image) if img_arr is not None else outs


def decoder_heads(f, n, V, avg, params_head, root_head, coord_head):
    """pdf_decoder_heads: f [n*V, C] rows -> (params [n,3], root [n,3], verts [n,V,3]); the heads are (weight, bias)."""
    L.require_cuda(f)
    M, C = _rows(f).shape
    assert M == n * V
    dev = f.device
    params = torch.empty((n, 3), dtype=torch.float32, device=dev)
    root = torch.empty((n, 3), dtype=torch.float32, device=dev)
    verts = torch.empty((n, V, 3), dtype=torch.float32, device=dev)
    c = lambda t: L.f32c(t.detach())
    (aw, ab), (pw, pb), (rw, rb), (cw, cb) = [(c(w).reshape(-1) if i == 0 else c(w), c(b)) for i, (w, b) in
                                              enumerate((avg, params_head, root_head, coord_head))]
    L.call("pdf_decoder_heads", L.ptr(f), f.stride(0), n, V, C, L.ptr(aw), L.ptr(ab), L.ptr(pw), L.ptr(pb), L.ptr(rw),
           L.ptr(rb), L.ptr(cw), L.ptr(cb), L.ptr(params), L.ptr(root), L.ptr(verts), L.stream())
    return params, root, verts


def decoder_project(v_coarse, v_dense, params, img_size, rev, rep):
    """-> (coarse2d [B,Vc,2], dense2d [B,Vd,2], mano3d [B,Vd,3], mano2d [B,Vd,2]); see pdf_decoder_project."""
    L.require_cuda(v_coarse, v_dense, params, rev)
    B, Vc, _ = v_coarse.shape
    Vd = v_dense.shape[1]
    assert v_coarse.is_contiguous() and v_dense.is_contiguous() and params.stride(1) == 1 and rev.dtype == torch.int64
    dev = v_coarse.device
    c2 = torch.empty((B, Vc, 2), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, Vd, 2), dtype=torch.float32, device=dev)
    m3 = torch.empty((B, Vd, 3), dtype=torch.float32, device=dev)
    m2 = torch.empty((B, Vd, 2), dtype=torch.float32, device=dev)
    L.call("pdf_decoder_project", L.ptr(v_coarse), Vc, L.ptr(v_dense), Vd, L.ptr(params), params.stride(0),
           float(img_size), L.ptr(rev), rep, B, L.ptr(c2), L.ptr(d2), L.ptr(m3), L.ptr(m2), L.stream())
    return c2, d2, m3, m2
