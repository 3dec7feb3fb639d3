"""CPU oracle: a restatement of the reference's depth-branch / fusion / MANO hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker or the timed CPU baseline.  Nothing under
``pdfnet_b200/`` imports it; the product path fails loudly when the CUDA
library is missing.

Parity status: **pinned**.  The reference (zijinxuxu/PDFNet) ships no tests or
golden vectors, so the pin is the reference itself: every
function here is checked (``tests/test_oracle_golden.py``) against outputs of
the UNMODIFIED reference functions run on CPU in the authoring container and
frozen under ``tests/golden/`` by ``oracle/make_golden.py``.

Arithmetic is numpy / torch-CPU fp32, the same libraries the reference uses.
Reference citations are ``path:line`` relative to the reference root.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# distances + kNN-then-radius-mask ("ball query")
# ----------------------------------------------------------------------------


def sqdist(xyz, n_centroids):
    """Squared distances d2[b,i,j] = |p_j - p_i|^2 for the first ``n_centroids`` rows.

    lib/utils/utils.py:142-145 (and :169-172): ``diff = p_j - c_i``;
    ``mul(diff, diff)``; ``sum(2)`` over 3 channels.  torch-CPU / numpy evaluate
    the 3-term sum as fl(fl(dx2+dy2)+dz2) with no FMA (SURVEY.md section 7,
    "Bit-exact distances").
    """
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    c = xyz[:, :n_centroids, None, :]                       # [B,N1,1,3]
    d = xyz[:, None, :, :] - c                              # p_j - c_i
    d = d * d
    return (d[..., 0] + d[..., 1]) + d[..., 2]              # fp32 throughout


def knn_ball_indices(xyz, n_centroids, K, r2):
    """Indices of the K nearest points of every centroid, radius-masked.

    lib/utils/utils.py:146-151 / :173-179: ``topk(K, largest=False,
    sorted=False)`` then every neighbour with ``d2 > r2`` is replaced by the
    centroid's own index.  ``r2`` is compared in fp32 (a python scalar against a
    float32 tensor).  ``topk(sorted=False)`` leaves the order inside a group
    and the choice among exact ties at the K-th distance implementation-defined
    (SURVEY.md section 8c); this oracle returns each group SORTED ASCENDING BY
    INDEX and breaks ties towards the lowest index, which is the canonical form
    the parity tests compare in.
    """
    d2 = sqdist(xyz, n_centroids)                           # [B,N1,N]
    B, N1, N = d2.shape
    order = np.argsort(d2, axis=2, kind="stable")[:, :, :K]  # (distance, index) lexicographic
    dk = np.take_along_axis(d2, order, axis=2)
    own = np.arange(N1, dtype=np.int64)[None, :, None]
    idx = np.where(dk > np.float32(r2), own, order.astype(np.int64))
    return np.sort(idx, axis=2)


def canonicalize_indices(idx, xyz):
    """Map every index to the smallest index holding a bit-identical xyz row,
    then sort each group (SURVEY.md section 8c parity rule for duplicate points)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.empty_like(idx)
    for b in range(xyz.shape[0]):
        rows = xyz[b].view(np.uint32).reshape(xyz.shape[1], -1)
        _, first, inv = np.unique(rows, axis=0, return_index=True, return_inverse=True)
        canon = first[inv.reshape(-1)]
        out[b] = canon[idx[b]]
    return np.sort(out, axis=-1)


def group_points(points, opt, idx=None):
    """lib/utils/utils.py:134-162.  points [B,N,C>=3] f32 ->
    (x [B,C,N1,K] f32 with centroid-relative xyz, center [B,3,N1,1])."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    N1, K = opt.sample_num_level1, opt.knn_K
    C = opt.INPUT_FEATURE_NUM
    if idx is None:
        idx = knn_ball_indices(points[:, :, 0:3], N1, K, opt.ball_radius)
    B = points.shape[0]
    g = np.take_along_axis(points[:, :, None, :C], idx.reshape(B, N1 * K, 1, 1), axis=1)
    g = g.reshape(B, N1, K, C).copy()
    center = points[:, :N1, None, 0:3]
    g[..., 0:3] = g[..., 0:3] - center
    return g.transpose(0, 3, 1, 2), center.transpose(0, 3, 1, 2)


def group_points_2(points, N1, N2, K, r2, idx=None):
    """lib/utils/utils.py:165-187.  points [B,C,N1] f32 (channel-major, xyz =
    channels 0:3) -> (x [B,C,N2,K], center [B,3,N2,1])."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    B, C, _ = points.shape
    xyz = points[:, 0:3, :].transpose(0, 2, 1)
    if idx is None:
        idx = knn_ball_indices(xyz, N2, K, r2)
    g = np.take_along_axis(points, idx.reshape(B, 1, N2 * K), axis=2).reshape(B, C, N2, K).copy()
    center = points[:, 0:3, :N2, None]
    g[:, 0:3] = g[:, 0:3] - center
    return g, center


# ----------------------------------------------------------------------------
# pixel -> point pyramid gather + SFT modulation
# ----------------------------------------------------------------------------


def tranpose_and_gather_feat(feat, ind):
    """lib/models/utils.py:12-26: NCHW -> [B,HW,C] then row gather. -> [B,n,C]."""
    feat = torch.as_tensor(feat)
    ind = torch.as_tensor(ind).long()
    B, C = feat.shape[0], feat.shape[1]
    f = feat.reshape(B, C, -1)
    return torch.gather(f, 2, ind[:, None, :].expand(B, C, ind.shape[1])).transpose(1, 2).contiguous()


def pyramid_index(choose, R):
    """lib/models/networks/intaghand_encoder.py:125-126 (floor division, int64)."""
    choose = torch.as_tensor(choose).long()
    c2 = (choose // R // 2) * (R // 2) + choose % R // 2
    c4 = (choose // R // 4) * (R // 4) + choose % R // 4
    return c2, c4


def sft_layer(fea, cond, sd, prefix=""):
    """SFTLayer.forward, intaghand_encoder.py:205-219.
    fea [B,Cf,n], cond [B,n,Cc] -> [B,n,Cf];  1x1 convs with bias, leaky 0.1."""
    fea = torch.as_tensor(fea).unsqueeze(3)
    c = torch.as_tensor(cond).transpose(1, 2).unsqueeze(3)

    def conv(name, x):
        return F.conv2d(x, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"])

    scale = conv("SFT_scale_conv1", F.leaky_relu(conv("SFT_scale_conv0", c), 0.1))
    shift = conv("SFT_shift_conv1", F.leaky_relu(conv("SFT_shift_conv0", c), 0.1))
    return (fea * (scale + 1) + shift).transpose(1, 2).squeeze(-1)


def _conv_bn_relu(x, sd, prefix, i_conv, i_bn, eps=1e-5):
    x = F.conv2d(x, sd["%s.%d.weight" % (prefix, i_conv)], sd["%s.%d.bias" % (prefix, i_conv)])
    x = F.batch_norm(x, sd["%s.%d.running_mean" % (prefix, i_bn)], sd["%s.%d.running_var" % (prefix, i_bn)],
                     sd["%s.%d.weight" % (prefix, i_bn)], sd["%s.%d.bias" % (prefix, i_bn)], False, 0.0, eps)
    return F.relu(x)


def point_mlp_max(x, sd, prefix, pool_dim):
    """netR_1 / netR_2 / netR_3 (intaghand_encoder.py:48-103) in eval mode:
    (Conv1x1 -> BatchNorm2d -> ReLU) x3 -> max over ``pool_dim``."""
    x = torch.as_tensor(x)
    for i_conv, i_bn in ((0, 1), (3, 4), (6, 7)):
        x = _conv_bn_relu(x, sd, prefix, i_conv, i_bn)
    return x.max(dim=pool_dim, keepdim=True)[0]


def pointnet_plus_forward(sd, points, emb, choose, opt, return_intermediates=False):
    """PointNet_Plus.forward, intaghand_encoder.py:118-159 (eval mode).
    points [B,N,3], emb = [l0 [B,3,R,R], l1 [B,64,R/2,R/2], l2 [B,256,R/4,R/4]],
    choose [B,N] int64 -> [B,1,1024]."""
    with torch.no_grad():
        points = torch.as_tensor(points, dtype=torch.float32)
        choose = torch.as_tensor(choose).long()
        R = opt.default_resolution
        N1, N2, K = opt.sample_num_level1, opt.sample_num_level2, opt.knn_K
        e0 = tranpose_and_gather_feat(emb[0], choose)
        pts0 = sft_layer(points.transpose(1, 2), e0, sd, "sft0.")                 # [B,N,3]
        x, y = group_points(pts0.numpy(), opt)
        x, y = torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(y))
        c2, c4 = pyramid_index(choose, R)
        e1 = tranpose_and_gather_feat(emb[1], c2[:, :N1])
        e2 = tranpose_and_gather_feat(emb[2], c4[:, :N2])
        f1 = point_mlp_max(x, sd, "netR_1", 3)                                    # [B,128,N1,1]
        x1 = torch.cat((y, f1), 1).squeeze(-1)                                    # [B,131,N1]
        pts1 = sft_layer(x1, e1, sd, "sft1.").transpose(1, 2)                     # [B,131,N1]
        g2, c2xyz = group_points_2(pts1.numpy(), N1, N2, K, opt.ball_radius2)
        g2, c2xyz = torch.from_numpy(g2), torch.from_numpy(np.ascontiguousarray(c2xyz))
        f2 = point_mlp_max(g2, sd, "netR_2", 3)                                   # [B,256,N2,1]
        x2 = torch.cat((c2xyz, f2), 1)
        pts2 = sft_layer(x2.squeeze(-1), e2, sd, "sft2.").transpose(1, 2).unsqueeze(3)
        out = point_mlp_max(pts2, sd, "netR_3", 2).view(-1, 1, 1024)
        if return_intermediates:
            return out, dict(pts0=pts0, f1=f1.squeeze(-1), pts1=pts1, f2=f2.squeeze(-1), pts2=pts2.squeeze(-1))
        return out


# ----------------------------------------------------------------------------
# training mode (BASELINE cfg5): the same forward with batch-statistic BatchNorm,
# written with differentiable torch-CPU ops so torch.autograd yields the gradients
# the reference's autograd yields.  Pinned by tests/golden/train_step.npz (the
# unmodified reference PointNet_Plus in .train() mode, loss.backward()).
# ----------------------------------------------------------------------------


def _group_rows_torch(pts, idx):
    """Differentiable form of the gathers in utils.py:153-158 / :181-186.
    pts [B,N,C] tensor, idx [B,N1,K] int64 -> [B,N1,K,C] with centroid-relative xyz."""
    B, N, C = pts.shape
    N1, K = idx.shape[1], idx.shape[2]
    g = torch.gather(pts, 1, idx.reshape(B, N1 * K, 1).expand(B, N1 * K, C)).view(B, N1, K, C)
    center = pts[:, :N1, None, 0:3]
    return torch.cat((g[..., 0:3] - center, g[..., 3:]), -1)


def _mlp_max_train(x, sd, prefix, pool_dim, momentum=0.1, eps=1e-5):
    for i_conv, i_bn in ((0, 1), (3, 4), (6, 7)):
        x = F.conv2d(x, sd["%s.%d.weight" % (prefix, i_conv)], sd["%s.%d.bias" % (prefix, i_conv)])
        x = F.batch_norm(x, sd["%s.%d.running_mean" % (prefix, i_bn)], sd["%s.%d.running_var" % (prefix, i_bn)],
                         sd["%s.%d.weight" % (prefix, i_bn)], sd["%s.%d.bias" % (prefix, i_bn)], True, momentum, eps)
        x = F.relu(x)
    return x.max(dim=pool_dim, keepdim=True)[0]


def pointnet_plus_train(sd, points, emb, choose, opt, dtype=torch.float32):
    """PointNet_Plus.forward (intaghand_encoder.py:118-159) in .train() mode.  ``sd`` maps
    state-dict names to tensors (parameters with requires_grad=True; running buffers are
    updated in place as nn.BatchNorm2d does); ``emb`` tensors may require grad.  -> [B,1,1024].
    ``dtype=torch.float64`` (all tensors double) is the form the golden pins: fp32 autograd of
    this network carries ~1 % rounding noise in the early-layer gradients, fp64 does not."""
    points = torch.as_tensor(points, dtype=dtype)
    choose = torch.as_tensor(choose).long()
    R = opt.default_resolution
    N1, N2, K = opt.sample_num_level1, opt.sample_num_level2, opt.knn_K
    e0 = tranpose_and_gather_feat(emb[0], choose)
    pts0 = sft_layer(points.transpose(1, 2), e0, sd, "sft0.")                      # [B,N,3]
    idx1 = torch.from_numpy(knn_ball_indices(pts0.detach().numpy(), N1, K, opt.ball_radius))
    g1 = _group_rows_torch(pts0, idx1).permute(0, 3, 1, 2)                         # [B,3,N1,K]
    y = pts0[:, :N1, 0:3].transpose(1, 2).unsqueeze(-1)                            # [B,3,N1,1]
    c2, c4 = pyramid_index(choose, R)
    e1 = tranpose_and_gather_feat(emb[1], c2[:, :N1])
    e2 = tranpose_and_gather_feat(emb[2], c4[:, :N2])
    f1 = _mlp_max_train(g1, sd, "netR_1", 3)
    x1 = torch.cat((y, f1), 1).squeeze(-1)                                         # [B,131,N1]
    pts1 = sft_layer(x1, e1, sd, "sft1.")                                          # [B,N1,131]
    idx2 = torch.from_numpy(knn_ball_indices(pts1.detach().numpy()[:, :, 0:3], N2, K, opt.ball_radius2))
    g2 = _group_rows_torch(pts1, idx2).permute(0, 3, 1, 2)                         # [B,131,N2,K]
    c2xyz = pts1[:, :N2, 0:3].transpose(1, 2).unsqueeze(-1)
    f2 = _mlp_max_train(g2, sd, "netR_2", 3)
    x2 = torch.cat((c2xyz, f2), 1).squeeze(-1)                                     # [B,259,N2]
    pts2 = sft_layer(x2, e2, sd, "sft2.").transpose(1, 2).unsqueeze(3)
    return _mlp_max_train(pts2, sd, "netR_3", 2).view(-1, 1, 1024)


def fusion_tail(sd_pointnet, sd_sft, cloud, emb, choose, center_features, opt):
    """ResNetSimple.forward fusion tail, intaghand_encoder.py:805-809.
    cloud [B,2,N,3], choose [B,2,N], center_features [B,2,1024] -> fuse_feat [B,2,1024]."""
    with torch.no_grad():
        left = pointnet_plus_forward(sd_pointnet, cloud[:, 0], emb, choose[:, 0], opt)
        right = pointnet_plus_forward(sd_pointnet, cloud[:, 1], emb, choose[:, 1], opt)
        fuse = torch.cat((left, right), dim=1)
        return sft_layer(fuse.transpose(1, 2).contiguous(), torch.as_tensor(center_features), sd_sft, "")


def center_features(x0, w_up0, w_up1, ind):
    """ResNetSimple.forward centre features, intaghand_encoder.py:790-792:
    x0_up1 = center_feat_up1(center_feat_up0(x0)) (3x3, pad 1, no bias) over the WHOLE map, then
    _tranpose_and_gather_feat(x0_up1, ind).  x0 [B,C,H,W], ind [B,2] -> [B,2,1024]."""
    with torch.no_grad():
        up1 = F.conv2d(F.conv2d(torch.as_tensor(x0), w_up0, padding=1), w_up1, padding=1)
        return tranpose_and_gather_feat(up1, ind)


# ----------------------------------------------------------------------------
# farthest point sampling
# ----------------------------------------------------------------------------


def fps_order(points, n_sample, start_idx):
    """farthest_point_sampling_fast, lib/datasets/interhand.py:147-178, for
    pc_num > n_sample, with the random start (:159) injected.  Returns the index
    sequence IN SELECTION ORDER (the reference returns np.unique of it, :177).

    min_dist = |p - p_start|^2 (fp32, (dx2+dy2)+dz2); each round: argmax (first
    occurrence), then ONLY entries with min_dist > 1e-8 are lowered (:171-175).
    """
    pc = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros((n_sample,), dtype=np.int64)
    out[0] = start_idx
    diff = pc - pc[start_idx][None, :]
    min_dist = np.sum(diff * diff, 1)
    for s in range(1, n_sample):
        out[s] = np.argmax(min_dist)
        valid = min_dist > 1e-8
        diff = pc[valid] - pc[out[s]][None, :]
        min_dist[valid] = np.minimum(min_dist[valid], np.sum(diff * diff, 1))
    return out


def fps_batch(xyz, n_sample, start_idx):
    return np.stack([fps_order(xyz[b], n_sample, int(start_idx[b])) for b in range(xyz.shape[0])])


# ----------------------------------------------------------------------------
# depth back-projection and per-hand cloud construction
# ----------------------------------------------------------------------------


def backproject(depth, K):
    """get_normal(with_normal=False) -> get_points_coordinate,
    lib/utils/utils.py:264-275, :251-262.  depth [H,W] (any float), K [3,3] ->
    xyz [3,H,W] f32 = (inv(K)[:3,:3] @ [u,v,1]) * z, u = column, v = row.
    ``np.linalg.inv`` runs in the dtype of K; depth2pcl passes float32
    (intaghand_encoder.py:373)."""
    depth = np.asarray(depth)
    H, W = depth.shape
    Kinv = torch.from_numpy(np.linalg.inv(np.asarray(K)))[:3, :3].unsqueeze(0)
    d = torch.from_numpy(depth).unsqueeze(0).unsqueeze(-1).float()
    y, x = torch.meshgrid([torch.arange(0, H, dtype=torch.float32),
                           torch.arange(0, W, dtype=torch.float32)], indexing="ij")
    uv1 = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W)))[None]
    xyz = torch.matmul(Kinv, uv1) * d.view(1, 1, -1)
    return xyz.view(3, H, W).numpy()


def hand_candidates(xyz, z_min=0.2, z_max=2.5, half_window=0.08):
    """Candidate flat pixel indices of one hand, intaghand_encoder.py:406-411:
    mean z over non-zero pixels, window mean+-0.08 clipped to [z_min,z_max]."""
    z = xyz.reshape(3, -1)[2]
    nz = z[z != 0]
    if len(nz) == 0:
        return None
    mean_dis = nz.mean()
    lo, hi = max(z_min, mean_dis - half_window), min(z_max, mean_dis + half_window)
    return ((z > lo) & (z < hi)).nonzero()[0]


def depth2pcl(depth, mask, K, valid, subset_keys, perm, num_points=1024, min_pixels=10):
    """depth2pcl, intaghand_encoder.py:369-491, batch-1, with the two uses of
    ``np.random.shuffle`` replaced by injected randomness:

    * >num_points candidates (:418-422): the reference shuffles a 0/1 mask with
      ``num_points`` ones and keeps candidates where it is 1 (order preserved).
      Here the kept candidates are those with the ``num_points`` smallest
      ``subset_keys[h][pixel]`` (ties -> lower pixel), order preserved.
    * final shuffle (:427): ``choose = choose[perm[h]]``.

    depth [H,W] f32 metres, mask [1,2,H,W] (channel 0 = right, 1 = left, :376-377),
    valid [1,2] (0 = left, 1 = right, :401,:439).  Returns choose int64 [2,num_points]
    (row 0 = left) and cloud f32 [2,num_points,3].
    The mask here is already at depth resolution, so cv2.resize (:376) is identity.
    """
    depth = np.asarray(depth, dtype=np.float32)
    m = (np.asarray(mask) > 0.5).astype(np.uint8)
    K = np.asarray(K).astype(np.float32)
    noise = ((0.2 < depth) & (2.5 > depth)).astype(np.uint8)
    d = depth * noise
    chooses, clouds = [], []
    for h, mch in ((0, 1), (1, 0)):                     # left uses mask[0,1], right mask[0,0]
        if valid[0, h] == 1:
            xyz = backproject((d * m[0, mch]).squeeze(), K).reshape(3, -1)
            cand = hand_candidates(xyz)
            if cand is None or len(cand) < min_pixels:
                ch = np.zeros((num_points,), dtype=np.int64)
            elif len(cand) > num_points:
                keys = np.asarray(subset_keys[h])[cand]
                keep = np.sort(np.argsort(keys, kind="stable")[:num_points])
                ch = cand[keep]
            else:
                ch = np.pad(cand, (0, num_points - len(cand)), "wrap")
            ch = ch[np.asarray(perm[h])]
            pts = xyz.transpose(1, 0)[ch, :]
        else:
            ch = np.zeros((num_points,), dtype=np.int64)
            pts = np.zeros((num_points, 3), dtype=np.float32)
        chooses.append(ch)
        clouds.append(pts)
    return np.stack(chooses), np.stack(clouds)


def _mix32(h):
    """murmur3 finaliser on uint32 arrays."""
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def d2p_seeded_randomness(seed, n_clouds, npx):
    """The injected randomness pdf_depth2pcl_seeded generates from ``seed`` (include/pdfnet_b200.h), restated
    in numpy: keys int32 [n_clouds, npx] (signed order = the kernel's unsigned hash order) and perm int32
    [n_clouds, 1024] (4-round Feistel bijection on two 5-bit halves).  cloud = 2*frame + hand."""
    with np.errstate(over="ignore"):
        seed = np.uint32(seed & 0xFFFFFFFF)
        c = np.arange(n_clouds, dtype=np.uint32)[:, None]
        pix = np.arange(npx, dtype=np.uint32)[None, :]
        cs = _mix32(seed ^ (c * np.uint32(0x27D4EB2F) + np.uint32(0x165667B1)).astype(np.uint32))
        keys = _mix32((pix * np.uint32(0x9E3779B1)).astype(np.uint32) + cs)
        keys = (keys ^ np.uint32(0x80000000)).view(np.int32)
        k = _mix32((seed * np.uint32(0x9E3779B1)).astype(np.uint32) + c + np.uint32(0x7F4A7C15))
        i = np.arange(1024, dtype=np.uint32)[None, :]
        l, r = i >> np.uint32(5), i & np.uint32(31)
        for rnd in range(4):
            f = _mix32(r + np.uint32(32 * rnd) + k) & np.uint32(31)
            l, r = r, l ^ f
        perm = ((l << np.uint32(5)) | r).astype(np.int32)
    return keys, perm


# ----------------------------------------------------------------------------
# MANO head, coefficient split, linear blend skinning
# ----------------------------------------------------------------------------


def mano_head(x, sd, prefix="mano_head"):
    """mano_head, intaghand_encoder.py:630-643, eval mode:
    Linear(1024,512) BN1d ReLU Linear(512,256) BN1d ReLU Linear(256,122)."""
    x = torch.as_tensor(x)
    for i_fc, i_bn in ((0, 1), (3, 4)):
        x = F.linear(x, sd["%s.%d.weight" % (prefix, i_fc)], sd["%s.%d.bias" % (prefix, i_fc)])
        x = F.batch_norm(x, sd["%s.%d.running_mean" % (prefix, i_bn)], sd["%s.%d.running_var" % (prefix, i_bn)],
                         sd["%s.%d.weight" % (prefix, i_bn)], sd["%s.%d.bias" % (prefix, i_bn)], False, 0.0, 1e-5)
        x = F.relu(x)
    return F.linear(x, sd["%s.6.weight" % prefix], sd["%s.6.bias" % prefix])


def split_coeff(theta, index, K, input_res, down_ratio):
    """ManoRender.Split_coeff non-PCA branch, lib/models/hand3d/Mano_render.py:160-194.
    theta [B,122], index [B] (flat index on the input_res/down_ratio grid), K [B,3,3].
    Returns (orient_l, pose_l, betas_l, trans_l, orient_r, pose_r, betas_r, trans_r).
    Betas are multiplied by 0 (:163,:169); t_z += 0.6 (:165,:171); the SAME
    centre pixel (cx,cy) is used for both hands (:179-187)."""
    theta = torch.as_tensor(theta).clone()
    index = torch.as_tensor(index)
    K = torch.as_tensor(K)
    outs = []
    fx, fy, cw, ch = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    g = input_res // down_ratio
    cx = (index % g) * down_ratio
    cy = (index // g) * down_ratio
    for o in (0, 61):
        orient = theta[:, o:o + 3]
        pose = theta[:, o + 3:o + 48]
        betas = theta[:, o + 48:o + 58] * 0
        t = theta[:, o + 58:o + 61].clone()
        t[:, 2] = t[:, 2] + 0.6
        tx = t[:, 2] * (t[:, 0] + cx - cw) / fx
        ty = t[:, 2] * (t[:, 1] + cy - ch) / fy
        outs += [orient, pose, betas, torch.stack((tx, ty, t[:, 2]), 1)]
    return tuple(outs)


def rodrigues(axis):
    """rodrigues_batch, lib/models/networks/manolayer.py:32-48. [n,3] -> [n,3,3]."""
    axis = torch.as_tensor(axis)
    n = axis.shape[0]
    angle = torch.norm(axis, p=2, dim=1, keepdim=True) + 1e-8
    a = axis / angle
    s = torch.sin(angle).unsqueeze(2)
    c = torch.cos(angle).unsqueeze(2)
    L = torch.zeros((n, 3, 3), dtype=axis.dtype)
    L[:, 2, 1] = a[:, 0]
    L[:, 1, 2] = -a[:, 0]
    L[:, 0, 2] = a[:, 1]
    L[:, 2, 0] = -a[:, 1]
    L[:, 1, 0] = a[:, 2]
    L[:, 0, 1] = -a[:, 2]
    return torch.eye(3, dtype=axis.dtype).repeat(n, 1, 1) + s * L + (1 - c) * L.bmm(L)


MANO_PARENT = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]       # kintree_table[0]
MANO_NEW_ORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
MANO_TIPS = {"left": [745, 317, 445, 556, 673], "right": [745, 317, 444, 556, 673]}


def mano_lbs(tables, root_rotation, pose, shape, trans=None, scale=None, side="left",
             center_idx=None, new_skel=False, use_pca=False):
    """ManoLayer.forward, lib/models/networks/manolayer.py:257-334.
    tables: dict with v_template [778,3], shapedirs [778,3,10], posedirs [778,3,135],
    J_regressor [16,778] dense, weights [778,16] (f32) (+ hands_components [45,45], hands_mean [45]
    for use_pca).  use_pca=False (:268-272): axis-angle root [B,3] and pose [B,45].  use_pca=True
    (:266-267): root is a rotation MATRIX [B,3,3] used as is (:285) and pose holds PCA coefficients
    [B,ncomps] (pca2axis, :159-162).  shape [B,10].  Returns (v [B,778,3], j [B,21,3])."""
    T = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in tables.items()
         if k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "hands_components", "hands_mean")}
    root_rotation = torch.as_tensor(root_rotation, dtype=torch.float32)
    pose = torch.as_tensor(pose, dtype=torch.float32)
    shape = torch.as_tensor(shape, dtype=torch.float32)
    bs = root_rotation.shape[0]
    if use_pca:
        pose = pose.mm(T["hands_components"][:pose.shape[1]]) + T["hands_mean"]
        Rroot = root_rotation.reshape(bs, 3, 3)
    else:
        Rroot = rodrigues(root_rotation.reshape(-1, 3)).view(bs, 3, 3)
    Rpose = rodrigues(pose.reshape(-1, 3)).view(bs, 15, 3, 3)
    v_shaped = T["v_template"] + torch.matmul(T["shapedirs"], shape.permute(1, 0)).permute(2, 0, 1)
    j_tpose = torch.matmul(T["J_regressor"], v_shaped)
    pose_shape = Rpose.reshape(bs, -1) - torch.eye(3).repeat(bs, 15, 1, 1).view(bs, -1)
    v_tpose = v_shaped + torch.matmul(T["posedirs"], pose_shape.permute(1, 0)).permute(2, 0, 1)

    def se3(R, t):
        pad = torch.zeros((bs, 1, 4))
        pad[:, 0, 3] = 1.0
        return torch.cat([torch.cat([R, t], 2), pad], 1)

    eye = torch.eye(3).repeat(bs, 1, 1)
    G = [se3(Rroot, (eye - Rroot).bmm(j_tpose[:, 0].unsqueeze(2)))]
    for i in range(1, 16):
        R = Rpose[:, i - 1]
        G.append(torch.matmul(G[MANO_PARENT[i]], se3(R, (eye - R).bmm(j_tpose[:, i].unsqueeze(2)))))
    G = torch.stack(G, dim=1)
    joints = [j_tpose[:, 0]]
    one = torch.ones((bs, 1))
    for i in range(1, 16):
        joints.append(G[:, MANO_PARENT[i]].bmm(torch.cat([j_tpose[:, i], one], 1).unsqueeze(2))[:, :3, 0])
    Gv = torch.matmul(T["weights"], G.view(bs, 16, 16)).view(bs, -1, 4, 4)
    v = (Gv[:, :, :3, :3].matmul(v_tpose.unsqueeze(3)) + Gv[:, :, :3, 3:4])[:, :, :, 0]
    j = torch.stack(joints + [v[:, t] for t in MANO_TIPS[side]], dim=1)[:, MANO_NEW_ORDER]
    if center_idx is not None:
        center = j[:, center_idx:center_idx + 1]
        v, j = v - center, j - center
    if scale is not None:
        s = torch.as_tensor(scale, dtype=torch.float32).unsqueeze(1).unsqueeze(2)
        v, j = v * s, j * s
    if trans is not None:
        t = torch.as_tensor(trans, dtype=torch.float32).unsqueeze(1)
        v, j = v + t, j + t
    if new_skel:
        j = j.clone()
        j[:, 5] = (v[:, 63] + v[:, 144]) / 2
        j[:, 9] = (v[:, 271] + v[:, 220]) / 2
        j[:, 13] = (v[:, 148] + v[:, 290]) / 2
        j[:, 17] = (v[:, 770] + v[:, 83]) / 2
    return v, j


def full_regressor(J_regressor):
    """ManoModel.process_J_regressor, lib/models/hand3d/Mano_model.py:309-323: the 16-row rest-joint
    regressor plus one-hot rows for the five finger-tip vertices (745, 317, 444, 556, 673 - the SAME
    indices for both hands, unlike ManoLayer's tips), reordered to the 21-joint convention.  [21,778]."""
    J = torch.as_tensor(np.asarray(J_regressor), dtype=torch.float32)
    tips = torch.zeros((5, J.shape[1]))
    for r, v in enumerate((745, 317, 444, 556, 673)):
        tips[r, v] = 1.0
    return torch.cat([J, tips], 0)[MANO_NEW_ORDER].contiguous()


def regress_joints(reg, verts):
    """joints = full_regressor @ verts (demo.py:217-218, simplified.py:431-434): [21,778] x [B,778,3]."""
    return torch.matmul(torch.as_tensor(reg, dtype=torch.float32), torch.as_tensor(verts, dtype=torch.float32))


# ----------------------------------------------------------------------------
# GCN decoder: the consumer of fuse_feat.  resnet_mid hands
# fuse_feat[:, 0] / fuse_feat[:, 1] over as the per-hand global features
# (intaghand_encoder.py:873-874) and decoder.forward never touches fmaps (the
# img_ex calls are commented out, DualGraph.py:84-85), so the decoder is a pure
# function of fuse_feat, its parameters and the graph assets.
# ----------------------------------------------------------------------------


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def graph_conv_cheby(x, sd, p, L):
    """model_attn/gcn.py:34-69 for K = 2: features [x, Lx] interleaved (k fastest, :62-64) -> Linear."""
    B, V, Fin = x.shape
    x1 = torch.einsum("vu,buf->bvf", L, x)
    return _lin(torch.stack((x, x1), -1).reshape(B, V, Fin * 2), sd, p)


def gcn_resblock(x, sd, p, L):
    """GCN_ResBlock.forward, gcn.py:100-110 (eval: dropout = identity; the norm1/ReLU result of :104 is
    overwritten by :105, which convolves the un-normalised x)."""
    x1 = graph_conv_cheby(x, sd, p + ".fc1", L)
    x1 = F.relu(_ln(x1, sd, p + ".norm2"))
    x1 = graph_conv_cheby(x1, sd, p + ".fc2", L)
    return _ln(x1 + _lin(x, sd, p + ".shortcut"), sd, p + ".norm3")


def graph_layer(x, sd, p, L, n_blocks=4):
    """GraphLayer.forward, gcn.py:131-137."""
    for i in range(n_blocks):
        x = gcn_resblock(x, sd, "%s.GCN_blocks.%d" % (p, i), L)
        if i != n_blocks - 1:
            x = F.relu(x)
    return x


def _mlp_res(x, sd, p):
    """MLP_res_block, self_attn.py:17-33."""
    return x + _lin(F.relu(_lin(_ln(x, sd, p + ".layer_norm"), sd, p + ".fc1")), sd, p + ".fc2")


def _mha(xq, xkv, sd, p, heads=4):
    """softmax(q k^T / sqrt(d)) v with the projections of ``p`` (self_attn.py:60-72, inter_attn.py:84-108)."""
    B, V, f = xq.shape
    d = f // heads
    q = _lin(xq, sd, p + ".w_qs").view(B, V, heads, d).transpose(1, 2)
    k = _lin(xkv, sd, p + ".w_ks").view(B, V, heads, d).transpose(1, 2)
    v = _lin(xkv, sd, p + ".w_vs").view(B, V, heads, d).transpose(1, 2)
    a = F.softmax(torch.matmul(q, k.transpose(-1, -2)) / d ** 0.5, dim=-1)
    return _lin(torch.matmul(a, v).transpose(1, 2).reshape(B, V, f), sd, p + ".fc")


def self_attn(x, sd, p):
    """SelfAttn.forward, self_attn.py:75-84."""
    h = _ln(x, sd, p + ".layer_norm")
    return _mlp_res(x + _mha(h, h, sd, p), sd, p + ".ff")


def inter_attn(Lf, Rf, sd, p):
    """inter_attn.forward, inter_attn.py:113-125 (+ :72-111): two self-attention blocks, then the two
    cross-attention directions share w_qs / w_ks / w_vs / fc."""
    Lf = self_attn(Lf, sd, p + ".L_self_attn_layer")
    Rf = self_attn(Rf, sd, p + ".R_self_attn_layer")
    L2, R2 = _ln(Lf, sd, p + ".layer_norm1"), _ln(Rf, sd, p + ".layer_norm2")
    feat_R2L = _mha(L2, R2, sd, p)            # queries from the left hand, keys / values from the right
    feat_L2R = _mha(R2, L2, sd, p)
    return _mlp_res(Lf + feat_R2L, sd, p + ".ffL"), _mlp_res(Rf + feat_L2R, sd, p + ".ffR")


def hand_position_encoding(assets, side, n):
    """decoder.get_hand_pe, intaghand_decoder.py:169-178: dense colours -> GCN order -> average pool."""
    pe = torch.as_tensor(assets["dense_coor"], dtype=torch.float32) * 2 - 1
    pe = pe[torch.as_tensor(assets["graph_perm_" + side]).long()]
    return pe.view(n, pe.shape[0] // n, 3).mean(1)


def projection_batch(scale, trans2d, v, img_size):
    """lib/utils/utils.py:231-249."""
    return (scale * img_size)[:, None, None] * v[..., :2] + (trans2d * img_size / 2 + img_size / 2)[:, None]


def gcn_decoder_forward(sd, assets, fuse_feat, img_size=384):
    """decoder.forward, intaghand_decoder.py:180-242, fed as HandNET_GCN.forward does
    (intaghand_model.py:30-31 with resnet_mid.forward :873-874): fuse_feat [B,2,1024] ->
    dict(verts3d_{left,right} [B,778,3], verts2d_*, verts3d_gcn_* [B,252,3], verts2d_gcn_*,
    scale_*, trans2d_*, root_*, verts3d_mano_* / verts2d_mano_* [B,778,*])."""
    fuse_feat = torch.as_tensor(fuse_feat)
    dt = fuse_feat.dtype
    out = {}
    feats = {}
    for h, side in enumerate(("left", "right")):
        g = _ln(_lin(fuse_feat[:, h], sd, "gf_layer_%s.0" % side), sd, "gf_layer_%s.1" % side)
        pe = hand_position_encoding(assets, side, 63).to(dt)
        feats[side] = torch.cat((g[:, None, :].expand(-1, 63, -1), pe[None].expand(g.shape[0], -1, -1)), -1)
    Lf, Rf = feats["left"], feats["right"]
    for i in range(3):
        p = "dual_gcn.layers.%d" % i
        pos = sd[p + ".position_embeddings.weight"]
        Lf = graph_layer(Lf + pos, sd, p + ".graph_left", torch.as_tensor(assets["L_left_%d" % i]).to(dt))
        Rf = graph_layer(Rf + pos, sd, p + ".graph_right", torch.as_tensor(assets["L_right_%d" % i]).to(dt))
        Lf, Rf = inter_attn(Lf, Rf, sd, p + ".attn")
        if i != 2:
            Lf, Rf = Lf.repeat_interleave(2, dim=1), Rf.repeat_interleave(2, dim=1)    # graph_upsample(., 2)
    for side, f in (("left", Lf), ("right", Rf)):
        temp = _lin(f.transpose(-1, -2), sd, "avg_head")[..., 0]
        params, root = _lin(temp, sd, "params_head"), _lin(temp, sd, "root_head")
        v252 = _lin(f, sd, "coord_head")
        v778 = _lin(v252.transpose(1, 2), sd, "unsample_layer").transpose(1, 2)
        scale, trans2d = params[:, 0], params[:, 1:]
        rev = torch.as_tensor(assets["graph_perm_reverse_" + side]).long()[:778]
        up = lambda t: t.repeat_interleave(1008 // t.shape[1], dim=1)[:, rev]             # graph_upsample + GCN_to_vert
        v2_252 = projection_batch(scale, trans2d, v252, img_size)
        out.update({"verts3d_" + side: v778, "verts2d_" + side: projection_batch(scale, trans2d, v778, img_size),
                    "verts3d_gcn_" + side: v252, "verts2d_gcn_" + side: v2_252, "scale_" + side: scale,
                    "trans2d_" + side: trans2d, "root_" + side: root, "verts3d_mano_" + side: up(v252),
                    "verts2d_mano_" + side: up(v2_252)})
    return out
