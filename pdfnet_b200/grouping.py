"""Grouping front end: the reference's ``group_points`` / ``group_points_2`` and the
pointnet2-style aliases named by the north star.

Reference: lib/utils/utils.py:134-187 (kNN-then-radius-mask grouping) and
lib/datasets/interhand.py:147-178 (the only FPS in the repo).  Same names,
argument meaning and output shapes; the work is done by ``pdf_knn_ball``,
``pdf_group_gather`` and ``pdf_fps``.  Order of the K neighbours inside a group
is ascending point index (the reference's ``topk(sorted=False)`` order is
implementation-defined; every consumer is a max-pool over K).
"""
import torch

from . import ops


def group_points(points, opt):
    """lib/utils/utils.py:134-162.  points [B,SAMPLE_NUM,C>=INPUT_FEATURE_NUM] fp32 ->
    (inputs_level1 [B,C,N1,K] with centroid-relative xyz, center [B,3,N1,1])."""
    C = opt.INPUT_FEATURE_NUM
    N1, K = opt.sample_num_level1, opt.knn_K
    if points.shape[1] != opt.SAMPLE_NUM:
        raise RuntimeError("group_points: expected %d points, got %d" % (opt.SAMPLE_NUM, points.shape[1]))
    idx = ops.knn_ball(points, N1, K, opt.ball_radius)
    g, center = ops.group_gather(points[:, :, :C], idx)
    return g.permute(0, 3, 1, 2), center.transpose(1, 2).unsqueeze(3)


def group_points_2(points, sample_num_level1, sample_num_level2, knn_K, ball_radius):
    """lib/utils/utils.py:165-187.  points [B,C,N1] fp32 (xyz = channels 0:3) ->
    (inputs_level2 [B,C,N2,K], center [B,3,N2,1])."""
    if points.shape[2] != sample_num_level1:
        raise RuntimeError("group_points_2: expected %d points, got %d" % (sample_num_level1, points.shape[2]))
    idx = ops.knn_ball(points, sample_num_level2, knn_K, ball_radius, channel_major=True)
    g, center = ops.group_gather(points, idx, channel_major=True)
    return g.permute(0, 3, 1, 2), center.transpose(1, 2).unsqueeze(3)


# ---- pointnet2-style aliases (semantics = the functions above) ----


def farthest_point_sample(xyz, npoint, start_idx=None, generator=None):
    """xyz [B,N,3] -> idx [B,npoint] int64 in SELECTION order, rule of
    farthest_point_sampling_fast (interhand.py:159-175).  ``start_idx`` [B] injects
    the random first sample (:159); default = drawn from ``generator``.  For
    N <= npoint the reference returns arange + random repeats (:153-156)."""
    B, N, _ = xyz.shape
    dev = xyz.device
    if N <= npoint:
        extra = torch.randint(0, N, (B, npoint - N), generator=generator, device="cpu").to(dev)
        return torch.cat([torch.arange(N, device=dev).expand(B, N), extra], 1)
    if start_idx is None:
        start_idx = torch.randint(0, N, (B,), generator=generator, device="cpu").to(dev)
    return ops.fps(xyz, npoint, start_idx).long()


def query_ball_point(radius2, nsample, xyz, new_xyz=None):
    """kNN-then-mask neighbour indices [B,S,nsample] int64 (utils.py:146-151).
    ``radius2`` is the SQUARED radius, as in the reference (opts.py:229-230).
    Centroids are the first S rows of ``xyz`` (S = new_xyz.shape[1]), which is the
    only case the reference has; a ``new_xyz`` that is not that prefix is rejected."""
    S = xyz.shape[1] if new_xyz is None else new_xyz.shape[1]
    if new_xyz is not None and not torch.equal(new_xyz, xyz[:, :S, :3]):
        raise RuntimeError("query_ball_point: new_xyz must be the first S rows of xyz (reference semantics)")
    return ops.knn_ball(xyz, S, nsample, radius2).long()


def index_points(points, idx):
    """points [B,N,C], idx [B,...] -> [B,...,C] row gather (utils.py:153-154)."""
    B = points.shape[0]
    flat = idx.reshape(B, -1)
    out = torch.gather(points, 1, flat[..., None].expand(-1, -1, points.shape[2]).long())
    return out.reshape(*idx.shape, points.shape[2])


def sample_and_group(npoint, radius2, nsample, xyz, points=None):
    """First-``npoint`` centroids + query_ball_point + gather + centroid subtraction on
    xyz: (new_xyz [B,npoint,3], new_points [B,npoint,nsample,3+D])."""
    src = xyz if points is None else torch.cat([xyz, points], 2)
    idx = ops.knn_ball(src, npoint, nsample, radius2)
    g, center = ops.group_gather(src, idx)
    return center, g
