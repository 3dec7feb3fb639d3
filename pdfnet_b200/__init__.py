"""pdfnet_b200 — B200-native (sm_100a) depth-branch / fusion / MANO hot path of PDFNet.

Drop-in surface (same names as the reference):
    group_points, group_points_2, _tranpose_and_gather_feat, SFTLayer, PointNet_Plus,
    depth2pcl, get_points_coordinate, ManoLayer, Split_coeff
plus the pointnet2-style aliases farthest_point_sample, query_ball_point, index_points,
sample_and_group, the fused two-hand tail HandFusion, patch_reference(), CapturedStep / PipelinedStep
(CUDA-graph replay of a step / of two overlapped stages of consecutive batches) and the training-mode entry points (PointNet_Plus / HandFusion in .train()).
Compute modules need the CUDA library (pdfnet_b200/lib/libpdfnet_b200.so); there is no
CPU fallback.  ``pdfnet_b200.synth`` and ``pdfnet_b200.parallel`` are pure host code.
"""
__version__ = "0.1.0"

_COMPUTE = {
    "group_points": "grouping", "group_points_2": "grouping", "farthest_point_sample": "grouping",
    "query_ball_point": "grouping", "index_points": "grouping", "sample_and_group": "grouping",
    "_tranpose_and_gather_feat": "encoder", "SFTLayer": "encoder", "PointNet_Plus": "encoder",
    "HandFusion": "encoder", "CenterFeatures": "encoder", "depth2pcl": "encoder", "depth2pcl_batched": "encoder",
    "get_points_coordinate": "encoder", "ManoLayer": "manolayer", "Split_coeff": "manolayer",
    "mano_tail": "manolayer", "mano_tail_pair": "manolayer", "patch_reference": "patch",
    "rodrigues_batch": "manolayer", "process_J_regressor": "manolayer", "regress_joints": "manolayer",
    "decoder": "decoder", "load_decoder": "decoder", "CapturedStep": "graph", "PipelinedStep": "graph", "pointnet_plus_train": "training", "hand_fusion_train": "training",
    "allreduce_gradients": "training", "BucketedAllReduce": "training",
}


def __getattr__(name):
    if name in _COMPUTE:
        import importlib
        return getattr(importlib.import_module("." + _COMPUTE[name], __name__), name)
    raise AttributeError(name)
