"""Monkey-patch the reference so its own forward runs on the B200 kernels.

After ``patch_reference()`` the reference's modules resolve the hot-path names to
this package: ``HandNET_GCN.forward(img, choose, cloud,
depth, ind, K_new, valid)`` and everything above it are unchanged.
"""
import importlib


def patch_reference():
    """Requires the reference to be importable (``lib`` on sys.path).  Returns the list of patched names."""
    from . import decoder, encoder, grouping, manolayer
    patched = []
    ru = importlib.import_module("lib.utils.utils")
    rmu = importlib.import_module("lib.models.utils")
    renc = importlib.import_module("lib.models.networks.intaghand_encoder")
    rml = importlib.import_module("lib.models.networks.manolayer")
    rdec = importlib.import_module("lib.models.networks.intaghand_decoder")
    rmodel = importlib.import_module("lib.models.networks.intaghand_model")
    for mod, name, new in (
        (rdec, "load_decoder", decoder.load_decoder), (rmodel, "load_decoder", decoder.load_decoder),
        (ru, "group_points", grouping.group_points), (ru, "group_points_2", grouping.group_points_2),
        (ru, "get_points_coordinate", encoder.get_points_coordinate),
        (rmu, "_tranpose_and_gather_feat", encoder._tranpose_and_gather_feat),
        (renc, "group_points", grouping.group_points), (renc, "group_points_2", grouping.group_points_2),
        (renc, "_tranpose_and_gather_feat", encoder._tranpose_and_gather_feat),
        (renc, "SFTLayer", encoder.SFTLayer), (renc, "PointNet_Plus", encoder.PointNet_Plus),
        (renc, "depth2pcl", encoder.depth2pcl), (rml, "ManoLayer", manolayer.ManoLayer),
    ):
        setattr(mod, name, new)
        patched.append("%s.%s" % (mod.__name__, name))
    return patched
