"""Monkey-patch the reference so its own forward runs on the B200 kernels.

After ``patch_reference()`` the reference's modules resolve the hot-path names to
this package: ``HandNET_GCN.forward(img, choose, cloud,
depth, ind, K_new, valid)`` and everything above it are unchanged.

Call it BEFORE the reference's trainer / dataset / demo modules are imported: those bind
``ManoLayer``, ``rodrigues_batch`` and ``_tranpose_and_gather_feat`` by name at import time
(``from lib.models.networks.manolayer import ManoLayer``), so only modules imported afterwards see
the replacements.
"""
import importlib

# names whose replacement has no backward: they stay the reference's in mode="training"
_INFERENCE_ONLY = ("load_decoder", "ManoLayer", "rodrigues_batch")


def patch_reference(mode="inference"):
    """Requires the reference to be importable (``lib`` on sys.path).  Returns the list of patched names.

    mode="inference" (default): every hot-path name is rebound, including the GCN decoder and the MANO
    layer, which are inference-only here (they raise if a gradient is requested).
    mode="training": the differentiable stages (grouping, gathers, SFTLayer, PointNet_Plus - forward and
    backward on this library's kernels) are rebound; ``load_decoder``, ``ManoLayer`` and
    ``rodrigues_batch`` stay the reference's own autograd implementations, so ``loss.backward()`` reaches
    every parameter exactly as in the unpatched model."""
    if mode not in ("inference", "training"):
        raise ValueError("patch_reference: mode must be 'inference' or 'training'")
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
        (rml, "rodrigues_batch", manolayer.rodrigues_batch),
    ):
        if mode == "training" and name in _INFERENCE_ONLY:
            continue
        setattr(mod, name, new)
        patched.append("%s.%s" % (mod.__name__, name))
    return patched
