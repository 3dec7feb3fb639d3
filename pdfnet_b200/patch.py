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

# names whose replacement runs its BACKWARD (and its train-mode forward) on differentiable torch ops rather than on
# this library's kernels (decoder._forward_autograd, ManoLayer._forward_autograd): mode="training" leaves them the
# reference's own modules, mode="training-all" rebinds them too
_INFERENCE_ONLY = ("load_decoder", "ManoLayer", "rodrigues_batch")


def patch_reference(mode="inference"):
    """Requires the reference to be importable (``lib`` on sys.path).  Returns the list of patched names.

    mode="inference" (default) / "training-all": every hot-path name is rebound, including the GCN decoder and the
    MANO layer.  Those two run their kernels under ``torch.no_grad()`` / in ``.eval()`` and switch to differentiable
    torch ops whenever a gradient is wanted (``decoder._forward_autograd`` - same dropout calls in the same order as
    the reference, tests/test_decoder_autograd.py - and ``ManoLayer._forward_autograd``).
    mode="training": only the stages whose forward AND backward run on this library's kernels (grouping, gathers,
    SFTLayer, PointNet_Plus) are rebound; ``load_decoder``, ``ManoLayer`` and ``rodrigues_batch`` stay the
    reference's own modules.  Either way ``loss.backward()`` reaches every parameter as in the unpatched model."""
    if mode not in ("inference", "training", "training-all"):
        raise ValueError("patch_reference: mode must be 'inference', 'training' or 'training-all'")
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
