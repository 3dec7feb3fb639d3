"""Import harness for the UNMODIFIED reference (zijinxuxu/PDFNet) on CPU.

TEST INFRASTRUCTURE ONLY.  This module is used in the authoring container
(where /root/reference is mounted) by ``oracle/make_golden.py`` to freeze
golden vectors under ``tests/golden/``.  Nothing in ``pdfnet_b200/``,
``bench.py`` or the ``-m gpu`` tests imports it, and /root/reference does not
exist on the GPU box.

The reference imports a few optional packages at module import time that are
not installed here.  They are not used on the hot path,
so they are replaced by inert ``sys.modules`` stubs:

* ``matplotlib``, ``matplotlib.pyplot``, ``matplotlib.patches``,
  ``mpl_toolkits.mplot3d``            (lib/utils/utils.py:14-15)
* ``tkinter.messagebox``, ``progress.bar``  (lib/datasets/interhand.py:4)
* ``chumpy``: MANO_*.pkl['shapedirs'] is a pickled ``chumpy.reordering.Select``
  (lib/models/networks/manolayer.py:141-144).  The stub class accepts any
  pickled state and exposes ``.r`` the way chumpy's Select does:
  ``a.x.ravel()[idxs].reshape(preferred_shape)``.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PDFNET_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib"))


class _ChStub(object):
    """Generic stand-in for any pickled chumpy object."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:  # pragma: no cover
            self._state = state

    @property
    def r(self):
        d = self.__dict__
        if "x" in d and not isinstance(d["x"], _ChStub):
            return np.asarray(d["x"])
        if "a" in d and "idxs" in d:
            base = d["a"].r if isinstance(d["a"], _ChStub) else np.asarray(d["a"])
            out = np.asarray(base).ravel()[np.asarray(d["idxs"])]
            shape = d.get("preferred_shape", None)
            return out.reshape(shape) if shape is not None else out
        raise AttributeError("chumpy stub cannot evaluate %r" % sorted(d))

    def __array__(self, dtype=None, copy=None):
        a = self.r
        return a.astype(dtype) if dtype is not None else a


class _StubModule(types.ModuleType):
    """Module whose every attribute is a harmless placeholder class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        # __module__ = the stubbed package: Mano_model.to_np (:543) recognises chumpy objects by their type string
        cls = type(name, (_ChStub,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


def _install_stubs():
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits",
        "mpl_toolkits.mplot3d", "tkinter", "tkinter.messagebox", "progress",
        "progress.bar", "chumpy", "chumpy.ch", "chumpy.reordering",
        "chumpy.utils", "chumpy.logic", "chumpy.ch_ops",
    ]
    for n in names:
        try:
            if n.split(".")[0] in ("tkinter",):
                raise ImportError
            __import__(n)
        except Exception:
            if n not in sys.modules:
                sys.modules[n] = _StubModule(n)
    for n in names:
        if "." in n:
            parent, child = n.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _StubModule):
                setattr(sys.modules[parent], child, sys.modules[n])
    # numpy>=1.24 removed these aliases; the dataset code still uses them.
    for alias, typ in (("int", int), ("float", float), ("bool", bool)):
        if alias not in np.__dict__:
            setattr(np, alias, typ)


_loaded = {}


def load_reference():
    """Return a namespace with the reference's hot-path callables."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    u = importlib.import_module("lib.utils.utils")
    mu = importlib.import_module("lib.models.utils")
    enc = importlib.import_module("lib.models.networks.intaghand_encoder")
    ml = importlib.import_module("lib.models.networks.manolayer")
    ns.utils = u
    ns.model_utils = mu
    ns.encoder = enc
    ns.manolayer = ml
    ns.group_points = u.group_points
    ns.group_points_2 = u.group_points_2
    ns.get_normal = u.get_normal
    ns.get_points_coordinate = u.get_points_coordinate
    ns.gather = mu._tranpose_and_gather_feat
    ns.SFTLayer = enc.SFTLayer
    ns.PointNet_Plus = enc.PointNet_Plus
    ns.depth2pcl = enc.depth2pcl
    ns.ManoLayer = ml.ManoLayer
    ns.rodrigues_batch = ml.rodrigues_batch
    ns.mano_dir = os.path.join(REFERENCE_ROOT, "lib", "models", "hand3d", "mano_core")
    _loaded["ns"] = ns
    return ns


def load_fps():
    """The only FPS in the reference is a dataset method
    (lib/datasets/interhand.py:147-178); it does not touch ``self``."""
    load_reference()
    import importlib

    ih = importlib.import_module("lib.datasets.interhand")
    cls = ih.InterHandDataset if hasattr(ih, "InterHandDataset") else None
    if cls is None:
        for v in vars(ih).values():
            if isinstance(v, type) and hasattr(v, "farthest_point_sampling_fast"):
                cls = v
                break
    return lambda pc, n: cls.farthest_point_sampling_fast(None, pc, n)


def load_split_coeff():
    """ManoRender.Split_coeff (lib/models/hand3d/Mano_render.py:145-194) needs
    only ``self.opt.using_pca``, ``self.opt.down_ratio`` and ``self.input_res``;
    pytorch3d-dependent parts of the class are never touched."""
    load_reference()
    for n in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.structures", "pytorch3d.io",
              "pytorch3d.renderer.mesh", "pytorch3d.renderer.mesh.shader",
              "pytorch3d.renderer.mesh.textures", "pytorch3d.transforms",
              "pytorch3d.renderer.cameras", "pytorch3d.renderer.lighting",
              "pytorch3d.renderer.materials", "pytorch3d.renderer.blending", "pytorch3d.ops"):
        if n not in sys.modules:
            sys.modules[n] = _StubModule(n)
    import importlib

    mr = importlib.import_module("lib.models.hand3d.Mano_render")
    return mr.ManoRender.Split_coeff


def default_opt(**over):
    """Hot-path knobs with the reference defaults (lib/opts.py:213-231)."""
    opt = types.SimpleNamespace(
        SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512,
        sample_num_level2=128, ball_radius=0.015, ball_radius2=0.04, PCA_SZ=63,
        default_resolution=384, sample_strategy="random", using_pca=False,
        down_ratio=4, input_res=384,
    )
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def load_decoder(cfg=None, global_feature_dim=1024):
    """The reference's GCN decoder (lib/models/networks/intaghand_decoder.py:244-277, load_decoder)
    built exactly as load_model_intag does: cfg defaults from lib/opts.py:235-239, encoder_info from
    resnet_mid.get_info (global_feature_dim 1024, intaghand_encoder.py:851).  Returns (module, dec_mod)."""
    load_reference()
    import importlib

    dec = importlib.import_module("lib.models.networks.intaghand_decoder")
    if cfg is None:
        cfg = types.SimpleNamespace(IMG_DIMS=[256, 128, 64], GCN_IN_DIM=[512, 256, 128], GCN_OUT_DIM=[256, 128, 64],
                                    graph_k=2, graph_layer_num=4)
    model = dec.load_decoder(cfg, {"global_feature_dim": global_feature_dim, "fmaps_dim": [256, 256, 256, 256]})
    return model, dec
