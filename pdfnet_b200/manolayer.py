"""MANO tail: ``ManoLayer`` (linear blend skinning) and ``Split_coeff``.

Reference: lib/models/networks/manolayer.py:32-48,100-334, lib/models/hand3d/Mano_render.py:145-194
and the 21x778 joint regressor of lib/models/hand3d/Mano_model.py:309-323.  Same constructor and forward signature;
the forward is one ``pdf_mano_lbs`` launch (one CTA per hand) instead of ~100
small bmm/cat launches.
"""
import io
import os
import pickle

import numpy as np
import torch
from torch.nn import Module

from . import ops

NEW_ORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]   # manolayer.py:110-115
TIPS = {"left": (745, 317, 445, 556, 673), "right": (745, 317, 444, 556, 673)}          # :305-308


class _ChumpyStub(object):
    """Unpickle stand-in for chumpy objects (MANO_*.pkl['shapedirs'] is a
    chumpy.reordering.Select; manolayer.py:141-144 evaluates it with ``.r``)."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})

    @property
    def r(self):
        d = self.__dict__
        if "x" in d and not isinstance(d["x"], _ChumpyStub):
            return np.asarray(d["x"])
        base = d["a"].r if isinstance(d["a"], _ChumpyStub) else np.asarray(d["a"])
        out = np.asarray(base).ravel()[np.asarray(d["idxs"])]
        return out.reshape(d["preferred_shape"]) if d.get("preferred_shape") is not None else out


class _ManoUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "chumpy":
            return type(name, (_ChumpyStub,), {})
        return super(_ManoUnpickler, self).find_class(module, name)


def load_mano_data(path):
    """MANO tables as float32 numpy from the reference's .pkl or from an .npz export."""
    if path.endswith(".npz"):
        d = dict(np.load(path))
        return {k: d[k] for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "hands_components",
                                  "hands_mean", "faces") if k in d}
    with open(path, "rb") as fh:
        raw = _ManoUnpickler(io.BytesIO(fh.read()), encoding="latin1").load()
    sd = raw["shapedirs"]
    sd = np.asarray(sd) if isinstance(sd, np.ndarray) else np.asarray(sd.r)
    J = raw["J_regressor"]
    J = np.asarray(J.todense()) if hasattr(J, "todense") else np.asarray(J)
    return dict(v_template=np.asarray(raw["v_template"], np.float32), shapedirs=sd.astype(np.float32),
                posedirs=np.asarray(raw["posedirs"], np.float32), J_regressor=J.astype(np.float32),
                weights=np.asarray(raw["weights"], np.float32),
                hands_components=np.asarray(raw["hands_components"], np.float32),
                hands_mean=np.asarray(raw["hands_mean"], np.float32), faces=np.asarray(raw["f"]))


def kernel_tables(v_template, shapedirs, posedirs, J_regressor, weights, device):
    """Re-lay the MANO constants for pdf_mano_lbs (see include/pdfnet_b200.h): basis index
    outermost, rest-joint regressor pre-applied to template and shape basis in fp64."""
    vt = v_template.double().reshape(778, 3)
    sd = shapedirs.double().reshape(778, 3, 10)
    Jr = J_regressor.double().reshape(16, 778)
    t = dict(
        v_template=vt.reshape(-1).float(),
        shapedirs_t=sd.reshape(2334, 10).t().contiguous().float(),
        posedirs_t=posedirs.float().reshape(2334, 135).t().contiguous(),
        j_template=(Jr @ vt).reshape(-1).float(),
        j_shapedirs=torch.einsum("jv,vck->jck", Jr, sd).reshape(48, 10).contiguous().float(),
        weights_t=weights.float().reshape(778, 16).t().contiguous(),
        # [2334, 145] = [shapedirs | posedirs]: the blend shapes of a batch of hands are one GEMM
        blend_w=torch.cat([sd.reshape(2334, 10).float(), posedirs.float().reshape(2334, 135)], 1).contiguous(),
    )
    return {k: v.contiguous().to(device) for k, v in t.items()}


class ManoLayer(Module):
    """manolayer.py:100-334.  Buffers keep the reference's names and shapes
    (``shapedirs`` [778,3,10], ``posedirs`` [778,3,135], ``J_regressor`` [16,778], ...), so
    code that edits them (e.g. ``fix_shape``, interhand.py:120-123) keeps working: the
    kernel-layout copy is rebuilt whenever a buffer changes."""

    def __init__(self, manoPath, center_idx=9, use_pca=False, new_skel=False):
        super(ManoLayer, self).__init__()
        self.center_idx = center_idx
        self.use_pca = use_pca
        self.new_skel = new_skel
        data = manoPath if isinstance(manoPath, dict) else load_mano_data(manoPath)
        self.new_order = list(NEW_ORDER)
        g = lambda k: torch.tensor(np.asarray(data[k]), dtype=torch.float32)   # copy: never alias caller arrays
        if "hands_components" in data:
            self.register_buffer("hands_components", g("hands_components"))
            self.register_buffer("hands_components_inv", torch.inverse(self.hands_components))
            self.register_buffer("hands_mean", g("hands_mean"), persistent=False)
        self.register_buffer("J_regressor", g("J_regressor"), persistent=False)
        self.register_buffer("weights", g("weights"), persistent=False)
        self.register_buffer("posedirs", g("posedirs"), persistent=False)
        self.register_buffer("v_template", g("v_template"), persistent=False)
        self.register_buffer("shapedirs", g("shapedirs"), persistent=False)
        self.faces = data.get("faces")
        self.parent = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
        self._tables = None
        self._tables_key = None

    def get_faces(self):
        return self.faces

    def train(self, mode=True):
        self.is_train = mode
        return self

    def eval(self):
        return self.train(False)

    def pca2axis(self, pca):
        """:159-162"""
        return pca.mm(self.hands_components[:pca.shape[1]]) + self.hands_mean

    def axis2pca(self, axis):
        """:180-184"""
        return (axis - self.hands_mean).mm(self.hands_components_inv)

    def axis2Rmat(self, axis):
        """:167-171: axis [bs,45] -> [bs,15,3,3]"""
        return rodrigues_batch(axis.reshape(-1, 3)).view(-1, 15, 3, 3)

    def myaxis2Rmat(self, axis):
        """:173-178: axis [bs,3] or [bs,45] -> [bs,-1,3,3]"""
        return rodrigues_batch(axis.reshape(-1, 3)).view(axis.shape[0], -1, 3, 3)

    def pca2Rmat(self, pca):
        """:164-165"""
        return self.axis2Rmat(self.pca2axis(pca))

    def _kernel_tables(self, device):
        bufs = (self.v_template, self.shapedirs, self.posedirs, self.J_regressor, self.weights)
        key = tuple((b.data_ptr(), b._version) for b in bufs) + (str(device),)
        if self._tables is None or key != self._tables_key:
            self._tables = kernel_tables(*[b.detach().cpu() for b in bufs], device)
            self._tables_key = key
        return self._tables

    # ---- differentiable path (losses that back-propagate through MANO, e.g. CtdetLoss: simplified.py:730-736) ----
    # pdf_mano_lbs has no backward.  When a gradient is wanted the same skinning is evaluated with differentiable
    # torch ops on the tensors' own device (library kernels under torch.autograd, not this repo's): joint k maps a
    # rest-pose point x to R_k (x - j_k) + t_k with R_k = R_parent R_local, t_k = R_parent (j_k - j_parent) + t_parent
    # (the 4x4 chain of manolayer.py:287-303 written as rotation / position pairs); t_k is the posed joint.
    def _forward_autograd(self, root_rotation, pose, shape, trans, scale, side):
        bs = root_rotation.shape[0]
        dev = root_rotation.device
        f = lambda t: t.to(dev)
        pose, shape = pose.float(), shape.float()
        if self.use_pca:
            pose = pose.mm(f(self.hands_components)[:pose.shape[1]]) + f(self.hands_mean)             # pca2axis
            R_root = root_rotation.float().reshape(bs, 3, 3)
        else:
            R_root = _rodrigues_autograd(root_rotation.float().reshape(bs, 3))
        R_pose = _rodrigues_autograd(pose.reshape(-1, 3)).view(bs, 15, 3, 3)
        v_shaped = f(self.v_template) + torch.einsum("vck,bk->bvc", f(self.shapedirs), shape.reshape(bs, 10))
        j_rest = torch.einsum("jv,bvc->bjc", f(self.J_regressor), v_shaped)
        pose_feat = (R_pose - torch.eye(3, device=dev)).reshape(bs, 135)
        v_tpose = v_shaped + torch.einsum("vck,bk->bvc", f(self.posedirs), pose_feat)
        Rw, tw = [R_root], [j_rest[:, 0]]
        for i in range(1, 16):
            p = self.parent[i]
            Rw.append(Rw[p].bmm(R_pose[:, i - 1]))
            tw.append(Rw[p].bmm((j_rest[:, i] - j_rest[:, p]).unsqueeze(2))[:, :, 0] + tw[p])
        Rw, tw = torch.stack(Rw, 1), torch.stack(tw, 1)                                              # [bs,16,3,3], [bs,16,3]
        off = tw - torch.matmul(Rw, j_rest.unsqueeze(3))[..., 0]
        w = f(self.weights)
        v = torch.matmul(torch.einsum("vk,bkij->bvij", w, Rw), v_tpose.unsqueeze(3))[..., 0] + torch.einsum("vk,bki->bvi", w, off)
        j = torch.cat([tw, v[:, TIPS[side]]], 1)[:, NEW_ORDER]
        if self.center_idx is not None:                                                              # :310-313
            center = j[:, self.center_idx:self.center_idx + 1]
            v, j = v - center, j - center
        if scale is not None:
            v, j = v * scale.float().reshape(bs, 1, 1), j * scale.float().reshape(bs, 1, 1)
        if trans is not None:
            v, j = v + trans.float().reshape(bs, 1, 3), j + trans.float().reshape(bs, 1, 3)
        if self.new_skel:                                                                            # :320-330
            j = j.clone()
            j[:, 5] = (v[:, 63] + v[:, 144]) / 2
            j[:, 9] = (v[:, 271] + v[:, 220]) / 2
            j[:, 13] = (v[:, 148] + v[:, 290]) / 2
            j[:, 17] = (v[:, 770] + v[:, 83]) / 2
        return v, j

    def forward(self, root_rotation, pose, shape, trans=None, scale=None, side="left"):
        """use_pca=False (:268-272): root_rotation [bs,3] and pose [bs,45] axis-angle.
        use_pca=True (:266-267, the dataset layers of interhand.py:192,220-223): root_rotation is a rotation
        MATRIX [bs,3,3] used as is (:285) and pose holds PCA coefficients [bs,ncomps].
        shape [bs,10], trans [bs,3] or None, scale [bs] or None -> (v [bs,778,3], j [bs,21,3]).
        Host tensors (the reference's dataset / demo code calls the layer on CPU tensors, demo.py:155,
        interhand.py:220,568) are staged to the current CUDA device and the result is returned on the host:
        the skinning runs in pdf_mano_lbs.  A call whose inputs require grad takes the differentiable torch path
        (``_forward_autograd``) on the tensors' own device instead."""
        if side not in TIPS:
            raise ValueError("side must be 'left' or 'right'")
        args = [root_rotation, pose, shape, trans, scale]
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in args):
            if self.use_pca and root_rotation.numel() != root_rotation.shape[0] * 9:
                raise RuntimeError("ManoLayer(use_pca=True): root_rotation must be rotation matrices [bs,3,3]")
            return self._forward_autograd(root_rotation, pose, shape, trans, scale, side)
        host = not root_rotation.is_cuda
        if host:
            if not torch.cuda.is_available():
                raise RuntimeError("pdfnet_b200.ManoLayer needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
            root_rotation, pose, shape, trans, scale = [t.to(dev) if t is not None else None for t in args]
        bs = root_rotation.shape[0]
        with torch.no_grad():
            dev = root_rotation.device
            if self.use_pca:
                if root_rotation.numel() != bs * 9:
                    raise RuntimeError("ManoLayer(use_pca=True): root_rotation must be rotation matrices [bs,3,3] "
                                       "(manolayer.py:258-260,285), got %s" % (tuple(root_rotation.shape),))
                pose = pose.float()
                pose = pose.mm(self.hands_components[:pose.shape[1]].to(dev)) + self.hands_mean.to(dev)   # pca2axis
            elif root_rotation.numel() != bs * 3:
                raise RuntimeError("ManoLayer(use_pca=False): root_rotation must be axis-angle [bs,3] "
                                   "(manolayer.py:270), got %s" % (tuple(root_rotation.shape),))
            v, j = ops.mano_lbs(self._kernel_tables(dev), root_rotation.reshape(bs, -1), pose.reshape(bs, 45),
                                shape.reshape(bs, 10), trans, scale, TIPS[side], self.center_idx, self.new_skel,
                                root_is_matrix=self.use_pca)
        return (v.cpu(), j.cpu()) if host else (v, j)


def _rodrigues_autograd(axis):
    """manolayer.py:32-48 with differentiable torch ops: R = I + sin(t) K + (1 - cos(t)) K^2, t = |a| + 1e-8."""
    angle = torch.norm(axis, p=2, dim=1, keepdim=True) + 1e-8
    a = axis / angle
    z = torch.zeros_like(a[:, 0])
    K = torch.stack((z, -a[:, 2], a[:, 1], a[:, 2], z, -a[:, 0], -a[:, 1], a[:, 0], z), 1).view(-1, 3, 3)
    s, c = torch.sin(angle).unsqueeze(2), torch.cos(angle).unsqueeze(2)
    return torch.eye(3, dtype=axis.dtype, device=axis.device) + s * K + (1 - c) * K.bmm(K)


def rodrigues_batch(axis):
    """manolayer.py:32-48: axis-angle [bs,3] -> rotation matrices [bs,3,3] (pdf_rodrigues); host tensors are
    staged to the GPU and returned on the host, as in ManoLayer.forward.  An input that requires grad takes the
    differentiable torch formulation on its own device."""
    if torch.is_grad_enabled() and axis.requires_grad:
        return _rodrigues_autograd(axis)
    if axis.is_cuda:
        return ops.rodrigues(axis)
    if not torch.cuda.is_available():
        raise RuntimeError("pdfnet_b200.rodrigues_batch needs a CUDA device (no CPU fallback)")
    return ops.rodrigues(axis.cuda()).cpu()


def process_J_regressor(J_regressor):
    """ManoModel.process_J_regressor (lib/models/hand3d/Mano_model.py:309-323): [16,778] rest-joint
    regressor + one-hot finger-tip rows (745, 317, 444, 556, 673 for BOTH hands) in the 21-joint order."""
    J = J_regressor.detach().float()
    tips = torch.zeros((5, J.shape[1]), dtype=J.dtype, device=J.device)
    tips[torch.arange(5), torch.tensor([745, 317, 444, 556, 673])] = 1.0
    return torch.cat([J, tips], 0)[NEW_ORDER].contiguous()


def regress_joints(full_regressor, verts):
    """joints [B,21,3] = full_regressor [21,778] @ verts [B,778,3] (demo.py:217-218, simplified.py:431-434)."""
    return ops.joint_regress(full_regressor, verts)


def Split_coeff(theta, index, K, input_res=384, down_ratio=4):
    """ManoRender.Split_coeff, non-PCA branch (Mano_render.py:160-194): theta [B,122],
    index [B], K [B,3,3] -> (orient_l, pose_l, betas_l, trans_l, orient_r, pose_r, betas_r,
    trans_r).  Unlike the reference it does not modify ``theta`` in place (:165,:171)."""
    left = ops.split_coeff(theta, 0, index, K, input_res, down_ratio)
    right = ops.split_coeff(theta, 61, index, K, input_res, down_ratio)
    return left + right


def mano_tail_pair(theta, ind, K, layer_left, layer_right, input_res=384, down_ratio=4):
    """Same as ``mano_tail`` for theta [B,2,122] / ind [B,2] ((frame, side) layout) with ONE launch per
    stage for both hands: Split_coeff, blend-shape coefficients and skinning.  Returns verts [B,2,778,3],
    joints [B,2,21,3], trans [B,2,3]."""
    if layer_left.use_pca or layer_right.use_pca or layer_left.center_idx != layer_right.center_idx:
        raise RuntimeError("mano_tail_pair: both layers must be axis-angle layers with the same center_idx")
    root, pose, shape, trans = ops.split_coeff_pair(theta, ind, K, input_res, down_ratio)
    dev = theta.device
    v, j = ops.mano_lbs_pair(layer_left._kernel_tables(dev), layer_right._kernel_tables(dev), root, pose, shape, None,
                             None, TIPS["left"], TIPS["right"], layer_left.center_idx, layer_left.new_skel)
    return v, j, trans


def mano_tail(theta_left, theta_right, ind_left, ind_right, K, layer_left, layer_right, input_res=384,
              down_ratio=4):
    """MANO tail as CtdetLoss.origforward runs it (lib/trains/simplified.py:722-736): each hand has
    its own 122-vector (point2mano_left / point2mano_right, [B,122]); the left slice of the left
    vector and the right slice of the right vector are split with that hand's centre index, and the
    MANO layers are called WITHOUT translation (:733-734).  Returns verts [B,2,778,3],
    joints [B,2,21,3] and the metric translations (trans_l, trans_r) [B,3]."""
    ol, pl, bl, tl = ops.split_coeff(theta_left, 0, ind_left, K, input_res, down_ratio)
    orr, pr, br, tr = ops.split_coeff(theta_right, 61, ind_right, K, input_res, down_ratio)
    v_l, j_l = layer_left(ol, pl, bl, side="left")
    v_r, j_r = layer_right(orr, pr, br, side="right")
    return torch.stack((v_l, v_r), 1), torch.stack((j_l, j_r), 1), tl, tr
