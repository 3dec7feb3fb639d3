"""GCN decoder - the consumer of fuse_feat in the reference's live path.

Mirrors ``lib/models/networks/intaghand_decoder.py:74-242`` (class ``decoder``) with the same
state-dict keys (``dual_gcn.layers.{i}.graph_{left,right}.GCN_blocks.{j}.*``,
``dual_gcn.layers.{i}.attn.*``, ``gf_layer_{left,right}.*``, ``unsample_layer``, ``coord_head``,
``avg_head``, ``params_head``, ``root_head``) and the same call:

    result, paramsDict, handDictList, otherInfo = decoder(global_feature_left, global_feature_right, fmaps)

``fmaps`` is accepted and ignored, as in the reference (the ``img_ex`` calls are commented out,
``model_attn/DualGraph.py:84-85``; the ``img_ex_*`` parameters of a reference checkpoint are carried as
inert holders so state dicts round-trip between the two models).  Dense layers run on the GEMM kernels (``precision='fp32'``: FFMA; ``'bf16x3'`` (alias
``'bf16'``): tcgen05 with split-bf16 operands, fp32-accurate - both hold the 1e-4 parity.  Plain bf16
operands were measured and dropped: after ~40 chained layers the scale / translation heads drift by
5-9 % for a 5 % shorter step); each run of glue between two GEMMs is one fused kernel
(``csrc/gcn_decoder.cu``): Chebyshev graph term (sparse L) + bias + shortcut + LayerNorm + ReLU,
residual + LayerNorm, attention, projection.  The kernels are the inference path (``.eval()`` under
``torch.no_grad()``); with a gradient wanted (``.train()``, or grad enabled and parameters / inputs that require
it) the same function runs as differentiable torch ops in the reference's order (``_forward_autograd``).
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops

IMG_SIZE = 384                                             # intaghand_decoder.py:17


class GCN_vert_convert(object):
    """intaghand_decoder.py:32-43"""

    def __init__(self, vertex_num=1, graph_perm_reverse=(0,), graph_perm=(0,)):
        self.graph_perm_reverse = graph_perm_reverse[:vertex_num]
        self.graph_perm = graph_perm

    def vert_to_GCN(self, x):
        return x[:, self.graph_perm]

    def GCN_to_vert(self, x):
        return x[:, self.graph_perm_reverse]


class _GCNResBlock(nn.Module):
    def __init__(self, cin, cout, k):
        super(_GCNResBlock, self).__init__()
        self.norm1 = nn.LayerNorm(cin, eps=1e-6)               # present in checkpoints; its output is discarded (gcn.py:104-105)
        self.fc1 = nn.Linear(cin * k, cout)
        self.norm2 = nn.LayerNorm(cout, eps=1e-6)
        self.fc2 = nn.Linear(cout * k, cout)
        self.shortcut = nn.Linear(cin, cout)
        self.norm3 = nn.LayerNorm(cout, eps=1e-6)


class _GraphLayer(nn.Module):
    def __init__(self, cin, cout, k, n):
        super(_GraphLayer, self).__init__()
        self.GCN_blocks = nn.ModuleList([_GCNResBlock(cin if i == 0 else cout, cout, k) for i in range(n)])


class _MLPRes(nn.Module):
    def __init__(self, f):
        super(_MLPRes, self).__init__()
        self.layer_norm = nn.LayerNorm(f, eps=1e-6)
        self.fc1 = nn.Linear(f, f)
        self.fc2 = nn.Linear(f, f)


class _SelfAttn(nn.Module):
    def __init__(self, f):
        super(_SelfAttn, self).__init__()
        self.w_qs, self.w_ks, self.w_vs = nn.Linear(f, f), nn.Linear(f, f), nn.Linear(f, f)
        self.layer_norm = nn.LayerNorm(f, eps=1e-6)
        self.fc = nn.Linear(f, f)
        self.ff = _MLPRes(f)


class _InterAttn(nn.Module):
    def __init__(self, f):
        super(_InterAttn, self).__init__()
        self.L_self_attn_layer, self.R_self_attn_layer = _SelfAttn(f), _SelfAttn(f)
        self.w_qs, self.w_ks, self.w_vs, self.fc = nn.Linear(f, f), nn.Linear(f, f), nn.Linear(f, f), nn.Linear(f, f)
        self.layer_norm1, self.layer_norm2 = nn.LayerNorm(f, eps=1e-6), nn.LayerNorm(f, eps=1e-6)
        self.ffL, self.ffR = _MLPRes(f), _MLPRes(f)


class _ImgFeatToGrid(nn.Module):
    """model_attn/img_attn.py:38-51 - parameters only (see _ImgEx)."""

    def __init__(self, img_size, img_f_dim, grid_size, grid_f_dim):
        super(_ImgFeatToGrid, self).__init__()
        self.position_embeddings = nn.Embedding(grid_size * grid_size, grid_f_dim)
        patch = img_size // grid_size
        self.proj = nn.Conv2d(img_f_dim, grid_f_dim, kernel_size=patch, stride=patch)
        self.self_attn = _SelfAttn(grid_f_dim)


class _ImgAttn(nn.Module):
    """model_attn/img_attn.py:72-79 - parameters only."""

    def __init__(self, verts_f_dim, img_f_dim):
        super(_ImgAttn, self).__init__()
        self.fc = nn.Linear(img_f_dim, verts_f_dim)
        self.Attn = _SelfAttn(verts_f_dim)


class _ImgEx(nn.Module):
    """model_attn/img_attn.py:97-111.  The reference constructs ``img_ex_left`` / ``img_ex_right`` in every
    DualGraphLayer but never calls them (DualGraph.py:84-85), so their tensors live in every reference checkpoint.
    They are kept here as inert parameter holders: a state dict saved from either model loads into the other
    with no missing or unexpected key."""

    def __init__(self, img_size, img_f_dim, grid_size, grid_f_dim, verts_f_dim):
        super(_ImgEx, self).__init__()
        self.encoder = _ImgFeatToGrid(img_size, img_f_dim, grid_size, grid_f_dim)
        self.attn = _ImgAttn(verts_f_dim, grid_f_dim)


class _DualGraphLayer(nn.Module):
    def __init__(self, V, cin, cout, k, n, img=None):
        super(_DualGraphLayer, self).__init__()
        self.position_embeddings = nn.Embedding(V, cin)
        self.graph_left, self.graph_right = _GraphLayer(cin, cout, k, n), _GraphLayer(cin, cout, k, n)
        if img is not None:                                    # (img_size, img_f_dim, grid_size, grid_f_dim)
            self.img_ex_left, self.img_ex_right = _ImgEx(*img, cout), _ImgEx(*img, cout)
        self.attn = _InterAttn(cout)


class _DualGraph(nn.Module):
    def __init__(self, verts, cins, couts, k, n, img=None):
        super(_DualGraph, self).__init__()
        img = img or [None] * len(verts)
        self.layers = nn.ModuleList([_DualGraphLayer(V, ci, co, k, n, im)
                                     for V, ci, co, im in zip(verts, cins, couts, img)])


def _csr(dense):
    """Dense Laplacian (as GCN_ResBlock registers it, gcn.py:83-87) -> (rowptr, colidx, vals) int32/int32/fp32."""
    dense = np.asarray(dense, dtype=np.float32)
    rows, cols = np.nonzero(dense)
    rowptr = np.zeros(dense.shape[0] + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    return (torch.from_numpy(np.cumsum(rowptr).astype(np.int32)), torch.from_numpy(cols.astype(np.int32)),
            torch.from_numpy(dense[rows, cols].copy()))


class decoder(nn.Module):
    """``assets``: dict with L_{left,right}_{0,1,2} (dense Laplacians, coarse to fine), graph_perm_{side},
    graph_perm_reverse_{side}, dense_coor [778,3], upsample [778,252] - the contents of
    ``gcn_core/*.pkl`` (``oracle/make_golden.export_gcn_assets`` writes them as ``gcn_assets.npz``)."""

    def __init__(self, assets, global_feature_dim=1024, gcn_in_dim=(512, 256, 128), gcn_out_dim=(256, 128, 64),
                 graph_k=2, graph_layer_num=4, vertex_num=778, num_attn_heads=4, precision="fp32",
                 f_in_Dim=(256, 256, 256, 256), f_out_Dim=(256, 128, 64), img_ex_params=True, dropout=0.05):
        super(decoder, self).__init__()
        if precision not in ("fp32", "bf16x3", "bf16"):
            raise ValueError("decoder: precision must be 'fp32' or 'bf16x3' ('bf16' is an alias of the latter)")
        if graph_k != 2:
            raise NotImplementedError("decoder: only Chebyshev order graph_k=2 (the reference default) is built")
        self.precision = precision
        self.heads = num_attn_heads
        self.verts = [int(np.asarray(assets["L_left_%d" % i]).shape[0]) for i in range(3)]
        self.vNum_in, self.vNum_out = self.verts[0], self.verts[2]
        self.vNum_all = len(assets["graph_perm_left"])
        self.vNum_mano = vertex_num
        self.gf_dim = global_feature_dim
        self.gcn_in_dim, self.gcn_out_dim = list(gcn_in_dim), list(gcn_out_dim)
        self.register_buffer("dense_coor", torch.as_tensor(np.asarray(assets["dense_coor"]), dtype=torch.float32))
        self.converter = {}
        for side in ("left", "right"):
            perm = np.asarray(assets["graph_perm_" + side]).astype(np.int64)
            rev = np.asarray(assets["graph_perm_reverse_" + side]).astype(np.int64)
            self.converter[side] = GCN_vert_convert(vertex_num, rev, perm)
            self.register_buffer("_rev_" + side, torch.from_numpy(rev[:vertex_num].copy()), persistent=False)
            pe = (self.dense_coor * 2 - 1)[torch.from_numpy(perm)]                     # get_hand_pe (:169-178)
            self.register_buffer("_pe_" + side, pe.view(self.vNum_in, -1, 3).mean(1), persistent=False)
            for i in range(3):
                for name, t in zip(("rowptr", "colidx", "vals"), _csr(assets["L_%s_%d" % (side, i)])):
                    self.register_buffer("_L_%s_%d_%s" % (side, i, name), t, persistent=False)
                # dense copy for the differentiable (training) path, which multiplies it like the reference does
                self.register_buffer("_Ld_%s_%d" % (side, i), torch.as_tensor(np.asarray(assets["L_%s_%d" % (side, i)]),
                                                                               dtype=torch.float32), persistent=False)
        # inert img_ex_* parameter holders with the reference's shapes (intaghand_decoder.py:121-138:
        # img_size [12,24,48], grid_size 6, img_f_dim = f_in_Dim[:3], grid_f_dim = f_out_Dim)
        img = [(s, fi, 6, fo) for s, fi, fo in zip((12, 24, 48), list(f_in_Dim)[:3], f_out_Dim)] if img_ex_params else None
        self.dual_gcn = _DualGraph(self.verts, gcn_in_dim, gcn_out_dim, graph_k, graph_layer_num, img)
        for side in ("left", "right"):
            setattr(self, "gf_layer_" + side, nn.Sequential(nn.Linear(global_feature_dim, gcn_in_dim[0] - 3),
                                                            nn.LayerNorm(gcn_in_dim[0] - 3, eps=1e-6)))
        self.unsample_layer = nn.Linear(self.vNum_out, vertex_num, bias=False)
        self.coord_head = nn.Linear(gcn_out_dim[-1], 3)
        self.avg_head = nn.Linear(self.vNum_out, 1)
        self.params_head = nn.Linear(gcn_out_dim[-1], 3)
        self.root_head = nn.Linear(gcn_out_dim[-1], 3)
        if "upsample" in assets:
            self.unsample_layer.weight.data.copy_(torch.as_tensor(np.asarray(assets["upsample"])))
        self._cache, self._cache_key = {}, None
        self._side = None
        self.dropout = float(dropout)                          # decoder(dropout=0.05), intaghand_decoder.py:90,275; training only
        # both hands share one graph (true for the shipped gcn_core pickles): the grouped path can then run the two
        # hands' Chebyshev kernels as one launch over one CSR
        self._same_graph = all(np.array_equal(np.asarray(assets["L_left_%d" % i]), np.asarray(assets["L_right_%d" % i]))
                               for i in range(3))
        # Grouped path (both hands per launch) is OFF by default: measured on B200 at 128 frames it is SLOWER than the
        # two-stream path (2.39 vs 2.16 ms per decoder pass) - the two streams' half-footprint GEMMs (two CTAs per SM)
        # and glue kernels do overlap, while a grouped launch pays the L2-bound part of both hands back to back.
        # PDF_DECODER_GROUPED=1 (or .grouped = True) selects it; tests keep both paths pinned to the reference.
        self.grouped = os.environ.get("PDF_DECODER_GROUPED", "0") == "1"

    def get_upsample_weight(self):
        return self.unsample_layer.weight.data

    def set_joint_regressors(self, J_regressor_left, J_regressor_right):
        """Optional epilogue: with the MANO rest-joint regressors [16,778] of both hands set, forward
        also returns the 21 joints ``full_regressor @ verts3d`` (Mano_model.py:309-323, what demo.py:217-218 and
        simplified.py:431-434 compute right after the decoder) as ``otherInfo['joints3d'][side]`` [B,21,3]."""
        from .manolayer import process_J_regressor
        for side, J in (("left", J_regressor_left), ("right", J_regressor_right)):
            reg = process_J_regressor(torch.as_tensor(np.asarray(J), dtype=torch.float32))
            self.register_buffer("_full_regressor_" + side, reg.to(self.dense_coor.device), persistent=False)
        return self

    def get_converter(self):
        return self.converter

    def load_state_dict(self, state_dict, strict=True, **kw):
        state_dict = dict(state_dict)
        have = any("img_ex_" in k for k in self.state_dict())
        if not have:                                                # built with img_ex_params=False
            state_dict = {k: v for k, v in state_dict.items() if "img_ex_" not in k}
        elif not any("img_ex_" in k for k in state_dict):           # a hot-path-only state: keep the inert holders
            state_dict.update({k: v for k, v in self.state_dict().items() if "img_ex_" in k})
        state_dict.setdefault("dense_coor", self.dense_coor)        # an asset, already set by the constructor
        return super(decoder, self).load_state_dict(state_dict, strict=strict, **kw)

    # -- concatenated / re-laid weights, rebuilt when a parameter changes --
    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for n, p in self.named_parameters() if "img_ex_" not in n)
        if key == self._cache_key:
            return self._cache
        c = {}
        w = lambda m: m.weight.detach()
        for li, layer in enumerate(self.dual_gcn.layers):
            for side in ("left", "right"):
                for bi, blk in enumerate(getattr(layer, "graph_" + side).GCN_blocks):
                    # Chebyshev features are interleaved (x_k fastest, gcn.py:62-64): W0 = even, W1 = odd columns
                    c[(li, side, bi, "in")] = torch.cat([w(blk.fc1)[:, 0::2], w(blk.fc1)[:, 1::2], w(blk.shortcut)], 0).contiguous()
                    c[(li, side, bi, "mid")] = torch.cat([w(blk.fc2)[:, 0::2], w(blk.fc2)[:, 1::2]], 0).contiguous()
            a = layer.attn
            for name, m in (("L", a.L_self_attn_layer), ("R", a.R_self_attn_layer), ("X", a)):
                c[(li, name, "qkv_w")] = torch.cat([w(m.w_qs), w(m.w_ks), w(m.w_vs)], 0).contiguous()
                c[(li, name, "qkv_b")] = torch.cat([m.w_qs.bias, m.w_ks.bias, m.w_vs.bias]).detach().contiguous()
        for side in ("left", "right"):
            pos0 = self.dual_gcn.layers[0].position_embeddings.weight.detach()
            row0 = pos0.clone()
            row0[:, -3:] += getattr(self, "_pe_" + side)          # cat([g, pe]) + pos = pad(g) + (pos + [0 | pe])
            c[("row0", side)] = row0.contiguous()
        self._cache, self._cache_key = c, key
        return c

    def _tc(self, M):
        """Rows enough for the tensor-core GEMM path (whose operand images the producers write directly)."""
        return self.precision != "fp32" and M >= 1024

    def _linear(self, x, w, b=None, act=L.ACT_NONE, x_img=None, M=None, out_image=False, tc_min_rows=None):
        """out_image (tensor-core path only): the result feeds nothing but the next GEMM, so it is written only as
        that GEMM's split operand image (returned instead of rows)."""
        M = x.shape[0] if x is not None else M
        tc = self._tc(M) if tc_min_rows is None else (self.precision != "fp32" and M >= tc_min_rows)
        if tc and min(w.shape) >= 16:
            key = ("tc", w.data_ptr(), b.data_ptr() if b is not None else 0)     # weight image packed once
            packed = self._cache.get(key)
            if packed is None:
                packed = self._cache[key] = ops.pack_linear_tc(w, b, split=True) + (w, b)   # keep w / b alive
            return ops.linear_tc(x, None, None, act=act, packed=packed[:5], x_img=x_img, M=M, out_image=out_image,
                                 light=True)
        assert not out_image
        return ops.linear(x, w, b, act=act)

    def _graph_layer(self, x, x_img, li, side, V, M):
        """x: fp32 rows or None, x_img: split tile image or None (tensor-core path) -> fp32 rows."""
        c = self._weights()
        tc = self._tc(M)
        layer = getattr(self.dual_gcn.layers[li], "graph_" + side)
        csr = tuple(getattr(self, "_L_%s_%d_%s" % (side, li, n)) for n in ("rowptr", "colidx", "vals"))
        nb = len(layer.GCN_blocks)
        for bi, blk in enumerate(layer.GCN_blocks):
            co, last = blk.fc1.out_features, bi == nb - 1
            U = self._linear(x, c[(li, side, bi, "in")], x_img=x_img, M=M)                  # [M, 3*co] = [U0 | U1 | shortcut]
            res = ops.graph_cheby_ln(U[:, :co], U[:, co:2 * co], blk.fc1.bias.detach(), csr, V,
                                     (blk.norm2.weight.detach(), blk.norm2.bias.detach()), True, want_rows=not tc,
                                     want_img=tc)
            y, y_img = res if tc else (res, None)
            U2 = self._linear(y, c[(li, side, bi, "mid")], x_img=y_img, M=M)                # [M, 2*co]
            img = tc and not last                              # the last block feeds the attention residual: fp32 rows
            res = ops.graph_cheby_ln(U2[:, :co], U2[:, co:], blk.fc2.bias.detach(), csr, V,
                                     (blk.norm3.weight.detach(), blk.norm3.bias.detach()), not last,
                                     R=U[:, 2 * co:], bias_r=blk.shortcut.bias.detach(), want_rows=not img,
                                     want_img=img)
            x, x_img = res if img else (res, None)
        return x

    @staticmethod
    def _ln(m):
        return (m.weight.detach(), m.bias.detach())

    def _attention_fc(self, problems, n, V, fc):
        """fc(softmax(q k^T / sqrt(d)) v) for 1-2 problems (stacked along the rows).  Tensor-core precisions: the
        attention runs on mma.sync with split-bf16 operands and writes its result straight as the fc GEMM's operand
        image; precision='fp32' (and tiny batches): FFMA attention kernel + FFMA linear."""
        M, f = problems[0][0].shape
        w, b = fc.weight.detach(), fc.bias.detach()
        if self._tc(n * V) and f % 64 == 0:
            _, img = ops.mha_tc([(q, k, v, None) for q, k, v in problems], n, V, self.heads, rows=False, image=True)
            return self._linear(None, w, b, x_img=img, M=len(problems) * M)
        att = torch.empty((len(problems) * M, f), dtype=torch.float32, device=problems[0][0].device)
        for i, (q, k, v) in enumerate(problems):
            ops.mha(q, k, v, n, V, self.heads, out=att[i * M:(i + 1) * M])
        return self._linear(att, w, b)

    def _mlp(self, h, ff, h_img=None, M=None):
        M = h.shape[0] if h is not None else M
        if self._tc(M) and ff.fc1.out_features % 64 == 0:      # hidden activations only exist as fc2's operand image
            f1 = self._linear(h, ff.fc1.weight.detach(), ff.fc1.bias.detach(), L.ACT_RELU, x_img=h_img, M=M, out_image=True)
            return self._linear(None, ff.fc2.weight.detach(), ff.fc2.bias.detach(), x_img=f1, M=M)
        f1 = self._linear(h, ff.fc1.weight.detach(), ff.fc1.bias.detach(), L.ACT_RELU, x_img=h_img, M=M)
        return self._linear(f1, ff.fc2.weight.detach(), ff.fc2.bias.detach())

    def _ln_for_gemm(self, a, b, ln, want_sum):
        """row_combine whose LayerNorm output only feeds a GEMM: fp32 rows on the FFMA path, the split tile
        image on the tensor-core path.  -> (sum rows or None, ln rows or None, ln image or None)"""
        if self._tc(a.shape[0]):
            s_out, _, _, l_img = ops.row_combine(a, b, ln=ln, want_sum=want_sum, ln_rows=False, ln_img=True)
            return s_out, None, l_img
        s_out, l_out = ops.row_combine(a, b, ln=ln, want_sum=want_sum)
        return s_out, l_out, None

    def _self_attn(self, x, li, name, sa, n, V):
        """-> (x2, f2) with SelfAttn(x) = x2 + f2; the add is fused into the caller's next kernel."""
        c = self._weights()
        f, M = x.shape[1], x.shape[0]
        _, h, h_img = self._ln_for_gemm(x, None, self._ln(sa.layer_norm), False)
        qkv = self._linear(h, c[(li, name, "qkv_w")], c[(li, name, "qkv_b")], x_img=h_img, M=M)
        g = self._attention_fc([(qkv[:, :f], qkv[:, f:2 * f], qkv[:, 2 * f:])], n, V, sa.fc)
        x2, h2, h2_img = self._ln_for_gemm(x, g, self._ln(sa.ff.layer_norm), True)
        return x2, self._mlp(h2, sa.ff, h2_img, M)

    def _inter_attn(self, l2, lf, r2, rf, li, n, V):
        """Cross attention after the two SelfAttn blocks (given as deferred sums l2 + lf, r2 + rf)
        -> ((xL, fL), (xR, fR)) with outputs Lf = xL + fL, Rf = xR + fR (adds deferred)."""
        c = self._weights()
        a = self.dual_gcn.layers[li].attn
        f, M = l2.shape[1], l2.shape[0]
        if self._tc(2 * M) and M % 128 == 0:                   # [LN1(Lf) ; LN2(Rf)] written as ONE operand image
            both_img = ops.split_image_empty(2 * M, f, l2.device)
            half = (M // 128) * 3 * (f // 64) * 16384
            Lf = ops.row_combine(l2, lf, ln=self._ln(a.layer_norm1), want_sum=True, ln_rows=False, ln_img=both_img[:half])[0]
            Rf = ops.row_combine(r2, rf, ln=self._ln(a.layer_norm2), want_sum=True, ln_rows=False, ln_img=both_img[half:])[0]
            qkv = self._linear(None, c[(li, "X", "qkv_w")], c[(li, "X", "qkv_b")], x_img=both_img, M=2 * M)
        else:
            both = torch.empty((2 * M, f), dtype=torch.float32, device=l2.device)
            Lf, _ = ops.row_combine(l2, lf, ln=self._ln(a.layer_norm1), want_sum=True, ln_out=both[:M])
            Rf, _ = ops.row_combine(r2, rf, ln=self._ln(a.layer_norm2), want_sum=True, ln_out=both[M:])
            qkv = self._linear(both, c[(li, "X", "qkv_w")], c[(li, "X", "qkv_b")])        # shared projections
        q, k, v = qkv[:, :f], qkv[:, f:2 * f], qkv[:, 2 * f:]
        # R2L (left queries, right keys / values) and L2R in one launch, then the shared fc on the stacked result
        feat = self._attention_fc([(q[:M], k[M:], v[M:]), (q[M:], k[:M], v[:M])], n, V, a.fc)
        # the two feed-forward tails are independent again: the right hand's goes to the side stream (forked here,
        # before the left hand's kernels are queued, and NOT joined - the right hand stays there for the next level)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            x4r, h4r, h4r_img = self._ln_for_gemm(Rf, feat[M:], self._ln(a.ffR.layer_norm), True)
            outR = (x4r, self._mlp(h4r, a.ffR, h4r_img, M))
        for t in (Rf, feat):
            t.record_stream(self._side)
        x4l, h4l, h4l_img = self._ln_for_gemm(Lf, feat[:M], self._ln(a.ffL.layer_norm), True)
        return (x4l, self._mlp(h4l, a.ffL, h4l_img, M)), outR

    def _level_input(self, a, b, rowvec, V_out, up):
        """Input of a DualGraph level (sum + position embedding, x2 up-sampled): it only feeds the first GEMM of
        the level, so on the tensor-core path it is written as that GEMM's operand image.  -> (rows, image)"""
        if self._tc(a.shape[0] * up):
            _, _, s_img, _ = ops.row_combine(a, b, rowvec=rowvec, V_out=V_out, up=up, sum_img=True)
            return None, s_img
        return ops.row_combine(a, b, rowvec=rowvec, V_out=V_out, up=up, want_sum=True)[0], None

    # ---- grouped path: both hands as two row groups of every launch --------------------------------------------
    # The left and the right hand run the same layers with different parameters (graph_left / graph_right,
    # L_/R_self_attn_layer, ffL / ffR).  On two streams their kernels mostly serialise (each one fills the GPU), so
    # every launch's fixed cost is paid twice.  Here the activations of both hands are stacked as two groups of Mp =
    # ceil(B*V / 128) * 128 rows and every kernel takes both groups with their own parameters
    # (pdf_gemm_bf16_grouped, pdf_graph_cheby_ln_grouped, pdf_row_combine_grouped, two problems per pdf_mha_tc):
    # half the launches on ONE stream, same arithmetic per row as the two-stream path.
    def _grouped_ok(self, B):
        return (self.grouped and self.precision != "fp32" and self._same_graph and B * self.verts[0] >= 1024 and
                self.gcn_in_dim[0] in (64, 128, 256, 512) and all(c in (64, 128, 256, 512) for c in self.gcn_out_dim) and
                all(ci == co for ci, co in zip(self.gcn_in_dim[1:], self.gcn_out_dim[:-1])))

    def _gpack(self, key, ws, bs):
        key = ("gpack",) + tuple(key)
        p = self._cache.get(key)
        if p is None:
            p = self._cache[key] = ops.pack_linear_tc_grouped([w.detach() for w in ws],
                                                              [b.detach() if b is not None else None for b in bs])
        return p

    def _gstack(self, key, ts):
        key = ("gstack",) + tuple(key)
        p = self._cache.get(key)
        if p is None:
            p = self._cache[key] = torch.stack([t.detach() for t in ts]).contiguous()
        return p

    def _gln(self, key, ml, mr):
        return (self._gstack(tuple(key) + ("g",), [ml.weight, mr.weight]), self._gstack(tuple(key) + ("b",), [ml.bias, mr.bias]))

    def _glinear(self, key, ml, mr, x_img, Mp, M, act=L.ACT_NONE, out_image=False):
        return ops.linear_tc_grouped(self._gpack(key, [ml.weight, mr.weight], [ml.bias, mr.bias]), x_img, Mp, M, act=act,
                                     out_image=out_image)

    def _graph_layer_grouped(self, x_img, li, V, Mp, M):
        c = self._weights()
        gl, gr = self.dual_gcn.layers[li].graph_left, self.dual_gcn.layers[li].graph_right
        csr = tuple(getattr(self, "_L_left_%d_%s" % (li, n)) for n in ("rowptr", "colidx", "vals"))
        nb = len(gl.GCN_blocks)
        x = None
        for bi, (bl, br) in enumerate(zip(gl.GCN_blocks, gr.GCN_blocks)):
            co, last = bl.fc1.out_features, bi == nb - 1
            U = ops.linear_tc_grouped(self._gpack((li, bi, "in"), [c[(li, "left", bi, "in")], c[(li, "right", bi, "in")]],
                                                  [None, None]), x_img, Mp, M)              # [U0 | U1 | shortcut]
            _, y_img = ops.graph_cheby_ln_grouped(U[:, :co], U[:, co:2 * co], self._gstack((li, bi, "b1"), [bl.fc1.bias, br.fc1.bias]),
                                                  csr, V, Mp, M, self._gln((li, bi, "n2"), bl.norm2, br.norm2), True,
                                                  want_rows=False, want_img=True)
            U2 = ops.linear_tc_grouped(self._gpack((li, bi, "mid"), [c[(li, "left", bi, "mid")], c[(li, "right", bi, "mid")]],
                                                   [None, None]), y_img, Mp, M)
            x, x_img = ops.graph_cheby_ln_grouped(U2[:, :co], U2[:, co:], self._gstack((li, bi, "b2"), [bl.fc2.bias, br.fc2.bias]),
                                                  csr, V, Mp, M, self._gln((li, bi, "n3"), bl.norm3, br.norm3), not last,
                                                  R=U[:, 2 * co:], bias_r2=self._gstack((li, bi, "bs"), [bl.shortcut.bias, br.shortcut.bias]),
                                                  want_rows=last, want_img=not last)
        return x                                               # fp32 rows [2*Mp, co]: the attention residual

    def _features_grouped(self, global_feature_left, global_feature_right):
        """-> fboth [2, B*V, C]: both hands' final vertex features."""
        c = self._weights()
        B = global_feature_left.shape[0]
        cin0, dev = self.gcn_in_dim[0], global_feature_left.device
        pad = lambda m: (m + 127) // 128 * 128
        gp2 = torch.zeros((2, B, cin0), dtype=torch.float32, device=dev)
        for si, (side, gfeat) in enumerate((("left", global_feature_left), ("right", global_feature_right))):
            gf = getattr(self, "gf_layer_" + side)
            g = self._linear(L.f32c(gfeat), gf[0].weight.detach(), gf[0].bias.detach(), tc_min_rows=128)
            ops.row_combine(g, ln=self._ln(gf[1]), ln_out=gp2[si, :, :cin0 - 3])
        V = self.verts[0]
        M, Mp = B * V, pad(B * V)
        row0 = self._gstack(("row0",), [c[("row0", "left")], c[("row0", "right")]])                  # [2, V, cin0]
        # Lf = cat([g repeated over the 63 vertices, pe], -1) + position embedding (:197-198, DualGraph.py:76-80)
        _, x_img, _ = ops.row_combine_grouped(gp2.view(2 * B, cin0), None, Mp, M, B, rowvec=row0, rowvec_gstride=V * cin0,
                                              V_out=V, up=V, sum_img=True)
        for li, V in enumerate(self.verts):
            M, Mp = B * V, pad(B * V)
            a = self.dual_gcn.layers[li].attn
            sl, sr = a.L_self_attn_layer, a.R_self_attn_layer
            x = self._graph_layer_grouped(x_img, li, V, Mp, M)
            f = x.shape[1]
            grp = lambda t, g, c0, c1: t[g * Mp:g * Mp + M, c0:c1]
            # SelfAttn of each hand (self_attn.py:60-84): x2 = x + fc(attention(LN(x))), f2 = MLP(LN(x2))
            _, _, h_img = ops.row_combine_grouped(x, None, Mp, M, Mp, ln=self._gln((li, "sa_ln"), sl.layer_norm, sr.layer_norm),
                                                  ln_img=True)
            qkv = ops.linear_tc_grouped(self._gpack((li, "sa_qkv"), [c[(li, "L", "qkv_w")], c[(li, "R", "qkv_w")]],
                                                    [c[(li, "L", "qkv_b")], c[(li, "R", "qkv_b")]]), h_img, Mp, M)
            _, att_img = ops.mha_tc([(grp(qkv, g, 0, f), grp(qkv, g, f, 2 * f), grp(qkv, g, 2 * f, 3 * f), None) for g in (0, 1)],
                                    B, V, self.heads, rows=False, image=ops.split_image_empty(2 * Mp, f, dev), image_rows=[0, Mp])
            g1 = self._glinear((li, "sa_fc"), sl.fc, sr.fc, att_img, Mp, M)
            x2, _, h2_img = ops.row_combine_grouped(x, g1, Mp, M, Mp, ln=self._gln((li, "sa_ffln"), sl.ff.layer_norm, sr.ff.layer_norm),
                                                    want_sum=True, ln_img=True)
            f1_img = self._glinear((li, "sa_fc1"), sl.ff.fc1, sr.ff.fc1, h2_img, Mp, M, act=L.ACT_RELU, out_image=True)
            f2 = self._glinear((li, "sa_fc2"), sl.ff.fc2, sr.ff.fc2, f1_img, Mp, M)
            # cross attention (inter_attn.py:72-125): shared projections on [LN1(Lf) ; LN2(Rf)], R2L and L2R in one launch
            Xf, _, both_img = ops.row_combine_grouped(x2, f2, Mp, M, Mp, ln=self._gln((li, "x_ln"), a.layer_norm1, a.layer_norm2),
                                                      want_sum=True, ln_img=True)
            qkvx = self._linear(None, c[(li, "X", "qkv_w")], c[(li, "X", "qkv_b")], x_img=both_img, M=2 * Mp)
            _, x_att = ops.mha_tc([(grp(qkvx, 0, 0, f), grp(qkvx, 1, f, 2 * f), grp(qkvx, 1, 2 * f, 3 * f), None),
                                   (grp(qkvx, 1, 0, f), grp(qkvx, 0, f, 2 * f), grp(qkvx, 0, 2 * f, 3 * f), None)],
                                  B, V, self.heads, rows=False, image=ops.split_image_empty(2 * Mp, f, dev), image_rows=[0, Mp])
            feat = self._linear(None, a.fc.weight.detach(), a.fc.bias.detach(), x_img=x_att, M=2 * Mp)
            x4, _, h4_img = ops.row_combine_grouped(Xf, feat, Mp, M, Mp, ln=self._gln((li, "x_ffln"), a.ffL.layer_norm, a.ffR.layer_norm),
                                                    want_sum=True, ln_img=True)
            f3_img = self._glinear((li, "x_fc1"), a.ffL.fc1, a.ffR.fc1, h4_img, Mp, M, act=L.ACT_RELU, out_image=True)
            f4 = self._glinear((li, "x_fc2"), a.ffL.fc2, a.ffR.fc2, f3_img, Mp, M)
            if li != 2:                                        # add + graph_upsample(., 2) + next position embedding
                pos = self.dual_gcn.layers[li + 1].position_embeddings.weight.detach()
                _, x_img, _ = ops.row_combine_grouped(x4, f4, pad(2 * M), 2 * M, Mp, rowvec=pos, V_out=2 * V, up=2, sum_img=True)
            else:
                fb, _, _ = ops.row_combine_grouped(x4, f4, Mp, M, Mp, V_out=V, want_sum=True)
                fboth = fb.view(2, Mp, f)[:, :M]
                return fboth if Mp == M else fboth.contiguous()

    # ---- differentiable path (training / fine-tuning) -------------------------------------------------------------
    # The fused kernels above have no backward.  Whenever a gradient is wanted (module in .train(), or grad enabled and
    # an input / parameter requires it) the SAME function is evaluated with differentiable torch ops on whatever
    # device the tensors live on (ATen / cuBLAS kernels under torch.autograd: library code, not this repo's kernels),
    # in the reference's operation ORDER - including its nn.Dropout calls (gcn.py:107, self_attn.py:69-73,
    # inter_attn.py:26-31,96-103; p = 0.05 as load_decoder passes it, intaghand_decoder.py:275; active only in .train()) - so that a reference training step with the same
    # RNG state sees the same masks.  Same parameters, same state dict, same return structure as the kernel path.
    def _drop(self, t):
        return F.dropout(t, self.dropout, True) if (self.training and self.dropout > 0) else t

    def _ag_cheby(self, x, fc, Ld):
        B, V, fin = x.shape                                    # gcn.py:34-69, K = 2: [x, Lx] interleaved, k fastest
        x1 = torch.einsum("vu,buf->bvf", Ld, x)
        return fc(torch.stack((x, x1), -1).reshape(B, V, fin * 2))

    def _ag_graph_layer(self, x, layer, Ld):
        nb = len(layer.GCN_blocks)
        for bi, blk in enumerate(layer.GCN_blocks):            # GCN_ResBlock.forward, gcn.py:100-110
            x1 = F.relu(blk.norm2(self._ag_cheby(x, blk.fc1, Ld)))
            x1 = self._drop(self._ag_cheby(x1, blk.fc2, Ld))
            x = blk.norm3(x1 + blk.shortcut(x))
            if bi != nb - 1:
                x = F.relu(x)                                  # GraphLayer.forward, gcn.py:131-137
        return x

    def _ag_mha(self, xq, xkv, m):
        B, V, f = xq.shape
        h, d = self.heads, f // self.heads
        q = m.w_qs(xq).view(B, V, h, d).transpose(1, 2)
        k = m.w_ks(xkv).view(B, V, h, d).transpose(1, 2)
        v = m.w_vs(xkv).view(B, V, h, d).transpose(1, 2)
        return F.softmax(torch.matmul(q, k.transpose(-1, -2)) / d ** 0.5, dim=-1), v

    def _ag_mlp(self, x, ff):                                  # MLP_res_block, self_attn.py:17-33
        return x + self._drop(ff.fc2(self._drop(F.relu(ff.fc1(ff.layer_norm(x))))))

    def _ag_self_attn(self, x, sa):                            # SelfAttn.forward, self_attn.py:60-84
        B, V, f = x.shape
        hn = sa.layer_norm(x)
        a, v = self._ag_mha(hn, hn, sa)
        o = torch.matmul(self._drop(a), v).transpose(1, 2).contiguous().view(B, V, f)
        return self._ag_mlp(x + self._drop(sa.fc(o)), sa.ff)

    def _ag_inter_attn(self, Lf, Rf, a):                       # inter_attn.forward, inter_attn.py:72-125
        B, V, f = Lf.shape
        Lf, Rf = self._ag_self_attn(Lf, a.L_self_attn_layer), self._ag_self_attn(Rf, a.R_self_attn_layer)
        L2, R2 = a.layer_norm1(Lf), a.layer_norm2(Rf)
        a_r2l, Rv = self._ag_mha(L2, R2, a)                    # left queries, right keys / values
        a_l2r, Lv = self._ag_mha(R2, L2, a)
        a_r2l, a_l2r = self._drop(a_r2l), self._drop(a_l2r)    # dropout1 in the reference's order (:96-97)
        f_l2r = torch.matmul(a_l2r, Lv).transpose(1, 2).contiguous().view(B, V, f)
        f_r2l = torch.matmul(a_r2l, Rv).transpose(1, 2).contiguous().view(B, V, f)
        f_l2r, f_r2l = self._drop(a.fc(f_l2r)), self._drop(a.fc(f_r2l))
        return self._ag_mlp(Lf + f_r2l, a.ffL), self._ag_mlp(Rf + f_l2r, a.ffR)

    def _forward_autograd(self, gl, gr):
        B = gl.shape[0]
        V0 = self.verts[0]
        feats = {}
        for side, gfeat in (("left", gl), ("right", gr)):
            g = getattr(self, "gf_layer_" + side)(gfeat.float())
            pe = getattr(self, "_pe_" + side).to(g.dtype)
            feats[side] = torch.cat((g[:, None, :].expand(-1, V0, -1), pe[None].expand(B, -1, -1)), -1)   # :197-198
        Lf, Rf = feats["left"], feats["right"]
        for li, layer in enumerate(self.dual_gcn.layers):      # DualGraphLayer.forward, DualGraph.py:72-95
            pos = layer.position_embeddings.weight
            Lf = self._ag_graph_layer(Lf + pos, layer.graph_left, getattr(self, "_Ld_left_%d" % li))
            Rf = self._ag_graph_layer(Rf + pos, layer.graph_right, getattr(self, "_Ld_right_%d" % li))
            Lf, Rf = self._ag_inter_attn(Lf, Rf, layer.attn)
            if li != len(self.dual_gcn.layers) - 1:
                Lf, Rf = Lf.repeat_interleave(2, dim=1), Rf.repeat_interleave(2, dim=1)                # graph_upsample(., 2)
        scale, trans2d, root, verts3d, verts2d = {}, {}, {}, {}, {}
        result = {"verts3d": {}, "verts2d": {}}
        other = {"verts3d_MANO_list": {"left": [], "right": []}, "verts2d_MANO_list": {"left": [], "right": []}}
        proj = lambda sc, tr, v: (sc * IMG_SIZE)[:, None, None] * v[..., :2] + (tr * IMG_SIZE / 2 + IMG_SIZE / 2)[:, None]
        for side, f in (("left", Lf), ("right", Rf)):          # intaghand_decoder.py:213-240
            temp = self.avg_head(f.transpose(-1, -2))[..., 0]
            params, root[side] = self.params_head(temp), self.root_head(temp)
            v252 = self.coord_head(f)
            v778 = self.unsample_layer(v252.transpose(1, 2)).transpose(1, 2)
            scale[side], trans2d[side] = params[:, 0], params[:, 1:]
            rev = getattr(self, "_rev_" + side)
            up = lambda t: t.repeat_interleave(self.vNum_all // t.shape[1], dim=1)[:, rev]              # graph_upsample + GCN_to_vert
            verts3d[side], verts2d[side] = v252, proj(scale[side], trans2d[side], v252)
            result["verts3d"][side], result["verts2d"][side] = v778, proj(scale[side], trans2d[side], v778)
            other["verts3d_MANO_list"][side].append(up(verts3d[side]))
            other["verts2d_MANO_list"][side].append(up(verts2d[side]))
            reg = getattr(self, "_full_regressor_" + side, None)
            if reg is not None:
                other.setdefault("joints3d", {})[side] = torch.matmul(reg, v778)
        return (result, {"scale": scale, "trans2d": trans2d, "root": root},
                [{"verts3d": verts3d, "verts2d": verts2d}], other)

    def _wants_grad(self, *inputs):
        if not torch.is_grad_enabled():
            return False
        return any(t.requires_grad for t in inputs) or any(p.requires_grad for n, p in self.named_parameters()
                                                            if "img_ex_" not in n)

    def forward(self, global_feature_left, global_feature_right, fmaps=None):
        if self.training or self._wants_grad(global_feature_left, global_feature_right):
            return self._forward_autograd(global_feature_left, global_feature_right)
        L.require_cuda(global_feature_left, global_feature_right)
        with torch.no_grad():
            c = self._weights()
            B = global_feature_left.shape[0]
            assert global_feature_left.shape[1] == self.gf_dim and global_feature_right.shape[1] == self.gf_dim
            cin0 = self.gcn_in_dim[0]
            x = {}
            grouped = self._grouped_ok(B)
            if grouped:
                fboth = self._features_grouped(global_feature_left, global_feature_right)
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            if not grouped:
                self._side.wait_stream(cur)                    # fork: the right hand lives on the side stream from its gf layer on
                global_feature_right.record_stream(self._side)
            for side, gfeat in (() if grouped else (("left", global_feature_left), ("right", global_feature_right))):
                with torch.cuda.stream(self._side if side == "right" else cur):
                    gf = getattr(self, "gf_layer_" + side)
                    g = self._linear(L.f32c(gfeat), gf[0].weight.detach(), gf[0].bias.detach(), tc_min_rows=128)
                    gpad = torch.zeros((B, cin0), dtype=torch.float32, device=g.device)
                    ops.row_combine(g, ln=self._ln(gf[1]), ln_out=gpad[:, :cin0 - 3])
                    # Lf = cat([g repeated over the 63 vertices, pe], -1) + position embedding (:197-198, DualGraph.py:76-80)
                    x[side] = self._level_input(gpad, None, c[("row0", side)], self.verts[0], self.verts[0])
            for li, V in (() if grouped else enumerate(self.verts)):
                # the two hands are independent up to the cross attention: the right hand's GraphLayer and
                # SelfAttn run on a second stream (captured as a fork / join inside a CUDA graph); from the
                # cross attention of level 0 on the right hand's tensors are produced there (_inter_attn)
                att = self.dual_gcn.layers[li].attn
                with torch.cuda.stream(self._side):
                    xr = self._graph_layer(x["right"][0], x["right"][1], li, "right", V, B * V)
                    r2, rf = self._self_attn(xr, li, "R", att.R_self_attn_layer, B, V)
                xl = self._graph_layer(x["left"][0], x["left"][1], li, "left", V, B * V)
                l2, lf = self._self_attn(xl, li, "L", att.L_self_attn_layer, B, V)
                cur.wait_stream(self._side)
                for t in (r2, rf):
                    t.record_stream(cur)
                (al, bl), (ar, br) = self._inter_attn(l2, lf, r2, rf, li, B, V)   # (ar, br): on the side stream
                if li != 2:                                    # add + graph_upsample(., 2) + next position embedding
                    pos = self.dual_gcn.layers[li + 1].position_embeddings.weight.detach()
                    x["left"] = self._level_input(al, bl, pos, 2 * V, 2)
                    with torch.cuda.stream(self._side):
                        x["right"] = self._level_input(ar, br, pos, 2 * V, 2)
                else:
                    # both hands' final features stacked: the output heads are shared modules -> one pass for both
                    fboth = torch.empty((2, B * V, self.gcn_out_dim[-1]), dtype=torch.float32, device=al.device)
                    ops.row_combine(al, bl, V_out=V, sum_out=fboth[0])
                    fboth.record_stream(self._side)
                    with torch.cuda.stream(self._side):
                        ops.row_combine(ar, br, V_out=V, sum_out=fboth[1])
                    cur.wait_stream(self._side)                # join: the output heads read both halves
            V, fo = self.verts[2], self.gcn_out_dim[-1]
            scale, trans2d, root, verts3d, verts2d = {}, {}, {}, {}, {}
            result = {"verts3d": {}, "verts2d": {}}
            other = {"verts3d_MANO_list": {"left": [], "right": []}, "verts2d_MANO_list": {"left": [], "right": []}}
            hd = lambda m: (m.weight, m.bias)
            params2, root2, v252_2 = ops.decoder_heads(fboth.view(2 * B * V, fo), 2 * B, V, hd(self.avg_head),
                                                       hd(self.params_head), hd(self.root_head), hd(self.coord_head))
            # 252 -> 778 up-sampling of both hands as ONE GEMM over rows (sample, xyz)
            vt = v252_2.transpose(1, 2).contiguous().view(2 * B * 3, V)
            # with the joint regressors set (a15) the 21 joints ride along as 2 x 21 extra output columns of the same
            # GEMM: joints = full_regressor @ (U @ v252) = (full_regressor @ U) @ v252, the product of the two constant
            # matrices taken once in fp64 - no separate regression launch per hand
            Wup, nv = self.unsample_layer.weight.detach(), self.vNum_mano
            have_reg = all(getattr(self, "_full_regressor_" + sd, None) is not None for sd in ("left", "right"))
            if have_reg:
                key = ("up+joints", Wup.data_ptr(), Wup._version, self._full_regressor_left.data_ptr(),
                       self._full_regressor_right.data_ptr())
                Wcat = self._cache.get(key)
                if Wcat is None:
                    Wcat = self._cache[key] = torch.cat(
                        [Wup] + [(getattr(self, "_full_regressor_" + sd).double() @ Wup.double()).float()
                                 for sd in ("left", "right")], 0).contiguous()
                up_out = self._linear(vt, Wcat, tc_min_rows=256)
            else:
                up_out = self._linear(vt, Wup, tc_min_rows=256)
            v778_2 = up_out[:, :nv].reshape(2 * B, 3, nv).transpose(1, 2).contiguous()
            for si, side in enumerate(("left", "right")):
                params, v252, v778 = params2[si * B:(si + 1) * B], v252_2[si * B:(si + 1) * B], v778_2[si * B:(si + 1) * B]
                root[side] = root2[si * B:(si + 1) * B]
                c2, d2, m3, m2 = ops.decoder_project(v252, v778, params, IMG_SIZE, getattr(self, "_rev_" + side),
                                                     self.vNum_all // V)
                scale[side], trans2d[side] = params[:, 0], params[:, 1:]
                verts3d[side], verts2d[side] = v252, c2
                result["verts3d"][side], result["verts2d"][side] = v778, d2
                other["verts3d_MANO_list"][side].append(m3)
                other["verts2d_MANO_list"][side].append(m2)
                if have_reg:
                    jc = up_out[si * B * 3:(si + 1) * B * 3, nv + 21 * si:nv + 21 * (si + 1)]
                    other.setdefault("joints3d", {})[side] = jc.reshape(B, 3, 21).transpose(1, 2).contiguous()
                elif getattr(self, "_full_regressor_" + side, None) is not None:
                    other.setdefault("joints3d", {})[side] = ops.joint_regress(getattr(self, "_full_regressor_" + side), v778)
            paramsDict = {"scale": scale, "trans2d": trans2d, "root": root}
            handDictList = [{"verts3d": verts3d, "verts2d": verts2d}]
            return result, paramsDict, handDictList, other


def assets_from_graph_dicts(left_graph_dict, right_graph_dict, dense_coor, upsample_weight=None, vertex_num=778):
    """The ``assets`` dict from the objects ``load_decoder`` un-pickles (intaghand_decoder.py:244-258):
    ``gcn_core/graph_{left,right}.pkl`` dicts (scipy-sparse ``coarsen_graphs_L`` fine-to-coarse,
    ``graph_perm``, ``graph_perm_reverse``), ``v_color.pkl`` and ``upsample.pkl``."""
    assets = {"dense_coor": np.asarray(dense_coor, dtype=np.float32)}
    if upsample_weight is not None:
        assets["upsample"] = np.asarray(upsample_weight, dtype=np.float32)
    for side, gd in (("left", left_graph_dict), ("right", right_graph_dict)):
        coarse_first = list(gd["coarsen_graphs_L"])[::-1]                 # decoder.__init__ reverses the list (:96-97)
        for i in range(3):
            Lm = coarse_first[i]
            assets["L_%s_%d" % (side, i)] = np.asarray(Lm.todense() if hasattr(Lm, "todense") else Lm, dtype=np.float32)
        assets["graph_perm_" + side] = np.asarray(gd["graph_perm"], dtype=np.int64)
        assets["graph_perm_reverse_" + side] = np.asarray(gd["graph_perm_reverse"], dtype=np.int64)[:vertex_num]
    return assets


def load_decoder(cfg, encoder_info, precision="bf16x3"):
    """Drop-in for ``intaghand_decoder.load_decoder(cfg, encoder_info)`` (:244-277): reads the reference's own
    ``gcn_core`` pickles through the reference module's path helpers and returns the B200 decoder."""
    import importlib
    import pickle
    rdec = importlib.import_module("lib.models.networks.intaghand_decoder")
    paths = rdec.get_graph_dict_path()
    with open(paths["left"], "rb") as f:
        left = pickle.load(f)
    with open(paths["right"], "rb") as f:
        right = pickle.load(f)
    with open(rdec.get_dense_color_path(), "rb") as f:
        dense = pickle.load(f)
    with open(rdec.get_upsample_path(), "rb") as f:
        up = pickle.load(f)
    return decoder(assets_from_graph_dicts(left, right, dense, up), global_feature_dim=encoder_info["global_feature_dim"],
                   gcn_in_dim=cfg.GCN_IN_DIM, gcn_out_dim=cfg.GCN_OUT_DIM, graph_k=cfg.graph_k,
                   graph_layer_num=cfg.graph_layer_num, precision=precision,
                   f_in_Dim=encoder_info.get("fmaps_dim", (256, 256, 256, 256)),
                   f_out_Dim=getattr(cfg, "IMG_DIMS", (256, 128, 64)))
