"""Depth branch + point/pixel fusion: host-side mirror of the reference modules.

Reference: lib/models/networks/intaghand_encoder.py — ``PointNet_Plus`` :32-159,
``SFTLayer`` :205-219, ``depth2pcl`` :369-491, the fusion tail of
``ResNetSimple.forward`` :805-813 — and lib/models/utils.py:22-26.  Class names,
constructor arguments, parameter names (state-dict keys) and forward signatures
are the reference's, so a reference checkpoint loads unchanged; the bodies call
the sm_100a kernels through the C ABI.  Under ``torch.no_grad()`` (or when nothing
requires grad) the modules run the fused inference kernels; when an autograd graph
is expected (``.train()``, or ``.eval()`` with grad enabled and trainable inputs /
parameters) they dispatch to the differentiable path in ``training.py``.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .profiling import stage

nstates_plus_1 = [64, 64, 128]
nstates_plus_2 = [128, 128, 256]
nstates_plus_3 = [512, 512, 1024, 1024, 512]


def _grad_needed(module, *tensors):
    """True when the caller expects an autograd graph: grad mode is on and a parameter of ``module`` or one of
    ``tensors`` requires grad.  The fused inference kernels have no backward, so such calls are routed to
    the differentiable path (training.py) instead of silently returning detached tensors."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    return module is not None and any(p.requires_grad for p in module.parameters())


def _tranpose_and_gather_feat(feat, ind):
    """lib/models/utils.py:22-26: feat [B,C,H,W], ind [B,n] int64 -> [B,n,C] (no full-map copy).
    Differentiable w.r.t. ``feat`` (the reference gathers from maps that carry gradients:
    intaghand_encoder.py:792,816-817, simplified.py:698-699): the backward is pdf_gather_nchw_bwd."""
    if _grad_needed(None, feat):
        from .training import GatherNCHWFn
        return GatherNCHWFn.apply(feat, ind)
    return ops.gather_nchw(feat, ind)


def get_points_coordinate(depth, instrinsic_inv, device="cuda"):
    """lib/utils/utils.py:251-262: depth [B,H,W,1], inverse intrinsics [B,3,3] -> [B,3,H,W]."""
    B, H, W, _ = depth.shape
    return ops.backproject(depth.reshape(B, H, W).to(device), instrinsic_inv.to(device))


class SFTLayer(nn.Module):
    """intaghand_encoder.py:205-219.  forward((fea [B,Cf,n], cond [B,n,Cc])) -> [B,n,Cf]."""

    def __init__(self, c_fea, c_cond):
        super(SFTLayer, self).__init__()
        self.SFT_scale_conv0 = nn.Conv2d(c_cond, c_cond, 1)
        self.SFT_scale_conv1 = nn.Conv2d(c_cond, c_fea, 1)
        self.SFT_shift_conv0 = nn.Conv2d(c_cond, c_cond, 1)
        self.SFT_shift_conv1 = nn.Conv2d(c_cond, c_fea, 1)

    def _w(self, conv):
        return conv.weight.detach().view(conv.out_channels, conv.in_channels), conv.bias.detach()

    def weights(self, pad_at=None):
        """(ws0,bs0,ws1,bs1,wh0,bh0,wh1,bh1); with ``pad_at`` an all-zero output channel is
        inserted at that index of the second convs (scale=0, shift=0 -> the padded feature
        column passes through unchanged), matching the 16 B-aligned internal row layout."""
        out = []
        for c0, c1 in ((self.SFT_scale_conv0, self.SFT_scale_conv1), (self.SFT_shift_conv0, self.SFT_shift_conv1)):
            w0, b0 = self._w(c0)
            w1, b1 = self._w(c1)
            if pad_at is not None:
                w1 = torch.cat([w1[:pad_at], torch.zeros_like(w1[:1]), w1[pad_at:]], 0).contiguous()
                b1 = torch.cat([b1[:pad_at], torch.zeros_like(b1[:1]), b1[pad_at:]], 0).contiguous()
            out += [w0, b0, w1, b1]
        return tuple(out)

    def apply_rows(self, fea_rows, cond_rows, out=None, weights=None):
        """fea_rows [M,Cf] (row pitch free), cond_rows [M,Cc] -> fea*(scale+1)+shift, [M,Cf].
        ``out`` may alias ``fea_rows`` (in-place modulation)."""
        ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = weights if weights is not None else self.weights()
        hs = ops.linear(cond_rows, ws0, bs0, act=L.ACT_LEAKY01)
        hh = ops.linear(cond_rows, wh0, bh0, act=L.ACT_LEAKY01)
        if out is None:
            out = torch.empty((fea_rows.shape[0], fea_rows.shape[1]), dtype=torch.float32, device=fea_rows.device)
        ops.linear(hs, ws1, bs1, epilogue=L.EPI_SFT_SCALE, f=fea_rows, out=out)
        ops.linear(hh, wh1, bh1, epilogue=L.EPI_ACCUM, out=out)
        return out

    def forward(self, x):
        fea, cond = x[0], x[1]
        B, Cf, n = fea.shape
        fea_rows = L.f32c(fea.transpose(1, 2)).view(B * n, Cf)
        cond_rows = L.f32c(cond).view(B * n, -1)
        if _grad_needed(self, fea, cond):                   # train or eval mode alike (no BatchNorm in this layer)
            from .training import sft_rows
            return sft_rows(self, fea_rows, cond_rows).view(B, n, Cf)
        with torch.no_grad():
            return self.apply_rows(fea_rows, cond_rows).view(B, n, Cf)

    def packed_sft0(self):
        """48 floats in the layout pdf_pyramid_gather expects (only for c_fea=c_cond=3)."""
        parts = []
        for conv in (self.SFT_scale_conv0, self.SFT_scale_conv1, self.SFT_shift_conv0, self.SFT_shift_conv1):
            parts += [conv.weight.detach().reshape(-1), conv.bias.detach().reshape(-1)]
        return torch.cat(parts).float().contiguous()


def _fold_bn(conv, bn):
    """conv -> BN(eval) folded into (W [out,in], b [out]) in fp64, returned fp32."""
    w = conv.weight.detach().double().view(conv.out_channels, -1)
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(conv.out_channels, dtype=torch.float64,
                                                                               device=w.device)
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return (w * s[:, None]).float().contiguous(), ((b - bn.running_mean.detach().double()) * s +
                                                   bn.bias.detach().double()).float().contiguous()


def _mlp(cin, chans, pool):
    layers = []
    for c in chans:
        layers += [nn.Conv2d(cin, c, kernel_size=(1, 1)), nn.BatchNorm2d(c), nn.ReLU(inplace=True)]
        cin = c
    layers.append(nn.MaxPool2d(pool, stride=1))
    return nn.Sequential(*layers)


class PointNet_Plus(nn.Module):
    """intaghand_encoder.py:32-159.  forward(points [B,N,3], emb [l0,l1,l2], choose [B,N])
    -> [B,1,1024].  ``precision``: 'fp32' (FFMA kernels, 1e-4 parity) or 'bf16' (the two
    set-abstraction MLPs on tcgen05 with bf16 operands / fp32 accumulate, 2e-2 parity)."""

    def __init__(self, opt, precision="fp32"):
        super(PointNet_Plus, self).__init__()
        self.num_outputs = opt.PCA_SZ
        self.knn_K = opt.knn_K
        self.ball_radius2 = opt.ball_radius2
        self.sample_num_level1 = opt.sample_num_level1
        self.sample_num_level2 = opt.sample_num_level2
        self.INPUT_FEATURE_NUM = opt.INPUT_FEATURE_NUM
        self.opt = opt
        self.precision = precision
        self.sft0 = SFTLayer(3, 3)
        self.sft1 = SFTLayer(131, 64)
        self.sft2 = SFTLayer(259, 256)
        self.netR_1 = _mlp(self.INPUT_FEATURE_NUM, nstates_plus_1, (1, self.knn_K))
        self.netR_2 = _mlp(3 + nstates_plus_1[2], nstates_plus_2, (1, self.knn_K))
        self.netR_3 = _mlp(3 + nstates_plus_2[2], nstates_plus_3[:3], (self.sample_num_level2, 1))
        self.netR_FC = nn.Sequential(                      # present for state-dict parity; unused (:156)
            nn.Linear(nstates_plus_3[2], nstates_plus_3[3]), nn.BatchNorm1d(nstates_plus_3[3]), nn.ReLU(inplace=True),
            nn.Linear(nstates_plus_3[3], nstates_plus_3[4]), nn.BatchNorm1d(nstates_plus_3[4]), nn.ReLU(inplace=True),
            nn.Linear(nstates_plus_3[4], self.num_outputs))
        self._folded = None
        self._folded_key = None
        self.chunk_clouds = None      # clouds per internal pass (None: 256 in fp32 mode, 2048 in bf16 mode)

    # -- folded / packed parameters, rebuilt when any parameter or buffer changes --
    def _params_key(self):
        ts = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts) + (self.precision,)

    def folded(self):
        key = self._params_key()
        if self._folded is None or key != self._folded_key:
            f = {}
            for name in ("netR_1", "netR_2", "netR_3"):
                net = getattr(self, name)
                f[name] = [_fold_bn(net[i], net[i + 1]) for i in (0, 3, 6)]
            f["sft0"] = self.sft0.packed_sft0()
            # internal rows are [x y z 0 | features]: one zero pad column at index 3 keeps the
            # feature block 16 B aligned (pitch 132 / 260 floats)
            f["sft1"] = self.sft1.weights(pad_at=3)
            f["sft2"] = self.sft2.weights(pad_at=3)
            for name in ("netR_2", "netR_3"):
                w1, b1 = f[name][0]
                f[name + "_w1pad"] = torch.cat([w1[:, :3], torch.zeros_like(w1[:, :1]), w1[:, 3:]], 1).contiguous()
            if self.precision == "bf16":
                dev = f["sft0"].device
                for name in ("netR_1", "netR_2"):
                    (w1, b1), (w2, b2), (w3, b3) = f[name]
                    f[name + "_pack"] = ops.sa_pack_weights(w1, b1, w2, b2, w3, b3).to(dev)
                f["tc"] = self._tensor_core_images(f, dev)
            self._folded, self._folded_key = f, key
        return self._folded

    def _tensor_core_images(self, f, dev):
        """bf16 tile images (pdf_pack_image_host) of the dense layers that run on the streaming
        tcgen05 GEMM.  Channel order inside the images: [features..., xyz, zero pad]."""
        img = lambda w: ops.pack_image(w).to(dev)
        z = lambda n, like: torch.zeros((n,) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)
        tc = {}
        # SFT1: hidden = lrelu([Ws0;Wh0] cond), features 3..130 modulated on tensor cores (xyz in fp32)
        ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = self.sft1.weights()
        tc["sft1_plain"] = (ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1)
        # hidden layer as a split (fp32-accurate) GEMM: its fp32 accumulator also feeds the xyz channels
        tc["sft1_w0"] = ops.pack_image(torch.cat([ws0, wh0], 0), split=True).to(dev)
        tc["sft1_b0"] = torch.cat([bs0, bh0]).contiguous()
        tc["sft1_xyz"] = torch.cat([ws1[:3].reshape(-1), wh1[:3].reshape(-1), bs1[:3], bh1[:3]]).contiguous()
        tc["sft1_w1"] = img(torch.cat([ws1[3:], wh1[3:]], 1))
        tc["sft1_b1s"], tc["sft1_b1h"] = bs1[3:].contiguous(), bh1[3:].contiguous()
        # SFT2: N-tiles = features 0..127 | features 128..255 | [xyz, pad]
        ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = self.sft2.weights()
        tc["sft2_w0"], tc["sft2_b0"] = img(torch.cat([ws0, wh0], 0)), torch.cat([bs0, bh0]).contiguous()
        reorder = lambda t: torch.cat([t[3:], t[:3], z(125, t)], 0)
        tc["sft2_w1"] = img(torch.cat([reorder(ws1), reorder(wh1)], 1))
        tc["sft2_b1s"], tc["sft2_b1h"] = reorder(bs1).contiguous(), reorder(bh1).contiguous()
        # netR_3: layer-1 columns follow the image written by SFT2 ([256 features, xyz])
        (w1, b1), (w2, b2), (w3, b3) = f["netR_3"]
        tc["n3_w1"], tc["n3_b1"] = img(torch.cat([w1[:, 3:], w1[:, :3]], 1)), b1
        tc["n3_w2"], tc["n3_b2"] = img(w2), b2
        tc["n3_w3"], tc["n3_b3"] = img(w3), b3
        return tc

    def gather(self, points, emb, choose, clouds_per_frame=1):
        """Stage 1 of the bf16 inference path on its own: pixel -> point gather of the bf16 channels-last pyramid
        (+ SFT0) -> ``(pts0, cond1_image, cond2_image)`` to be handed to ``forward(..., gathered=)``.  With a
        host-resident pyramid (zero-copy) this is the only part of the path that waits on the PCIe link, so a caller
        can run it for batch i+1 on a second stream while batch i computes (bench.py e2e)."""
        if self.precision != "bf16" or not all(ops._is_bf16_nhwc(e) for e in emb):
            raise RuntimeError("PointNet_Plus.gather: bf16 mode with a bf16 channels-last pyramid only")
        L.require_cuda(points, choose)
        N1, N2 = self.sample_num_level1, self.sample_num_level2
        with torch.no_grad(), stage("pyramid_gather"):
            return ops.pyramid_gather_bf16(points, choose, emb, self.folded()["sft0"], N1, N2,
                                           self.opt.default_resolution, clouds_per_frame)

    def forward(self, points, emb, choose, clouds_per_frame=1, gathered=None):
        if gathered is not None:                       # inference, stage 1 already done (self.gather)
            if self.training or self.precision != "bf16":
                raise RuntimeError("PointNet_Plus.forward(gathered=...): bf16 inference path only")
            L.require_cuda(points, *gathered)
            with torch.no_grad():
                return self._forward_chunk_bf16(points, None, None, clouds_per_frame, 0, gathered=gathered)
        if self.training or _grad_needed(self, points, *emb):
            # autograd through every stage (training.py): train-mode BatchNorm statistics in .train(), folded
            # running statistics in .eval() with grad enabled (fine-tuning with frozen BatchNorm)
            from .training import pointnet_plus_train
            if clouds_per_frame != 1:
                if self.training:
                    raise RuntimeError("PointNet_Plus (training): one cloud per frame per call, as the reference "
                                       "(intaghand_encoder.py:805-806); BatchNorm statistics are per call")
                f = torch.arange(points.shape[0], device=points.device) // clouds_per_frame
                emb = [e[f] for e in emb]
            return pointnet_plus_train(self, points, emb, choose)
        # maps of a bf16 channels-last pyramid may stay in page-locked HOST memory: the gather reads them in place
        # (zero-copy, ops.pyramid_gather_bf16 checks that they are pinned)
        in_place_ok = self.precision == "bf16" and all(ops._is_bf16_nhwc(e) for e in emb)
        L.require_cuda(points, choose, *(() if in_place_ok else emb))
        with torch.no_grad():
            B = points.shape[0]
            chunk = self.chunk_clouds or (256 if self.precision == "fp32" else 2048)
            if B <= chunk:
                return self._forward_chunk(points, emb, choose, clouds_per_frame, 0)
            assert chunk % clouds_per_frame == 0
            outs = [self._forward_chunk(points[s:s + chunk], emb, choose[s:s + chunk], clouds_per_frame,
                                        s // clouds_per_frame) for s in range(0, B, chunk)]
            return torch.cat(outs, 0)

    def _forward_chunk(self, points, emb, choose, cpf, frame0):
        if self.precision == "bf16":
            return self._forward_chunk_bf16(points, emb, choose, cpf, frame0)
        opt, f = self.opt, self.folded()
        N1, N2, K = self.sample_num_level1, self.sample_num_level2, self.knn_K
        B, N = points.shape[0], points.shape[1]
        dev = points.device
        if frame0:
            nf = (B + cpf - 1) // cpf
            emb = [e[frame0:frame0 + nf] for e in emb]
        emb = [e.float() if e.dtype != torch.float32 else e for e in emb]
        # level 0: pixel->point gather at three pyramid levels + SFT0 on xyz (:120-128), fp32
        with stage("pyramid_gather"):
            pts0, cond1, cond2 = ops.pyramid_gather(points, choose, emb, f["sft0"], N1, N2, opt.default_resolution,
                                                    cpf)
        # SA1: neighbour search + grouping + netR_1 + max over K (:123,:132)
        with stage("knn1"):
            idx1 = ops.knn_ball(pts0, N1, K, opt.ball_radius)
        C1p, C2p = 4 + nstates_plus_1[2], 4 + nstates_plus_2[2]              # 132, 260
        bf16 = self.precision == "bf16"
        x1 = (torch.empty if bf16 else torch.zeros)((B, N1, C1p), dtype=torch.float32, device=dev)
        if not bf16:
            x1[:, :, 0:3] = pts0[:, :N1]                      # torch.cat((y, x), 1) (:134)
        with stage("sa1"):
            self._sa(pts0, idx1, "netR_1", f, x1, None)
        # SFT1 in place, then SA2 (:137-143)
        x1r = x1.view(B * N1, C1p)
        with stage("sft1"):
            self.sft1.apply_rows(x1r, cond1.view(B * N1, -1), out=x1r, weights=f["sft1"])
        with stage("knn2"):
            idx2 = ops.knn_ball(x1, N2, K, self.ball_radius2)
        x2 = (torch.empty if bf16 else torch.zeros)((B, N2, C2p), dtype=torch.float32, device=dev)
        if not bf16:
            x2[:, :, 0:3] = x1[:, :N2, 0:3]
        with stage("sa2"):
            self._sa(x1, idx2, "netR_2", f, x2, f["netR_2_w1pad"])
        # SFT2 in place, global MLP + max over the N2 points (:147-154)
        x2r = x2.view(B * N2, C2p)
        with stage("sft2"):
            self.sft2.apply_rows(x2r, cond2.view(B * N2, -1), out=x2r, weights=f["sft2"])
        (_, b1), (w2, b2), (w3, b3) = f["netR_3"]
        with stage("global_mlp"):
            h = ops.linear(x2r, f["netR_3_w1pad"], b1, act=L.ACT_RELU)
            h = ops.linear(h, w2, b2, act=L.ACT_RELU)
            out = ops.linear(h, w3, b3, act=L.ACT_RELU, epilogue=L.EPI_GROUP_MAX, group=N2)
        return out.view(B, 1, nstates_plus_3[2])

    def _forward_chunk_bf16(self, points, emb, choose, cpf, frame0, gathered=None):
        """Tensor-core pipeline: fused tcgen05 set-abstraction kernels + streaming tcgen05 GEMMs over
        bf16 tile images for SFT1 / SFT2 / netR_3.  Everything that decides neighbour indices (SFT0,
        the xyz channels of SFT1, both neighbour searches) stays fp32."""
        opt, f = self.opt, self.folded()
        tc = f["tc"]
        N1, N2, K = self.sample_num_level1, self.sample_num_level2, self.knn_K
        if N2 != 128 or N1 % 128 != 0 or K != 64:
            raise RuntimeError("PointNet_Plus(precision='bf16') is built for K=64, N2=128, N1 %% 128 == 0 "
                               "(got K=%d N1=%d N2=%d); use precision='fp32'" % (K, N1, N2))
        B = points.shape[0]
        dev = points.device
        if frame0:
            nf = (B + cpf - 1) // cpf
            emb = [e[frame0:frame0 + nf] for e in emb]
        u8 = lambda n: torch.empty((n,), dtype=torch.uint8, device=dev)
        RELU, LEAKY, BLK = L.ACT_RELU, L.ACT_LEAKY01, 16384
        # a bf16 channels-last pyramid (autocast RGB neck) is gathered straight into the SFT GEMMs' operand images
        bf16_emb = gathered is not None or (all(ops._is_bf16_nhwc(e) for e in emb) and (B * N1) % 128 == 0
                                            and (B * N2) % 128 == 0)
        with stage("pyramid_gather"):
            if gathered is not None:
                pts0, c1img, c2img = gathered
                if tuple(pts0.shape) != (B, points.shape[1], 3) or c1img.numel() != ops.image_bytes(B * N1, 64) or \
                        c2img.numel() != ops.image_bytes(B * N2, 256):
                    raise RuntimeError("PointNet_Plus.forward(gathered=...): not the output of gather() for these clouds")
            elif bf16_emb:
                pts0, c1img, c2img = ops.pyramid_gather_bf16(points, choose, emb, f["sft0"], N1, N2,
                                                             opt.default_resolution, cpf)
            else:
                emb = [e.float() if e.dtype != torch.float32 else e for e in emb]
                pts0, cond1, cond2 = ops.pyramid_gather(points, choose, emb, f["sft0"], N1, N2,
                                                        opt.default_resolution, cpf)
        with stage("knn1"):
            idx1 = ops.knn_ball(pts0, N1, K, opt.ball_radius)
        x1 = torch.empty((B, N1, 132), dtype=torch.float32, device=dev)
        with stage("sa1"):
            self._sa(pts0, idx1, "netR_1", f, x1, None)
        M1, M2 = B * N1, B * N2
        t1, t2 = M1 // 128, M2 // 128
        x1r = x1.view(M1, 132)
        with stage("sft1"):
            # hidden = lrelu([Ws0;Wh0] cond): split-bf16 GEMM (fp32-accurate); the same epilogue modulates
            # the 3 xyz channels of x1 in fp32 (they drive the level-2 neighbour search)
            h1img = u8(t1 * 2 * BLK)
            if bf16_emb:
                # cond is exactly representable in bf16: its image has ONE k-block, read for both products
                # cond x W_hi + cond x W_lo (the [hi | lo] head of the split weight image)
                ops.gemm_bf16(c1img, t1, 1, tc["sft1_w0"], 1, 3, 2, tc["sft1_b0"], act=LEAKY, out_img=h1img, out_kb=2,
                              rows_valid=M1, tile_desc=[(0, 128, 0)], xyz_w=tc["sft1_xyz"], xyz_x=x1r)
            else:
                c1img = ops.rows_to_image(cond1.view(M1, 64), 0, 64, split=True)
                ops.gemm_bf16(c1img, t1, 3, tc["sft1_w0"], 1, 3, 3, tc["sft1_b0"], act=LEAKY, out_img=h1img, out_kb=2,
                              rows_valid=M1, tile_desc=[(0, 128, 0)], xyz_w=tc["sft1_xyz"], xyz_x=x1r)
            # modulated features leave as bf16 rows (the level-2 gather copies them with cp.async);
            # the xyz channels stay fp32 in x1 (they drive the level-2 neighbour search)
            x1h = torch.empty((M1, 128), dtype=torch.bfloat16, device=dev)
            ops.gemm_bf16(h1img, t1, 2, tc["sft1_w1"], 1, 2, 2, tc["sft1_b1s"], kb_split=1, bias1=tc["sft1_b1h"],
                          F=x1r, out_bf16=x1h, bf16_col_off=4, rows_valid=M1, tile_desc=[(4, 128, 0)])
        with stage("knn2"):
            idx2 = ops.knn_ball(x1, N2, K, self.ball_radius2)
        x2 = torch.empty((B, N2, 260), dtype=torch.float32, device=dev)
        with stage("sa2"):
            (w1, _), (w2, _), (w3, _) = f["netR_2"]
            ops.sa_mlp_max_bf16(x1, idx2, f["netR_2_pack"], w1.shape[1], w1.shape[0], w2.shape[0], w3.shape[0], x2, 4,
                                feat_bf16=x1h.view(B, N1, 128))
        x2r = x2.view(M2, 260)
        four = [(0, 128, 0), (0, 128, 2), (0, 128, 4), (0, 128, 6)]
        with stage("sft2"):
            if not bf16_emb:
                c2img = ops.rows_to_image(cond2.view(M2, 256), 0, 256)
            h2img = u8(t2 * 8 * BLK)
            ops.gemm_bf16(c2img, t2, 4, tc["sft2_w0"], 4, 4, 4, tc["sft2_b0"], act=LEAKY, out_img=h2img, out_kb=8,
                          rows_valid=M2, tile_desc=four)
            x3img = u8(t2 * 6 * BLK)
            ops.gemm_bf16(h2img, t2, 8, tc["sft2_w1"], 3, 8, 8, tc["sft2_b1s"], kb_split=4, bias1=tc["sft2_b1h"],
                          F=x2r, out_img=x3img, out_kb=6, rows_valid=M2,
                          tile_desc=[(4, 128, 0), (132, 128, 2), (0, 4, 4)])
        with stage("global_mlp"):
            g1img, g2img = u8(t2 * 8 * BLK), u8(t2 * 8 * BLK)
            ops.gemm_bf16(x3img, t2, 6, tc["n3_w1"], 4, 5, 5, tc["n3_b1"], act=RELU, out_img=g1img, out_kb=8,
                          rows_valid=M2, tile_desc=four)
            ops.gemm_bf16(g1img, t2, 8, tc["n3_w2"], 4, 8, 8, tc["n3_b2"], act=RELU, out_img=g2img, out_kb=8,
                          rows_valid=M2, tile_desc=four)
            out = torch.empty((B, nstates_plus_3[2]), dtype=torch.float32, device=dev)
            ops.gemm_bf16(tc["n3_w3"], 8, 8, g2img, t2, 8, 8, tc["n3_b3"], out_max=out)
        return out.view(B, 1, nstates_plus_3[2])

    def _sa(self, pts, idx, name, f, xout, w1pad):
        """One set-abstraction stage: max-pooled features to xout[:, :, 4:], centroid xyz to
        xout[:, :, 0:3] (written here by the tensor-core kernel, by the caller in fp32 mode)."""
        B, N1, K = idx.shape
        (w1, b1), (w2, b2), (w3, b3) = f[name]
        if self.precision == "bf16":
            ops.sa_mlp_max_bf16(pts, idx, f[name + "_pack"], w1.shape[1], w1.shape[0], w2.shape[0], w3.shape[0],
                                xout, 4)
            return
        g, _ = ops.group_gather(pts, idx, want_center=False)              # [B,N1,K,C]
        h = ops.linear(g.view(B * N1 * K, -1), w1 if w1pad is None else w1pad, b1, act=L.ACT_RELU)
        h = ops.linear(h, w2, b2, act=L.ACT_RELU)
        ops.linear(h, w3, b3, act=L.ACT_RELU, epilogue=L.EPI_GROUP_MAX, group=K,
                   out=xout.view(B * N1, -1)[:, 4:])


class HandFusion(nn.Module):
    """Fusion tail of ResNetSimple.forward (intaghand_encoder.py:805-813): both hands of
    every frame as ONE batch of 2B clouds through PointNet_Plus, the final
    SFTLayer(1024,1024) with the centre-pixel features, and (optionally, the branch the
    reference disables at :811) the MANO head.  Parameter names follow ResNetSimple:
    ``pointnet_plus.*``, ``sft.*``, ``mano_head.*``."""

    def __init__(self, opt, precision="fp32"):
        super(HandFusion, self).__init__()
        self.opt = opt
        self.pointnet_plus = PointNet_Plus(opt, precision)
        self.sft = SFTLayer(1024, 1024)
        self.mano_head = nn.Sequential(
            nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Linear(256, 122))

    def gather(self, cloud, point_wise_emb, choose):
        """``PointNet_Plus.gather`` for both hands of every frame (cloud [B,2,N,3], choose [B,2,N])."""
        B, H, N, _ = cloud.shape
        return self.pointnet_plus.gather(cloud.reshape(B * H, N, 3), point_wise_emb, choose.reshape(B * H, N),
                                         clouds_per_frame=H)

    def forward(self, cloud, point_wise_emb, choose, center_features, with_mano=False, mano_stream=None, gathered=None):
        """cloud [B,2,N,3], choose [B,2,N], center_features [B,2,1024] ->
        fuse_feat [B,2,1024] (and theta [B,2,122] = (point2mano_left, point2mano_right)).
        ``mano_stream`` (inference): the MANO head - which only reads the un-fused per-hand features - is enqueued
        on that stream, concurrently with the fusion SFT (and whatever the caller runs on the current stream
        afterwards, e.g. the GCN decoder); the caller joins with ``current_stream().wait_stream(mano_stream)``
        before reading theta on another stream."""
        B, H, N, _ = cloud.shape
        if gathered is not None:                       # inference with the gather stage done beforehand (self.gather)
            feat = self.pointnet_plus(cloud.reshape(B * H, N, 3), None, None, clouds_per_frame=H, gathered=gathered)
        elif self.training or _grad_needed(self, cloud, center_features, *point_wise_emb):
            from .training import hand_fusion_train
            return hand_fusion_train(self, cloud, point_wise_emb, choose, center_features, with_mano=with_mano)
        else:
            feat = self.pointnet_plus(cloud.reshape(B * H, N, 3), point_wise_emb, choose.reshape(B * H, N),
                                      clouds_per_frame=H)                   # [2B,1,1024]
        rows = feat.view(B * H, 1024)
        with torch.no_grad():
            theta = None
            if with_mano and mano_stream is not None:      # fork: the MANO head runs beside the fusion SFT
                mano_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(mano_stream):
                    theta = self.mano_head_forward(rows).reshape(B, H, 122)
                rows.record_stream(mano_stream)
            with stage("fusion_sft"):
                cen = L.f32c(center_features).view(B * H, 1024)
                if self.pointnet_plus.precision == "bf16":
                    fused = self._fusion_sft_bf16(rows, cen).view(B, H, 1024)
                else:
                    fused = self.sft.apply_rows(rows, cen).view(B, H, 1024)
            if not with_mano:
                return fused
            if theta is not None:
                return fused, theta
            with stage("mano_head"):
                theta = self.mano_head_forward(rows).reshape(B, H, 122)
            return fused, theta

    def _fusion_sft_bf16(self, rows, cen):
        """SFTLayer(1024,1024) (:809) on the streaming tcgen05 GEMM: hidden = lrelu([Ws0;Wh0] c) as a
        bf16 image, then the dual-accumulator SFT epilogue writes fp32 rows."""
        key = tuple((p.data_ptr(), p._version) for p in self.sft.parameters())
        if getattr(self, "_sft_tc_key", None) != key:
            ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = self.sft.weights()
            dev = rows.device
            self._sft_tc = dict(w0=ops.pack_image(torch.cat([ws0, wh0], 0)).to(dev),
                                b0=torch.cat([bs0, bh0]).contiguous(),
                                w1=ops.pack_image(torch.cat([ws1, wh1], 1)).to(dev), b1s=bs1.contiguous(),
                                b1h=bh1.contiguous())
            self._sft_tc_key = key
        t = self._sft_tc
        M = rows.shape[0]
        mt = (M + 127) // 128
        cimg = ops.rows_to_image(cen, 0, 1024)
        himg = torch.empty((mt * 32 * 16384,), dtype=torch.uint8, device=rows.device)
        ops.gemm_bf16(cimg, mt, 16, t["w0"], 16, 16, 16, t["b0"], act=L.ACT_LEAKY01, out_img=himg, out_kb=32,
                      rows_valid=M, tile_desc=[(0, 128, 2 * i) for i in range(16)])
        out = torch.empty((M, 1024), dtype=torch.float32, device=rows.device)
        ops.gemm_bf16(himg, mt, 32, t["w1"], 8, 32, 32, t["b1s"], kb_split=16, bias1=t["b1h"], F=rows, out_f32=out,
                      rows_valid=M, tile_desc=[(128 * i, 128, 0) for i in range(8)])
        return out

    def mano_head_forward(self, x):
        """mano_head (:630-643) in eval mode: Linear+BN1d folded, ReLU, on the FFMA linear kernel."""
        m = self.mano_head
        key = tuple((t.data_ptr(), t._version) for t in list(m.parameters()) + list(m.buffers()))
        if getattr(self, "_mano_fold_key", None) != key:
            self._mano_fold = _fold_bn_linear(m[0], m[1]) + _fold_bn_linear(m[3], m[4])
            self._mano_fold_key = key
        w1, b1, w2, b2 = self._mano_fold
        if self.pointnet_plus.precision == "bf16":
            return self._mano_head_tc(x, key, (w1, b1, w2, b2, m[6].weight.detach(), m[6].bias.detach()))
        h = ops.linear(x, w1, b1, act=L.ACT_RELU)
        h = ops.linear(h, w2, b2, act=L.ACT_RELU)
        return ops.linear(h, m[6].weight.detach(), m[6].bias.detach())


def _mano_head_tc_impl(self, x, key, w):
    """mano_head on the streaming tcgen05 GEMM with split-bf16 operands ([hi|hi|lo] x [hi|lo|hi]):
    fp32-accurate products (the MANO pose parameters feed rotations), fp32 accumulate."""
    if getattr(self, "_mano_tc_key", None) != key:
        w1, b1, w2, b2, w3, b3 = w
        dev = x.device
        pad = lambda b: torch.cat([b, torch.zeros(128 - b.shape[0] % 128 if b.shape[0] % 128 else 0, device=b.device)])
        self._mano_tc = dict(w1=ops.pack_image(w1, split=True).to(dev), b1=b1.contiguous(),
                             w2=ops.pack_image(w2, split=True).to(dev), b2=b2.contiguous(),
                             w3=ops.pack_image(w3, split=True).to(dev), b3=pad(b3.float()).contiguous())
        self._mano_tc_key = key
    t = self._mano_tc
    M = x.shape[0]
    mt = (M + 127) // 128
    dev = x.device
    h1 = torch.empty((M, 512), dtype=torch.float32, device=dev)
    ops.gemm_bf16(ops.rows_to_image(x, 0, 1024, split=True), mt, 48, t["w1"], 4, 48, 48, t["b1"], act=L.ACT_RELU,
                  out_f32=h1, rows_valid=M, tile_desc=[(128 * i, 128, 0) for i in range(4)])
    h2 = torch.empty((M, 256), dtype=torch.float32, device=dev)
    ops.gemm_bf16(ops.rows_to_image(h1, 0, 512, split=True), mt, 24, t["w2"], 2, 24, 24, t["b2"], act=L.ACT_RELU,
                  out_f32=h2, rows_valid=M, tile_desc=[(128 * i, 128, 0) for i in range(2)])
    theta = torch.empty((M, 124), dtype=torch.float32, device=dev)
    ops.gemm_bf16(ops.rows_to_image(h2, 0, 256, split=True), mt, 12, t["w3"], 1, 12, 12, t["b3"], out_f32=theta,
                  rows_valid=M, tile_desc=[(0, 122, 0)])
    return theta[:, :122]


HandFusion._mano_head_tc = _mano_head_tc_impl


class CenterFeatures(nn.Module):
    """center_feat_up0 / center_feat_up1 + the centre gather of ResNetSimple.forward
    (intaghand_encoder.py:627-628, 790-792), evaluated only at ``ind``: numerically the
    same as convolving the whole map and gathering 2 pixels, at 1/2000 of the work.  Parameter names
    follow ResNetSimple (``center_feat_up0.weight`` [512,256,3,3], ``center_feat_up1.weight`` [1024,512,3,3])."""

    def __init__(self, c_in=256, c_mid=512, c_out=1024, precision="fp32"):
        super(CenterFeatures, self).__init__()
        self.center_feat_up0 = nn.Conv2d(c_in, c_mid, kernel_size=3, stride=1, padding=1, bias=False)
        self.center_feat_up1 = nn.Conv2d(c_mid, c_out, kernel_size=3, stride=1, padding=1, bias=False)
        self.precision = precision
        self._key = None

    def _weights(self, dev):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (self.precision, str(dev))
        if self._key != key:
            # K order of the im2col rows: tap-major, channel-minor
            w0 = self.center_feat_up0.weight.detach().permute(0, 2, 3, 1).reshape(self.center_feat_up0.out_channels, -1)
            w1 = self.center_feat_up1.weight.detach().permute(0, 2, 3, 1).reshape(self.center_feat_up1.out_channels, -1)
            if self.precision == "bf16":
                self._w = (ops.pack_image(w0, split=True).to(dev), ops.pack_image(w1, split=True).to(dev))
            else:
                self._w = (w0.contiguous().to(dev), w1.contiguous().to(dev))
            self._zero = torch.zeros((max(w0.shape[0], w1.shape[0]),), dtype=torch.float32, device=dev)
            self._key = key
        return self._w

    def forward(self, x0, ind):
        """x0 [B,C,H,W] fp32 (the 1/4-resolution feature map), ind [B,2] -> center_features [B,2,c_out]."""
        L.require_cuda(x0, ind)
        if _grad_needed(self, x0):
            # training: the reference's own two full-map convolutions (they belong to the RGB network, which
            # stays stock PyTorch) followed by the differentiable centre gather (intaghand_encoder.py:790-792)
            return _tranpose_and_gather_feat(self.center_feat_up1(self.center_feat_up0(x0)), ind)
        with torch.no_grad():
            B = x0.shape[0]
            w0, w1 = self._weights(x0.device)
            cm, co = self.center_feat_up0.out_channels, self.center_feat_up1.out_channels
            rows = ops.center_im2col(x0, ind)                                  # [B*18, 9*C]
            if self.precision != "bf16":
                mid = ops.linear(rows, w0)                                     # conv0 at the 9 positions
                return ops.linear(mid.view(B * 2, 9 * cm), w1).view(B, 2, co)  # conv1 at the centre
            # split-bf16 tensor-core GEMMs (fp32-accurate products, fp32 accumulate)
            M0, K0 = rows.shape
            kb0 = (K0 + 63) // 64 * 3
            mid = torch.empty((M0, cm), dtype=torch.float32, device=x0.device)
            ops.gemm_bf16(ops.rows_to_image(rows, 0, K0, split=True), (M0 + 127) // 128, kb0, w0, cm // 128, kb0, kb0,
                          self._zero, out_f32=mid, rows_valid=M0, tile_desc=[(128 * i, 128, 0) for i in range(cm // 128)])
            mid2 = mid.view(B * 2, 9 * cm)
            M1, K1 = mid2.shape
            kb1 = (K1 + 63) // 64 * 3
            out = torch.empty((M1, co), dtype=torch.float32, device=x0.device)
            ops.gemm_bf16(ops.rows_to_image(mid2, 0, K1, split=True), (M1 + 127) // 128, kb1, w1, co // 128, kb1, kb1,
                          self._zero, out_f32=out, rows_valid=M1, tile_desc=[(128 * i, 128, 0) for i in range(co // 128)])
            return out.view(B, 2, co)


def _fold_bn_linear(fc, bn):
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    w = (fc.weight.detach().double() * s[:, None]).float().contiguous()
    b = ((fc.bias.detach().double() - bn.running_mean.detach().double()) * s + bn.bias.detach().double())
    return w, b.float().contiguous()


def depth2pcl_batched(depth, mask, K_img, valid, subset_keys=None, perm=None, generator=None, min_pixels=10,
                      seed=None):
    """Device-side depth2pcl for a whole batch.  depth [B,H,W] or [B,1,H,W],
    mask [B,2,H,W] at depth resolution (fp32, or uint8 / bool), K_img [B,3,3], valid [B,2] -> choose int64
    [B,2,1024] (row 0 = left), cloud fp32 [B,2,1024,3].  Randomness (the two
    np.random.shuffle calls, intaghand_encoder.py:421,427) is injected: ``subset_keys``
    int32 [B,2,H*W] and ``perm`` int32 [B,2,1024]; with ``seed`` they are generated INSIDE the kernel
    (counter-based, no key tensors at all); otherwise both are drawn from ``generator`` on the device."""
    if depth.dim() == 4:
        depth = depth[:, 0]
    B, H, W = depth.shape
    dev = depth.device
    if mask.shape[-2:] != (H, W):
        raise RuntimeError("depth2pcl_batched: mask must already be at depth resolution")
    Kinv = torch.linalg.inv(K_img.float())
    if seed is None:
        if subset_keys is None:
            subset_keys = torch.argsort(torch.rand((B, 2, H * W), device=dev, generator=generator), dim=2).int()
        if perm is None:
            perm = torch.argsort(torch.rand((B, 2, 1024), device=dev, generator=generator), dim=2).int()
    choose, cloud, _ = ops.depth2pcl(depth, mask, Kinv, valid, subset_keys, perm, 1024, min_pixels, seed=seed)
    return choose, cloud


def depth2pcl(depth_256, mask, K_img, valid, subset_keys=None, perm=None):
    """Drop-in for depth2pcl (intaghand_encoder.py:369-491): batch-1 tensors in, numpy
    (choose int64 [2,1024], cloud float32 [2,1024,3]) out, as the reference returns them.
    The inverse intrinsics are computed on the host with numpy in float32 exactly as the
    reference does (utils.py:269)."""
    import cv2  # the reference resizes the mask with cv2 (:376-377)
    d = depth_256.detach()
    d = d.reshape(d.shape[-2], d.shape[-1]).float()
    H, W = d.shape
    m = (mask.detach().float().cpu().numpy() > 0.5).astype(np.uint8)
    mr = np.stack([cv2.resize(m[0, 0], (W, H)), cv2.resize(m[0, 1], (W, H))]).astype(np.float32)
    Kinv = np.linalg.inv(K_img.detach().cpu().numpy().astype(np.float32).reshape(3, 3))[:3, :3]
    dev = d.device if d.is_cuda else torch.device("cuda")
    rs = np.random
    if subset_keys is None:
        subset_keys = np.stack([rs.permutation(H * W) for _ in range(2)])
    if perm is None:
        perm = np.stack([rs.permutation(1024) for _ in range(2)])
    choose, cloud, _ = ops.depth2pcl(
        d.to(dev)[None], torch.from_numpy(mr).to(dev)[None], torch.from_numpy(Kinv.astype(np.float32)).to(dev)[None],
        valid.detach().float().reshape(1, 2).to(dev), torch.as_tensor(np.asarray(subset_keys)).to(dev)[None],
        torch.as_tensor(np.asarray(perm)).to(dev)[None])
    return choose[0].cpu().numpy(), cloud[0].cpu().numpy()
