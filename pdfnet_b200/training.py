"""Training-mode forward + backward of the fusion path (BASELINE cfg5), fp32.

The reference trains this path through torch.autograd over PointNet_Plus.forward
(intaghand_encoder.py:118-159) with the DDP gradient all-reduce of main.py:44-73.  Here every
differentiable stage is a ``torch.autograd.Function`` whose forward AND backward run on the
hand-written kernels of csrc/train_f32.cu / linear_f32.cu / gather.cu / knn_fps.cu; autograd only
chains them (and does the concatenations), so the module drops into the reference's training loop,
its optimizer and its DistributedDataParallel wrapper unchanged.

Semantics kept from the reference: train-mode BatchNorm2d over all rows of ONE call (the reference
calls PointNet_Plus once per hand, :805-806, so statistics are per hand and the running buffers are
updated twice per step); neighbour indices are constants of the graph; max-pool gradient goes to the
first maximum; gradients flow into the pyramid maps l0/l1/l2, the SFT and MLP parameters, but not
into the input cloud.
"""
import torch
from torch.autograd import Function

from . import _lib as L
from . import ops


def _c(t):
    return t if t.stride(-1) == 1 and t.dtype == torch.float32 else t.float().contiguous()


def _use_tc(M, N, K):
    """Dense layers large enough for the tensor-core path; tiny ones (SFT0's 3x3 convs, the K=3 first
    layer) stay on the FFMA kernels."""
    return TENSOR_CORE_GEMMS and M >= 1024 and min(N, K) >= 16


TENSOR_CORE_GEMMS = True     # tcgen05 GEMMs for forward, dX and dW (False: FFMA kernels everywhere)

# GEMM operand precision per mode: "fp32" -> split-bf16 (fp32-accurate), "bf16" -> plain bf16 operands
# with fp32 accumulation for the forward, data-gradient AND weight-gradient GEMMs (BASELINE cfg5 "training
# step bf16": one third of the operand bytes).  The layers that decide neighbour indices (SFT0 and the xyz
# channels through SFT1) are small / FFMA or split in both modes.
FP32, BF16 = "fp32", "bf16"


class LinearFn(Function):
    """y = act(x @ w.T + b);  x [M,K], w [N,K]."""

    @staticmethod
    def forward(ctx, x, w, b, act, bias_before_bn=False, precision=FP32):
        x, w = _c(x), _c(w)
        tc = _use_tc(x.shape[0], w.shape[0], w.shape[1])
        split = precision != BF16
        # K <= 4 (netR_1[0]: xyz -> 64) is pure streaming: bandwidth-shaped kernels for the layer and both gradients
        smallk = (not tc and w.shape[1] <= 4 and w.shape[0] % 4 == 0 and act == L.ACT_NONE and x.shape[0] >= 4096
                  and w.is_contiguous())
        x_img = None
        if tc:
            # the operand image of x is built here and kept: the weight gradient reads the same image again
            # (as an MN-major operand, pdf_gemm_tn_bf16) instead of a transposed copy
            x_img = ops.rows_to_image(x, 0, x.shape[1], split=1 if split else 0)
            y = ops.linear_tc(None, w, b, act=act, split=split, x_img=x_img, M=x.shape[0])
        elif smallk:
            y = ops.linear_smallk(0, x, w, b)
        else:
            y = ops.linear(x, w, b, act=act)
        ctx.act, ctx.tc, ctx.bias_before_bn, ctx.split, ctx.smallk = act, tc, bias_before_bn, split, smallk
        ctx.save_for_backward(x, w, y if act != L.ACT_NONE else None, x_img if tc else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y, x_img = ctx.saved_tensors
        dy = _c(dy)
        if ctx.act != L.ACT_NONE:
            dy = ops.act_bwd(dy, y, ctx.act)
        dx = dw = None
        M, N, K = dy.shape[0], w.shape[0], w.shape[1]
        # fp32 mode: split images (fp32-accurate gradient GEMMs); bf16 mode: ONE plain bf16 image of dY feeds both
        # gradient GEMMs, against the plain image of x kept from the forward pass (what "bf16 training" means in
        # every framework: bf16 operands, fp32 accumulation, for forward, dX and dW alike)
        sp = ctx.split
        dy_img = ops.rows_to_image(dy, 0, N, split=1 if sp else 0) if ctx.tc else None   # shared by the dX and dW GEMMs
        if ctx.needs_input_grad[0]:
            if ctx.smallk and dy.stride(0) % 4 == 0 and dy.data_ptr() % 16 == 0:
                dx = ops.linear_smallk(1, dy, w)
            else:
                wt = w.t().contiguous()
                dx = ops.linear_tc(None, wt, split=sp, x_img=dy_img, M=M) if ctx.tc else ops.linear(dy, wt)
        if ctx.needs_input_grad[1]:
            if ctx.smallk:
                dw = ops.linear_smallk(2, dy, x)
            else:
                dw = ops.linear_tn_mn(dy_img, N, x_img, K, M, split=sp) if ctx.tc else ops.linear_tn(dy, x)
        db = None
        if ctx.needs_input_grad[2]:
            # a bias in front of train-mode BatchNorm has an identically zero gradient (BatchNorm removes
            # the mean); autograd would return the rounding residue of sum(dy), we return the exact zeros
            db = torch.zeros((w.shape[0],), dtype=torch.float32, device=w.device) if ctx.bias_before_bn \
                else ops.col_sum(dy)
        return dx, dw, db, None, None, None


class BatchNormActFn(Function):
    """Train-mode BatchNorm over rows + optional ReLU; updates running_mean / running_var in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu):
        x = _c(x)
        mean, rstd = ops.bn_batch_stats(x, eps, momentum, running_mean, running_var)
        y = ops.bn_act_fwd(x, mean, rstd, gamma, beta, relu)
        ctx.relu = relu
        # y is not kept for the backward pass: the ReLU mask is recomputed from x (bit-identical expression)
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        dx, dgamma, dbeta = ops.bn_act_bwd(_c(dy), None, x, mean, rstd, gamma, ctx.relu, beta=beta)
        return dx, dgamma, dbeta, None, None, None, None, None


class LinearBNReLUFn(Function):
    """Conv1x1 -> BatchNorm2d(train) -> ReLU as ONE autograd node on the tensor-core path, so that the
    gradient between the BatchNorm and the convolution never exists as fp32 rows: the BatchNorm backward
    writes it as the split-bf16 tile image which both gradient GEMMs read (dX through pdf_gemm_bf16, dW
    through pdf_gemm_tn_bf16 together with the image of x kept from the forward pass)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, momentum, eps, precision, group=0,
                image_only=False):
        """group > 0: the nn.MaxPool2d over ``group`` consecutive rows that ends the stack is part of the node;
        its gradient (dOut at the argmax row, zero elsewhere) is then never materialised.
        image_only: the output is consumed by another node of this kind only, so it is written ONLY as that
        node's operand image, returned as the second (non-differentiable) output; the fp32 tensor returned
        first is then a one-element placeholder expanded to the right shape.  Without image_only the second
        output is an empty tensor."""
        w = _c(w)
        split = precision != BF16
        M = x.shape[0]
        x_img = getattr(x, "_pdf_split_img" if split else "_pdf_plain_img", None)
        if x_img is None:
            x = _c(x)
            x_img = ops.rows_to_image(x, 0, x.shape[1], split=1 if split else 0)
        pre = ops.linear_tc(None, w, b, split=split, x_img=x_img, M=M)
        mean, rstd = ops.bn_batch_stats(pre, eps, momentum, running_mean, running_var)
        arg = None
        if image_only and not group:
            _, y_img = ops.bn_act_fwd(pre, mean, rstd, gamma, beta, True, rows=False, image=True, plain=not split)
            y = torch.empty((1, 1), dtype=torch.float32, device=pre.device).expand(M, w.shape[0])
        else:
            y_img = None
            y = ops.bn_act_fwd(pre, mean, rstd, gamma, beta, True)
            if group:
                y, arg = ops.group_max(y, group, want_arg=True)
        ctx.group, ctx.split = group, split
        ctx.save_for_backward(x if x.stride(-1) == 1 else None, w, pre, mean, rstd, gamma, beta, x_img, arg)
        if y_img is None:
            y_img = torch.empty((0,), dtype=torch.uint8, device=pre.device)
        ctx.mark_non_differentiable(y_img)
        return y, y_img

    @staticmethod
    def backward(ctx, dy, _dimg=None):
        x, w, pre, mean, rstd, gamma, beta, x_img, arg = ctx.saved_tensors
        M, N, K = pre.shape[0], w.shape[0], w.shape[1]
        dy = _c(dy)
        if dy.stride(0) % 4 or dy.data_ptr() % 16:
            dy = dy.contiguous()
        sp = ctx.split                                  # bf16 mode: plain bf16 gradient image, plain-bf16 gradient GEMMs
        if ctx.group:
            dpre_img, dgamma, dbeta = ops.bn_maxpool_bwd(dy, arg, ctx.group, pre, mean, rstd, gamma, beta, True, plain=not sp)
        else:
            dpre_img, dgamma, dbeta = ops.bn_act_bwd(dy, None, pre, mean, rstd, gamma, True, beta=beta, image=True,
                                                     plain=not sp)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_tc(None, w.t().contiguous(), split=sp, x_img=dpre_img, M=M)
        if ctx.needs_input_grad[1]:
            dw = ops.linear_tn_mn(dpre_img, N, x_img, K, M, split=sp)
        if ctx.needs_input_grad[2]:                     # bias in front of BatchNorm: identically zero gradient
            db = torch.zeros((N,), dtype=torch.float32, device=w.device)
        return dx, dw, db, dgamma, dbeta, None, None, None, None, None, None, None


class GroupMaxFn(Function):
    """nn.MaxPool2d over groups of G consecutive rows."""

    @staticmethod
    def forward(ctx, y, G):
        y = _c(y)
        ctx.G = G
        ctx.save_for_backward(y)
        return ops.group_max(y, G)

    @staticmethod
    def backward(ctx, dout):
        (y,) = ctx.saved_tensors
        return ops.group_max_bwd(y, _c(dout), ctx.G), None


class GroupGatherFn(Function):
    """group_points / group_points_2 gather (utils.py:153-158, :181-186) with constant indices."""

    @staticmethod
    def forward(ctx, pts, idx):
        g, _ = ops.group_gather(pts, idx, want_center=False)
        ctx.n_points = pts.shape[1]
        ctx.save_for_backward(idx)
        return g

    @staticmethod
    def backward(ctx, dg):
        (idx,) = ctx.saved_tensors
        return ops.group_scatter_add(dg, idx, ctx.n_points), None


class SFTModulateFn(Function):
    """fea * (scale + 1) + shift (intaghand_encoder.py:219)."""

    @staticmethod
    def forward(ctx, fea, scale, shift):
        fea, scale, shift = _c(fea), _c(scale), _c(shift)
        ctx.save_for_backward(fea, scale)
        return ops.sft_modulate(fea, scale, shift)

    @staticmethod
    def backward(ctx, dout):
        fea, scale = ctx.saved_tensors
        dout = _c(dout)
        dfea, dscale = ops.sft_modulate_bwd(dout, fea, scale)
        return dfea, dscale, dout


class GatherNCHWFn(Function):
    """_tranpose_and_gather_feat (models/utils.py:22-26): feat [B,C,H,W], ind [B,n] -> [B,n,C]."""

    @staticmethod
    def forward(ctx, feat, ind):
        ctx.shape = feat.shape
        ctx.save_for_backward(ind)
        return ops.gather_nchw(feat, ind)

    @staticmethod
    def backward(ctx, dout):
        (ind,) = ctx.saved_tensors
        return ops.gather_nchw_bwd(dout, ind, ctx.shape), None


def _conv_w(conv):
    return conv.weight.view(conv.out_channels, conv.in_channels)


def sft_rows(sft, fea_rows, cond_rows, precision=FP32):
    """SFTLayer on rows: fea [M,Cf], cond [M,Cc] -> [M,Cf] (differentiable)."""
    lin = lambda x, conv, act: LinearFn.apply(x, _conv_w(conv), conv.bias, act, False, precision)
    scale = lin(lin(cond_rows, sft.SFT_scale_conv0, L.ACT_LEAKY01), sft.SFT_scale_conv1, L.ACT_NONE)
    shift = lin(lin(cond_rows, sft.SFT_shift_conv0, L.ACT_LEAKY01), sft.SFT_shift_conv1, L.ACT_NONE)
    return SFTModulateFn.apply(fea_rows, scale, shift)


def mlp_max_rows(net, rows, group, precision=FP32):
    """(Conv1x1 -> BatchNorm2d(train) -> ReLU) x3 -> max over ``group`` consecutive rows."""
    h = rows
    M = rows.shape[0]
    if not net[1].training:
        # eval-mode BatchNorm with an autograd graph (fine-tuning with frozen statistics): the running statistics
        # are folded into the convolution with differentiable parameter-sized torch ops; the layers themselves
        # (forward, dX, dW) stay on LinearFn's kernels
        for i in (0, 3, 6):
            conv, bn = net[i], net[i + 1]
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            h = LinearFn.apply(h, _conv_w(conv) * s[:, None], (conv.bias - bn.running_mean) * s + bn.bias,
                               L.ACT_RELU, False, precision)
        return GroupMaxFn.apply(h, group)
    fused = [_use_tc(M, _conv_w(net[i]).shape[0], _conv_w(net[i]).shape[1]) and net[i].out_channels % 64 == 0
             for i in (0, 3, 6)]
    for li, i in enumerate((0, 3, 6)):
        conv, bn = net[i], net[i + 1]
        momentum = bn.momentum if bn.momentum is not None else 0.1
        w = _conv_w(conv)
        if fused[li]:
            pool = group if (i == 6 and group <= 256 and M % group == 0) else 0
            # an intermediate layer whose only consumer is the next fused node hands over its operand image
            image_only = li < 2 and fused[li + 1]
            h, h_img = LinearBNReLUFn.apply(h, w, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                            momentum, bn.eps, precision, pool, image_only)
            if h_img.numel():                           # h's fp32 storage is a placeholder: the values live in the image
                setattr(h, "_pdf_plain_img" if precision == BF16 else "_pdf_split_img", h_img)
            pooled = pool > 0
        else:
            pooled = False
            h = LinearFn.apply(h, w, conv.bias, L.ACT_NONE, True, precision)
            h = BatchNormActFn.apply(h, bn.weight, bn.bias, bn.running_mean, bn.running_var, momentum, bn.eps, True)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return h if pooled else GroupMaxFn.apply(h, group)


def pyramid_indices(choose, R):
    """intaghand_encoder.py:125-126"""
    c2 = (choose // R // 2) * (R // 2) + choose % R // 2
    c4 = (choose // R // 4) * (R // 4) + choose % R // 4
    return c2, c4


def pointnet_plus_train(net, points, emb, choose):
    """PointNet_Plus.forward in training mode: points [B,N,3], emb [l0,l1,l2], choose [B,N]
    -> [B,1,1024] with an autograd graph through every stage."""
    L.require_cuda(points, choose, *emb)
    opt = net.opt
    prec = BF16 if getattr(net, "precision", FP32) == BF16 else FP32
    N1, N2, K = net.sample_num_level1, net.sample_num_level2, net.knn_K
    B, N, _ = points.shape
    choose = choose.long()
    points = _c(points)
    e0 = GatherNCHWFn.apply(_c(emb[0]), choose)                                         # [B,N,3]
    pts0 = sft_rows(net.sft0, points.reshape(B * N, 3), e0.reshape(B * N, 3)).view(B, N, 3)
    idx1 = ops.knn_ball(pts0.detach(), N1, K, opt.ball_radius)
    g1 = GroupGatherFn.apply(pts0, idx1)                                                # [B,N1,K,3]
    c2, c4 = pyramid_indices(choose, opt.default_resolution)
    e1 = GatherNCHWFn.apply(_c(emb[1]), c2[:, :N1].contiguous())
    e2 = GatherNCHWFn.apply(_c(emb[2]), c4[:, :N2].contiguous())
    f1 = mlp_max_rows(net.netR_1, g1.view(B * N1 * K, -1), K, prec)                     # [B*N1,128]
    x1 = torch.cat((pts0[:, :N1].reshape(B * N1, 3), f1), 1)                            # cat((y, x), 1) (:134)
    x1 = sft_rows(net.sft1, x1, e1.reshape(B * N1, -1))          # fp32-accurate: its xyz channels pick level-2 neighbours
    C1 = x1.shape[1]
    x1b = x1.view(B, N1, C1)
    idx2 = ops.knn_ball(x1b.detach(), N2, K, net.ball_radius2)
    g2 = GroupGatherFn.apply(x1b, idx2)                                                 # [B,N2,K,131]
    f2 = mlp_max_rows(net.netR_2, g2.view(B * N2 * K, C1), K, prec)                     # [B*N2,256]
    x2 = torch.cat((x1b[:, :N2, :3].reshape(B * N2, 3), f2), 1)
    x2 = sft_rows(net.sft2, x2, e2.reshape(B * N2, -1), prec)
    out = mlp_max_rows(net.netR_3, x2, N2, prec)                                        # [B,1024]
    return out.view(B, 1, -1)


def mano_head_rows(head, rows):
    """mano_head (intaghand_encoder.py:630-643: Linear-BN1d-ReLU, Linear-BN1d-ReLU, Linear) with an autograd
    graph: batch statistics in .train(), folded running statistics in .eval().  Always fp32-accurate GEMMs
    (the pose parameters feed rotations)."""
    h = rows
    for i in (0, 3):
        fc, bn = head[i], head[i + 1]
        if bn.training:
            momentum = bn.momentum if bn.momentum is not None else 0.1
            h = LinearFn.apply(h, fc.weight, fc.bias, L.ACT_NONE, True, FP32)
            h = BatchNormActFn.apply(h, bn.weight, bn.bias, bn.running_mean, bn.running_var, momentum, bn.eps, True)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        else:
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            h = LinearFn.apply(h, fc.weight * s[:, None], (fc.bias - bn.running_mean) * s + bn.bias, L.ACT_RELU,
                               False, FP32)
    return LinearFn.apply(h, head[6].weight, head[6].bias, L.ACT_NONE, False, FP32)


def hand_fusion_train(fusion, cloud, point_wise_emb, choose, center_features, with_mano=False):
    """ResNetSimple.forward :805-813 with an autograd graph: one PointNet_Plus call per hand (per-hand
    BatchNorm statistics, as in the reference), concatenation, final SFTLayer(1024,1024) and, with
    ``with_mano``, the mano_head branch on each hand's un-fused feature (:812-813)."""
    B = cloud.shape[0]
    left = pointnet_plus_train(fusion.pointnet_plus, cloud[:, 0], point_wise_emb, choose[:, 0])
    right = pointnet_plus_train(fusion.pointnet_plus, cloud[:, 1], point_wise_emb, choose[:, 1])
    feat = torch.cat((left, right), 1)                                                  # [B,2,1024]
    prec = BF16 if fusion.pointnet_plus.precision == BF16 else FP32
    fused = sft_rows(fusion.sft, feat.reshape(B * 2, -1), _c(center_features).reshape(B * 2, -1), prec)
    fused = fused.view(B, 2, -1)
    if not with_mano:
        return fused
    # one mano_head call per hand, as the reference (train-mode BatchNorm1d statistics are per call)
    theta = torch.stack((mano_head_rows(fusion.mano_head, left.reshape(B, -1)),
                         mano_head_rows(fusion.mano_head, right.reshape(B, -1))), 1)
    return fused, theta


class BucketedAllReduce(object):
    """DDP's exchange step (main.py:44-73, base_trainer.py:94-95,147) for stand-alone use: gradients live as
    VIEWS of a few flat buckets (filled back to front, the order in which the backward pass produces them); a
    post-accumulate hook counts the ready gradients of a bucket and launches its NCCL all-reduce (SUM of
    pre-divided values = mean) asynchronously the moment the last one lands, so the exchange of the late layers
    overlaps the backward kernels of the early ones.  ``finish()`` makes the current stream wait for every
    bucket (and exchanges buckets whose parameters received no gradient this step).

        sync = BucketedAllReduce(params, world_size)
        sync.zero_grad(); loss.backward(); sync.finish(); optimizer.step()
    """

    def __init__(self, params, world_size, bucket_bytes=4 << 20, group=None):
        self.world, self.group = int(world_size), group
        self.params = [p for p in params if p.requires_grad]
        self.buckets, self._of = [], {}
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self._close(cur)
                cur, size = [], 0
        if cur:
            self._close(cur)
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._ready))
        self.bytes_per_step = sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)

    def _close(self, plist):
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        b = dict(flat=flat, params=plist, pending=len(plist), work=None)
        for p in plist:
            self._of[p] = b
        self.buckets.append(b)

    def _launch(self, b):
        import torch.distributed as dist
        b["flat"].mul_(1.0 / self.world)
        b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _ready(self, p):
        b = self._of[p]
        b["pending"] -= 1
        if b["pending"] == 0 and b["work"] is None:
            self._launch(b)

    def zero_grad(self):
        """Zero the buckets (the gradients stay views of them) and re-arm the hooks."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["work"] = len(b["params"]), None
            off = 0
            for p in b["params"]:                         # a caller may have replaced / dropped .grad
                if p.grad is None or p.grad.data_ptr() != b["flat"].data_ptr() + off * b["flat"].element_size():
                    p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += p.numel()

    def finish(self):
        """-> bytes exchanged this step.  After it returns the current stream sees the averaged gradients."""
        import torch.distributed as dist
        if self.world <= 1 or not dist.is_initialized():
            return 0
        for b in self.buckets:
            if b["work"] is None:                         # some parameter of the bucket got no gradient this step
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
        return self.bytes_per_step

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(params, world_size, group=None):
    """DDP's exchange step (main.py:44-73): SUM over ranks then divide by world size, one flat
    bucket over NCCL / NVLink.  No-op outside an initialised process group."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world_size <= 1 or not dist.is_initialized():
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * 4
