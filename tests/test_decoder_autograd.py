"""The decoder's differentiable path (``pdfnet_b200.decoder._forward_autograd``: training / fine-tuning; torch ops
under autograd, any device) - CPU tests.  The kernel path is pinned by tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from pdfnet_b200 import synth
from pdfnet_b200.decoder import decoder

REF = os.environ.get("PDFNET_REFERENCE_ROOT", "/root/reference")


def _ours(dropout=0.05):
    assets = load_golden("gcn_assets")
    m = decoder(assets, precision="fp32", dropout=dropout)
    m.load_state_dict(synth.decoder_state(seed=317, upsample_weight=assets["upsample"]), strict=True)
    return m


def _flat(res):
    result, params, hands, other = res
    out = {}
    for side in ("left", "right"):
        out["verts3d_" + side], out["verts2d_" + side] = result["verts3d"][side], result["verts2d"][side]
        out["verts3d_gcn_" + side], out["verts2d_gcn_" + side] = hands[0]["verts3d"][side], hands[0]["verts2d"][side]
        out["scale_" + side], out["trans2d_" + side], out["root_" + side] = (params["scale"][side], params["trans2d"][side],
                                                                             params["root"][side])
        out["verts3d_mano_" + side] = other["verts3d_MANO_list"][side][0]
        out["verts2d_mano_" + side] = other["verts2d_MANO_list"][side][0]
    return out


def _rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_eval_with_grad_matches_reference_golden_and_backpropagates():
    """eval mode, grad enabled, parameters requiring grad (fine-tuning): every returned tensor equals the unmodified
    reference's (tests/golden/gcn_decoder.npz, 1e-5) and a loss reaches the inputs and every live parameter."""
    g = load_golden("gcn_decoder")
    m = _ours().eval()
    fuse = torch.from_numpy(g["fuse_feat"]).clone().requires_grad_(True)
    out = _flat(m(fuse[:, 0], fuse[:, 1], None))
    for k, v in out.items():
        assert v.requires_grad, k
        assert _rel(v.detach().numpy(), g[k]) < 1e-5, (k, _rel(v.detach().numpy(), g[k]))
    loss = sum((v * v).mean() for v in out.values())
    loss.backward()
    assert torch.isfinite(fuse.grad).all() and float(fuse.grad.abs().max()) > 0
    for n, p in m.named_parameters():
        if "img_ex_" in n or ("GCN_blocks" in n and ".norm1." in n):    # GCN_ResBlock.norm1: result discarded (gcn.py:104-105)
            assert p.grad is None, n
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_no_grad_inference_still_needs_cuda():
    """The kernel path has no CPU fallback: eval + no_grad on CPU tensors fails loudly (unchanged)."""
    m = _ours().eval()
    fuse = torch.zeros((1, 2, 1024))
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(fuse[:, 0], fuse[:, 1], None)


def test_train_mode_dropout():
    """.train(): dropout (p = 0.05, intaghand_decoder.py:275) is active - two passes differ, a fixed seed repeats - and with
    p = 0 the training pass equals the eval pass."""
    fuse = torch.randn((2, 2, 1024), generator=torch.Generator().manual_seed(5))
    m = _ours().train()
    torch.manual_seed(11)
    a = _flat(m(fuse[:, 0], fuse[:, 1], None))["verts3d_left"]
    b = _flat(m(fuse[:, 0], fuse[:, 1], None))["verts3d_left"]
    torch.manual_seed(11)
    c = _flat(m(fuse[:, 0], fuse[:, 1], None))["verts3d_left"]
    assert not torch.equal(a, b) and torch.equal(a, c)
    m0 = _ours(dropout=0.0)
    t = _flat(m0.train()(fuse[:, 0], fuse[:, 1], None))
    e = _flat(m0.eval()(fuse[:, 0], fuse[:, 1], None))
    assert all(torch.equal(t[k], e[k]) for k in t)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib")), reason="reference tree absent")
def test_train_step_equals_reference_decoder_same_seed():
    """A training-mode forward + backward of the UNMODIFIED reference decoder and of ours from the same RNG state:
    same dropout masks (same call order and shapes), so outputs and parameter gradients agree to fp32 rounding."""
    from oracle import ref_import
    ref, _ = ref_import.load_decoder()
    ours = _ours()
    state = {k: v for k, v in ours.state_dict().items() if "img_ex_" not in k and k != "dense_coor"}
    res = ref.load_state_dict(state, strict=False)
    assert not res.unexpected_keys
    B = 2
    fuse = torch.randn((B, 2, 1024), generator=torch.Generator().manual_seed(9))
    fmaps = [torch.zeros((B, 256, r, r)) for r in (12, 24, 48)] + [None]
    outs, grads = [], []
    for m, args in ((ref, (fmaps,)), (ours, (None,))):
        m.train()
        m.zero_grad()
        torch.manual_seed(123)
        o = _flat(m(fuse[:, 0], fuse[:, 1], *args))
        loss = sum((v * v).mean() for v in o.values())
        loss.backward()
        outs.append({k: v.detach().numpy() for k, v in o.items()})
        grads.append({n: p.grad.detach().numpy() for n, p in m.named_parameters() if p.grad is not None and "img_ex_" not in n})
    for k in outs[0]:
        assert _rel(outs[1][k], outs[0][k]) < 1e-5, (k, _rel(outs[1][k], outs[0][k]))
    assert set(grads[0]) == set(grads[1])
    for n in grads[0]:
        if n.endswith("w_ks.bias"):
            # softmax is invariant to a constant added to every key of a row: this gradient is identically zero and
            # both implementations return rounding noise - compare it with the scale of the weight's gradient instead
            scale = np.abs(grads[0][n.replace(".bias", ".weight")]).max()
            assert np.abs(grads[1][n]).max() < 1e-3 * scale and np.abs(grads[0][n]).max() < 1e-3 * scale, n
            continue
        assert _rel(grads[1][n], grads[0][n]) < 1e-4, (n, _rel(grads[1][n], grads[0][n]))


def test_mano_layer_autograd_path_vs_oracle():
    """ManoLayer with inputs that require grad (CtdetLoss differentiates through MANO, simplified.py:730-736): the
    torch formulation equals the oracle (itself pinned to the reference's ManoLayer golden) to 1e-6 m for the
    axis-angle, PCA + matrix-root and new_skel variants, and its input gradients equal the oracle's autograd."""
    from oracle import pdf_oracle as O
    from pdfnet_b200.manolayer import ManoLayer, rodrigues_batch
    for side in ("left", "right"):
        T = dict(load_golden("mano_" + side))
        for kw in (dict(center_idx=9), dict(center_idx=None, new_skel=True), dict(center_idx=9, use_pca=True)):
            layer = ManoLayer(T, **kw)
            g = torch.Generator().manual_seed(3)
            bs = 4
            root, pose = torch.randn((bs, 3), generator=g) * 0.5, torch.randn((bs, 45), generator=g) * 0.3
            shape, trans = torch.randn((bs, 10), generator=g) * 0.5, torch.randn((bs, 3), generator=g) * 0.1
            scale = torch.rand((bs,), generator=g) + 0.5
            if kw.get("use_pca"):
                root, pose = O.rodrigues(root), pose[:, :30]
            grads = []
            for fn in ("ours", "oracle"):
                r, p, s = (t.clone().requires_grad_(True) for t in (root, pose, shape))
                if fn == "ours":
                    v, j = layer(r, p, s, trans, scale, side=side)
                else:
                    v, j = O.mano_lbs(T, r, p, s, trans, scale, side=side, center_idx=kw.get("center_idx"),
                                      new_skel=kw.get("new_skel", False), use_pca=kw.get("use_pca", False))
                ((v * v).sum() + j.sum()).backward()
                grads.append((v.detach(), j.detach(), r.grad, p.grad, s.grad))
            for a, b in zip(*grads):
                assert float((a - b).abs().max()) <= 1e-6 + 1e-4 * float(b.abs().max()), (side, kw)
    a = torch.randn((5, 3), generator=torch.Generator().manual_seed(1)).requires_grad_(True)
    assert float((rodrigues_batch(a) - O.rodrigues(a.detach())).abs().max()) < 1e-6
