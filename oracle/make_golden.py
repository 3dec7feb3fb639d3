"""Freeze golden vectors from the UNMODIFIED reference run on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container, where
/root/reference is mounted:

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference ships no tests or fixtures of its own, so
these files ARE the pin for the oracle (``oracle/pdf_oracle.py``) and, through
it, for the CUDA path.  Every array below is produced by calling a reference
function imported from /root/reference through ``oracle/ref_import.py`` (import
stubs only; no reference source is modified or copied).  Randomness inside the
reference (``np.random``) is injected from outside by seeding or by temporarily
replacing ``np.random.shuffle`` with a deterministic stand-in, never by editing
the reference.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from pdfnet_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def golden_knn_level1(ref):
    """group_points (lib/utils/utils.py:134-162) with a 4th feature channel that
    carries the point index, so the gathered output exposes the reference's own
    neighbour indices (order and tie choice as torch.topk left them)."""
    pts = torch.cat([
        synth.clouds(1, seed=1, sigma=0.05),
        synth.clouds(1, seed=2, sigma=0.20),                 # many neighbours beyond the radius
        synth.clouds(1, seed=3, sigma=0.05, wrap_from=300),  # wrap-padded: exact ties
        synth.clouds(1, seed=4, sigma=0.05, wrap_from=40),   # >K copies... every group is all-duplicates
    ])
    B, N, _ = pts.shape
    idx_ch = torch.arange(N, dtype=torch.float32).view(1, N, 1).expand(B, N, 1)
    p4 = torch.cat([pts, idx_ch], 2).contiguous()
    out = {}
    for tag, r2 in (("r015", 0.015), ("r010", 0.01)):
        opt = ref_import.default_opt(INPUT_FEATURE_NUM=4, ball_radius=r2)
        x, y = ref.group_points(p4.clone(), opt)
        out["idx_" + tag] = x[:, 3].round().to(torch.int16).numpy()          # [B,512,64]
        if tag == "r015":
            out["xyz_" + tag] = x[:, 0:3].contiguous().numpy()               # [B,3,512,64]
            out["center_" + tag] = y.contiguous().numpy()
    save("knn_level1", points=pts.numpy(), **out)


def golden_knn_level2(ref):
    """group_points_2 (lib/utils/utils.py:165-187) on channel-major input with an
    index channel appended."""
    C = 6
    pts = synth.clouds(2, n_points=512, seed=5, sigma=0.08)
    pts[1] = synth.clouds(1, n_points=512, seed=6, sigma=0.06, wrap_from=200)[0]
    feat = torch.randn((2, 512, C - 4), generator=torch.Generator().manual_seed(7))
    idx_ch = torch.arange(512, dtype=torch.float32).view(1, 512, 1).expand(2, 512, 1)
    p = torch.cat([pts, feat, idx_ch], 2).transpose(1, 2).contiguous()       # [2,C,512]
    x, y = ref.group_points_2(p.clone(), 512, 128, 64, 0.04)
    save("knn_level2", points=p.numpy(), idx=x[:, C - 1].round().to(torch.int16).numpy(),
         grouped=x.contiguous().numpy(), center=y.contiguous().numpy())


def golden_gather(ref):
    g = torch.Generator().manual_seed(11)
    feat = torch.randn((2, 5, 12, 10), generator=g)
    ind = torch.randint(0, 120, (2, 7), generator=g)
    save("gather", feat=feat.numpy(), ind=ind.numpy(), out=ref.gather(feat, ind).numpy())


def golden_sft(ref):
    g = torch.Generator().manual_seed(12)
    out = {}
    for name, (cf, cc), n in (("a", (131, 64), 16), ("b", (3, 3), 33)):
        sd = synth.sft_state("", cf, cc, seed=20)
        m = ref.SFTLayer(cf, cc)
        m.load_state_dict(sd)
        m.eval()
        fea = torch.randn((2, cf, n), generator=g)
        cond = torch.randn((2, n, cc), generator=g)
        with torch.no_grad():
            o = m((fea, cond))
        out.update({"fea_" + name: fea.numpy(), "cond_" + name: cond.numpy(), "out_" + name: o.numpy()})
    save("sft", **out)


def golden_pointnet_plus(ref):
    """PointNet_Plus.forward (intaghand_encoder.py:118-159), eval mode, synthetic
    weights from pdfnet_b200.synth (loaded into the reference module)."""
    R, B = 64, 3
    opt = ref_import.default_opt(default_resolution=R)
    m = ref.PointNet_Plus(opt)
    missing = m.load_state_dict(synth.pointnet_plus_state(seed=317), strict=False)
    assert all(k.startswith("netR_FC") for k in missing.missing_keys), missing
    m.eval()
    pts = synth.clouds(B, seed=31)
    pts[2] = synth.clouds(1, seed=32, wrap_from=500)[0]
    choose = synth.choose_indices(B, R, seed=31)
    emb = synth.pyramid(B, R, seed=31)
    with torch.no_grad():
        out = m(pts.clone(), emb, choose)
        e0 = ref.gather(emb[0], choose)
        pts0 = m.sft0((pts.transpose(1, 2), e0))
    save("pointnet_plus", out=out.numpy(), pts0=pts0.numpy(), R=np.int64(R), B=np.int64(B))


def golden_fps(ref):
    fps = ref_import.load_fps()
    out = {}
    cases = [("a", synth.clouds(1, seed=41)[0].numpy(), 512), ("b", synth.clouds(1, n_points=512, seed=42)[0].numpy(), 128),
             ("c", synth.clouds(1, seed=43, wrap_from=700)[0].numpy(), 512)]
    for tag, pc, n in cases:
        s = 1000 + zlib.crc32(tag.encode()) % 1000
        np.random.seed(s)
        start = np.random.randint(pc.shape[0])           # first draw inside the reference (:159)
        np.random.seed(s)
        res = fps(pc, n)
        out.update({"pc_" + tag: pc, "start_" + tag: np.int64(start), "n_" + tag: np.int64(n),
                    "unique_" + tag: np.asarray(res, dtype=np.int64)})
    save("fps", **out)


def golden_backproject(ref):
    g = torch.Generator().manual_seed(51)
    depth = (0.3 + torch.rand((40, 56), generator=g)).numpy().astype(np.float32)
    depth[depth < 0.5] = 0.0
    K = np.array([[210.0, 0, 27.5], [0, 205.0, 20.25], [0, 0, 1]], dtype=np.float32)
    xyz, _ = ref.get_normal(depth, K, False)
    save("backproject", depth=depth, K=K, xyz=xyz)


class _InjectedShuffle(object):
    """Deterministic stand-in for np.random.shuffle while depth2pcl runs (see the
    module docstring).  Behaviour is documented in oracle.pdf_oracle.depth2pcl."""

    def __init__(self, subset_keys, perm, num_points=1024):
        self.keys, self.perm, self.n = subset_keys, perm, num_points

    def __call__(self, arr):
        loc = sys._getframe(1).f_locals
        hand = 1 if "points_xyz_right" in loc else 0
        if len(arr) > self.n:                                # c_mask: pick the kept subset
            cand = loc["choose_right" if hand else "choose_left"]
            keep = np.argsort(np.asarray(self.keys[hand])[cand], kind="stable")[: self.n]
            arr[:] = 0
            arr[keep] = 1
        else:                                                # final shuffle of choose
            arr[:] = arr[np.asarray(self.perm[hand])]


def _real_depth_frame(R):
    """One of the reference's own RGB-D assets (assets/H2O/depth, uint16 mm),
    centre-cropped and nearest-resized to RxR metres; hand masks = left/right
    halves of the near-range (<0.7 m) pixels."""
    import cv2
    p = os.path.join(ref_import.REFERENCE_ROOT, "assets", "H2O", "depth", "000094.png")
    d = cv2.imread(p, cv2.IMREAD_UNCHANGED).astype(np.float32) / 1000.0
    h, w = d.shape
    d = d[:, (w - h) // 2:(w - h) // 2 + h]
    d = cv2.resize(d, (R, R), interpolation=cv2.INTER_NEAREST)
    near = ((d > 0.2) & (d < 0.7)).astype(np.float32)
    mask = np.zeros((1, 2, R, R), dtype=np.float32)
    mask[0, 1, :, : R // 2] = near[:, : R // 2]
    mask[0, 0, :, R // 2:] = near[:, R // 2:]
    f = 636.66 * R / h
    K = np.array([[f, 0, R / 2.0], [0, f, R / 2.0], [0, 0, 1]], dtype=np.float32)
    return d, mask, K


def golden_depth2pcl(ref):
    R = 96
    out = {}
    depth_t, mask_t, K_t, _ = synth.rgbd_frames(1, R, seed=61)
    cases = {}
    d0, m0 = depth_t[0].numpy().copy(), mask_t.numpy().copy()
    cases["full"] = (d0, m0, K_t[0].numpy(), np.array([[1, 1]]))                 # both hands > 1024 px
    d1, m1 = d0.copy(), m0.copy()
    m1[0, 1, :, :] = 0
    m1[0, 1, 30:50, 10:25] = 1                                                   # left: 300 px -> wrap pad
    m1[0, 0, :, :] = 0
    m1[0, 0, 30:32, 60:64] = 1                                                   # right: 8 px -> zeros
    cases["wrap_tiny"] = (d1, m1, K_t[0].numpy(), np.array([[1, 1]]))
    cases["invalid"] = (d0, m0, K_t[0].numpy(), np.array([[0, 1]]))              # left flagged invalid
    d3 = d0.copy()
    d3[30:40, :] = 3.0                                                           # beyond Z_max: noise gate
    d3[45:50, :] = 0.1
    cases["noise"] = (d3, m0, K_t[0].numpy(), np.array([[1, 1]]))
    dr, mr, Kr = _real_depth_frame(128)
    cases["h2o"] = (dr, mr, Kr, np.array([[1, 1]]))
    orig = np.random.shuffle
    for tag, (d, m, K, valid) in cases.items():
        Rr = d.shape[0]
        rs = np.random.RandomState(zlib.crc32(tag.encode()) % 10000)
        keys = np.stack([rs.permutation(Rr * Rr) for _ in range(2)]).astype(np.int32)
        perm = np.stack([rs.permutation(1024) for _ in range(2)]).astype(np.int32)
        np.random.shuffle = _InjectedShuffle(keys, perm)
        try:
            choose, cloud = ref.depth2pcl(torch.from_numpy(d), torch.from_numpy(m), torch.from_numpy(K),
                                          torch.from_numpy(valid))
        finally:
            np.random.shuffle = orig
        out.update({"depth_" + tag: d, "mask_" + tag: m, "K_" + tag: K, "valid_" + tag: valid,
                    "keys_" + tag: keys, "perm_" + tag: perm,
                    "choose_" + tag: choose.astype(np.int64), "cloud_" + tag: cloud.astype(np.float32)})
    save("depth2pcl", **out)


def export_mano_tables(ref):
    """MANO constants exactly as ManoLayer registers them (manolayer.py:117-152)."""
    layers = {}
    for side in ("left", "right"):
        m = ref.ManoLayer(os.path.join(ref.mano_dir, "MANO_%s.pkl" % side.upper()), center_idx=None, use_pca=False)
        layers[side] = m
        save("mano_" + side,
             v_template=m.v_template.numpy(), shapedirs=m.shapedirs.numpy(), posedirs=m.posedirs.numpy(),
             J_regressor=m.J_regressor.numpy(), weights=m.weights.numpy(),
             hands_components=m.hands_components.numpy(), hands_mean=m.hands_mean.numpy(),
             parent=np.asarray([int(p) for p in m.parent], dtype=np.int64),
             faces=np.asarray(m.faces, dtype=np.int32))
    return layers


def golden_mano(ref):
    out = {}
    for side in ("left", "right"):
        path = os.path.join(ref.mano_dir, "MANO_%s.pkl" % side.upper())
        rot, pose, shape, trans = synth.mano_inputs(6, seed=71 if side == "left" else 72)
        scale = torch.rand((6,), generator=torch.Generator().manual_seed(73)) + 0.5
        out.update({"rot_" + side: rot.numpy(), "pose_" + side: pose.numpy(), "shape_" + side: shape.numpy(),
                    "trans_" + side: trans.numpy(), "scale_" + side: scale.numpy()})
        for tag, kw, ci, ns in (("plain", {}, None, False), ("full", dict(trans=trans, scale=scale), 9, False),
                                ("newskel", dict(trans=trans), None, True)):
            m = ref.ManoLayer(path, center_idx=ci, use_pca=False, new_skel=ns)
            with torch.no_grad():
                v, j = m(rot.clone(), pose.clone(), shape.clone(), side=side, **kw)
            out["v_%s_%s" % (tag, side)] = v.numpy()
            out["j_%s_%s" % (tag, side)] = j.numpy()
    save("mano_lbs", **out)


def golden_split_coeff(ref):
    split = ref_import.load_split_coeff()
    g = torch.Generator().manual_seed(81)
    theta = torch.randn((5, 122), generator=g) * 0.2
    index = torch.randint(0, 96 * 96, (5,), generator=g)
    K = torch.tensor([[300.0, 0, 192.0], [0, 310.0, 190.0], [0, 0, 1]]).repeat(5, 1, 1)
    import types
    fake_self = types.SimpleNamespace(opt=types.SimpleNamespace(using_pca=False, down_ratio=4), input_res=384)
    outs = split(fake_self, theta.clone(), index, K)
    save("split_coeff", theta=theta.numpy(), index=index.numpy(), K=K.numpy(),
         **{"out%d" % i: o.numpy() for i, o in enumerate(outs)})


def golden_mano_head(ref):
    """mano_head is a plain nn.Sequential (intaghand_encoder.py:630-643); rebuild it
    with torch.nn exactly as the reference constructs it and freeze eval outputs."""
    import torch.nn as nn
    head = nn.Sequential(nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                         nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Linear(256, 122))
    sd = {k[len("mano_head."):]: v for k, v in synth.mano_head_state(seed=317, std=0.05).items()}
    head.load_state_dict(sd)
    head.eval()
    x = torch.randn((4, 1024), generator=torch.Generator().manual_seed(91))
    with torch.no_grad():
        y = head(x)
    save("mano_head", x=x.numpy(), y=y.numpy())


def grad_digest(name, g):
    """Full tensor when small, otherwise (sum, L2 norm, every 97th element)."""
    g = g.detach().reshape(-1).double()
    if g.numel() <= 20000:
        return {name: g.numpy()}
    return {name + "@sum": np.float64(g.sum().item()), name + "@norm": np.float64(g.norm().item()),
            name + "@s97": g[::97].numpy()}


def golden_train_step(ref):
    """One training-mode forward + backward of the UNMODIFIED reference PointNet_Plus
    (intaghand_encoder.py:118-159, nn.BatchNorm2d batch statistics, torch.autograd):
    loss = sum(out * gdir).  Records the output, every parameter gradient, the gradients of
    the three pyramid maps and the updated BatchNorm running buffers.  The module is run in
    float64 (``m.double()``): in fp32 the early-layer gradients of this network carry ~1 %
    rounding noise, which would hide semantic differences; the fp32 forward output is kept too."""
    R, B = 64, 2
    opt = ref_import.default_opt(default_resolution=R)
    m = ref.PointNet_Plus(opt)
    m.load_state_dict(synth.pointnet_plus_state(seed=317), strict=False)
    m.train()
    pts, choose, emb, gdir = synth.train_inputs(B, R)
    with torch.no_grad():
        out32 = m(pts.clone(), [e.clone() for e in emb], choose).numpy()
    m.load_state_dict(synth.pointnet_plus_state(seed=317), strict=False)     # undo the running-stat update
    m.double()
    emb = [e.double().requires_grad_(True) for e in emb]
    out = m(pts.double(), emb, choose)
    (out * gdir.double()).sum().backward()
    rec = dict(out=out.detach().numpy(), out_fp32=out32, R=np.int64(R), B=np.int64(B))
    for k, p in m.named_parameters():
        if k.startswith("netR_FC"):
            continue
        rec.update(grad_digest("grad:" + k, p.grad))
    for i, e in enumerate(emb):
        rec.update(grad_digest("grad:emb%d" % i, e.grad))
    for k, b in m.named_buffers():
        if k.startswith("netR_FC") or k.endswith("num_batches_tracked"):
            continue
        rec["buf:" + k] = b.detach().numpy()
    save("train_step", **rec)


def export_gcn_assets(ref_dec):
    """Graph assets of the GCN decoder (lib/models/networks/gcn_core/*.pkl) as plain arrays: the three
    coarsened Laplacians per hand (dense, as GCN_ResBlock registers them, gcn.py:83-87), the vertex
    permutations, the dense colour table and the 252 -> 778 upsampling matrix.  Data, not code."""
    m, _ = ref_dec
    rec = {"dense_coor": m.dense_coor.numpy(), "upsample": m.unsample_layer.weight.detach().numpy()}
    for side in ("left", "right"):
        rec["graph_perm_" + side] = np.asarray(m.converter[side].graph_perm, dtype=np.int64)
        rec["graph_perm_reverse_" + side] = np.asarray(m.converter[side].graph_perm_reverse, dtype=np.int64)
        for i in range(3):
            blk = getattr(m.dual_gcn.layers[i], "graph_" + side).GCN_blocks[0]
            rec["L_%s_%d" % (side, i)] = blk.graph_L.numpy()
    save("gcn_assets", **rec)
    return rec


def golden_gcn_decoder(ref_dec):
    """decoder.forward (intaghand_decoder.py:180-242) of the UNMODIFIED reference, eval mode, synthetic
    weights from pdfnet_b200.synth.decoder_state loaded into the reference module (strict for every
    parameter forward uses), driven the way HandNET_GCN.forward drives it: global features =
    fuse_feat[:, 0] / fuse_feat[:, 1] (intaghand_encoder.py:873-874), fmaps = zero placeholders (shape-checked, never read)."""
    m, dec = ref_dec
    state = synth.decoder_state(seed=317, upsample_weight=m.unsample_layer.weight.detach())
    res = m.load_state_dict(state, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("img_ex" in k or k == "dense_coor" for k in res.missing_keys), res.missing_keys      # buffer: an asset
    m.eval()
    B = 3
    fuse = torch.randn((B, 2, 1024), generator=torch.Generator().manual_seed(71))
    with torch.no_grad():
        # fmaps only pass DualGraphLayer's shape asserts (DualGraph.py:69-73); their values are never read
        fmaps = [torch.zeros((B, 256, r, r)) for r in (12, 24, 48)] + [None]
        result, params, hand_list, other = m(fuse[:, 0], fuse[:, 1], fmaps)
    rec = {"fuse_feat": fuse.numpy(), "img_size": np.int64(dec.IMG_SIZE)}
    for side in ("left", "right"):
        rec["verts3d_" + side] = result["verts3d"][side].numpy()
        rec["verts2d_" + side] = result["verts2d"][side].numpy()
        rec["verts3d_gcn_" + side] = hand_list[0]["verts3d"][side].numpy()
        rec["verts2d_gcn_" + side] = hand_list[0]["verts2d"][side].numpy()
        rec["scale_" + side] = params["scale"][side].numpy()
        rec["trans2d_" + side] = params["trans2d"][side].numpy()
        rec["root_" + side] = params["root"][side].numpy()
        rec["verts3d_mano_" + side] = other["verts3d_MANO_list"][side][0].numpy()
        rec["verts2d_mano_" + side] = other["verts2d_MANO_list"][side][0].numpy()
    save("gcn_decoder", **rec)


def golden_mano_extra(ref):
    """Round-2 additions, all from the UNMODIFIED reference:
    * rodrigues_batch (manolayer.py:32-48) stand-alone, incl. a zero and a tiny rotation;
    * ManoLayer(use_pca=True): root rotation as a 3x3 matrix, pose as PCA coefficients (:266-267), the way
      the InterHand dataset drives it (interhand.py:192,220-223), 45 and 30 components;
    * ManoModel (lib/models/hand3d/Mano_model.py): full_regressor (:309-323), joints = full_regressor @ verts
      (demo.py:217-218) and the second LBS implementation (``lbs``, :560-647) on the mano_lbs.npz inputs."""
    import importlib
    ref_import.load_split_coeff()                       # installs the pytorch3d import stubs Mano_model needs
    MM = importlib.import_module("lib.models.hand3d.Mano_model")
    out = {}
    g = torch.Generator().manual_seed(101)
    axis = torch.randn((64, 3), generator=g) * 0.8
    axis[0] = 0.0
    axis[1] = torch.tensor([1e-6, -2e-6, 5e-7])
    axis[2] = torch.tensor([3.1, 0.0, 0.0])
    out["rod_axis"] = axis.numpy()
    out["rod_R"] = ref.rodrigues_batch(axis.clone()).numpy()
    for side in ("left", "right"):
        path = os.path.join(ref.mano_dir, "MANO_%s.pkl" % side.upper())
        rot, pose, shape, trans = synth.mano_inputs(6, seed=171 if side == "left" else 172)
        Rroot = ref.rodrigues_batch(rot.clone())
        scale = torch.rand((6,), generator=torch.Generator().manual_seed(173)) + 0.5
        out.update({"pca_root_" + side: Rroot.numpy(), "pca_shape_" + side: shape.numpy(),
                    "pca_trans_" + side: trans.numpy(), "pca_scale_" + side: scale.numpy()})
        for nc in (45, 30):
            coef = torch.randn((6, nc), generator=torch.Generator().manual_seed(180 + nc)) * 0.7
            out["pca_coef%d_%s" % (nc, side)] = coef.numpy()
            for tag, kw, ci in (("plain", {}, None), ("full", dict(trans=trans, scale=scale), 9)):
                m = ref.ManoLayer(path, center_idx=ci, use_pca=True)
                with torch.no_grad():
                    v, j = m(Rroot.clone(), coef.clone(), shape.clone(), side=side, **kw)
                out["pca_v%d_%s_%s" % (nc, tag, side)] = v.numpy()
                out["pca_j%d_%s_%s" % (nc, tag, side)] = j.numpy()
        mm = MM.ManoModel(model_path=path, is_rhand=(side == "right"), use_pca=False, flat_hand_mean=True,
                          num_pca_comps=45)
        out["full_regressor_" + side] = mm.full_regressor.numpy()
        rot2, pose2, shape2, trans2 = synth.mano_inputs(6, seed=71 if side == "left" else 72)   # = mano_lbs.npz inputs
        with torch.no_grad():
            o = mm(betas=shape2.clone(), global_orient=rot2.clone(), hand_pose=pose2.clone(), transl=trans2.clone(),
                   using_wrist_rotate=True)
            out["model_v_" + side] = o.vertices.numpy()
            out["model_j16_" + side] = o.joints.numpy()
            out["model_j21_" + side] = torch.matmul(mm.full_regressor, o.vertices).numpy()
    save("mano_extra", **out)


def main():
    ref = ref_import.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if len(sys.argv) > 1:                               # python oracle/make_golden.py mano_extra ...
        for name in sys.argv[1:]:
            globals()["golden_" + name](ref)
        return
    golden_knn_level1(ref)
    golden_knn_level2(ref)
    golden_gather(ref)
    golden_sft(ref)
    golden_pointnet_plus(ref)
    golden_fps(ref)
    golden_backproject(ref)
    golden_depth2pcl(ref)
    export_mano_tables(ref)
    golden_mano(ref)
    golden_split_coeff(ref)
    golden_mano_head(ref)
    golden_train_step(ref)
    ref_dec = ref_import.load_decoder()
    export_gcn_assets(ref_dec)
    golden_gcn_decoder(ref_dec)
    golden_mano_extra(ref)


if __name__ == "__main__":
    main()
