"""Pin the CPU oracle against golden vectors frozen from the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import types

import numpy as np
import torch

from conftest import load_golden, mano_tables
from oracle import pdf_oracle as O
from pdfnet_b200 import synth


def _opt(**kw):
    d = dict(SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512, sample_num_level2=128,
             ball_radius=0.015, ball_radius2=0.04, default_resolution=384)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _assert_index_parity(idx, ref_idx, xyz, n_centroids, K, r2):
    """SURVEY 8c parity rule: sorted lists equal wherever the K-th / (K+1)-th
    distances differ; otherwise equal after canonicalising duplicate rows."""
    ref_sorted = np.sort(ref_idx.astype(np.int64), axis=-1)
    ours = np.sort(idx.astype(np.int64), axis=-1)
    exact = (ours == ref_sorted).all(-1)
    if exact.all():
        return
    d2 = np.sort(O.sqdist(xyz, n_centroids), axis=2)
    tie = d2[:, :, K - 1] == d2[:, :, K]
    assert (exact | tie).all(), "index mismatch in a group without a K-th distance tie"
    assert (O.canonicalize_indices(ours, xyz) == O.canonicalize_indices(ref_sorted, xyz)).all()


def test_knn_level1_indices():
    g = load_golden("knn_level1")
    pts = g["points"]
    for tag, r2 in (("r015", 0.015), ("r010", 0.01)):
        idx = O.knn_ball_indices(pts, 512, 64, r2)
        _assert_index_parity(idx, g["idx_" + tag], pts, 512, 64, r2)
    # clouds 0 and 1 have no duplicate rows: indices must be exactly the reference's
    idx = O.knn_ball_indices(pts[:2], 512, 64, 0.015)
    assert (idx == np.sort(g["idx_r015"][:2].astype(np.int64), -1)).all()


def test_group_points_output():
    g = load_golden("knn_level1")
    x, y = O.group_points(g["points"], _opt())
    assert x.shape == (4, 3, 512, 64) and y.shape == (4, 3, 512, 1)
    assert (y == g["center_r015"]).all()
    # order inside a group is implementation-defined: compare lexicographically sorted rows
    def canon(a):
        r = np.ascontiguousarray(a.transpose(0, 2, 3, 1))
        k = np.lexsort((r[..., 2], r[..., 1], r[..., 0]), axis=-1)
        return np.take_along_axis(r, k[..., None], axis=2)
    assert (canon(x) == canon(g["xyz_r015"])).all()


def test_knn_level2():
    g = load_golden("knn_level2")
    p = g["points"]
    xyz = np.ascontiguousarray(p[:, 0:3].transpose(0, 2, 1))
    idx = O.knn_ball_indices(xyz, 128, 64, 0.04)
    _assert_index_parity(idx, g["idx"], xyz, 128, 64, 0.04)
    x, y = O.group_points_2(p, 512, 128, 64, 0.04)
    assert (y == g["center"]).all()
    ch = p.shape[1] - 1
    order_o = np.argsort(x[:1, ch], axis=-1)
    order_r = np.argsort(g["grouped"][:1, ch], axis=-1)
    a = np.take_along_axis(x[:1], order_o[:, None].repeat(p.shape[1], 1), axis=3)
    b = np.take_along_axis(g["grouped"][:1], order_r[:, None].repeat(p.shape[1], 1), axis=3)
    assert (a == b).all()


def test_gather():
    g = load_golden("gather")
    assert (O.tranpose_and_gather_feat(g["feat"], g["ind"]).numpy() == g["out"]).all()


def test_sft():
    g = load_golden("sft")
    for name, (cf, cc) in (("a", (131, 64)), ("b", (3, 3))):
        sd = synth.sft_state("", cf, cc, seed=20)
        o = O.sft_layer(torch.from_numpy(g["fea_" + name]), torch.from_numpy(g["cond_" + name]), sd)
        np.testing.assert_allclose(o.numpy(), g["out_" + name], rtol=1e-6, atol=1e-6)


def test_pointnet_plus_forward():
    g = load_golden("pointnet_plus")
    R, B = int(g["R"]), int(g["B"])
    pts = synth.clouds(B, seed=31)
    pts[2] = synth.clouds(1, seed=32, wrap_from=500)[0]
    out, inter = O.pointnet_plus_forward(synth.pointnet_plus_state(seed=317), pts, synth.pyramid(B, R, seed=31),
                                         synth.choose_indices(B, R, seed=31), _opt(default_resolution=R), True)
    np.testing.assert_allclose(inter["pts0"].numpy(), g["pts0"], rtol=1e-6, atol=1e-7)
    err = np.abs(out.numpy() - g["out"]).max() / np.abs(g["out"]).max()
    assert err < 1e-5, err


def test_fps():
    g = load_golden("fps")
    for tag in "abc":
        order = O.fps_order(g["pc_" + tag], int(g["n_" + tag]), int(g["start_" + tag]))
        assert (np.unique(order) == g["unique_" + tag]).all()


def test_backproject():
    g = load_golden("backproject")
    assert (O.backproject(g["depth"], g["K"]) == g["xyz"]).all()


def test_depth2pcl():
    g = load_golden("depth2pcl")
    for tag in ("full", "wrap_tiny", "invalid", "noise", "h2o"):
        choose, cloud = O.depth2pcl(g["depth_" + tag], g["mask_" + tag], g["K_" + tag], g["valid_" + tag],
                                    g["keys_" + tag], g["perm_" + tag])
        assert (choose == g["choose_" + tag]).all(), tag
        assert (cloud == g["cloud_" + tag]).all(), tag


def test_mano_lbs():
    g = load_golden("mano_lbs")
    for side in ("left", "right"):
        T = mano_tables(side)
        a = {k: g["%s_%s" % (k, side)] for k in ("rot", "pose", "shape", "trans", "scale")}
        for tag, kw in (("plain", {}), ("full", dict(trans=a["trans"], scale=a["scale"], center_idx=9)),
                        ("newskel", dict(trans=a["trans"], new_skel=True))):
            v, j = O.mano_lbs(T, a["rot"], a["pose"], a["shape"], side=side, **kw)
            assert np.abs(v.numpy() - g["v_%s_%s" % (tag, side)]).max() < 1e-6
            assert np.abs(j.numpy() - g["j_%s_%s" % (tag, side)]).max() < 1e-6


def test_mano_extra_rodrigues_pca_and_full_regressor():
    """Round-2 goldens (oracle/make_golden.golden_mano_extra): stand-alone rodrigues_batch, the
    use_pca=True layer (matrix root + PCA coefficients) and ManoModel's 21x778 joint regressor / second LBS."""
    g = load_golden("mano_extra")
    assert np.abs(O.rodrigues(torch.from_numpy(g["rod_axis"])).numpy() - g["rod_R"]).max() < 1e-6
    lbs = load_golden("mano_lbs")
    for side in ("left", "right"):
        T = mano_tables(side)
        a = {k: g["pca_%s_%s" % (k, side)] for k in ("root", "shape", "trans", "scale")}
        for nc in (45, 30):
            coef = g["pca_coef%d_%s" % (nc, side)]
            for tag, kw in (("plain", {}), ("full", dict(trans=a["trans"], scale=a["scale"], center_idx=9))):
                v, j = O.mano_lbs(T, a["root"], coef, a["shape"], side=side, use_pca=True, **kw)
                assert np.abs(v.numpy() - g["pca_v%d_%s_%s" % (nc, tag, side)]).max() < 1e-6
                assert np.abs(j.numpy() - g["pca_j%d_%s_%s" % (nc, tag, side)]).max() < 1e-6
        reg = O.full_regressor(T["J_regressor"])
        assert (reg.numpy() == g["full_regressor_" + side]).all()
        j21 = O.regress_joints(reg, g["model_v_" + side])
        assert np.abs(j21.numpy() - g["model_j21_" + side]).max() < 1e-6
        # the reference's two LBS implementations agree: ManoModel.lbs == ManoLayer on the same inputs
        assert np.abs(g["model_v_" + side] - lbs["v_newskel_" + side]).max() < 1e-6


def test_split_coeff():
    g = load_golden("split_coeff")
    outs = O.split_coeff(g["theta"], g["index"], g["K"], 384, 4)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.numpy(), g["out%d" % i], rtol=1e-6, atol=1e-7)


def test_mano_head():
    g = load_golden("mano_head")
    y = O.mano_head(torch.from_numpy(g["x"]), synth.mano_head_state(seed=317, std=0.05))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-5, atol=1e-6)


def _train_state(dtype=torch.float64):
    sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone())
          for k, v in synth.pointnet_plus_state(seed=317).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return sd


def test_train_step_oracle_vs_reference_autograd():
    """The train-mode restatement reproduces the reference's output, every parameter gradient,
    the pyramid-map gradients and the BatchNorm running-buffer updates (cfg5 semantics).
    Compared in float64, where autograd rounding noise does not mask semantic differences."""
    from conftest import check_grad_digest
    g = load_golden("train_step")
    B, R = int(g["B"]), int(g["R"])
    opt = _opt(default_resolution=R)
    pts, choose, emb, gdir = synth.train_inputs(B, R)
    out32 = O.pointnet_plus_train(_train_state(torch.float32), pts, emb, choose, opt)
    np.testing.assert_allclose(out32.detach().numpy(), g["out_fp32"], rtol=1e-4, atol=1e-5)
    emb = [e.double().requires_grad_(True) for e in emb]
    sd = _train_state()
    out = O.pointnet_plus_train(sd, pts.double(), emb, choose, opt, dtype=torch.float64)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-9, atol=1e-10)
    (out * gdir.double()).sum().backward()
    n = 0
    for k, v in sd.items():
        if v.requires_grad:
            check_grad_digest(g, "grad:" + k, v.grad.numpy(), rtol=1e-6)
            n += 1
    assert n == 3 * 8 + 3 * 3 * 4          # 3 SFT layers x 8 tensors + 9 conv/BN pairs x 4 tensors
    for i, e in enumerate(emb):
        check_grad_digest(g, "grad:emb%d" % i, e.grad.numpy(), rtol=1e-6)
    for k, v in sd.items():
        if "running_" in k:
            np.testing.assert_allclose(v.numpy(), g["buf:" + k], rtol=1e-9, atol=1e-12)


def test_gcn_decoder_oracle_vs_reference():
    """the decoder restatement against the unmodified reference decoder.forward."""
    g, assets = load_golden("gcn_decoder"), load_golden("gcn_assets")
    sd = synth.decoder_state(seed=317, upsample_weight=assets["upsample"])
    with torch.no_grad():
        out = O.gcn_decoder_forward(sd, assets, torch.from_numpy(g["fuse_feat"]), int(g["img_size"]))
    assert out["verts3d_left"].shape == (3, 778, 3)
    for k, v in out.items():
        np.testing.assert_allclose(v.numpy(), g[k], rtol=1e-4, atol=2e-5 * max(1.0, np.abs(g[k]).max()), err_msg=k)
