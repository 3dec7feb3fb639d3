"""CPU-only checks: the C-ABI library loads and exports what include/pdfnet_b200.h declares,
host-side logic (BN folding, weight packing, padding, sharding), and loud failure without CUDA."""
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from pdfnet_b200 import synth


def _opt(**kw):
    d = dict(SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512, sample_num_level2=128,
             ball_radius=0.015, ball_radius2=0.04, default_resolution=64, PCA_SZ=63)
    d.update(kw)
    return types.SimpleNamespace(**d)


@pytest.fixture(scope="session")
def lib():
    from pdfnet_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from pdfnet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pdfnet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pdf_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.EXPORTS, (declared, _lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pdf_version() >= 100
    assert lib.pdf_sa_pack_size(3, 64, 64, 128) == 32768 + 128 * 4        # bf16 operand tiles + fp32 layer-3 bias
    assert lib.pdf_sa_pack_size(131, 128, 128, 256) == 147456 + 256 * 4
    assert lib.pdf_sa_pack_size(3, 64, 64, 64) == -1


def test_no_cpu_fallback(lib):
    from pdfnet_b200 import ManoLayer, ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn_ball(torch.zeros((1, 1024, 3)), 512, 64, 0.01)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.zeros((4, 4)), torch.zeros((4, 4)))
    layer = ManoLayer(synth.to_numpy(synth.synthetic_mano_tables()), center_idx=None)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros((1, 3)), torch.zeros((1, 45)), torch.zeros((1, 10)))


def test_host_resident_pyramid_interface_fails_loudly(lib):
    """Zero-copy hand-off (pdf_host_device_pointer / PointNet_Plus.gather / forward(gathered=)): argument errors and
    pageable memory are reported, nothing is dereferenced; the split pass exists for the bf16 inference path only."""
    import ctypes
    from pdfnet_b200 import HandFusion, PointNet_Plus, _lib
    assert lib.pdf_host_device_pointer(None, None) == -1
    assert b"null" in lib.pdf_last_error()
    out = ctypes.c_void_p()
    buf = torch.zeros(64)                                            # pageable (and, here, no CUDA device at all)
    assert lib.pdf_host_device_pointer(ctypes.c_void_p(buf.data_ptr()), ctypes.byref(out)) == -1 and not out.value
    with pytest.raises(RuntimeError, match="page-locked"):
        _lib.host_ptr(buf)
    emb = [e.bfloat16().contiguous(memory_format=torch.channels_last) for e in synth.pyramid(1, 64, seed=3)]
    cloud, choose = synth.clouds(2, seed=3), synth.choose_indices(2, 64, seed=3)
    with pytest.raises(RuntimeError, match="bf16"):
        PointNet_Plus(_opt(), "fp32").eval().gather(cloud, emb, choose, 2)
    with pytest.raises(RuntimeError, match="bf16"):
        PointNet_Plus(_opt(), "bf16").eval().gather(cloud, [e.float() for e in emb], choose, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        PointNet_Plus(_opt(), "bf16").eval().gather(cloud, emb, choose, 2)       # CPU clouds: no fallback
    g = (torch.zeros((2, 1024, 3)), torch.zeros(1, dtype=torch.uint8), torch.zeros(1, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="bf16 inference"):
        PointNet_Plus(_opt(), "fp32").eval()(cloud, None, None, 2, gathered=g)
    with pytest.raises(RuntimeError, match="bf16 inference"):
        PointNet_Plus(_opt(), "bf16").train()(cloud, None, None, 2, gathered=g)
    with pytest.raises(RuntimeError, match="CUDA"):
        HandFusion(_opt(), "bf16").eval()(cloud.view(1, 2, 1024, 3), None, None, torch.zeros((1, 2, 1024)), gathered=g)


def test_bad_arguments_return_errors_not_crashes(lib):
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.pdf_knn_ball(None, 1, 1024, 512, 64, 0.01, 0, 0, 0, None, None) == -1
    assert lib.pdf_knn_ball(None, 0, 1024, 512, 64, 0.01, 0, 0, 0, None, None) == 0     # empty batch: no-op
    assert b"null" in lib.pdf_last_error()
    assert lib.pdf_linear_f32(None, 0, None, 0, None, 1, 1, 1, 0, 0, 0, None, 0, None, 0, None) == -1
    assert lib.pdf_sa_pack_weights_host(None, None, None, None, None, None, 3, 64, 64, 128, None) == -1
    # training / decoder entry points: null pointers and unsupported shapes are reported, never dereferenced
    assert lib.pdf_bn_stats(None, 64, 10, 64, None, None) == -1
    assert lib.pdf_linear_tn_f32(None, 0, None, 0, 10, 4, 4, None, 4, None) == -1
    assert lib.pdf_gemm_tn_bf16(None, 64, None, 64, 128, 1, 4, None, 64, 0, None) == -1
    assert lib.pdf_gemm_tn_bf16(None, 64, None, 64, 0, 1, 4, None, 64, 0, None) == 0      # no rows: no-op
    assert lib.pdf_row_combine(None, 0, None, 0, None, 0, 1, 1, 64, 10, None, None, 1e-6, 0, None, 0, None, 0,
                               None, None, None) == -1
    assert lib.pdf_mha(None, 0, None, 0, None, 0, 0, 63, 4, 64, None, 0, None) == 0           # no samples: no-op
    assert lib.pdf_mha(None, 0, None, 0, None, 0, 1, 63, 4, 64, None, 0, None) == -1
    assert lib.pdf_linear_smallk_f32(0, None, 0, None, 0, None, 10, 64, 3, None, 64, None) == -1
    assert lib.pdf_group_scatter_add(None, None, 1, 1024, 512, 64, 3, None, 3, None) == -1


def test_state_dict_keys_match_reference_layout(lib):
    from pdfnet_b200 import HandFusion, PointNet_Plus
    m = PointNet_Plus(_opt())
    keys = set(m.state_dict().keys())
    want = set(synth.pointnet_plus_state().keys())
    assert want <= keys
    assert all(k.startswith("netR_FC") for k in keys - want)
    for k, v in synth.pointnet_plus_state().items():
        assert tuple(m.state_dict()[k].shape) == tuple(v.shape), k
    hf = HandFusion(_opt())
    assert set(synth.mano_head_state().keys()) <= set(hf.state_dict().keys())
    assert {"sft." + k for k in synth.fusion_sft_state()} <= set(hf.state_dict().keys())


def test_bn_folding_matches_torch_eval(lib):
    from pdfnet_b200 import PointNet_Plus
    from pdfnet_b200.encoder import _fold_bn
    m = PointNet_Plus(_opt())
    m.load_state_dict(synth.pointnet_plus_state(), strict=False)
    m.eval()
    x = torch.randn((2, 131, 5, 7))
    w, b = _fold_bn(m.netR_2[0], m.netR_2[1])
    with torch.no_grad():
        ref = m.netR_2[1](m.netR_2[0](x))
    got = torch.einsum("oc,bchw->bohw", w, x) + b[None, :, None, None]
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)


def test_padded_weights_are_equivalent(lib):
    from pdfnet_b200 import PointNet_Plus
    m = PointNet_Plus(_opt())
    m.load_state_dict(synth.pointnet_plus_state(), strict=False)
    m.eval()
    f = m.folded()
    w1, _ = f["netR_2"][0]
    x = torch.randn((9, 131))
    xp = torch.cat([x[:, :3], torch.full((9, 1), 123.0), x[:, 3:]], 1)   # pad column must be ignored
    assert torch.allclose(xp @ f["netR_2_w1pad"].t(), x @ w1.t(), rtol=1e-6, atol=1e-5)
    ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1 = f["sft1"]
    assert ws1.shape == (132, 64) and float(ws1[3].abs().sum()) == 0 and float(bs1[3]) == 0 and float(bh1[3]) == 0


def _unpack_image(img, off_feat, off_aux, rows, kfeat):
    """Read a packed layer back with the documented byte offsets (independent re-statement)."""
    u16 = lambda byte: int(img[byte]) | (int(img[byte + 1]) << 8)
    def bf(byte):
        return np.array([u16(byte) << 16], dtype=np.uint32).view(np.float32)[0]
    feat = np.zeros((rows, kfeat), np.float32)
    for r in range(rows):
        for k in range(kfeat):
            off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2
            feat[r, k] = bf(off_feat + (k >> 6) * rows * 128 + off)
    aux = np.zeros((rows, 16), np.float32)
    for r in range(rows):
        for k in range(16):
            aux[r, k] = bf(off_aux + (r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2)
    return feat, aux


def test_sa_weight_image_layout(lib):
    """The host packer writes bf16 weights at the SW128 / interleave offsets the kernel reads."""
    from pdfnet_b200 import ops
    g = torch.Generator().manual_seed(1)
    w1, b1 = torch.randn((64, 3), generator=g), torch.randn((64,), generator=g)
    w2, b2 = torch.randn((64, 64), generator=g), torch.randn((64,), generator=g)
    w3, b3 = torch.randn((128, 64), generator=g), torch.randn((128,), generator=g)
    img = ops.sa_pack_weights(w1, b1, w2, b2, w3, b3).numpy()
    assert img.shape == (32768 + 128 * 4,)
    assert (img[32768:].view(np.float32) == b3.numpy()).all()            # fp32 layer-3 bias after the operand tiles
    bf = lambda t: t.bfloat16().float().numpy()
    _, aux1 = _unpack_image(img, 0, 0, 64, 0)
    assert (aux1[:, 0:3] == bf(w1)).all() and (aux1[:, 4:7] == bf(w1)).all()
    assert np.allclose(aux1[:, 3] + aux1[:, 7], b1.numpy(), rtol=1e-4, atol=1e-6) and (aux1[:, 8:] == 0).all()
    feat2, aux2 = _unpack_image(img, 2048, 2048 + 8192, 64, 64)
    assert (feat2 == bf(w2)).all() and (aux2[:, :3] == 0).all() and (aux2[:, 3] == bf(b2)).all()
    feat3, aux3 = _unpack_image(img, 2048 + 8192 + 2048, 2048 + 8192 + 2048 + 16384, 128, 64)
    assert (feat3 == bf(w3)).all() and (aux3[:, 3] == bf(b3)).all()
    # level 2: feature k of the gathered row multiplies W1 column 3 + k
    w1b, b1b = torch.randn((128, 131), generator=g), torch.randn((128,), generator=g)
    w2b, b2b = torch.randn((128, 128), generator=g), torch.randn((128,), generator=g)
    w3b, b3b = torch.randn((256, 128), generator=g), torch.randn((256,), generator=g)
    img = ops.sa_pack_weights(w1b, b1b, w2b, b2b, w3b, b3b).numpy()
    feat1, aux1 = _unpack_image(img, 0, 32768, 128, 128)
    assert (feat1 == bf(w1b[:, 3:])).all() and (aux1[:, 0:3] == bf(w1b[:, :3])).all()
    feat3, _ = _unpack_image(img, 36864 * 2, 36864 * 2 + 65536, 256, 128)
    assert (feat3 == bf(w3b)).all()
    assert (img[147456:].view(np.float32) == b3b.numpy()).all()


def test_mano_pkl_loader_matches_npz_export(lib):
    from pdfnet_b200.manolayer import kernel_tables, load_mano_data
    d = load_mano_data(os.path.join(ROOT, "tests", "golden", "mano_left.npz"))
    t = kernel_tables(*[torch.from_numpy(d[k]) for k in ("v_template", "shapedirs", "posedirs", "J_regressor",
                                                           "weights")], "cpu")
    assert t["posedirs_t"].shape == (135, 2334) and t["shapedirs_t"].shape == (10, 2334)
    assert t["weights_t"].shape == (16, 778) and t["j_shapedirs"].shape == (48, 10)
    J = torch.from_numpy(d["J_regressor"]).double() @ torch.from_numpy(d["v_template"]).double()
    assert torch.allclose(t["j_template"].view(16, 3).double(), J, atol=1e-7)
    ref_pkl = "/root/reference/lib/models/hand3d/mano_core/MANO_LEFT.pkl"
    if os.path.exists(ref_pkl):                       # only in the authoring container
        p = load_mano_data(ref_pkl)
        for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights"):
            assert (p[k] == d[k]).all(), k


def test_seeded_depth2pcl_randomness_host_mirror_matches_numpy(lib):
    """pdf_depth2pcl_host_randomness (the C mirror of the kernel's counter-based keys / permutation) equals the
    oracle's independent numpy restatement; the permutation is a bijection of [0,1024)."""
    from oracle import pdf_oracle as O
    from pdfnet_b200 import ops
    for seed in (0, 1, 317, 2 ** 32 - 1):
        k, p = ops.d2p_host_randomness(seed, 5, 3000)
        k2, p2 = O.d2p_seeded_randomness(seed, 5, 3000)
        assert (k.numpy() == k2).all() and (p.numpy() == p2).all()
        for c in range(5):
            assert sorted(p2[c].tolist()) == list(range(1024))
    assert not (O.d2p_seeded_randomness(1, 2, 64)[0][0] == O.d2p_seeded_randomness(1, 2, 64)[0][1]).all()
    assert not (O.d2p_seeded_randomness(1, 1, 64)[1] == O.d2p_seeded_randomness(2, 1, 64)[1]).all()


def test_shard_range_covers_everything():
    from pdfnet_b200.parallel import shard_range
    for n in (0, 1, 7, 128, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_patch_reference_rebinds_every_boundary_name(monkeypatch):
    """patch_reference() swaps the names of SURVEY 8b inside (stand-ins for) the reference's modules."""
    import types
    import pdfnet_b200
    names = ["lib", "lib.utils", "lib.utils.utils", "lib.models", "lib.models.utils", "lib.models.networks",
             "lib.models.networks.intaghand_encoder", "lib.models.networks.manolayer",
             "lib.models.networks.intaghand_decoder", "lib.models.networks.intaghand_model"]
    for n in names:
        monkeypatch.setitem(sys.modules, n, types.ModuleType(n))
    patched = pdfnet_b200.patch_reference()
    enc = sys.modules["lib.models.networks.intaghand_encoder"]
    from pdfnet_b200 import decoder, encoder, grouping, manolayer
    assert enc.PointNet_Plus is encoder.PointNet_Plus and enc.SFTLayer is encoder.SFTLayer
    assert enc.group_points is grouping.group_points and enc.depth2pcl is encoder.depth2pcl
    assert sys.modules["lib.models.networks.manolayer"].ManoLayer is manolayer.ManoLayer
    assert sys.modules["lib.models.networks.intaghand_model"].load_decoder is decoder.load_decoder
    assert sys.modules["lib.models.networks.manolayer"].rodrigues_batch is manolayer.rodrigues_batch
    assert len(patched) == 14


def test_patch_reference_training_mode_keeps_the_reference_autograd_modules(monkeypatch):
    """mode='training': the replacements whose backward is torch autograd rather than this library's kernels (decoder,
    ManoLayer, rodrigues_batch) are NOT installed; mode='training-all' installs them too."""
    import types
    import pdfnet_b200
    names = ["lib", "lib.utils", "lib.utils.utils", "lib.models", "lib.models.utils", "lib.models.networks",
             "lib.models.networks.intaghand_encoder", "lib.models.networks.manolayer",
             "lib.models.networks.intaghand_decoder", "lib.models.networks.intaghand_model"]
    for n in names:
        monkeypatch.setitem(sys.modules, n, types.ModuleType(n))
    patched = pdfnet_b200.patch_reference(mode="training")
    assert len(patched) == 10 and not any(p.endswith(("load_decoder", "ManoLayer", "rodrigues_batch")) for p in patched)
    assert not hasattr(sys.modules["lib.models.networks.manolayer"], "ManoLayer")
    from pdfnet_b200 import encoder
    assert sys.modules["lib.models.networks.intaghand_encoder"].PointNet_Plus is encoder.PointNet_Plus
    assert len(pdfnet_b200.patch_reference(mode="training-all")) == 14      # + load_decoder (x2), ManoLayer, rodrigues_batch
    with pytest.raises(ValueError):
        pdfnet_b200.patch_reference(mode="bogus")


@pytest.mark.skipif(not os.path.isdir("/root/reference/lib/models/networks/gcn_core"), reason="reference assets absent")
def test_decoder_assets_from_reference_pickles_match_fixture():
    """assets_from_graph_dicts (what the drop-in load_decoder builds from gcn_core/*.pkl) == tests/golden/gcn_assets.npz."""
    import pickle
    import warnings
    from pdfnet_b200.decoder import assets_from_graph_dicts, decoder
    base = "/root/reference/lib/models/networks/gcn_core/"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        objs = [pickle.load(open(base + n, "rb")) for n in ("graph_left.pkl", "graph_right.pkl", "v_color.pkl", "upsample.pkl")]
    assets = assets_from_graph_dicts(*objs)
    fixture = dict(np.load(os.path.join(ROOT, "tests", "golden", "gcn_assets.npz")))
    assert sorted(assets) == sorted(fixture)
    for k in fixture:
        np.testing.assert_array_equal(assets[k], fixture[k], err_msg=k)
    m = decoder(assets)
    assert m.verts == [63, 126, 252] and m.vNum_all == 1008
    assert int(m._L_left_2_rowptr[-1]) == 1546            # nnz of the 252-vertex Laplacian
