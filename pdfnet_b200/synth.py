"""Deterministic synthetic weights and RGB-D inputs.

The reference ships no checkpoint and no dataset, so every test, golden vector
and benchmark runs on seeded synthetic tensors.  Everything here is a pure
function of (name, shape, seed) on the torch CPU generator, so the golden
generator (which loads these tensors into the UNMODIFIED reference modules),
the CPU oracle, and the CUDA path all see bit-identical parameters without
shipping multi-megabyte weight files.
"""
import math
import zlib

import numpy as np
import torch

POINTNET_CHANNELS = dict(
    netR_1=[(3, 64), (64, 64), (64, 128)],          # intaghand_encoder.py:27,48-65
    netR_2=[(131, 128), (128, 128), (128, 256)],    # :28,67-84
    netR_3=[(259, 512), (512, 512), (512, 1024)],   # :30,86-103
)
SFT_CHANNELS = dict(sft0=(3, 3), sft1=(131, 64), sft2=(259, 256))  # (c_fea, c_cond) :43-45


def _gen(name, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 31))
    return g


def _uniform(name, shape, bound, seed):
    return (torch.rand(shape, generator=_gen(name, seed)) * 2 - 1) * bound


def _normal(name, shape, std, seed):
    return torch.randn(shape, generator=_gen(name, seed)) * std


def sft_state(prefix, c_fea, c_cond, seed=317, out_gain=1.0):
    """Parameters of one SFTLayer (intaghand_encoder.py:205-212), conv weights [out,in,1,1]."""
    sd = {}
    for br in ("scale", "shift"):
        for i, (cin, cout) in enumerate(((c_cond, c_cond), (c_cond, c_fea))):
            b = 1.0 / math.sqrt(cin)
            gain = out_gain if i == 1 else 1.0
            k = "%sSFT_%s_conv%d" % (prefix, br, i)
            sd[k + ".weight"] = _uniform(k + ".weight", (cout, cin, 1, 1), b, seed) * gain
            sd[k + ".bias"] = _uniform(k + ".bias", (cout,), b, seed) * gain
    return sd


def _bn_state(k, c, seed):
    return {
        k + ".weight": torch.rand((c,), generator=_gen(k + ".weight", seed)) + 0.5,
        k + ".bias": _normal(k + ".bias", (c,), 0.1, seed),
        k + ".running_mean": _normal(k + ".running_mean", (c,), 0.1, seed),
        k + ".running_var": torch.rand((c,), generator=_gen(k + ".running_var", seed)) + 0.5,
        k + ".num_batches_tracked": torch.tensor(0, dtype=torch.long),
    }


def pointnet_plus_state(seed=317, sft0_gain=0.05):
    """State dict with exactly the reference's PointNet_Plus keys
    (netR_{1,2,3}.{0,1,3,4,6,7}.*, sft{0,1,2}.SFT_*; netR_FC is unused by
    forward, intaghand_encoder.py:156, and omitted).  BN running stats are
    perturbed so folding is exercised.  ``sft0_gain`` keeps the learned xyz
    modulation small enough that the radius mask keeps real neighbours."""
    sd = {}
    for name, (cf, cc) in SFT_CHANNELS.items():
        sd.update(sft_state(name + ".", cf, cc, seed, out_gain=sft0_gain if name == "sft0" else 0.3))
    for net, chans in POINTNET_CHANNELS.items():
        for li, (cin, cout) in enumerate(chans):
            b = math.sqrt(3.0 / cin)
            k = "%s.%d" % (net, 3 * li)
            sd[k + ".weight"] = _uniform(k + ".weight", (cout, cin, 1, 1), b, seed)
            sd[k + ".bias"] = _uniform(k + ".bias", (cout,), 0.1, seed)
            sd.update(_bn_state("%s.%d" % (net, 3 * li + 1), cout, seed))
    return sd


def fusion_sft_state(seed=317):
    """ResNetSimple.sft = SFTLayer(1024,1024), intaghand_encoder.py:673."""
    return sft_state("", 1024, 1024, seed, out_gain=0.3)


def mano_head_state(seed=317, std=1e-3):
    """mano_head (intaghand_encoder.py:630-643): Linear weights N(0,std) with zero
    bias as fill_fc_weights leaves them (:336-347); BN1d stats perturbed."""
    sd = {}
    for i, (cin, cout) in zip((0, 3, 6), ((1024, 512), (512, 256), (256, 122))):
        k = "mano_head.%d" % i
        sd[k + ".weight"] = _normal(k + ".weight", (cout, cin), std, seed)
        sd[k + ".bias"] = _normal(k + ".bias", (cout,), std, seed)
    for i, c in ((1, 512), (4, 256)):
        sd.update(_bn_state("mano_head.%d" % i, c, seed))
    return sd


def clouds(n_clouds, n_points=1024, seed=317, sigma=0.05, wrap_from=None):
    """cloud ~ N((0,0,0.5), sigma^2) metres.  ``wrap_from=m`` mimics np.pad(...,'wrap')
    of a hand with only m valid pixels (intaghand_encoder.py:424): rows repeat with
    period m and are then shuffled, which creates exact distance ties."""
    g = _gen("cloud", seed)
    c = torch.randn((n_clouds, n_points, 3), generator=g) * sigma
    c[..., 2] += 0.5
    if wrap_from is not None:
        base = c[:, :wrap_from]
        reps = (n_points + wrap_from - 1) // wrap_from
        c = base.repeat(1, reps, 1)[:, :n_points]
        perm = torch.stack([torch.randperm(n_points, generator=g) for _ in range(n_clouds)])
        c = torch.gather(c, 1, perm[..., None].expand(-1, -1, 3))
    return c.contiguous()


def choose_indices(n_clouds, R, n_points=1024, seed=317):
    """choose ~ U{1..R^2-1} int64 flat pixel indices."""
    return torch.randint(1, R * R, (n_clouds, n_points), generator=_gen("choose", seed), dtype=torch.long)


def pyramid(n_frames, R, seed=317, dtype=torch.float32):
    """Stand-in for the RGB neck outputs (backbone is out of scope):
    l0 [B,3,R,R] (relu(e_conv1)), l1 [B,64,R/2,R/2], l2 [B,256,R/4,R/4], all >= 0."""
    g = _gen("pyramid", seed)
    out = []
    for c, r in ((3, R), (64, R // 2), (256, R // 4)):
        out.append(torch.relu(torch.randn((n_frames, c, r, r), generator=g)).to(dtype))
    return out


def train_inputs(B=2, R=64, seed=41):
    """Inputs of the training-step golden (tests/golden/train_step.npz): clouds, choose,
    pyramid maps and the direction ``gdir`` of the scalar loss sum(out * gdir)."""
    gdir = torch.randn((B, 1, 1024), generator=_gen("gdir", seed))
    return clouds(B, seed=seed), choose_indices(B, R, seed=seed), pyramid(B, R, seed=seed), gdir


def rgbd_frames(n_frames, R, seed=317):
    """Synthetic depth [B,R,R] f32 metres (0.5 m + 2 cm noise inside two disjoint hand
    rectangles, 0 elsewhere), masks [B,2,R,R] f32 (channel 0 = right, 1 = left as
    depth2pcl reads them, intaghand_encoder.py:376-377), K [B,3,3] f32, valid [B,2]."""
    g = _gen("rgbd", seed)
    depth = torch.zeros((n_frames, R, R))
    mask = torch.zeros((n_frames, 2, R, R))
    h0, h1 = R // 4, R // 4 + R // 3
    boxes = ((R // 16, R // 16 + R // 3), (R // 2 + R // 16, R // 2 + R // 16 + R // 3))
    for ch, (w0, w1) in enumerate(boxes):
        mask[:, ch, h0:h1, w0:w1] = 1.0
        z = 0.5 + 0.05 * ch + 0.02 * torch.rand((n_frames, h1 - h0, w1 - w0), generator=g)
        depth[:, h0:h1, w0:w1] = z
    f = 300.0 * R / 384.0
    K = torch.tensor([[f, 0.0, R / 2.0], [0.0, f, R / 2.0], [0.0, 0.0, 1.0]]).repeat(n_frames, 1, 1)
    valid = torch.ones((n_frames, 2))
    return depth, mask, K, valid


def mano_inputs(n, seed=317):
    """rot~N(0,0.5^2), pose~N(0,0.3^2), shape~N(0,0.5^2), trans~N(0,0.1^2)."""
    g = _gen("mano", seed)
    return (torch.randn((n, 3), generator=g) * 0.5, torch.randn((n, 45), generator=g) * 0.3,
            torch.randn((n, 10), generator=g) * 0.5, torch.randn((n, 3), generator=g) * 0.1)


def synthetic_mano_tables(seed=317):
    """MANO-shaped random tables (same shapes / sparsity pattern class as the real
    model) for size-independent tests that must not depend on the licensed pkl."""
    g = _gen("mano_tables", seed)
    w = torch.rand((778, 16), generator=g) ** 8
    w = w / w.sum(1, keepdim=True)
    jr = torch.rand((16, 778), generator=g) ** 16
    jr = jr / jr.sum(1, keepdim=True)
    return dict(
        v_template=(torch.rand((778, 3), generator=g) - 0.5) * torch.tensor([0.19, 0.06, 0.17]),
        shapedirs=torch.randn((778, 3, 10), generator=g) * 2e-3,
        posedirs=torch.randn((778, 3, 135), generator=g) * 1e-3,
        J_regressor=jr, weights=w,
    )


def to_numpy(d):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


# ------------------------------------------------------------------------------------------------
# GCN decoder: deterministic stand-ins for a trained checkpoint
GCN_IN_DIM, GCN_OUT_DIM, GCN_VERTS = (512, 256, 128), (256, 128, 64), (63, 126, 252)


def decoder_param_shapes(gf_dim=1024, graph_k=2, n_blocks=4):
    """(name, shape) of every parameter decoder.forward uses (intaghand_decoder.py:74-147; the img_ex_*
    sub-modules exist in the reference state dict but their call is commented out, DualGraph.py:84-85)."""
    out = []
    lin = lambda n, o, i, bias=True: out.extend([(n + ".weight", (o, i))] + ([(n + ".bias", (o,))] if bias else []))
    ln = lambda n, c: out.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])
    for li, (cin, cout, V) in enumerate(zip(GCN_IN_DIM, GCN_OUT_DIM, GCN_VERTS)):
        base = "dual_gcn.layers.%d." % li
        out.append((base + "position_embeddings.weight", (V, cin)))
        for side in ("graph_left", "graph_right"):
            for b in range(n_blocks):
                p = "%s%s.GCN_blocks.%d." % (base, side, b)
                ci = cin if b == 0 else cout
                ln(p + "norm1", ci)
                lin(p + "fc1", cout, ci * graph_k)
                ln(p + "norm2", cout)
                lin(p + "fc2", cout, cout * graph_k)
                lin(p + "shortcut", cout, ci)
                ln(p + "norm3", cout)
        a = base + "attn."
        for sa in ("L_self_attn_layer.", "R_self_attn_layer."):
            for w in ("w_qs", "w_ks", "w_vs"):
                lin(a + sa + w, cout, cout)
            ln(a + sa + "layer_norm", cout)
            lin(a + sa + "fc", cout, cout)
            ln(a + sa + "ff.layer_norm", cout)
            lin(a + sa + "ff.fc1", cout, cout)
            lin(a + sa + "ff.fc2", cout, cout)
        for w in ("w_qs", "w_ks", "w_vs", "fc"):
            lin(a + w, cout, cout)
        ln(a + "layer_norm1", cout)
        ln(a + "layer_norm2", cout)
        for ff in ("ffL.", "ffR."):
            ln(a + ff + "layer_norm", cout)
            lin(a + ff + "fc1", cout, cout)
            lin(a + ff + "fc2", cout, cout)
    for side in ("left", "right"):
        lin("gf_layer_%s.0" % side, GCN_IN_DIM[0] - 3, gf_dim)
        ln("gf_layer_%s.1" % side, GCN_IN_DIM[0] - 3)
    lin("unsample_layer", 778, GCN_VERTS[-1], bias=False)
    lin("coord_head", 3, GCN_OUT_DIM[-1])
    lin("avg_head", 1, GCN_VERTS[-1])
    lin("params_head", 3, GCN_OUT_DIM[-1])
    lin("root_head", 3, GCN_OUT_DIM[-1])
    return out


def decoder_state(seed=317, upsample_weight=None):
    """Synthetic decoder weights with the reference's state-dict keys: Xavier-uniform Linear weights
    (gcn.py:8-16), small non-zero biases, LayerNorm gains around 1, N(0,1) position embeddings."""
    sd = {}
    for name, shape in decoder_param_shapes():
        if name == "unsample_layer.weight" and upsample_weight is not None:
            sd[name] = torch.as_tensor(upsample_weight, dtype=torch.float32).clone()
        elif name.endswith("position_embeddings.weight"):
            sd[name] = _normal(name, shape, 1.0, seed)
        elif len(shape) == 2:
            sd[name] = _uniform(name, shape, math.sqrt(6.0 / (shape[0] + shape[1])), seed)
        elif "norm" in name or (name.startswith("gf_layer") and name.split(".")[1] == "1"):
            sd[name] = (1.0 + _uniform(name, shape, 0.1, seed)) if name.endswith("weight") else _uniform(name, shape, 0.1, seed)
        else:
            sd[name] = _uniform(name, shape, 0.05, seed)
    return sd
