"""Cosine similarity of every bf16-mode training gradient with the fp64 reference golden (diagnostic)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden
from pdfnet_b200 import PointNet_Plus, synth
g = load_golden("train_step")
opt = types.SimpleNamespace(SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512, sample_num_level2=128,
                            ball_radius=0.015, ball_radius2=0.04, default_resolution=64, PCA_SZ=63)
for prec in ("fp32", "bf16"):
    m = PointNet_Plus(opt, prec); m.load_state_dict(synth.pointnet_plus_state(seed=317), strict=False); m = m.cuda().train()
    pts, choose, emb, gdir = synth.train_inputs(2, 64)
    emb = [e.cuda().requires_grad_(True) for e in emb]
    out = m(pts.cuda(), emb, choose.cuda()); (out * gdir.cuda()).sum().backward()
    res = []
    for k, p in list(m.named_parameters()) + [("emb%d" % i, e) for i, e in enumerate(emb)]:
        if k.startswith("netR_FC"): continue
        got = p.grad.double().reshape(-1).cpu().numpy(); name = "grad:" + k
        ref, got = (g[name].reshape(-1), got) if name in g else (g[name + "@s97"], got[::97])
        if np.abs(ref).max() < 1e-6: continue
        res.append((float(np.dot(got, ref) / (np.linalg.norm(got) * np.linalg.norm(ref))), k))
    res.sort()
    print(prec, "lowest cosines:", [(round(c, 4), k) for c, k in res[:14]])
