"""bf16-mode training step against the fp32-accurate mode on a 64-cloud batch: output error and the
cosine of every gradient tensor (diagnostic behind tests::test_train_step_bf16_mode_vs_fp32_mode_64_clouds)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pdfnet_b200 import PointNet_Plus, synth
B, R = 64, 64
opt = types.SimpleNamespace(SAMPLE_NUM=1024, INPUT_FEATURE_NUM=3, knn_K=64, sample_num_level1=512, sample_num_level2=128,
                            ball_radius=0.015, ball_radius2=0.04, default_resolution=R, PCA_SZ=63)
pts, choose, emb0, gdir = synth.clouds(B, seed=41), synth.choose_indices(B, R, seed=41), synth.pyramid(B, R, seed=41), \
    torch.randn((B, 1, 1024), generator=torch.Generator().manual_seed(41))
grads, outs = {}, {}
for prec in ("fp32", "bf16"):
    m = PointNet_Plus(opt, prec); m.load_state_dict(synth.pointnet_plus_state(seed=317), strict=False); m = m.cuda().train()
    emb = [e.cuda().requires_grad_(True) for e in emb0]
    out = m(pts.cuda(), emb, choose.cuda()); (out * gdir.cuda()).sum().backward()
    outs[prec] = out.detach().double().cpu()
    grads[prec] = {k: p.grad.double().reshape(-1).cpu() for k, p in list(m.named_parameters()) + [("emb%d" % i, e) for i, e in enumerate(emb)]
                   if not k.startswith("netR_FC") and p.grad is not None}
print("out rel err", float((outs["bf16"] - outs["fp32"]).abs().max() / outs["fp32"].abs().max()))
res = []
for k, a in grads["fp32"].items():
    b = grads["bf16"][k]
    if float(a.abs().max()) < 1e-6: continue
    res.append((float(torch.dot(a, b) / (a.norm() * b.norm())), float((a - b).norm() / a.norm()), k))
res.sort()
for c, e, k in res[:20]: print("%.5f  relL2 %.4f  %s" % (c, e, k))
print("n", len(res), "min cos", res[0][0], "median", res[len(res)//2][0])
a = torch.cat([grads["fp32"][k] for _, _, k in res]); b = torch.cat([grads["bf16"][k] for _, _, k in res])
print("global cos", float(torch.dot(a, b) / (a.norm() * b.norm())), "relL2", float((a - b).norm() / a.norm()))
big = [(float(grads["fp32"][k].norm()), c, k) for c, _, k in res]
big.sort(reverse=True)
print("largest-norm tensors:", [(round(n_, 3), round(c, 4), k) for n_, c, k in big[:8]])
print("cos >= 0.99:", sum(c >= 0.99 for c, _, _ in res), "cos >= 0.9:", sum(c >= 0.9 for c, _, _ in res), "of", len(res))
