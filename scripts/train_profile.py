"""One cfg5 training step under `ncu --profile-from-start off` (kernel launch list of the step).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/train_launches.csv python scripts/train_profile.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pdfnet_b200 import HandFusion  # noqa: E402

R, B = 256, int(os.environ.get("PDF_TRAIN_FRAMES", 64))
dev = torch.device("cuda", 0)
model = HandFusion(bench.make_opt(R), precision=os.environ.get("PDF_TRAIN_PRECISION", "fp32"))
st = bench.load_states()
sd = {"pointnet_plus." + k: v for k, v in st["pointnet"].items()}
sd.update({"sft." + k: v for k, v in st["sft"].items()})
model.load_state_dict(sd, strict=False)
model = model.to(dev).train()
d = {k: v.to(dev) for k, v in bench.cfg5_inputs(B, R, 317).items()}


def step():
    model.zero_grad(set_to_none=True)
    fused = model(d["cloud"], [d["l0"], d["l1"], d["l2"]], d["choose"], d["center"])
    ((fused - d["target"]) ** 2).mean().backward()


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
