"""The drop-in meets the REAL reference: the unmodified tree under /root/reference is
imported (import stubs only, oracle/ref_import.py), ``patch_reference()`` is applied and
``load_model_intag(opt)`` is built the way demo.py / main.py build it.  CPU only; skipped where the reference
tree is absent (the GPU box).  Runs in a subprocess: patching mutates the reference's modules."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("PDFNET_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib")), reason="reference tree absent")

SCRIPT = r'''
import importlib, json, sys, types, warnings
warnings.simplefilter("ignore")
sys.path.insert(0, %(root)r)
from oracle import ref_import as RI
RI.load_reference()
import torch
opts_mod = importlib.import_module("lib.opts")
# demo.sh / main.py flags that select this network (simplified task, depth branch, GCN decoder, MANO params head)
opt = opts_mod.opts().parse(["--task", "simplified", "--depth", "--gcn_decoder", "--reproj_loss"])
opt = opts_mod.opts.update_dataset_info_and_set_heads(opt, types.SimpleNamespace(mean=None, std=None, num_classes=1))
rmodel = importlib.import_module("lib.models.networks.intaghand_model")
renc = importlib.import_module("lib.models.networks.intaghand_encoder")
torch.manual_seed(317)
m0 = rmodel.load_model_intag(opt)
ref_classes = (renc.PointNet_Plus, renc.SFTLayer)
sd0 = {k: (list(v.shape), str(v.dtype)) for k, v in m0.state_dict().items()}
import pdfnet_b200
mode = %(mode)r
patched = pdfnet_b200.patch_reference(mode=mode)
torch.manual_seed(317)
m1 = rmodel.load_model_intag(opt)
sd1 = {k: (list(v.shape), str(v.dtype)) for k, v in m1.state_dict().items()}
out = {"patched": patched, "n_keys": len(sd0), "same_keys": sorted(sd0) == sorted(sd1),
       "same_shapes": sd0 == sd1, "missing": sorted(set(sd0) - set(sd1))[:5], "extra": sorted(set(sd1) - set(sd0))[:5]}
from pdfnet_b200 import decoder as D, encoder as E
out["pointnet_is_ours"] = type(m1.encoder.pointnet_plus) is E.PointNet_Plus and renc.PointNet_Plus is E.PointNet_Plus
out["sft_is_ours"] = type(m1.encoder.sft) is E.SFTLayer and type(m1.encoder.pointnet_plus.sft1) is E.SFTLayer
out["unpatched_was_reference"] = type(m0.encoder.pointnet_plus) is ref_classes[0] and type(m0.encoder.sft) is ref_classes[1]
out["decoder_is_ours"] = type(m1.decoder) is D.decoder
# checkpoints travel both ways, strictly
r = m1.load_state_dict(m0.state_dict(), strict=True)
out["load_ref_into_ours"] = not r.missing_keys and not r.unexpected_keys
r = m0.load_state_dict(m1.state_dict(), strict=True)
out["load_ours_into_ref"] = not r.missing_keys and not r.unexpected_keys
out["params"] = sum(p.numel() for p in m1.parameters())
out["same_param_names"] = [n for n, _ in m0.named_parameters()] == [n for n, _ in m1.named_parameters()]
# the wrapper and its forward signature are untouched
import inspect
out["forward_args"] = list(inspect.signature(m1.forward).parameters)
out["converter"] = sorted(m1.decoder.converter) == ["left", "right"] and hasattr(m1.decoder.converter["left"], "GCN_to_vert")
# no CPU fallback: the patched modules refuse host tensors loudly
try:
    with torch.no_grad():
        m1.encoder.pointnet_plus(torch.zeros(1, 1024, 3), [torch.zeros(1, 3, 8, 8)] * 3, torch.zeros(1, 1024, dtype=torch.long))
    out["cpu_refused"] = False
except RuntimeError as e:
    out["cpu_refused"] = "CUDA" in str(e)
print("RESULT" + json.dumps(out))
'''


def _run(mode):
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "mode": mode}], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])


def test_patched_real_reference_model_is_state_dict_compatible():
    r = _run("inference")
    assert len(r["patched"]) == 14
    assert r["n_keys"] > 1000 and r["same_keys"] and r["same_shapes"], (r["missing"], r["extra"])
    assert r["unpatched_was_reference"] and r["pointnet_is_ours"] and r["sft_is_ours"] and r["decoder_is_ours"]
    assert r["load_ref_into_ours"] and r["load_ours_into_ref"] and r["same_param_names"]
    assert r["params"] == 100985448
    assert r["forward_args"] == ["img", "choose", "cloud", "depth", "ind", "K_new", "valid"]
    assert r["converter"] and r["cpu_refused"]


def test_training_mode_patch_keeps_the_reference_decoder():
    r = _run("training")
    assert len(r["patched"]) == 10
    assert r["same_keys"] and r["same_shapes"] and r["pointnet_is_ours"] and r["sft_is_ours"]
    assert not r["decoder_is_ours"]                      # the reference's differentiable decoder stays
    assert r["load_ref_into_ours"] and r["load_ours_into_ref"]
