import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.set_grad_enabled(False)             # inference: the kernel path (a wanted gradient selects the torch autograd path)
from conftest import load_golden
from pdfnet_b200 import synth
from pdfnet_b200.decoder import decoder
B = int(os.environ.get("B", 128)); prec = os.environ.get("PREC", "bf16")
a = load_golden("gcn_assets")
fuse = torch.randn((B, 2, 1024), generator=torch.Generator().manual_seed(1)).cuda()
fl, fr = fuse[:, 0].contiguous(), fuse[:, 1].contiguous()
m = decoder(a, precision=prec); m.load_state_dict(synth.decoder_state(317, a["upsample"])); m = m.cuda().eval()
for _ in range(2): m(fl, fr, None)
torch.cuda.synchronize(); torch.cuda.profiler.start(); m(fl, fr, None); torch.cuda.synchronize(); torch.cuda.profiler.stop()
