"""Quick timing of the GCN decoder (eager vs CUDA-graph replay) at B frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.set_grad_enabled(False)             # inference: the kernel path (a wanted gradient selects the torch autograd path)
from conftest import load_golden
from pdfnet_b200 import synth, _lib
from pdfnet_b200.decoder import decoder
from pdfnet_b200.graph import CapturedStep
B = int(os.environ.get("B", 128))
a = load_golden("gcn_assets")
fuse = torch.randn((B, 2, 1024), generator=torch.Generator().manual_seed(1)).cuda()
fl, fr = fuse[:, 0].contiguous(), fuse[:, 1].contiguous()
for prec in ("fp32", "bf16x3"):
    m = decoder(a, precision=prec); m.load_state_dict(synth.decoder_state(317, a["upsample"])); m = m.cuda().eval()
    fn = lambda: m(fl, fr, None)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    n0 = _lib.launch_count(); fn(); nl = _lib.launch_count() - n0
    def timeit(f, n=20):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    t_e = timeit(fn)
    step = CapturedStep(fn)
    t_g = timeit(step.replay)
    print("%-7s B=%d launches=%d eager %.3f ms graph %.3f ms -> %.0f frames/s" % (prec, B, nl, t_e, t_g, B / t_g * 1e3))
