"""What does ONE kernel of the decoder chain cost inside a CUDA graph?  A serial chain of N identical launches is
captured and replayed; per-launch time = replay time / N.  Shapes: the three DualGraph levels at 128 frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pdfnet_b200 import ops, _lib as L
from pdfnet_b200.graph import CapturedStep

dev = torch.device("cuda")
N = 24


def timeit(fn, reps=10):
    step = CapturedStep(fn, warmup=2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        step.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / N * 1e3          # us per launch


def gemm_case(M, Nout, K, light=True, out_image=False):
    x = torch.randn((M, K), device=dev)
    w = torch.randn((Nout, K), device=dev) * 0.05
    b = torch.randn((Nout,), device=dev)
    packed = ops.pack_linear_tc(w, b, split=True)
    img = ops.rows_to_image(x, 0, K, split=1)

    def chain():
        for _ in range(N):
            ops.linear_tc(None, None, None, packed=packed, x_img=img, M=M, light=light, out_image=out_image)
    return timeit(chain)


print("GEMM (split-bf16, fp32 rows out unless img): us per launch inside a graph chain")
for (M, Nout, K) in [(8064, 768, 512), (8064, 768, 256), (8064, 512, 256), (8064, 256, 256), (8064, 256, 64),
                     (16128, 384, 128), (16128, 256, 128), (16128, 128, 128),
                     (32256, 192, 64), (32256, 128, 64), (32256, 64, 64), (1024, 128, 64)]:
    flop = 2.0 * M * Nout * K * 3
    for light in (True, False):
        t = gemm_case(M, Nout, K, light=light)
        print("  M=%5d N=%3d K=%3d %s: %6.2f us  (%5.0f TFLOP/s incl. x3)" % (M, Nout, K, "light" if light else "full ", t, flop / t / 1e6))
    t = gemm_case(M, Nout, K, light=True, out_image=True) if Nout % 64 == 0 else float("nan")
    print("                      image out: %6.2f us" % t)

# an (almost) empty kernel chain: the launch floor inside a graph
z = torch.zeros((256,), device=dev)
def empty_chain():
    for _ in range(N):
        ops.row_combine(z.view(1, 256), None, want_sum=True)
print("row_combine on 1 row (launch floor): %.2f us" % timeit(empty_chain))
for (M, C) in [(8064, 256), (16128, 128), (32256, 64)]:
    a = torch.randn((M, C), device=dev); g = torch.ones((C,), device=dev); be = torch.zeros((C,), device=dev)
    def rc_chain():
        for _ in range(N):
            ops.row_combine(a, a, ln=(g, be), want_sum=True, ln_rows=False, ln_img=True)
    print("row_combine+LN->image M=%d C=%d: %.2f us" % (M, C, timeit(rc_chain)))
