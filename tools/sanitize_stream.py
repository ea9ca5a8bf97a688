"""compute-sanitizer target for the two warp-specialised kernels (named barriers, mbarriers, TMA): the streamed strided-lines
kernel (pow2_stream_kernel.cuh) and the fused four-step kernel (fused4v2_kernel.cuh) on the smallest arrays they take.
    compute-sanitizer --tool racecheck python tools/sanitize_stream.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "stream"
if which == "stream":
    for shape in ((2, 1024, 4808), (5, 512, 1936)):
        x = torch.randn(*shape, dtype=torch.complex64, device=dev)
        want = torch.fft.fft(x, dim=1)
        R.launch_trace(True)
        R.c2c(x, x, [1], True, 1.0)
        torch.cuda.synchronize()
        print(R.launch_trace_get(), float((x - want).abs().max() / want.abs().max()))
else:
    x = torch.randn(16384, 520, dtype=torch.complex64, device=dev)
    want = torch.fft.fft(x, dim=0)
    R.launch_trace(True)
    R.c2c(x, x, [0], True, 1.0)
    torch.cuda.synchronize()
    print(R.launch_trace_get(), float((x - want).abs().max() / want.abs().max()))
print("done")
