import os, sys
sys.path.insert(0, '/root/repo')
import torch
import rocket_fft_b200 as R
dev = torch.device('cuda:0')
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n in (1000, 1920):
    rows = (1 << 26) // (n * 8)
    x = torch.randn(rows, n, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    R.c2c(x, y, [1], True, 1.0)
    ref = torch.fft.fft(x, dim=1)
    err = float(torch.linalg.vector_norm((y - ref).to(torch.complex128)) / torch.linalg.vector_norm(ref.to(torch.complex128)))
    ms = timeit(lambda: R.c2c(x, y, [1], True, 1.0))
    print(f"n={n} spec={os.environ.get('RFB200_SPEC_TEST')} err={err:.2e} {ms:.4f} ms {2*x.numel()*8/ms/1e6/6527.8*100:.1f}% of peak")
