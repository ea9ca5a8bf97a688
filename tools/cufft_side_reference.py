import torch, sys
sys.path.insert(0, '/root/repo')
dev = torch.device('cuda:0')
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
x = torch.randn(16384, 16384, dtype=torch.float32, device=dev)
nb = x.numel()*4 + 16384*8193*8
ms = timeit(lambda: torch.fft.rfft(x, dim=1))
print(f"cuFFT rfft rows 16384x16384: {ms:.4f} ms  {nb/ms/1e6:.0f} GB/s  {nb/ms/1e6/6527.8*100:.1f}%")
X = torch.randn(16384, 8193, dtype=torch.complex64, device=dev)
ms = timeit(lambda: torch.fft.fft(X, dim=0))
print(f"cuFFT fft cols (16384 x 8193, dim 0): {ms:.4f} ms  {2*X.numel()*8/ms/1e6:.0f} GB/s")
for n in (8192, 16384):
    z = torch.randn((1<<27)//(n*8)*2, n, dtype=torch.complex64, device=dev)
    ms = timeit(lambda: torch.fft.fft(z, dim=1))
    print(f"cuFFT c2c c64 n={n}: {ms:.4f} ms {2*z.numel()*8/ms/1e6/6527.8*100:.1f}%")
z = torch.randn(2048, 8192, dtype=torch.complex128, device=dev)
ms = timeit(lambda: torch.fft.fft(z, dim=1))
print(f"cuFFT c2c c128 n=8192: {ms:.4f} ms {2*z.numel()*16/ms/1e6/6527.8*100:.1f}%")
v = torch.randn(512, 1024, 1024, dtype=torch.complex64, device=dev)
for d in (0, 1, 2):
    ms = timeit(lambda: torch.fft.fft(v, dim=d), 3)
    print(f"cuFFT fft 512x1024x1024 dim {d}: {ms:.4f} ms {2*v.numel()*8/ms/1e6/6527.8*100:.1f}%")
import scipy  # noqa
xd = torch.randn(2048, 2048, 64, dtype=torch.float64, device=dev)
