"""Device-resident timings of the BASELINE configs and their stages (CUDA events, torch
current stream).  Usage: python tools/microbench.py [names...]"""
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rocket_fft_b200 as R

PEAK = 6527.8
dev = torch.device("cuda:0")


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, ms, nbytes, flops=None, launches=None):
    gbs = nbytes / ms / 1e6
    s = f"{name:58s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}% of measured peak"
    if flops:
        s += f"  {flops / ms / 1e6:9.1f} GFLOP/s"
    if launches is not None:
        s += f"  launches/call={launches}"
    print(s, flush=True)


def run(name, fn, nbytes, flops=None, reps=10):
    R.launch_count_reset()
    ms = timeit(fn, reps)
    report(name, ms, nbytes, flops, R.launch_count() // (reps + 3))


def cfg1():
    x = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    y = torch.empty_like(x)
    run("cfg1 c2c c128 (4096,4096) axes=[1]", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 16, 5 * 4096 * 12 * 4096, 20)
    x = torch.randn(8192, 4096, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    run("     c2c c64 (8192,4096) axes=[1]", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 5 * 4096 * 12 * 8192, 20)
    ms = timeit(lambda: torch.fft.fft(x, dim=1), 20)
    report("     cuFFT (torch.fft.fft) c64 (8192,4096) [side ref]", ms, 2 * x.numel() * 8)
    x = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    ms = timeit(lambda: torch.fft.fft(x, dim=1), 20)
    report("     cuFFT (torch.fft.fft) c128 (4096,4096) [side ref]", ms, 2 * x.numel() * 16)
    y = torch.empty_like(x)
    ms = timeit(lambda: y.copy_(x), 20)
    report("     copy_ c128 (4096,4096) [bandwidth ref]", ms, 2 * x.numel() * 16)


def cfg2():
    x = torch.randn(16384, 16384, dtype=torch.float32, device=dev)
    X = torch.empty(16384, 8193, dtype=torch.complex64, device=dev)
    nb = x.numel() * 4 + X.numel() * 8
    run("cfg2 rfft2 f32 16384^2 (whole)", lambda: R.r2c(x, X, [0, 1], True, 1.0), nb, 2.5 * 2**28 * 28, 5)
    run("     rows: r2c axes=[1]", lambda: R.r2c(x, X, [1], True, 1.0), nb, None, 5)
    run("     cols: c2c axes=[0] in place", lambda: R.c2c(X, X, [0], True, 1.0), 2 * X.numel() * 8, None, 5)
    y = torch.empty_like(x)
    run("     irfft2 (c2r axes=[0,1])", lambda: R.c2r(X, y, [0, 1], False, 1.0), nb, None, 5)
    ms = timeit(lambda: torch.fft.rfft2(x), 5)
    report("     cuFFT rfft2 [side ref]", ms, nb)


def cfg3():
    n = 512 if len(sys.argv) > 2 and sys.argv[2] == "small" else 1024
    v = torch.randn(n, n, n, dtype=torch.complex64, device=dev)
    V = torch.empty_like(v)
    nb = 2 * v.numel() * 8
    run(f"cfg3 fftn c64 {n}^3 (whole)", lambda: R.c2c(v, V, [0, 1, 2], True, 1.0), nb, 5 * n**3 * 3 * math.log2(n), 3)
    for ax in (0, 1, 2):
        run(f"     axis {ax}", lambda: R.c2c(v, V, [ax], True, 1.0), nb, None, 3)
    run("     axes (1,2)", lambda: R.c2c(v, V, [1, 2], True, 1.0), nb, None, 3)
    run("     in place (whole)", lambda: R.c2c(V, V, [0, 1, 2], True, 1.0), nb, None, 3)
    # host time to enqueue the whole call (launch-bound if it approaches the device time)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R.c2c(v, V, [0, 1, 2], True, 1.0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"     host enqueue time of one call: {(t1 - t0) * 1e3:.3f} ms", flush=True)
    ms = timeit(lambda: torch.fft.fftn(v), 3)
    report("     cuFFT fftn [side ref]", ms, nb)
    del v, V


def batch2d():
    x = torch.randn(64, 2048, 2048, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    nb = 2 * x.numel() * 8
    run("fft2 c64 64 x (2048,2048)", lambda: R.c2c(x, y, [1, 2], True, 1.0), nb, None, 3)
    run("     in place", lambda: R.c2c(y, y, [1, 2], True, 1.0), nb, None, 3)
    ms = timeit(lambda: torch.fft.fft2(x), 3)
    report("     cuFFT fft2 [side ref]", ms, nb)
    del x, y
    x = torch.randn(256, 1024, 1024, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    nb = 2 * x.numel() * 8
    run("fft2 c64 256 x (1024,1024)", lambda: R.c2c(x, y, [1, 2], True, 1.0), nb, None, 3)
    ms = timeit(lambda: torch.fft.fft2(x), 3)
    report("     cuFFT fft2 [side ref]", ms, nb)


def cfg4():
    x = torch.randn(256, 15015, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    run("cfg4a c2c c64 (256,15015)", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 5 * 15015 * math.log2(15015) * 256, 20)
    x = torch.randn(4096, 15015, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    run("      c2c c64 (4096,15015)", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 5 * 15015 * math.log2(15015) * 4096, 5)
    x = torch.randn(256, 1000003, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    run("cfg4b c2c c64 (256,1000003) Bluestein", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 5 * 1000003 * math.log2(1000003) * 256, 3)


def cfg5():
    x = torch.randn(2048, 2048, 64, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    nb = 2 * x.numel() * 8
    run("cfg5 dct2 f64 (2048,2048,64) axes=(0,1)", lambda: R.dct(x, y, [0, 1], 2, 1.0, False), nb, None, 2)
    run("     dct2 axis 1 only", lambda: R.dct(x, y, [1], 2, 1.0, False), nb, None, 2)
    run("     dst2 axes=(0,1)", lambda: R.dst(x, y, [0, 1], 2, 1.0, False), nb, None, 2)


def sizes():
    for dt, esz in ((torch.complex64, 8), (torch.complex128, 16)):
        for logn in range(4, 15):
            n = 1 << logn
            if esz == 16 and logn > 13:
                continue
            rows = (1 << 27) // (n * esz) * 2
            x = torch.randn(rows, n, dtype=dt, device=dev)
            y = torch.empty_like(x)
            run(f"c2c {str(dt)[6:]} ({rows},{n})", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * esz, None, 10)
            del x, y


def nonpow2():
    for n in (96, 100, 243, 360, 1000, 1080, 1536, 1920, 3000, 4000, 5000, 6561, 10000, 12288):
        rows = max(4, (1 << 26) // (n * 8))
        x = torch.randn(rows, n, dtype=torch.complex64, device=dev)
        y = torch.empty_like(x)
        run(f"c2c complex64 ({rows},{n})", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, None, 10)
        ms = timeit(lambda: torch.fft.fft(x, dim=1), 10)
        report(f"   cuFFT [side ref]", ms, 2 * x.numel() * 8)
        del x, y
    x = torch.randn(64, 1080, 1920, dtype=torch.float32, device=dev)
    X = torch.empty(64, 1080, 961, dtype=torch.complex64, device=dev)
    nb = x.numel() * 4 + X.numel() * 8
    run("rfft2 f32 64 x (1080,1920)", lambda: R.r2c(x, X, [1, 2], True, 1.0), nb, None, 5)
    ms = timeit(lambda: torch.fft.rfft2(x), 5)
    report("   cuFFT rfft2 [side ref]", ms, nb)
    y = torch.empty_like(x)
    run("irfft2 f32 64 x (1080,1920)", lambda: R.c2r(X, y, [1, 2], False, 1.0), nb, None, 5)
    ms = timeit(lambda: torch.fft.irfft2(X, s=(1080, 1920)), 5)
    report("   cuFFT irfft2 [side ref]", ms, nb)


def rfftn():
    x = torch.randn(1024, 1024, 1024, dtype=torch.float32, device=dev)
    X = torch.empty(1024, 1024, 513, dtype=torch.complex64, device=dev)
    nb = x.numel() * 4 + X.numel() * 8
    run("rfftn f32 1024^3 (r2c axes=[0,1,2])", lambda: R.r2c(x, X, [0, 1, 2], True, 1.0), nb, None, 3)
    y = torch.empty_like(x)
    run("irfftn (c2r axes=[0,1,2])", lambda: R.c2r(X, y, [0, 1, 2], False, 1.0), nb, None, 3)
    ms = timeit(lambda: torch.fft.rfftn(x), 3)
    report("     cuFFT rfftn [side ref]", ms, nb)


ALL = {"rfftn": rfftn, "batch2d": batch2d, "nonpow2": nonpow2, "cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "sizes": sizes}
if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in ALL] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    print(R.version(), torch.cuda.get_device_name(0))
    for nme in names:
        ALL[nme]()
        torch.cuda.empty_cache()
