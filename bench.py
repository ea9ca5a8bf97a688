#!/usr/bin/env python
"""Benchmark of the transform path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): np.fft.rfft2 of float32 16384x16384 images through
the low-level r2c(axes=[1,2]) call, IMAGES images per GPU, batch-sharded over N GPUs with no
data-path collective (weak scaling).  A step = one rfft2 pass over the rank's images.

  value   : GFLOP/s over all ranks, inputs resident in HBM, CUDA-event timed (max over ranks)
  e2e     : same metric through the numba_r2c C-ABI entry point with pinned HOST buffers
            (H2D + kernels + D2H inside the timed region)
  roofline: dominant kernel's algorithmic bytes / its event-timed duration vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference's own PocketFFT path (oracle/_ref, built from
            /root/reference by oracle/Makefile) with all host threads on the same workload.

flops convention: 2.5*N*log2(N) per real N-point image (the standard real-FFT count; the
BASELINE label "5N*log2N" applied verbatim to real data is exactly 2x this -- both arms use
the same formula so ratios are unaffected).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 16384
WORKLOAD = "cfg2: rfft2 float32 16384x16384 (r2c axes=[1,2]), batch-sharded images"


def flops_per_image():
    n = H * W
    return 2.5 * n * math.log2(n)


def alg_bytes_per_image():
    return H * W * 4 + H * (W // 2 + 1) * 8


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_rfft2(steps, warmup, sample_rows=None):
    """Times the reference's own CPU implementation (oracle/_ref) -- or, if that library did
    not travel, the NumPy oracle -- on the same workload with all host threads."""
    import numpy as np

    cores = os.cpu_count() or 1
    so = os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so")
    rng = np.random.default_rng(1)
    rows = sample_rows or H
    x = rng.standard_normal((rows, W), dtype=np.float32)
    out = np.empty((rows, W // 2 + 1), dtype=np.complex64)
    if os.path.exists(so):
        from rocket_fft_b200._abi import LowLevelLib

        ref = LowLevelLib(so)
        kind = "reference"

        def run():
            ref.r2c(x, out, [0, 1], True, 1.0, cores)
    else:
        from oracle import pocketfft_oracle as O

        kind = "port"
        cores = 1

        def run():
            O.r2c(x, out, [0, 1], True, 1.0)
    for _ in range(max(1, warmup)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    n = rows * W
    fl = 2.5 * n * math.log2(n)
    return {"value": fl / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": f"rfft2 of one float32 {rows}x{W} image via r2c(axes=[0,1]), nthreads={cores}, mean of {steps}",
            "ms_per_image": dt * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=4, help="images per GPU per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    config = {"workload": WORKLOAD, "images_per_gpu": args.images, "image": [H, W], "axes": [1, 2],
              "flops_per_image": flops_per_image(), "algorithmic_bytes_per_image": alg_bytes_per_image(),
              "l2_policy": "inputs (1 GiB per image) are far larger than the 126 MB L2; no flush needed",
              "parallelism": f"batch-sharded x{max(world, args.gpus)} (no collective)"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 5)
        r = cpu_reference_rfft2(steps, 1)
        line = {"impl": "reference", "metric": "rfft2 GFLOP/s (2.5*N*log2N per real image)", "value": r["value"],
                "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": r["ms_per_image"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, images_per_gpu=1),
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import rocket_fft_b200 as R

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.images
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(B, H, W, dtype=torch.float32, device=dev, generator=g)
    X = torch.empty(B, H, W // 2 + 1, dtype=torch.complex64, device=dev)

    def step():
        R.r2c(x, X, [1, 2], True, 1.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    R.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = R.launch_count()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B * flops_per_image() / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage timing (rank 0): which kernel dominates, and its roofline --------------------
    stages = []
    roofline = None
    if rank == 0:
        def timeit(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        reps = max(5, min(args.steps, 20))
        R.launch_count_reset()
        t_row = timeit(lambda: R.r2c(x, X, [2], True, 1.0), reps)
        row_launches = R.launch_count() // (reps + 3)
        R.launch_count_reset()
        t_col = timeit(lambda: R.c2c(X, X, [1], True, 1.0), reps)
        col_launches = R.launch_count() // (reps + 3)
        row_bytes = B * alg_bytes_per_image()
        col_bytes = B * 2 * H * (W // 2 + 1) * 8
        stages = [
            {"stage": "r2c rows (n=16384 real -> 8193 complex)", "ms": t_row, "launches": row_launches,
             "algorithmic_GBps": row_bytes / t_row / 1e6},
            {"stage": "c2c columns (n=16384, stride 65544 B, in place)", "ms": t_col, "launches": col_launches,
             "algorithmic_GBps": col_bytes / t_col / 1e6},
        ]
        peak, peak_src = load_peaks()
        # dominant kernel: the stage with the larger per-launch time
        dual = int(os.environ.get("RFB200_DUAL", "2"))
        row_kernel = ("fft_pow2_dual_kernel<12,1,%s> r2c rows" % ("true" if dual == 2 else "false")) if dual else \
            "fft_pow2_kernel<float,13,1,1> r2c rows"
        per_launch = [(t_row / max(row_launches, 1), row_bytes, row_kernel, row_launches),
                      (t_col / max(col_launches, 1), col_bytes,
                       "fft_fourstep_fused_kernel<float,7,32> columns (both four-step passes, intermediate in L2)" if col_launches == 1
                       else ("fft_pow2_pair_kernel<7,32> four-step column pass" if int(os.environ.get("RFB200_PAIR", "1"))
                             else "fft_pow2_kernel<float,7,32,0> four-step column pass"), col_launches)]
        dom = max(per_launch, key=lambda p: p[0])
        achieved = dom[1] / (dom[0] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of that kernel for ONE image, from the
        # committed `ncu --set full` capture (profiles/r01_traffic.json names the report); x images per launch
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                key = "rows" if dom[2].endswith("r2c rows") else ("cols_fused" if "fused" in dom[2] else "cols_pass")
                traffic = tj[key]["dram_bytes_per_image"] * B
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom[2], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[0],
                    "whole_step_frac_of_compulsory": (B * alg_bytes_per_image() / (ms_per_step * 1e-3) / 1e9) / peak}

    # ---- end to end through the C ABI with pinned host buffers (rank-local, max over ranks) -----
    e2e = None
    if not args.no_e2e:
        # the same step (B images) through the C ABI from pinned host memory; the library pipelines the
        # images (H2D of one overlaps kernels / D2H of the previous one)
        hx = torch.empty(B, H, W, dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hX = torch.empty(B, H, W // 2 + 1, dtype=torch.complex64, pin_memory=True)
        nx, nX = hx.numpy(), hX.numpy()
        for _ in range(2):
            R.r2c(nx, nX, [1, 2], True, 1.0)
        barrier()
        k = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            R.r2c(nx, nX, [1, 2], True, 1.0)
        dt = (time.perf_counter() - t0) / k
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * flops_per_image() / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": B * H * W * 4,
               "d2h_bytes_per_step": B * H * (W // 2 + 1) * 8, "ms_per_step": dt * 1e3,
               "call": f"numba_r2c via rocket_fft_b200.r2c(numpy pinned in/out), {B} images per step per rank, "
                       "image-pipelined H2D / kernels / D2H inside the call"}
        # cheap parity spot check of the e2e result against the device-resident result
        step()
        torch.cuda.synchronize()
        chk = float(torch.linalg.vector_norm(torch.view_as_real(hX[:, :4].to(dev) - X[:, :4])) /
                    torch.linalg.vector_norm(torch.view_as_real(X[:, :4])))
        e2e["matches_device_path_rel_l2"] = chk

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference_rfft2(3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": "rfft2 GFLOP/s (2.5*N*log2N per real image)", "value": value, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            # the BASELINE label's "5N*log2N" applied verbatim to real data is exactly twice `value`
            "value_5NlogN_label": 2.0 * value,
            "clocks": clocks, "stages": stages, "library": R.version()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
