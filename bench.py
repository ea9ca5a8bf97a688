#!/usr/bin/env python
"""Benchmark of the transform path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): np.fft.rfft2 of float32 16384x16384 images through the low-level
r2c(axes=[1,2]) call, --images images per GPU, batch-sharded over N GPUs with no data-path collective
(weak scaling).  A step = one rfft2 pass over the rank's images.

  value        GFLOP/s over all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline     the step's longest kernel (named from the library's launch trace): algorithmic bytes / its
               event-timed duration against MEASURED_PEAKS.json
  parity       the device result of one image against the reference's CPU result of the same input
  e2e          the same step through the numba_r2c C-ABI entry point with pinned HOST buffers (H2D + kernels +
               D2H inside the timed region), with the copies alone beside it (h2d_ms, d2h_ms, copy_only_ms)
               and the same call on plain pageable NumPy arrays (e2e_pageable)
  fftn         BASELINE.json configs[2]: fftn of a complex64 1024^3 volume on the same N GPUs (N = 1: one
               c2c call; N > 1: slab decomposition with the FFT + NVLink all-to-all fused in one kernel),
               total ms, the exchange alone, its bus bandwidth, parity against the reference at 512^3
  other_configs  (N = 1) the remaining BASELINE configs, device-resident, one line each
  cpu_baseline / --impl reference: the reference's own PocketFFT path (oracle/_ref, the UNMODIFIED reference
               compiled by oracle/Makefile) with all host threads on the same workload.  The reference arm
               imports nothing of the product: the only shared library it maps is oracle/_ref's.

flops convention: 2.5*N*log2(N) per real N-point image (the standard real-FFT count; the BASELINE label
"5N*log2N" applied verbatim to real data is exactly 2x this -- both arms use the same formula).
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
METRIC = "rfft2 GFLOP/s (2.5*N*log2N per real image)"


def flops_per_image():
    n = H * W
    return 2.5 * n * math.log2(n)


def alg_bytes_per_image():
    return H * W * 4 + H * (W // 2 + 1) * 8


def make_config(args, world):
    return {"workload": WORKLOAD, "images_per_gpu": args.images, "image": [H, W], "axes": [1, 2],
            "flops_per_image": flops_per_image(), "algorithmic_bytes_per_image": alg_bytes_per_image(),
            "l2_policy": "inputs (1 GiB per image) are far larger than the 126 MB L2; no flush needed",
            "parallelism": f"batch-sharded x{world} (no collective)"}


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


# ---------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref; the NumPy restatement if that library did not travel)
# ---------------------------------------------------------------------------------------------------
def reference_lib():
    """(callable r2c(x, out, axes, forward, fct, nthreads), c2c likewise, kind, cores) -- no product imports."""
    so = os.path.join(ROOT, "oracle", "_ref", "libpocketfft_ref.so")
    cores = os.cpu_count() or 1
    if os.path.exists(so):
        from oracle.abi_view import RefLib

        return RefLib(so), "reference", cores
    from oracle import pocketfft_oracle as O

    class Port:
        def r2c(self, a, b, axes, fwd, fct, nthreads=1):
            return O.r2c(a, b, axes, fwd, fct)

        def c2c(self, a, b, axes, fwd, fct, nthreads=1):
            return O.c2c(a, b, axes, fwd, fct)

    return Port(), "port", 1


def cpu_reference_rfft2(images, steps, warmup):
    """Times the reference on `images` float32 16384x16384 images per step (r2c axes=[1,2], nthreads = all cores)."""
    import numpy as np

    ref, kind, cores = reference_lib()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((images, H, W), dtype=np.float32)
    out = np.empty((images, H, W // 2 + 1), dtype=np.complex64)
    for _ in range(max(1, warmup)):
        ref.r2c(x, out, [1, 2], True, 1.0, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.r2c(x, out, [1, 2], True, 1.0, cores)
    dt = (time.perf_counter() - t0) / steps
    return {"value": images * flops_per_image() / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": f"rfft2 of {images} float32 {H}x{W} image(s) per step via r2c(axes=[1,2]), nthreads={cores}, "
                      f"mean of {steps} steps after {max(1, warmup)} warm-up",
            "ms_per_step": dt * 1e3}


def run_reference_arm(args, world):
    # a step is bounded: at most `images` images whatever N is (the host cores do not grow with the GPU count)
    r = cpu_reference_rfft2(args.images, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(args, world),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-rank runs: keep this rank's threads (and therefore the first touch of its pinned host buffers) on the NUMA node
    its GPU hangs off, so that eight ranks do not push their H2D / D2H traffic through one node's memory.  Best effort:
    returns what was found for the JSON line."""
    info = {"numa_node": None, "cpus_bound": None}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        info["cpus_allowed"] = len(allowed)
        mine = cpus & allowed
        if mine and mine != allowed:
            os.sched_setaffinity(0, mine)
            info["cpus_bound"] = len(mine)
    except Exception as e:  # pragma: no cover - depends on the box
        info["error"] = repr(e)[:80]
    return info


def rel_l2(a, b):
    import numpy as np

    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=4, help="images per GPU per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fftn", action="store_true")
    ap.add_argument("--no-other", action="store_true")
    ap.add_argument("--fftn-n", type=int, default=1024)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, max(world, args.gpus))
        return

    warmup = max(args.warmup, 3)
    config = make_config(args, world)

    import numpy as np
    import torch
    import torch.distributed as dist

    import rocket_fft_b200 as R

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
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

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

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
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B * flops_per_image() / (ms_per_step * 1e-3) / 1e9

    def timeit(fn, reps, warm=3):
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

    # ---- per-stage timing (rank 0): which kernel dominates, and its roofline ----------------------------
    stages, roofline, parity = [], None, None
    if rank == 0:
        reps = max(5, min(args.steps, 20))

        def traced(fn):
            R.launch_trace(True)
            fn()
            torch.cuda.synchronize()
            names = R.launch_trace_get()
            R.launch_trace(False)
            return names

        row_names = traced(lambda: R.r2c(x, X, [2], True, 1.0))
        t_row = timeit(lambda: R.r2c(x, X, [2], True, 1.0), reps)
        col_names = traced(lambda: R.c2c(X, X, [1], True, 1.0))
        t_col = timeit(lambda: R.c2c(X, X, [1], True, 1.0), reps)
        step_names = traced(step)
        row_bytes = B * alg_bytes_per_image()
        col_bytes = B * 2 * H * (W // 2 + 1) * 8
        stages = [
            {"stage": "r2c rows (n=16384 real -> 8193 complex)", "ms": t_row, "launches": len(row_names), "kernels": row_names,
             "algorithmic_GBps": row_bytes / t_row / 1e6},
            {"stage": "c2c columns (n=16384, stride 65544 B, in place)", "ms": t_col, "launches": len(col_names),
             "kernels": col_names, "algorithmic_GBps": col_bytes / t_col / 1e6},
        ]
        peak, peak_src = load_peaks()
        # the dominant kernel = the stage with the largest time per launch (a stage's launches are the same kernel or,
        # for the two-launch four-step, two instances of it)
        per_launch = [(t_row / max(len(row_names), 1), row_bytes, row_names), (t_col / max(len(col_names), 1), col_bytes, col_names)]
        dom = max(per_launch, key=lambda p: p[0])
        achieved = dom[1] / (dom[0] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of that kernel for ONE image, from the committed `ncu --set full`
        # capture (profiles/r02_traffic.json names the report); x images per launch
        traffic = None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                try:
                    tj = json.load(open(tpath))
                    ent = tj.get("kernels", {}).get(dom[2][0].split("<")[0]) if "kernels" in tj else None
                    if ent:
                        traffic = ent["dram_bytes_per_image"] * B
                        break
                except Exception:
                    pass
        roofline = {"bound": "hbm", "kernel": dom[2][0] if dom[2] else None, "kernels_of_one_step": step_names,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[0],
                    # the same launch on the bytes it really moved through HBM (ncu): how close the kernel runs to the copy peak
                    "frac_on_traffic": (traffic / (dom[0] * 1e-3) / 1e9) / peak if traffic else None,
                    "whole_step_frac_of_compulsory": (B * alg_bytes_per_image() / (ms_per_step * 1e-3) / 1e9) / peak}
        # ---- parity of the timed path against the reference's CPU result of the same input (image 0) -------
        try:
            ref, kind, cores = reference_lib()
            step()
            torch.cuda.synchronize()
            hx = x[0].cpu().numpy()
            want = np.empty((H, W // 2 + 1), dtype=np.complex64)
            ref.r2c(hx, want, [0, 1], True, 1.0, cores)
            err = rel_l2(X[0].cpu().numpy(), want)
            bound = 1e-5 * math.log2(H * W)
            parity = {"parity_rel_l2": err, "bound": bound, "ok": bool(err <= bound), "against": f"oracle/_ref ({kind}), image 0 of the step",
                      "tolerance": "1e-5*log2(n), n = 16384*16384"}
            del hx, want
        except Exception as e:  # pragma: no cover
            parity = {"parity_rel_l2": None, "error": repr(e)}

    # ---- end to end through the C ABI with HOST buffers (every rank; max over ranks) ---------------------------
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(B, H, W, dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hX = torch.empty(B, H, W // 2 + 1, dtype=torch.complex64, pin_memory=True)
        nx, nX = hx.numpy(), hX.numpy()
        k = max(3, min(args.steps, 5))

        def wall(fn, reps, warm=2):
            for _ in range(warm):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return max_over_ranks((time.perf_counter() - t0) / reps)

        dt = wall(lambda: R.r2c(nx, nX, [1, 2], True, 1.0), k)
        h2d_b, d2h_b = B * H * W * 4, B * H * (W // 2 + 1) * 8
        e2e = {"value": world * B * flops_per_image() / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d_b,
               "d2h_bytes_per_step": d2h_b, "ms_per_step": dt * 1e3,
               "call": f"numba_r2c via rocket_fft_b200.r2c(numpy pinned in/out), {B} images per step per rank, "
                       "image-pipelined H2D / kernels / D2H inside the call"}
        # parity of the e2e result against the device-resident result (itself checked against the reference above)
        step()
        torch.cuda.synchronize()
        e2e["matches_device_path_rel_l2"] = float(torch.linalg.vector_norm(torch.view_as_real(hX[:, :4].to(dev) - X[:, :4])) /
                                                  torch.linalg.vector_norm(torch.view_as_real(X[:, :4])))
        # the copies alone: what the link allows (no kernels)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def h2d_only():
            x.copy_(hx, non_blocking=True)
            torch.cuda.synchronize()

        def d2h_only():
            hX.copy_(X, non_blocking=True)
            torch.cuda.synchronize()

        def both():
            with torch.cuda.stream(s1):
                x.copy_(hx, non_blocking=True)
            with torch.cuda.stream(s2):
                hX.copy_(X, non_blocking=True)
            torch.cuda.synchronize()

        e2e["h2d_ms"] = wall(h2d_only, 3, 1) * 1e3
        e2e["d2h_ms"] = wall(d2h_only, 3, 1) * 1e3
        e2e["copy_only_ms"] = wall(both, 3, 1) * 1e3
        e2e["h2d_GBps"] = h2d_b / e2e["h2d_ms"] / 1e6
        e2e["d2h_GBps"] = d2h_b / e2e["d2h_ms"] / 1e6
        e2e["frac_of_copy_only"] = e2e["copy_only_ms"] / e2e["ms_per_step"]
        if numa is not None:
            e2e["numa_binding_rank0"] = numa
        del hx, hX, nx, nX
        # the drop-in case: plain (pageable) NumPy arrays, as an @njit caller has them
        try:
            px = np.empty((B, H, W), dtype=np.float32)
            px[...] = 0.5
            pX = np.empty((B, H, W // 2 + 1), dtype=np.complex64)
            dtp = wall(lambda: R.r2c(px, pX, [1, 2], True, 1.0), 3, 2)
            e2e["e2e_pageable"] = {"value": world * B * flops_per_image() / dtp / 1e9, "unit": "GFLOP/s", "ms_per_step": dtp * 1e3,
                                   "frac_of_pinned": dt / dtp,
                                   "call": "the same call on plain NumPy arrays (pageable host memory)"}
            # the same arrays page-locked by the caller for the duration (rfb200_host_pin / rocket_fft_b200.pinned): the
            # registration is paid once, the calls then run at the pinned rate
            if world > 1:
                raise StopIteration  # (one rank is enough for this leg; eight ranks would page-lock 8 x 8.6 GB more of one host)
            t0 = time.perf_counter()
            with R.pinned(px, pX):
                reg_ms = (time.perf_counter() - t0) * 1e3
                dtr = wall(lambda: R.r2c(px, pX, [1, 2], True, 1.0), 3, 1)
            e2e["e2e_pageable"]["pinned_by_caller"] = {"value": world * B * flops_per_image() / dtr / 1e9, "ms_per_step": dtr * 1e3,
                                                       "frac_of_pinned": dt / dtr, "host_pin_ms_once": reg_ms,
                                                       "call": "with rocket_fft_b200.pinned(x, out): r2c(x, out, ...)"}
            del px, pX
        except StopIteration:
            pass
        except Exception as e:  # pragma: no cover
            e2e.setdefault("e2e_pageable", {})["error"] = repr(e)

    del x, X
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2]: fftn complex64 1024^3 on the same N GPUs ---------------------------------------------
    fftn = None
    if not args.no_fftn:
        try:
            fftn = bench_fftn(args, R, torch, dist, dev, rank, world, barrier, max_over_ranks)
        except Exception as e:  # pragma: no cover
            fftn = {"error": repr(e)}
        torch.cuda.empty_cache()

    other = None
    if rank == 0 and world == 1 and not args.no_other:
        try:
            other = bench_other_configs(R, torch, dev, timeit)
        except Exception as e:  # pragma: no cover
            other = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference_rfft2(1, 10, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "parity": parity,
            "parity_rel_l2": parity.get("parity_rel_l2") if parity else None,
            # the BASELINE label's "5N*log2N" applied verbatim to real data is exactly twice `value`
            "value_5NlogN_label": 2.0 * value, "clocks": clocks, "stages": stages, "fftn": fftn, "other_configs": other,
            "library": R.version()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_fftn(args, R, torch, dist, dev, rank, world, barrier, max_over_ranks):
    """fftn of a complex64 n^3 volume: N = 1 one c2c(axes=[0,1,2]) call; N > 1 slab decomposition (SlabFFTN, default
    engine).  Parity: the same code path at 512^3 against the reference on the host (rank 0 gathers the shards)."""
    import numpy as np

    from rocket_fft_b200.distributed import SlabFFTN, shard_batch

    n = args.fftn_n
    steps = max(3, min(args.steps, 10))
    flops = 5.0 * n**3 * math.log2(n**3)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn, k):
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / k)

    res = {"workload": f"cfg3: fftn complex64 {n}^3, axes=[0,1,2]", "n_gpus": world, "scaling": "strong"}
    # ---- parity at 512^3 through the same path ---------------------------------------------------------------------------
    pn = min(512, n)
    rng = np.random.default_rng(2)
    if rank == 0:
        full = (rng.standard_normal((pn, pn, pn), dtype=np.float32) + 1j * rng.standard_normal((pn, pn, pn), dtype=np.float32)).astype(np.complex64)
    else:
        full = None
    if world == 1:
        v = torch.from_numpy(full).to(dev)
        V = torch.empty_like(v)
        R.c2c(v, V, [0, 1, 2], True, 1.0)
        got = V.cpu().numpy()
        del v, V
    else:
        ft = torch.from_numpy(full).to(dev) if rank == 0 else torch.empty(pn, pn, pn, dtype=torch.complex64, device=dev)
        dist.broadcast(torch.view_as_real(ft), 0)  # (NCCL has no complex type: the same bytes as pairs of floats)
        lo, hi = shard_batch(pn, rank, world)
        plan = SlabFFTN((pn, pn, pn), torch.complex64, dev)
        y = plan.forward(ft[lo:hi].clone(), True, 1.0).contiguous()  # (pn, pn/P, pn): axis 1 sharded
        parts = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.view_as_real(y), [torch.view_as_real(q) for q in parts] if rank == 0 else None, dst=0)
        got = torch.cat(parts, dim=1).cpu().numpy() if rank == 0 else None
        res["exchange_engine"] = plan.mode
        del ft, y, parts, plan
    if rank == 0:
        ref, kind, cores = reference_lib()
        want = np.empty_like(full)
        ref.c2c(full, want, [0, 1, 2], True, 1.0, cores)
        err = rel_l2(got, want)
        bound = 1e-5 * math.log2(pn**3)
        res.update({"parity_rel_l2": err, "parity_bound": bound, "parity_ok": bool(err <= bound),
                    "parity_case": f"{pn}^3 through the same path vs oracle/_ref ({kind})"})
        del want, got
    del full
    torch.cuda.empty_cache()
    # ---- timing at n^3 --------------------------------------------------------------------------------------------------
    if world == 1:
        x = torch.randn(n, n, n, dtype=torch.complex64, device=dev)
        y = torch.empty_like(x)
        for _ in range(3):
            R.c2c(x, y, [0, 1, 2], True, 1.0)
        ms = timed(lambda: R.c2c(x, y, [0, 1, 2], True, 1.0), steps)
        axes_ms = {f"axis{ax}": timed(lambda: R.c2c(x, y, [ax], True, 1.0), steps) for ax in (0, 1, 2)}
        res.update({"ms": ms, "alltoall_ms": None, "bus_GBps": None, "frac_of_900": None, "GFLOPs": flops / ms / 1e6,
                    "compulsory_GBps": 2 * 8 * n**3 / ms / 1e6, "stages_ms": axes_ms,
                    "layout": "single device: natural order in and out",
                    "limiter": "axis " + max(axes_ms, key=axes_ms.get)[-1] + " pass (strided lines)"})
        return res
    plan = SlabFFTN((n, n, n), torch.complex64, dev)
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    x = torch.randn(n // world, n, n, dtype=torch.complex64, device=dev, generator=g)
    for _ in range(3):
        plan.forward(x, True, 1.0)
    total_ms = timed(lambda: plan.forward(x, True, 1.0), steps)
    if plan.mode == "fused":
        a2a_ms = timed(lambda: plan.scatter_axis1(x, True), steps)  # axis-1 transform + NVLink push: one kernel
        planes_ms = timed(lambda: R.c2c(x, x, [2], True, 1.0), steps)
    else:
        a2a_ms = timed(lambda: plan.exchange(x), steps)
        planes_ms = timed(lambda: plan.local_planes(x, True, 1.0), steps)
    axis0_ms = timed(lambda: plan.local_axis0(True), steps)
    bus = plan.bytes_sent_per_rank / a2a_ms / 1e6
    st = {"local_planes": planes_ms, "alltoall": a2a_ms, "axis0": axis0_ms}
    res.update({"ms": total_ms, "alltoall_ms": a2a_ms, "bus_GBps": bus, "frac_of_900": bus / 900.0, "frac_of_measured_770": bus / 770.0,
                "GFLOPs": flops / total_ms / 1e6, "bytes_sent_per_gpu": plan.bytes_sent_per_rank, "stages_ms": st,
                "exchange": plan.mode + {"symm": " (pack fused into the push)", "nccl": " (pack + all_to_all_single)",
                                         "fused": " (the axis-1 FFT kernel stores straight into peer HBM over NVLink; alltoall_ms = that kernel incl. its butterflies)"}[plan.mode],
                "layout": "result left axis-1 sharded (transposed); transpose_back available",
                "limiter": max(st, key=st.get)})
    del x, plan
    torch.cuda.empty_cache()
    try:
        res["rfftn"] = bench_rfftn_slab(args, R, torch, dist, dev, rank, world, timed)
    except Exception as e:  # pragma: no cover
        res["rfftn"] = {"error": repr(e)}
    return res


def bench_rfftn_slab(args, R, torch, dist, dev, rank, world, timed):
    """rfftn of a real float32 n^3 volume, slab-decomposed (SlabRFFTN: r2c along the last axis, then the complex pipeline on the
    half spectrum, fused transform + push): total ms, and parity of the same path at 256^3 against the reference."""
    import numpy as np

    from rocket_fft_b200.distributed import SlabRFFTN, shard_batch

    n = args.fftn_n
    out = {"workload": f"rfftn float32 {n}^3 (r2c axes=[0,1,2]), slab-decomposed over {world} GPUs"}
    pn = min(256, n)
    full = torch.empty(pn, pn, pn, dtype=torch.float32, device=dev)
    if rank == 0:
        full.copy_(torch.from_numpy(np.random.default_rng(3).standard_normal((pn, pn, pn), dtype=np.float32)))
    dist.broadcast(full, 0)
    lo, hi = shard_batch(pn, rank, world)
    plan = SlabRFFTN((pn, pn, pn), torch.float32, dev)
    y = plan.forward(full[lo:hi].contiguous()).contiguous()
    parts = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.view_as_real(y), [torch.view_as_real(q) for q in parts] if rank == 0 else None, dst=0)
    if rank == 0:
        ref, kind, cores = reference_lib()
        fh = full.cpu().numpy()
        want = np.empty((pn, pn, pn // 2 + 1), dtype=np.complex64)
        ref.r2c(fh, want, [0, 1, 2], True, 1.0, cores)
        err = rel_l2(torch.cat(parts, dim=1).cpu().numpy(), want)
        bound = 1e-5 * math.log2(pn**3)
        out.update({"parity_rel_l2": err, "parity_bound": bound, "parity_ok": bool(err <= bound),
                    "parity_case": f"{pn}^3 through the same path vs oracle/_ref ({kind})"})
    del plan, y, parts, full
    torch.cuda.empty_cache()
    plan = SlabRFFTN((n, n, n), torch.float32, dev)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    x = torch.randn(n // world, n, n, dtype=torch.float32, device=dev, generator=g)
    for _ in range(3):
        plan.forward(x)
    ms = timed(lambda: plan.forward(x), max(3, min(args.steps, 10)))
    out.update({"ms": ms, "exchange_engine": plan.mode, "bytes_sent_per_gpu": plan.bytes_sent_per_rank,
                "GFLOPs": 2.5 * n**3 * math.log2(n**3) / ms / 1e6})
    return out


def bench_other_configs(R, torch, dev, timeit):
    """The remaining BASELINE configs, device-resident, algorithmic (compulsory) bytes / time (CUDA events)."""
    peak, _ = load_peaks()
    out = []

    def row(name, fn, nbytes, reps=10):
        R.launch_count_reset()
        ms = timeit(fn, reps)
        out.append({"config": name, "ms": ms, "algorithmic_GBps": nbytes / ms / 1e6, "frac_of_peak": nbytes / ms / 1e6 / peak,
                    "launches_per_call": R.launch_count() // (reps + 3)})

    x = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    y = torch.empty_like(x)
    row("cfg1 c2c complex128 (4096,4096) axes=[1]", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 16, 20)
    del x, y
    x = torch.randn(256, 15015, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    row("cfg4a c2c complex64 (256,15015) axes=[1]", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 20)
    del x, y
    x = torch.randn(256, 1000003, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    row("cfg4b c2c complex64 (256,1000003) axes=[1] (Bluestein)", lambda: R.c2c(x, y, [1], True, 1.0), 2 * x.numel() * 8, 3)
    del x, y
    x = torch.randn(2048, 2048, 64, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    row("cfg5 dct type 2 float64 (2048,2048,64) axes=[0,1]", lambda: R.dct(x, y, [0, 1], 2, 1.0, False), 2 * x.numel() * 8, 3)
    row("cfg5 dst type 2 float64 (2048,2048,64) axes=[0,1]", lambda: R.dst(x, y, [0, 1], 2, 1.0, False), 2 * x.numel() * 8, 3)
    del x, y
    x = torch.randn(4, 16384, 8193, dtype=torch.complex64, device=dev)
    y = torch.empty(4, 16384, 16384, dtype=torch.float32, device=dev)
    row("cfg2 inverse: irfft2 (c2r axes=[1,2]) of 4 half spectra 16384x8193", lambda: R.c2r(x, y, [1, 2], False, 1.0),
        x.numel() * 8 + y.numel() * 4, 5)
    del x, y
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
