"""Run under torchrun (one rank per GPU): the slab-decomposed fftn / rfftn (all three exchange engines: the fused
FFT + NVLink push kernel, symmetric-memory pushes, NCCL all_to_all) against the reference's CPU result (oracle/_ref,
or the NumPy restatement when that library did not travel) of the gathered volume; also the batch-sharded path.
Exits non-zero on mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import torch.distributed as dist

import parity
import rocket_fft_b200 as R
from rocket_fft_b200.distributed import SlabFFTN, SlabRFFTN, shard_batch

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
T = parity.reflib()
if T is None:
    from oracle import pocketfft_oracle as T  # noqa: N812
NT = max(1, (os.cpu_count() or 1) // world)
ok = True


def ref_c2c(full):
    want = np.empty_like(full)
    T.c2c(full, want, [0, 1, 2], True, 1.0, NT) if parity.reflib() is not None else T.c2c(full, want, [0, 1, 2], True, 1.0)
    return want


def ref_r2c(full):
    want = np.empty(full.shape[:2] + (full.shape[2] // 2 + 1,), dtype=np.complex64)
    T.r2c(full, want, [0, 1, 2], True, 1.0, NT) if parity.reflib() is not None else T.r2c(full, want, [0, 1, 2], True, 1.0)
    return want


for shape in ((128, 128, 128), (64, 96, 40), (64, 1024, 1024), (256, 256, 64)):
    rng = np.random.default_rng(7)  # same volume on every rank
    fullh = (rng.standard_normal(shape, dtype=np.float32) + 1j * rng.standard_normal(shape, dtype=np.float32)).astype(np.complex64)
    want = ref_c2c(fullh)
    full = torch.from_numpy(fullh).to(dev)
    lo, hi = shard_batch(shape[0], rank, world)
    j0, j1 = shard_batch(shape[1], rank, world)
    n = shape[0] * shape[1] * shape[2]
    pow2 = (shape[1] & (shape[1] - 1)) == 0 and shape[1] >= 16
    for engine in (("fused",) if pow2 else ()) + ("symm", "nccl"):
        plan = SlabFFTN(shape, torch.complex64, dev, exchange=engine)
        y = plan.forward(full[lo:hi].clone(), True, 1.0)
        torch.cuda.synchronize()
        err = parity.l2err(y.cpu().numpy(), want[:, j0:j1])
        z = plan.forward(full[lo:hi].clone(), True, 1.0, transpose_back=True)
        torch.cuda.synchronize()
        err2 = parity.l2err(z.cpu().numpy(), want[lo:hi])
        tol = parity.tol(np.float32, n)
        print(f"rank {rank} shape {shape} engine {plan.mode} err {err:.3e} {err2:.3e} (tol {tol:.1e})", flush=True)
        ok = ok and plan.mode == engine and err <= tol and err2 <= tol
        del plan, y, z
    del full
# real volumes: rfftn into the transposed distribution and irfftn back
for shape in ((128, 128, 128), (64, 96, 41), (64, 512, 1024)):
    rng = np.random.default_rng(8)
    fullh = rng.standard_normal(shape, dtype=np.float32)
    want = ref_r2c(fullh)
    full = torch.from_numpy(fullh).to(dev)
    lo, hi = shard_batch(shape[0], rank, world)
    j0, j1 = shard_batch(shape[1], rank, world)
    n = shape[0] * shape[1] * shape[2]
    pow2 = (shape[1] & (shape[1] - 1)) == 0 and shape[1] >= 16
    for engine in (("auto",) if pow2 else ()) + ("symm", "nccl"):
        plan = SlabRFFTN(shape, torch.float32, dev, exchange=engine)
        y = plan.forward(full[lo:hi].contiguous())
        torch.cuda.synchronize()
        err = parity.l2err(y.cpu().numpy(), want[:, j0:j1])
        back = plan.inverse(y.clone())
        torch.cuda.synchronize()
        err2 = parity.l2err(back.cpu().numpy(), fullh[lo:hi])
        tol = parity.tol(np.float32, n)
        print(f"rank {rank} real {shape} engine {plan.mode} err {err:.3e} {err2:.3e} (tol {tol:.1e})", flush=True)
        ok = ok and err <= tol and err2 <= tol
        del plan, y, back
    del full
# batch sharding: every rank transforms its rows; concatenation equals the full transform
rows = 64
rng = np.random.default_rng(9)
fullh = rng.standard_normal((rows, 4096)) + 1j * rng.standard_normal((rows, 4096))
want = np.empty_like(fullh)
T.c2c(fullh, want, [1], True, 1.0)
full = torch.from_numpy(fullh).to(dev)
lo, hi = shard_batch(rows, rank, world)
mine = torch.empty_like(full[lo:hi])
R.c2c(full[lo:hi], mine, [1], True, 1.0)
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
err = parity.l2err(torch.cat(gathered).cpu().numpy(), want)
print(f"rank {rank} batch-sharded err {err:.3e}", flush=True)
ok = ok and err <= parity.tol(np.float64, 4096)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
