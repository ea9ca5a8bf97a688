"""Run under torchrun (one rank per GPU): slab-decomposed fftn vs torch.fft.fftn of the
gathered volume; also the batch-sharded path.  Exits non-zero on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import rocket_fft_b200 as R
from rocket_fft_b200.distributed import SlabFFTN, SlabRFFTN, shard_batch

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True
for shape in ((128, 128, 128), (64, 96, 40)):
    g = torch.Generator(device=dev).manual_seed(7)
    full = torch.randn(*shape, dtype=torch.complex64, device=dev, generator=g)  # same on every rank
    lo, hi = shard_batch(shape[0], rank, world)
    x = full[lo:hi].clone()
    plan = SlabFFTN(shape, torch.complex64, dev, exchange="fused" if shape[1] == 128 else "symm")
    print(f"rank {rank} exchange mode {plan.mode}", flush=True)
    y = plan.forward(x, True, 1.0)
    want = torch.fft.fftn(full)
    j0, j1 = shard_batch(shape[1], rank, world)
    err = float(torch.linalg.vector_norm((y - want[:, j0:j1]).to(torch.complex128)) / torch.linalg.vector_norm(want[:, j0:j1].to(torch.complex128)))
    x = full[lo:hi].clone()
    z = plan.forward(x, True, 1.0, transpose_back=True)
    err2 = float(torch.linalg.vector_norm((z - want[lo:hi]).to(torch.complex128)) / torch.linalg.vector_norm(want[lo:hi].to(torch.complex128)))
    tol = 1e-5 * 21
    print(f"rank {rank} shape {shape} err {err:.3e} {err2:.3e}", flush=True)
    ok = ok and err < tol and err2 < tol
# real volumes: rfftn into the transposed distribution and irfftn back
for shape in ((128, 128, 128), (64, 96, 41)):
    g = torch.Generator(device=dev).manual_seed(8)
    full = torch.randn(*shape, dtype=torch.float32, device=dev, generator=g)
    lo, hi = shard_batch(shape[0], rank, world)
    plan = SlabRFFTN(shape, torch.float32, dev, exchange="auto" if shape[1] == 128 else "symm")
    y = plan.forward(full[lo:hi].contiguous())
    want = torch.fft.rfftn(full)
    j0, j1 = shard_batch(shape[1], rank, world)
    err = float(torch.linalg.vector_norm((y - want[:, j0:j1]).to(torch.complex128)) / torch.linalg.vector_norm(want[:, j0:j1].to(torch.complex128)))
    back = plan.inverse(y.clone())
    err2 = float(torch.linalg.vector_norm((back - full[lo:hi]).double()) / torch.linalg.vector_norm(full[lo:hi].double()))
    print(f"rank {rank} real {shape} mode {plan.mode} err {err:.3e} {err2:.3e}", flush=True)
    ok = ok and err < 1e-5 * 21 and err2 < 1e-5 * 21
# batch sharding: every rank transforms its rows; concatenation equals the full transform
rows = 64
g = torch.Generator(device=dev).manual_seed(9)
full = torch.randn(rows, 4096, dtype=torch.complex128, device=dev, generator=g)
lo, hi = shard_batch(rows, rank, world)
mine = torch.empty_like(full[lo:hi])
R.c2c(full[lo:hi], mine, [1], True, 1.0)
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
err = float(torch.linalg.vector_norm(torch.cat(gathered) - torch.fft.fft(full, dim=1)) / torch.linalg.vector_norm(torch.fft.fft(full, dim=1)))
print(f"rank {rank} batch-sharded err {err:.3e}", flush=True)
ok = ok and err < 1e-13 * 12
dist.destroy_process_group()
sys.exit(0 if ok else 1)
