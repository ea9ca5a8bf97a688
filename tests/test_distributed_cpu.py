"""Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo process group.
The slab driver is exercised with the NumPy oracle standing in for the CUDA transform
(test infrastructure only), so pack / all-to-all / layout logic is checked without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rocket_fft_b200.distributed import SlabFFTN, SlabRFFTN, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_c2c(a, b, axes, fwd, fct):
    from oracle import pocketfft_oracle as O

    out = np.empty(tuple(a.shape), dtype=np.complex64 if a.dtype == torch.complex64 else np.complex128)
    O.c2c(a.numpy(), out, axes, fwd, fct)
    b.copy_(torch.from_numpy(out))
    return b


def _worker(rank, world, port, shape, transpose_back, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        full = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128)
        n0 = shape[0]
        lo, hi = shard_batch(n0, rank, world)
        x = torch.from_numpy(full[lo:hi].copy())
        plan = SlabFFTN(shape, torch.complex128, "cpu", local_c2c=_oracle_c2c)
        y = plan.forward(x, True, 1.0, transpose_back=transpose_back)
        want = np.fft.fftn(full)
        if transpose_back:
            mine = want[lo:hi]
        else:
            j0, j1 = shard_batch(shape[1], rank, world)
            mine = want[:, j0:j1]
        err = np.linalg.norm(y.numpy() - mine) / np.linalg.norm(mine)
        q.put((rank, float(err), plan.bytes_sent_per_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transpose_back", [False, True])
def test_slab_fftn_world2_gloo(transpose_back):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shape = (8, 6, 10)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, transpose_back, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, sent in res:
        assert err < 1e-13, (rank, err)
        assert sent == 8 * 6 * 10 * 16 // 2 // 2


def _oracle_real(kind):
    def run(a, b, axes, fwd, fct):
        from oracle import pocketfft_oracle as O

        out = np.empty(tuple(b.shape), dtype=b.numpy().dtype)
        getattr(O, kind)(np.ascontiguousarray(a.numpy()), out, axes, fwd, fct)
        b.copy_(torch.from_numpy(out))
        return b

    return run


def _worker_real(rank, world, port, shape, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(6)
        full = rng.standard_normal(shape)
        lo, hi = shard_batch(shape[0], rank, world)
        x = torch.from_numpy(full[lo:hi].copy())
        plan = SlabRFFTN(shape, torch.float64, "cpu", local=(_oracle_real("r2c"), _oracle_real("c2r"), _oracle_c2c))
        y = plan.forward(x)
        want = np.fft.rfftn(full)
        j0, j1 = shard_batch(shape[1], rank, world)
        err = np.linalg.norm(y.numpy() - want[:, j0:j1]) / np.linalg.norm(want[:, j0:j1])
        back = plan.inverse(y.clone())
        err2 = np.linalg.norm(back.numpy() - full[lo:hi]) / np.linalg.norm(full[lo:hi])
        q.put((rank, float(err), float(err2), tuple(y.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 6, 10), (4, 10, 9)])
def test_slab_rfftn_irfftn_world2_gloo(shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_real, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err2, yshape in res:
        assert yshape == (shape[0], shape[1] // 2, shape[2] // 2 + 1)
        assert err < 1e-13 and err2 < 1e-13, (rank, err, err2)


def test_shard_batch_covers_everything():
    for n in (0, 1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
