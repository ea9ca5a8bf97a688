"""Multi-GPU drivers (one process per GPU, torch.distributed over NCCL/NVLink).

The reference is single-process CPU code: nothing to mirror here.
Two ways the transform path shards:

* batched transforms (configs 1, 2, 4, 5): independent lines / images -> every rank runs
  the single-GPU plan on its own slice, no collective (`shard_batch`);
* one large N-D transform (config 3, fftn of a 1024^3 volume): slab decomposition with ONE
  exchange step.  Rank g owns planes [g*N0/P, (g+1)*N0/P) of axis 0:
      1. local transforms over axes (1, 2) of the slab           (our kernels)
      2. all-to-all: block (g -> h) = my planes x axis-1 range of h  (NCCL over NVLink)
      3. local transform along axis 0 on (N0, N1/P, N2)          (our kernels)
  The result is left in the TRANSPOSED distribution (axis 1 sharded, axis 0 complete),
  which is what a following inverse transform wants; `transpose_back=True` adds the second
  all-to-all that restores the input distribution.
* the real counterpart (`SlabRFFTN`): rfftn of a real volume = r2c along axis 2 on the
  slab, then the complex pipeline above on the half-spectrum volume (N0, N1, N2/2+1); irfftn runs it backwards
  from the transposed distribution (c2c along axis 0, exchange, c2r over axes (1, 2)).
"""
from __future__ import annotations

import math


def shard_batch(n_items: int, rank: int, world: int):
    """Half-open range of batch items owned by `rank` (even split, remainder to the first ranks)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SlabFFTN:
    """fftn/ifftn of a 3-D complex volume sharded by slabs of axis 0 over a process group.

    `local_c2c(ain, aout, axes, forward, fct)` is the single-device transform
    (default: rocket_fft_b200.c2c on CUDA tensors).  All buffers are allocated once.
    """

    def __init__(self, shape, dtype, device, group=None, local_c2c=None, dist=None, exchange="auto"):
        import torch

        if dist is None:
            import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.group = group
        self.P = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n0, n1, n2 = (int(s) for s in shape)
        if n0 % self.P or n1 % self.P:
            raise ValueError("axis 0 and axis 1 extents must be divisible by the number of ranks")
        self.shape = (n0, n1, n2)
        self.dtype = dtype
        self.device = device
        local_c2c_is_default = local_c2c is None
        if local_c2c is None:
            from . import lowlevel

            local_c2c = lambda a, b, axes, fwd, fct: lowlevel.c2c(a, b, axes, fwd, fct)  # noqa: E731
        self.c2c = local_c2c
        P = self.P
        self.local_in_shape = (n0 // P, n1, n2)
        self.local_out_shape = (n0, n1 // P, n2)
        blk = (P, n0 // P, n1 // P, n2)
        # Peer-mapped receive buffers are the plan's own: their rows are padded to a multiple of 128 bytes, so that the kernels
        # that write and read them (the fused transform + push, the axis-0 pass) see 16-byte aligned rows whatever n2 is --
        # the half spectrum of a real volume has n2/2 + 1 = 513 points per row (measured: rfftn 1024^3 on two GPUs 7.0 -> see
        # profiles/README.md).  The NCCL engine keeps dense buffers (all_to_all_single wants them contiguous).
        esz = torch.empty((), dtype=dtype).element_size()
        per128 = max(1, 128 // esz)
        self.n2_alloc = -(-n2 // per128) * per128
        blk_alloc = (P, n0 // P, n1 // P, self.n2_alloc)
        self.bytes_sent_per_rank = n0 // P * n1 * n2 * torch.empty((), dtype=dtype).element_size() * (P - 1) // P
        # Exchange engine.  "fused": symm buffers + the axis-1 transform writes its output
        # straight into the peers' receive buffers (rfb200_c2c_scatter): transform and all-to-all
        # are ONE kernel, the NVLink traffic overlaps the butterflies.  "symm": peer-mapped symmetric receive buffers (torch symmetric memory
        # over NVLink P2P): every rank PUSHES its strided blocks straight into the peers' receive
        # buffers -- the pack pass disappears and no NCCL staging is involved.  "nccl": pack +
        # all_to_all_single.  "auto": symm on CUDA when it can be set up, else nccl.
        self.mode = "nccl"
        self.hdl = None
        if exchange in ("auto", "symm", "fused") and str(device).startswith("cuda"):
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self.recv_alloc = symm_mem.empty(blk_alloc, dtype=dtype, device=device)
                self.recv = self.recv_alloc[..., :n2]
                gname = (group if group is not None else dist.group.WORLD).group_name
                self.hdl = symm_mem.rendezvous(self.recv_alloc, gname)
                self.peer_recv = [self.hdl.get_buffer(h, blk_alloc, dtype)[..., :n2] for h in range(P)]
                self.streams = [torch.cuda.Stream(device=device) for _ in range(min(P, 4))]
                self.mode = "symm"
                n1_ok = (n1 & (n1 - 1)) == 0 and 16 <= n1 <= 2048  # lengths the strided power-of-two kernel takes
                if exchange == "fused" or (exchange == "auto" and n1_ok and local_c2c_is_default):
                    if not n1_ok:
                        raise ValueError("fused exchange needs a power-of-two axis-1 length (16..2048)")
                    self.mode = "fused"
            except Exception as e:  # pragma: no cover - depends on the box
                if exchange in ("symm", "fused"):
                    raise
                self.symm_error = repr(e)
                self.hdl = None
        if self.mode == "nccl":
            self.n2_alloc = n2
            self.recv_alloc = self.recv = torch.empty(blk, dtype=dtype, device=device)
            self.send = torch.empty(blk, dtype=dtype, device=device)
        self.timings = {}

    # -- pieces (exposed so the bench can time the exchange on its own) ---------------------
    def local_planes(self, x, forward=True, fct=1.0):
        """In-place transforms over axes (1, 2) of the local slab (n0/P, n1, n2)."""
        self.c2c(x, x, [1, 2], forward, fct)
        return x

    def pack(self, x):
        """(n0/P, n1, n2) -> send[h, i0, j, i2] = x[i0, h*n1/P + j, i2]  (nccl mode only)."""
        P = self.P
        n0p, n1, n2 = self.local_in_shape
        self.send.copy_(x.view(n0p, P, n1 // P, n2).permute(1, 0, 2, 3))
        return self.send

    def exchange(self, x=None):
        """Move block (g -> h) for all h.  symm mode reads the strided slab `x` directly."""
        if self.mode == "fused":
            raise RuntimeError("fused mode has no stand-alone exchange (it is part of the axis-1 transform)")
        if self.mode == "nccl":
            if x is not None:
                self.pack(x)
            self.dist.all_to_all_single(self.recv.view(-1), self.send.view(-1), group=self.group)
            return self.recv
        torch = self.torch
        P, g = self.P, self.rank
        n0p, n1, n2 = self.local_in_shape
        xv = x.view(n0p, P, n1 // P, n2)
        cur = torch.cuda.current_stream()
        self.hdl.barrier(channel=0)  # peers are done reading their receive buffers
        ready = torch.cuda.Event()
        ready.record(cur)
        for off in range(P):
            h = (g + off) % P  # staggered so that no peer is hit by everyone at once
            st = self.streams[off % len(self.streams)]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                self.peer_recv[h][g].copy_(xv[:, h])
        for st in self.streams:
            cur.wait_stream(st)
        self.hdl.barrier(channel=1)  # everybody's pushes have landed
        return self.recv

    def planes_and_exchange_fused(self, x, forward=True, fct=1.0):
        """Axis-2 transform in place, then the axis-1 transform whose stores land in the peers'
        receive buffers (one kernel = transform + all-to-all push over NVLink)."""
        self.c2c(x, x, [2], forward, fct)
        return self.scatter_axis1(x, forward)

    def scatter_axis1(self, x, forward=True):
        from . import lowlevel

        self.hdl.barrier(channel=0)  # peers are done reading their receive buffers
        parts = [self.peer_recv[h][self.rank] for h in range(self.P)]
        lowlevel.c2c_scatter(x, parts, 1, forward, 1.0)
        self.hdl.barrier(channel=1)  # everybody's pushes have landed
        return self.recv

    def local_axis0(self, forward=True):
        """recv viewed as (n0, n1/P, n2): transform along axis 0 in place."""
        n0, n1p, n2 = self.local_out_shape
        y = self.recv_alloc.view(n0, n1p, self.n2_alloc)[..., :n2]
        self.c2c(y, y, [0], forward, 1.0)
        return y

    def forward(self, x, forward=True, fct=1.0, transpose_back=False):
        """x: this rank's slab (n0/P, n1, n2), overwritten.  Returns the local part of the
        result: (n0, n1/P, n2) [axis-1 sharded] or, with transpose_back, (n0/P, n1, n2)."""
        if self.mode == "fused":
            self.planes_and_exchange_fused(x, forward, fct)
        else:
            self.local_planes(x, forward, fct)
            self.exchange(x)
        y = self.local_axis0(forward)
        if not transpose_back:
            return y
        P = self.P
        n0, n1p, n2 = self.local_out_shape
        # second exchange (NCCL): block (h -> g) = axis-0 range of g x my axis-1 range
        send = y.reshape(P, n0 // P, n1p, n2).contiguous()
        back = self.torch.empty_like(send)
        self.dist.all_to_all_single(back.view(-1), send.view(-1), group=self.group)
        # back[h, i0, j, i2] holds X[i0, h*n1/P + j, i2]
        x.view(n0 // P, P, n1p, n2).copy_(back.permute(1, 0, 2, 3))
        return x

    @staticmethod
    def flops(shape):
        n = 1
        for s in shape:
            n *= int(s)
        return 5.0 * n * math.log2(n)


class SlabRFFTN:
    """rfftn / irfftn of a 3-D real volume sharded by slabs of axis 0.

    forward(x):  x = this rank's real slab (n0/P, n1, n2)  ->  its part (n0, n1/P, n2//2+1) of rfftn(X),
                 axis 1 sharded (the transposed distribution, as SlabFFTN leaves it).
    inverse(Y):  Y = (n0, n1/P, n2//2+1) in that distribution  ->  real slab (n0/P, n1, n2) of irfftn.
    Semantics per axis are those of the single-device path (r2c / c2r, _pocketfft_hdronly.h:3955-4023):
    the real transform runs along the LAST axis, c2c along the others.
    `local` = (r2c, c2r, c2c) single-device callables, default: rocket_fft_b200 on CUDA tensors.
    """

    def __init__(self, shape, dtype, device, group=None, local=None, dist=None, exchange="auto"):
        import torch

        n0, n1, n2 = (int(s) for s in shape)
        self.shape = (n0, n1, n2)
        self.n2h = n2 // 2 + 1
        self.rdtype = dtype
        self.cdtype = {torch.float32: torch.complex64, torch.float64: torch.complex128}[dtype]
        if local is None:
            from . import lowlevel

            local = (lowlevel.r2c, lowlevel.c2r, None)
        self.r2c, self.c2r, c2c = local
        self.inner = SlabFFTN((n0, n1, self.n2h), self.cdtype, device, group=group, local_c2c=c2c, dist=dist,
                              exchange=exchange)
        self.P, self.rank = self.inner.P, self.inner.rank
        self.torch = torch
        # (rows of the half spectrum padded to 128 bytes like the receive buffers)
        self.n2h_alloc = self.inner.n2_alloc if self.inner.mode != "nccl" else self.n2h
        self.spec_alloc = torch.empty((n0 // self.P, n1, self.n2h_alloc), dtype=self.cdtype, device=device)
        self.spec = self.spec_alloc[..., :self.n2h]
        self.bytes_sent_per_rank = self.inner.bytes_sent_per_rank

    @property
    def mode(self):
        return self.inner.mode

    def forward(self, x, forward=True, fct=1.0):
        """Real slab in (left untouched), complex (n0, n1/P, n2//2+1) out (a view of the plan's receive buffer)."""
        inner = self.inner
        self.r2c(x, self.spec, [2], forward, fct)                      # real transform along the last axis
        if inner.mode == "fused":
            inner.scatter_axis1(self.spec, forward)                    # axis-1 transform + all-to-all push
        else:
            inner.c2c(self.spec, self.spec, [1], forward, 1.0)
            inner.exchange(self.spec)
        return inner.local_axis0(forward)

    def inverse(self, y, out=None, forward=False, fct=None):
        """y: (n0, n1/P, n2//2+1) complex, overwritten.  Returns the real slab (n0/P, n1, n2);
        fct defaults to 1/(n0 n1 n2) (numpy's irfftn)."""
        inner, P = self.inner, self.P
        n0, n1, n2 = self.shape
        n1p = n1 // P
        if fct is None:
            fct = 1.0 / (float(n0) * n1 * n2)
        inner.c2c(y, y, [0], forward, 1.0)
        # exchange back: block (h -> g) = axis-0 range of g x axis-1 range of h
        send = y.reshape(P, n0 // P, n1p, self.n2h)
        if not send.is_contiguous():
            send = send.contiguous()
        back = self.torch.empty_like(send)
        inner.dist.all_to_all_single(back.view(-1), send.view(-1), group=inner.group)
        self.spec_alloc.view(n0 // P, P, n1p, self.n2h_alloc)[..., :self.n2h].copy_(back.permute(1, 0, 2, 3))
        if out is None:
            out = self.torch.empty((n0 // P, n1, n2), dtype=self.rdtype, device=self.spec.device)
        self.c2r(self.spec, out, [1, 2], forward, fct)                # c2c along axis 1, Hermitian -> real along 2
        return out

    @staticmethod
    def flops(shape):
        n = 1
        for s in shape:
            n *= int(s)
        return 2.5 * n * math.log2(n)
