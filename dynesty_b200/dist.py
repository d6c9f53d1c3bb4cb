"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).

The hot path shards naturally: the chains of one queue fill are
independent (the reference maps them over a pool, sampler.py:717), so rank r runs chains
[r*Q/W, (r+1)*Q/W) of every fill -- chain ids, and therefore Philox streams, are GLOBAL, so
the gathered queue is bit-identical to a single-GPU fill.  The one exchange step is an
all-gather of the finished chains (u, v, logl, counters) -- Q*(2n+1)*8 + O(Q) bytes -- after
which every rank holds the full queue and advances the identical host state; the live points
are therefore replicated and the bound update needs no further communication.

On GPUs the all-gather is FUSED into the chain kernels (``attach_peer``): every rank maps the
exchange windows of all peers (CUDA IPC over NVLink/NVSwitch) and the kernels store finished
chains into every window, with an in-kernel arrive/wait at the end (csrc/b2n_peer.cu).
torch.distributed then only carries the 64-byte IPC handles at start-up; ``allgather`` below
(one collective per output array) remains as the transport for the gloo/CPU tests.
"""
import numpy as np


class Comm:
    peer_ctx = None      # _lib.Context whose exchange windows are mapped (attach_peer)

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = device if device is not None else (
            torch.device('cuda', torch.cuda.current_device()) if self.backend == 'nccl' else torch.device('cpu'))

    def attach_peer(self, ctx, max_rows, ndim):
        """Set up the fused exchange for fills of up to `max_rows` chains of dimension `ndim`:
        allocate this rank's window, swap IPC handles, map the peers' windows."""
        handle = ctx.peer_export(ctx.peer_window_bytes(max_rows, ndim))
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        ctx.peer_import(self.rank, self.world, handles)
        self.dist.barrier()
        self.peer_ctx = ctx

    def shard(self, Q):
        """Rows [lo, hi) of a Q-row fill owned by this rank (Q is a multiple of world)."""
        assert Q % self.world == 0
        per = Q // self.world
        return self.rank * per, (self.rank + 1) * per

    def allgather(self, local, Q):
        """dict of per-rank arrays (first axis = local rows) -> dict of full arrays."""
        torch, dist = self.torch, self.dist
        out = {}
        for k in sorted(local):
            a = np.ascontiguousarray(local[k])
            # torch.distributed has no unsigned 16/32/64-bit collectives (gloo raises "Invalid scalar type"):
            # ship the bits as the signed type of the same width and view them back
            dt = a.dtype
            if dt.kind == 'u' and dt.itemsize > 1:
                a = a.view(np.dtype('i%d' % dt.itemsize))
            t = torch.from_numpy(a).to(self.device)
            full = torch.empty((Q,) + tuple(a.shape[1:]), dtype=t.dtype, device=self.device)
            dist.all_gather_into_tensor(full, t)
            out[k] = full.cpu().numpy().view(dt)
        return out

    def allreduce_sum(self, a):
        """Element-wise sum over ranks of a float64 array (the moment exchange of a sharded bound update)."""
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def gather_rows(self, a, dst=0):
        """Gather per-rank arrays (same shape on every rank) on rank `dst` only: the other ranks' hosts
        receive nothing (rank 0 owns the nested-sampling bookkeeping)."""
        a = np.ascontiguousarray(a)
        dt = a.dtype
        if dt.kind == 'u' and dt.itemsize > 1:
            a = a.view(np.dtype('i%d' % dt.itemsize))
        t = self.torch.from_numpy(a).to(self.device)
        if self.rank == dst:
            parts = [self.torch.empty_like(t) for _ in range(self.world)]
            self.dist.gather(t, parts, dst=dst)
            return self.torch.cat(parts).cpu().numpy().view(dt)
        self.dist.gather(t, None, dst=dst)
        return None

    def max(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])
