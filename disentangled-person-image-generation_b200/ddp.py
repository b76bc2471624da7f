"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the
exchange steps of the path -- one sum-all-reduce of the flat fp32 gradient arena per optimiser step and the
(2 x C) fp64 BatchNorm statistics of the discriminator (sync-BN).  The reference has no distributed code
(N_GPUS = 1, wgan_gp.py:114); per-image work is embarrassingly parallel."""
import os
import threading

import torch
import torch.distributed as dist


class Dist:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend, rank=self.rank, world_size=self.world_size, **kw)

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX)

    def broadcast(self, t, src=0):
        dist.broadcast(t, src)

    def barrier(self):
        dist.barrier()


def shard(batch, rank, world):
    """Rank r takes images [r*B/N, (r+1)*B/N) of a global batch (dict of arrays)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0] // world
        out[k] = v[rank * n:(rank + 1) * n]
    return out


class LocalGroup:
    """N data-parallel ranks inside ONE process on ONE GPU: one host thread + one CUDA stream + one dpig_ctx per rank,
    exchanging through device memory.  The DP-equivalence check of SURVEY.md section 4(iii) -- one engine on a batch of B
    against k engines on B/k with the sync-BN sums and the gradient all-reduce -- then runs wherever one GPU is, through
    exactly the hooks (`all_reduce_sum`, `world_size`) the NCCL path uses.  Not a performance path."""

    def __init__(self, world_size):
        self.world_size = int(world_size)
        self._barrier = threading.Barrier(self.world_size)
        self._slots = [None] * self.world_size

    def rank(self, r):
        return LocalDist(self, r)

    def run(self, fn):
        """fn(dist) on one thread per rank, each under its own CUDA stream; returns the per-rank results (rank order)
        and re-raises the first exception."""
        out, err = [None] * self.world_size, [None] * self.world_size

        def work(r):
            try:
                with torch.cuda.stream(torch.cuda.Stream()):
                    out[r] = fn(self.rank(r))
                    torch.cuda.current_stream().synchronize()
            except BaseException as e:  # noqa: BLE001 - re-raised below
                err[r] = e
                self._barrier.abort()

        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world_size)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out


class LocalDist:
    """The `dist` object of one LocalGroup rank (same surface as Dist)."""

    capturable = False      # the exchange synchronises host threads: it cannot sit inside a CUDA graph capture

    def __init__(self, group, rank):
        self.group, self.rank, self.world_size, self.local_rank = group, rank, group.world_size, 0

    def _exchange(self, t, op):
        g = self.group
        torch.cuda.current_stream().synchronize()      # this rank's contribution is complete
        g._slots[self.rank] = t
        g._barrier.wait()
        acc = g._slots[0].clone()
        for other in g._slots[1:]:                      # fixed (rank) order: every rank computes the same bits
            acc = op(acc, other)
        torch.cuda.current_stream().synchronize()
        g._barrier.wait()                               # everyone has read every slot
        t.copy_(acc)
        torch.cuda.current_stream().synchronize()
        g._barrier.wait()

    def all_reduce_sum(self, t):
        self._exchange(t, torch.add)

    def all_reduce_max(self, t):
        self._exchange(t, torch.maximum)

    def broadcast(self, t, src=0):
        g = self.group
        torch.cuda.current_stream().synchronize()
        g._slots[self.rank] = t
        g._barrier.wait()
        if self.rank != src:
            t.copy_(g._slots[src])
        torch.cuda.current_stream().synchronize()
        g._barrier.wait()

    def barrier(self):
        torch.cuda.current_stream().synchronize()
        self.group._barrier.wait()
