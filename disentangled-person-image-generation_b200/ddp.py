"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the
exchange steps of the path -- one sum-all-reduce of the flat fp32 gradient arena per optimiser step and the
(2 x C) fp64 BatchNorm statistics of the discriminator (sync-BN).  The reference has no distributed code
(N_GPUS = 1, wgan_gp.py:114); per-image work is embarrassingly parallel."""
import os

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
