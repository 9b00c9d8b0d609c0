"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the set-up exchange only.

Objects are sharded by contiguous ranges; each rank culls and compacts its shard; the per-rank draw lists are concatenated
in shard order on the presenting rank by blz_cull_gather_push -- one kernel per rank that stores straight into the
presenter's buffer over NVLink peer memory (CUDA IPC mapping).  torch.distributed only broadcasts the 128-byte IPC handle
blob once; there is no collective on the data path.
"""
import numpy as np

from . import capi


def shard_range(n, rank, world):
    """Contiguous object range of `rank`: [rank*n/world, (rank+1)*n/world)  (SURVEY.md 8e)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def exclusive_scan(counts):
    counts = np.asarray(counts, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(counts)[:-1]])


def concat_in_shard_order(lists):
    """Host-side reference of what the gather produces (used by the gloo CPU tests)."""
    return np.concatenate([np.asarray(l) for l in lists]) if lists else np.zeros(0)


class DrawListGather:
    """Presenter = rank 0.  Works with any initialised torch.distributed backend for the handle broadcast."""

    def __init__(self, ctx, rank, world, capacity_records, fmt=capi.REC_VK24, presenter=0):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.fmt, self.presenter = ctx, rank, world, fmt, presenter
        self.capacity = int(capacity_records)
        if rank == presenter:
            blob = ctx.gather_export(self.capacity, fmt)
            t = torch.from_numpy(blob.copy())
        else:
            t = torch.zeros(128, dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=presenter)
        blob = t.cpu().numpy()
        ctx.gather_import(None if rank == presenter else blob, rank, world, self.capacity, fmt)
        dist.barrier()

    def push(self, epoch):
        self.ctx.gather_push(epoch)

    def push_async(self, epoch):
        """Push on a side stream; the context flips to its second draw buffer so the next pass overlaps the transfer."""
        self.ctx.gather_push_async(epoch)

    def read(self, epoch):
        if self.rank != self.presenter:
            raise RuntimeError("only the presenting rank reads the gathered list")
        return self.ctx.gather_read(epoch, self.world, self.fmt)
