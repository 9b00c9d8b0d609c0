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


class InstanceListGather:
    """Indirect instancing over object shards: the per-LOD buckets of the ranks are concatenated in rank order on rank 0 (the presenter, whose
    own buckets start the global ones).  The only collective is an NCCL all-gather of the per-rank per-LOD counts (lod_count words per rank),
    issued on the context's stream; the instance ids travel by peer stores (blz_cull_instances_push).  `global_offset` / `global_cap`: bucket
    layout of the presenter's instance buffer (numpy uint32 [lod_count]); every rank's LOCAL layout must start bucket l at global_offset[l] on
    rank 0 (it is the same buffer there)."""

    def __init__(self, ctx, rank, world, n_lods, global_offset, global_cap, stream):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.n_lods, self.stream = ctx, rank, world, n_lods, stream
        t = torch.from_numpy(ctx.instances_export().copy()).cuda() if rank == 0 else torch.zeros(64, dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        ctx.instances_import(None if rank == 0 else t.cpu().numpy(), rank, world)
        self.mine = torch.zeros(n_lods, dtype=torch.int32, device="cuda")
        self.all = torch.zeros(world * n_lods, dtype=torch.int32, device="cuda")
        self.done = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.goff = torch.from_numpy(np.ascontiguousarray(global_offset, dtype=np.uint32).view(np.int32)).cuda()
        self.gcap = torch.from_numpy(np.ascontiguousarray(global_cap, dtype=np.uint32).view(np.int32)).cuda()
        dist.barrier()

    def push(self):
        """After ctx.instanced() on every rank; stream-ordered, no host synchronisation.  Ordering protocol (both ends are collectives on the
        context's stream): the all-gather of the counts cannot complete before the PRESENTER has reached this push on its own stream, i.e.
        behind whatever it enqueued to consume the previous frame's buckets -- no rank overwrites buckets that are still being read; the
        one-word all-reduce at the end completes on the presenter only after every rank's peer stores have been issued and fenced
        (__threadfence_system at the end of the push kernel) -- work enqueued behind push() on the presenter sees complete buckets.  Both
        collectives are inside what the benchmarks time."""
        import torch
        import torch.distributed as dist
        with torch.cuda.stream(self.stream):
            self.ctx.instances_counts(self.mine.data_ptr())
            dist.all_gather_into_tensor(self.all, self.mine)
            self.ctx.instances_push(self.all.data_ptr(), self.goff.data_ptr(), self.gcap.data_ptr())
            dist.all_reduce(self.done)

    def totals(self):
        """Per-LOD number of ids stored over all ranks (each rank's count clamped to its bucket capacity; host; synchronises)."""
        return self.all.view(self.world, self.n_lods).sum(dim=0).cpu().numpy().astype(np.uint32)
