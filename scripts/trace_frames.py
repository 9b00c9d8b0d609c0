#!/usr/bin/env python
"""Kernel timeline of the bench frame with the draw-list gather, one file per rank (there is no nsys in the image: CUPTI through
torch.profiler).  gpurun_out/trace_n<world>_rank<r>.json = [[kernel name, start us, duration us, stream], ...] for the traced frames.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29620 scripts/trace_frames.py [--frames 8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--objects", type=int, default=B.N_OBJECTS)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from torch.profiler import profile, ProfilerActivity
    from blitzen_b200 import capi, dist as bdist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = B.build_workload(a.objects, rank, world)
    stream = torch.cuda.Stream()
    ctx = capi.CullContext(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], object_id_base=w["object_id_base"], transform_id_base=w["transform_id_base"])
    ctx.set_view(w["view"]); ctx.set_depth(w["depth"])
    gather = bdist.DrawListGather(ctx, rank, world, capacity_records=a.objects * world // 4, fmt=capi.REC_VK24) if world > 1 else None
    epoch = [0]

    def frame():
        ctx.early(capi.REC_VK24)
        ctx.build_pyramid(capi.HIZ_VK)
        if gather:
            epoch[0] += 1; gather.push_async(epoch[0])
        ctx.late(capi.REC_VK24, capi.HIZ_VK)
        if gather:
            epoch[0] += 1; gather.push_async(epoch[0])

    ctx.clear_pyramid(capi.HIZ_VK, B.DEPTH_W, B.DEPTH_H); ctx.late(capi.REC_VK24, capi.HIZ_VK)
    for _ in range(5):
        frame()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.frames):
            frame()
        if gather:
            ctx.gather_join()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type is not None and "cuda" in str(e.device_type).lower():
            ev.append([e.name[:60], e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "device_resource_id", -1) if hasattr(e, "device_resource_id") else -1])
    ev.sort(key=lambda x: x[1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"trace_n{world}_rank{rank}.json"), "w") as f:
        json.dump(ev, f)
    if rank == 0:
        print("traced", len(ev), "device events")
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
