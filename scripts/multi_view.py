#!/usr/bin/env python
"""BASELINE config 5, multi-GPU form: G views (main camera + shadow-cascade-like views: same position, yaw steps of 45 degrees, zFar doubled
per view) of the SAME 16 777 216-object scene, one view per GPU, each against its own 3840x2160 depth image (2048x2048x11 pyramid).
"Replicas + shard by view" (SURVEY 8e): every rank holds the whole scene and runs the two-phase frame for its view; there is no exchange
on the data path.  Prints one JSON line (rank 0): views x objects culled per second, max over ranks of the device time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 scripts/multi_view.py [--steps K]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--objects", type=int, default=B.N_OBJECTS)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from blitzen_b200 import capi, scene
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = B.build_workload(a.objects, 0, 1)                   # every rank: the WHOLE scene
    n = len(w["objs"])
    cube = scene.cube_side(a.objects)
    view = scene.make_view((cube / 2, cube / 2, cube / 2), yaw=float(np.deg2rad(45.0 * rank)), z_far=650.0 * (2.0 ** min(rank, 4)), width=3840, height=2160)
    depth = scene.synthetic_depth(3840, 2160, seed=0x00B1172E + rank)
    stream = torch.cuda.Stream()
    ctx = capi.CullContext(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], transform_id_base=w["transform_id_base"])
    ctx.set_view(view); ctx.set_depth(depth)

    def frame():
        ctx.early(capi.REC_VK24); ctx.build_pyramid(capi.HIZ_VK); ctx.late(capi.REC_VK24, capi.HIZ_VK)

    ctx.clear_pyramid(capi.HIZ_VK, 3840, 2160); ctx.late(capi.REC_VK24, capi.HIZ_VK)
    for _ in range(max(a.warmup, 3)):
        frame()
    ctx.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(a.steps):
            frame()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    vis = int(ctx.read_visibility().sum())
    ctx.early(capi.REC_VK24); _, early = ctx.read_count()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        v = torch.tensor([vis, early], device="cuda", dtype=torch.int64); gl = [torch.zeros_like(v) for _ in range(world)]; dist.all_gather(gl, v)
        per_view = [[int(x[0]), int(x[1])] for x in gl]
    else:
        per_view = [[vis, early]]
    if rank == 0:
        print(json.dumps({"metric": "view_objects_culled_per_s", "value": world * n / (ms * 1e-3), "unit": "objects x views / s", "n_gpus": world, "views": world,
                          "objects_per_view": n, "ms_per_frame_max_over_ranks": ms, "depth": [3840, 2160], "parallelism": f"view-sharded x{world} (scene replicated, no exchange)",
                          "per_view_visible_and_early_draws": per_view}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
