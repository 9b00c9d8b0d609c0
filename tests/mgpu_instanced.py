#!/usr/bin/env python
"""BASELINE config 3, multi-GPU form: indirect instancing over an object-sharded scene (default 67 108 864 objects in total), per-LOD instance
buckets gathered on rank 0 (NCCL all-gather of the per-rank per-LOD counts + peer stores of the ids, blitzen_b200/dist.py InstanceListGather).
First a verification run (3 M objects): rank 0's gathered buckets + totals == the oracle's for the whole scene; then the timed run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/mgpu_instanced.py
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=67_108_864, help="objects in total (sharded over the ranks)")
    ap.add_argument("--verify-objects", type=int, default=3_000_000)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from blitzen_b200 import capi, scene, dist as bdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tables = scene.mesh_tables()
    nl = len(tables["lods"])
    stream = torch.cuda.Stream()

    def run(total, verify):
        groups = scene.scaled_groups(total - 1001)
        mult = scene.cube_side(total)
        lo, hi = bdist.shard_range(total, rank, world)
        objs, xf = scene.generate(groups, mult, True, "counter", seed=6, first=lo, count=hi - lo, threads=os.cpu_count() or 8)
        transforms, tbase = scene.assemble_transforms(objs, xf)
        view = scene.make_view((mult / 2, mult / 2, mult / 2), z_far=3000.0, width=1920, height=1080)
        # global bucket layout (rank 0's buffer): capacity = a generous share of the scene per LOD; local layouts use the same offsets
        gcap = np.full(nl, max(total // 3, 1), dtype=np.uint32)
        goff = np.concatenate([[0], np.cumsum(gcap.astype(np.uint64))[:-1]]).astype(np.uint32)
        li = tables["lodInstances"].copy()
        li["instanceOffset"] = goff
        lcap = gcap if rank == 0 else np.full(nl, hi - lo, dtype=np.uint32)
        if rank != 0:
            li["instanceOffset"] = (np.arange(nl, dtype=np.uint64) * (hi - lo)).astype(np.uint32)
        out = None
        with capi.CullContext(local) as ctx:
            ctx.set_stream(stream.cuda_stream)
            ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], lod_instances=li, bucket_capacity=lcap, object_id_base=lo, transform_id_base=tbase)
            ctx.set_view(view)
            g = bdist.InstanceListGather(ctx, rank, world, nl, goff, gcap, stream)
            ts = []
            for it in range(a.iters + 2):
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    ctx.instanced()
                g.push()
                with torch.cuda.stream(stream):
                    e1.record(stream)
                torch.cuda.synchronize(); dist.barrier()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.mean(ts))
            t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
            totals = g.totals()
            res = {"case": "instanced_sharded", "objects_total": total, "n_gpus": world, "ms_instancing_plus_gather_max_over_ranks": round(ms, 4),
                   "objects_per_s": total / (ms * 1e-3), "instances_total": int(totals.sum())}
            if verify and rank == 0:
                import oracle_lib as O
                fo, fx = scene.generate(groups, mult, True, "counter", seed=6, threads=os.cpu_count() or 8)
                ft, fb = scene.assemble_transforms(fo, fx)
                gli = tables["lodInstances"].copy(); gli["instanceOffset"] = goff
                idx_e, cnt_e, _ = O.cull_instanced(fo, ft, tables["surfaces"], tables["lods"], gli, gcap, view, threads=O.hardware_threads(), transform_id_base=fb)
                idx, _ = ctx.read_instances(int(goff[-1]) + int(gcap[-1]))
                same = bool(np.array_equal(totals, cnt_e))
                for l in range(nl):
                    o, c = int(goff[l]), int(min(cnt_e[l], gcap[l]))
                    same = same and bool(np.array_equal(idx[o:o + c], idx_e[o:o + c]))
                res["gathered_buckets_equal_oracle"] = same
            dist.barrier()
            out = res
        return out

    if a.verify_objects:
        r = run(a.verify_objects, True)
        if rank == 0:
            print(json.dumps(dict(r, what="verification run")), flush=True)
    r = run(a.objects, False)
    if rank == 0:
        print(json.dumps(r), flush=True)
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
