#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank culls its contiguous shard, pushes its draw list
into the presenting rank's buffer over NVLink peer memory (blz_cull_gather_push), and rank 0 compares the concatenated list
with the oracle's list for the WHOLE scene: byte-identical, for a frustum pass and for a two-phase frame (early + late lists).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_verify_gather.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from blitzen_b200 import capi, scene, dist as bdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_stress = int(os.environ.get("VERIFY_OBJECTS", "3000000"))
    total = 1001 + n_stress
    groups = scene.scaled_groups(n_stress)
    mult = scene.cube_side(total)
    a, b = bdist.shard_range(total, rank, world)
    objs, xf = scene.generate(groups, mult, True, "counter", seed=4, first=a, count=b - a, threads=8)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    view = scene.make_view((mult / 2, mult / 2, mult / 2), z_far=3000.0, width=1280, height=720)
    depth = scene.synthetic_depth(1280, 720, n_rects=48, z_min=30.0, z_max=900.0)
    ctx = capi.CullContext(local)
    ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], object_id_base=a, transform_id_base=tbase)
    ctx.set_view(view); ctx.set_depth(depth)
    gather = bdist.DrawListGather(ctx, rank, world, capacity_records=total, fmt=capi.REC_VK24)
    results = {}
    epoch = 0
    # pass 1: frustum + LOD
    ctx.frustum_lod(); epoch += 1; gather.push(epoch); ctx.synchronize(); dist.barrier()
    if rank == 0:
        results["frustum"] = gather.read(epoch)
    dist.barrier()
    # frame 0 (cleared pyramid) + frame 1 (real pyramid): late lists of both, early list of frame 1
    ctx.clear_pyramid(capi.HIZ_VK, 1280, 720)
    ctx.late(); epoch += 1; gather.push(epoch); ctx.synchronize(); dist.barrier()
    if rank == 0:
        results["late0"] = gather.read(epoch)
    dist.barrier()
    ctx.early(); epoch += 1; gather.push(epoch); ctx.synchronize(); dist.barrier()
    if rank == 0:
        results["early1"] = gather.read(epoch)
    dist.barrier()
    ctx.build_pyramid(capi.HIZ_VK); ctx.late(); epoch += 1; gather.push(epoch); ctx.synchronize(); dist.barrier()
    if rank == 0:
        results["late1"] = gather.read(epoch)
    dist.barrier()
    # the same frame again with ASYNCHRONOUS pushes (side stream, double-buffered draw lists): early list of epoch e must still be
    # intact on the presenter after the late list of epoch e + 1 has arrived (two halves of the gather buffer)
    ctx.early(); epoch += 1; e_early = epoch; gather.push_async(epoch)
    ctx.build_pyramid(capi.HIZ_VK); ctx.late(); epoch += 1; e_late = epoch; gather.push_async(epoch)
    ctx.synchronize(); dist.barrier()
    if rank == 0:
        results["early2_async"] = gather.read(e_early)
        results["late2_async"] = gather.read(e_late)
    dist.barrier()
    ok = True
    if rank == 0:
        import oracle_lib as O
        fo, fx = scene.generate(groups, mult, True, "counter", seed=4, threads=8)
        ft, fb = scene.assemble_transforms(fo, fx)
        kw = dict(threads=O.hardware_threads(), transform_id_base=fb)
        S = (fo, ft, tables["surfaces"], tables["lods"], view)
        exp = {}
        exp["frustum"], _, _ = O.cull(*S, O.PASS_FRUSTUM, **kw)
        vis = np.zeros(total, dtype=np.uint32)
        exp["late0"], _, vis = O.cull(*S, O.PASS_LATE, pyramid=O.cleared_pyramid(1280, 720, O.HIZ_VK), vis=vis, **kw)
        exp["early1"], _, _ = O.cull(*S, O.PASS_EARLY, vis=vis, **kw)
        exp["late1"], _, vis = O.cull(*S, O.PASS_LATE, pyramid=O.build_pyramid(depth, O.HIZ_VK), vis=vis, **kw)
        exp["early2_async"], _, _ = O.cull(*S, O.PASS_EARLY, vis=vis, **kw)
        exp["late2_async"], _, vis = O.cull(*S, O.PASS_LATE, pyramid=O.build_pyramid(depth, O.HIZ_VK), vis=vis, **kw)
        for k in ("frustum", "late0", "early1", "late1", "early2_async", "late2_async"):
            recs, counts = results[k]
            got = recs.view(np.uint32).reshape(-1, 6)
            same = got.shape == exp[k].shape and np.array_equal(got, exp[k])
            ok = ok and same
            print(f"[verify_gather] world={world} {k}: per-rank counts {list(counts)} total {len(got)} expected {len(exp[k])} -> {'IDENTICAL' if same else 'MISMATCH'}", flush=True)
    ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
