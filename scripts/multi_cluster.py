#!/usr/bin/env python
"""BASELINE config 4, multi-GPU form: M cluster dispatch records (default 268 435 456) sharded by contiguous record range across the ranks
("sharded by record range").  Every rank holds the scene (the records reference global object ids), produces the dispatch list
with cluster_expand, keeps its range [r M / G, (r + 1) M / G) and runs the cluster cull on it; the per-rank draw lists are concatenated in
shard order on rank 0 over NVLink peer memory (blz_cull_gather_push) -- ascending record index, so the gathered list is byte-identical to
the single-GPU list, which rank 0 also computes and compares (sha256 of both).  One JSON line per mode (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/multi_cluster.py
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=268_435_456)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--verify-records", type=int, default=33_554_432, help="size of the byte-for-byte comparison with the single-GPU list (0 = skip)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from blitzen_b200 import capi, scene, dist as bdist, types as T
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = B.build_workload(B.N_OBJECTS, 0, 1)
    n = len(w["objs"])
    cube = scene.cube_side(n); half = cube / 2
    view = scene.make_view((half, half, -1.2 * cube), z_far=1e9, width=1920, height=1080)
    stream = torch.cuda.Stream()

    def run(records, verify):
        lo, hi = bdist.shard_range(records, rank, world)
        m = hi - lo
        out = {}
        # context A: the whole dispatch list (every rank expands redundantly: 1.5 ms, no exchange); context B: this rank's range
        with capi.CullContext(local) as A, capi.CullContext(local) as Bc:
            for ctx, cap in ((A, records), (Bc, max(m, 1))):
                ctx.set_stream(stream.cuda_stream)
                ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], clusters=w["clusters"], transform_id_base=w["transform_id_base"],
                                 cluster_dispatch_capacity=cap, draw_capacity=cap if ctx is Bc or verify else 16)
                ctx.set_view(view); ctx.set_depth(w["depth"]); ctx.build_pyramid(capi.HIZ_VK)
            A.cluster_expand()
            mw, mt = C.c_uint32(), C.c_uint32()
            A._check(A._lib.blz_cull_read_cluster_dispatch(A._h, None, 0, C.byref(mw), C.byref(mt)))
            assert mw.value == records, (mw.value, records)
            src = A.outputs().cluster_dispatch + lo * 12
            Bc._check(Bc._lib.blz_cull_set_cluster_dispatch(Bc._h, C.c_void_p(src), m, 1))
            gather = bdist.DrawListGather(Bc, rank, world, capacity_records=records, fmt=capi.REC_VK24)
            epoch = 0
            for mode, mname in ((capi.CLUSTER_PASSTHROUGH, "passthrough"), (capi.CLUSTER_SPHERE, "sphere"), (capi.CLUSTER_SPHERE_HIZ, "sphere_hiz")):
                ts = []
                for it in range(a.iters + 2):
                    dist.barrier(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(stream):
                        e0.record(stream)
                        Bc.cluster_cull(mode, capi.REC_VK24, capi.HIZ_VK)
                        epoch += 1; gather.push(epoch)
                        e1.record(stream)
                    torch.cuda.synchronize(); dist.barrier()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1))
                ms = float(np.mean(ts))
                t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
                _, tot = Bc.read_count()
                tt = torch.tensor([tot], device="cuda", dtype=torch.int64); dist.all_reduce(tt); total = int(tt.item())
                res = {"case": f"cluster_cull_{mname}", "records": records, "n_gpus": world, "ms_cull_plus_gather_max_over_ranks": round(ms, 4),
                       "records_per_s": records / (ms * 1e-3), "draws_total": total}
                if verify and rank == 0:
                    got, counts = gather.read(epoch)
                    A.cluster_cull(mode, capi.REC_VK24, capi.HIZ_VK)
                    ref, rtot = A.read_draws(capi.REC_VK24)
                    res["gathered_equals_single_gpu"] = bool(rtot == total and hashlib.sha256(got.tobytes()).digest() == hashlib.sha256(ref.tobytes()).digest())
                    del got, ref
                dist.barrier()
                out[mname] = res
        return out

    if a.verify_records:
        v = run(a.verify_records, True)
        if rank == 0:
            for r in v.values():
                print(json.dumps(dict(r, what="verification run")), flush=True)
    full = run(a.records, False)
    if rank == 0:
        for r in full.values():
            print(json.dumps(r), flush=True)
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
