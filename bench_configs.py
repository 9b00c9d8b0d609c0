#!/usr/bin/env python
"""BASELINE.json configs[2..4] as first-class bench workloads (bench.py --workload cfg3 | cfg4 | cfg5), at 1/2/4/8 GPUs:

  cfg3  InstancingStressTest: 67 108 864 instances of the bundled meshes, indirect instancing (drawInst* shaders), OBJECT-SHARDED over the
        ranks, per-LOD instance buckets gathered on rank 0 (NCCL all-gather of the counts + peer stores).        scaling: strong
  cfg4  cluster mode: 268 435 456 cluster dispatch records (cluster_expand over an all-visible 16.7 M-object scene), bounding-sphere frustum +
        Hi-Z cull with per-cluster command compaction, sharded by RECORD RANGE, draw lists gathered on rank 0.        scaling: strong
  cfg5  multi-view: 8 cameras / cascades x 16 777 216 objects against a 3840x2160 depth image each (2048x2048x11 pyramid), VIEWS sharded over the
        ranks (scene replicated, no exchange).                                                                     scaling: strong

Same timing rules as bench.py: W >= 3 warm-up steps, exactly K timed steps between barrier + synchronize, CUDA events on the launching stream,
max over ranks, inputs far larger than L2, clocks sampled during the timed region.  Every line carries a device-side correctness summary
(csrc/consume.cu reductions; at N > 1 the gathered result against the all-reduced per-rank reductions) -- the full parity tests are tests/.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

INSTANCES = 67_108_864
CLUSTER_RECORDS = 268_435_456
VIEWS = 8


def _env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class Harness:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank, self.world, self.local = _env()
        if not torch.cuda.is_available():
            raise SystemExit("bench_configs.py needs a CUDA device: the cull path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.stream = torch.cuda.Stream()
        self.args = args
        self.peak, self.peak_src = B.load_peaks()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor([int(x) for x in xs], device="cuda", dtype=self.torch.int64)
        if self.dist is not None:
            self.dist.all_reduce(t)
        return [int(v) for v in t.tolist()]

    def xor_over_ranks(self, x):
        if self.dist is None:
            return int(x)
        torch = self.torch
        v = torch.tensor([int(x) & 0x7FFFFFFFFFFFFFFF, int(x) >> 63], device="cuda", dtype=torch.int64)
        vs = [torch.zeros_like(v) for _ in range(self.world)]
        self.dist.all_gather(vs, v)
        out = 0
        for q in vs:
            out ^= int(q[0].item()) | (int(q[1].item()) << 63)
        return out

    def timed(self, step, launches_of):
        """W warm-up steps, then exactly K steps: (ms per step as max over ranks, clocks, launches in the timed region on this rank)."""
        torch, a = self.torch, self.args
        for _ in range(max(a.warmup, 3)):
            step()
        self.barrier()
        sampler = B.ClockSampler(self.local)
        l0 = launches_of()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        sampler.start()
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
        for _ in range(a.steps):
            step()
        with torch.cuda.stream(self.stream):
            e1.record(self.stream)
        self.barrier()
        clocks = sampler.stop()
        return self.max_over_ranks(e0.elapsed_time(e1) / a.steps), clocks, launches_of() - l0

    def finish(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()

    def line(self, metric, unit, value, ms, config, roofline, clocks, launches, detail):
        a = self.args
        return {"metric": metric, "value": value, "unit": unit, "n_gpus": self.world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "roofline": roofline, "cpu_baseline": None, "e2e": None,
                "e2e_note": "secondary workload: device-timed only (the end-to-end and CPU-baseline legs belong to the contract line, --workload cfg2)",
                "gpu_launches": int(launches), "clocks": clocks, "detail": detail}


def summary(cs):
    return {"records": int(cs.records), "index_sum": int(cs.index_sum), "instance_sum": int(cs.instance_sum), "id_sum": int(cs.id_sum), "id_xor": "%016x" % int(cs.id_xor),
            "bad_object": int(cs.bad_object), "bad_lod": int(cs.bad_lod), "unsorted": int(cs.unsorted)}


# ------------------------------------------------------------------------------------------------------------------------------------------------
def run_cfg3(h):
    """Indirect instancing, objects sharded by contiguous ranges, buckets gathered on rank 0."""
    from blitzen_b200 import capi, scene, dist as bdist
    a, torch = h.args, h.torch
    total = a.objects if a.objects != B.N_OBJECTS else INSTANCES
    tables = scene.mesh_tables()
    nl = len(tables["lods"])
    groups = scene.scaled_groups(total - 1001)
    mult = scene.cube_side(total)
    lo, hi = bdist.shard_range(total, h.rank, h.world)
    objs, xf = scene.generate(groups, mult, True, "counter", seed=6, first=lo, count=hi - lo, threads=os.cpu_count() or 8)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    view = scene.make_view((mult / 2, mult / 2, mult / 2), z_far=3000.0, width=1920, height=1080)
    # the reference's 100 000-entry buckets (blitzenMeshes.cpp:159-161) would overflow at this size: the harness sizes them
    gcap = np.full(nl, max(total // 3, 1), dtype=np.uint32)
    goff = np.concatenate([[0], np.cumsum(gcap.astype(np.uint64))[:-1]]).astype(np.uint32)
    li = tables["lodInstances"].copy()
    li["instanceOffset"] = goff
    lcap = gcap if h.rank == 0 else np.full(nl, hi - lo, dtype=np.uint32)
    if h.rank != 0:
        li["instanceOffset"] = (np.arange(nl, dtype=np.uint64) * (hi - lo)).astype(np.uint32)
    with capi.CullContext(h.local) as ctx:
        ctx.set_stream(h.stream.cuda_stream)
        ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], lod_instances=li, bucket_capacity=lcap, object_id_base=lo, transform_id_base=tbase)
        ctx.set_view(view)
        g = bdist.InstanceListGather(ctx, h.rank, h.world, nl, goff, gcap, h.stream) if h.world > 1 else None

        def step():
            with torch.cuda.stream(h.stream):
                ctx.instanced()
            if g:
                g.push()
        ms, clocks, launches = h.timed(step, ctx.launch_count)
        # correctness on the device: every rank reduces its own buckets; after the gather the presenter reduces the global buckets
        step(); torch.cuda.synchronize()
        mine = ctx.consume_instances()
        if g:
            # the presenter's command list still describes its OWN buckets: rebuild the global counts from the all-gathered words for the reduction
            totals = g.totals()
            local_before = summary(mine)
            s = h.sum_over_ranks([local_before["records"], local_before["id_sum"], local_before["bad_object"]])
            x = h.xor_over_ranks(int(mine.id_xor))
            ok = None
            if h.rank == 0:
                idx, _ = ctx.read_instances(int(goff[-1]) + int(gcap[-1]))
                rec, ids, xr, uns = 0, 0, 0, 0
                for l in range(nl):
                    c = int(min(totals[l], gcap[l]))
                    seg = idx[int(goff[l]):int(goff[l]) + c].astype(np.uint64)
                    rec += c; ids += int(seg.sum())
                    if c:
                        xr ^= int(np.bitwise_xor.reduce(seg * np.uint64(0x9E3779B97F4A7C15)))
                        uns += int((seg[1:] <= seg[:-1]).sum())
                ok = bool(rec == s[0] and ids == s[1] and xr == x and uns == 0 and s[2] == 0)
            detail = {"instances_total": int(totals.sum()), "gathered_buckets_equal_sum_of_ranks": ok}
        else:
            sm = summary(mine)
            detail = {"instances_total": sm["records"], "bucket_checksum": sm}
        nbytes = total * 40 + detail["instances_total"] * 4 + nl * (8 + 32)
        roof = {"bound": "hbm", "kernel": "instancing step: stream_cull_kernel<PASS_FRUSTUM> -> survivor list -> counting sort by LOD (cull_list.cu)" + (" + bucket gather" if g else ""),
                "achieved": nbytes / (ms * 1e-3) / 1e9 / h.world, "peak": h.peak, "peak_source": h.peak_src, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / h.world / h.peak,
                "algorithmic_bytes_per_launch": nbytes // h.world, "bytes_per_object": 40, "objects_per_launch": total // h.world, "launch_ms": ms, "traffic": None,
                "note": "per GPU: the whole step (several kernels) against the algorithmic bytes of SURVEY 8d: N*40 + s*4 + L*(8 + R)"}
        cfg = {"workload": f"configs[2]: InstancingStressTest scaled to {total} instances of the bundled meshes, indirect instancing, object-sharded x{h.world}" +
                           (", per-LOD buckets gathered on rank 0 (NCCL all-gather of counts + NVLink peer stores + completion all-reduce)" if g else ""),
               "objects_total": total, "objects_per_gpu": hi - lo, "bucket_capacity_per_lod": int(gcap[0]), "l2_policy": "inputs larger than L2", "parallelism": f"object-sharded x{h.world}"}
        if h.rank == 0:
            print(json.dumps(h.line("objects_culled_per_s", "objects/s", total / (ms * 1e-3), ms, cfg, roof, clocks, launches, detail)), flush=True)
    h.finish()
    return 0


# ------------------------------------------------------------------------------------------------------------------------------------------------
def run_cfg4(h):
    """Cluster mode: M dispatch records sharded by record range; sphere + Hi-Z cull with per-cluster command compaction; lists gathered on rank 0."""
    import ctypes as C
    from blitzen_b200 import capi, scene, dist as bdist
    a, torch = h.args, h.torch
    records = a.records or CLUSTER_RECORDS
    w = B.build_workload(B.N_OBJECTS, 0, 1)
    n = len(w["objs"])
    cube = scene.cube_side(n); half = cube / 2
    view = scene.make_view((half, half, -1.2 * cube), z_far=1e9, width=1920, height=1080)      # everything in the frustum: the expand pass reaches M records
    lo, hi = bdist.shard_range(records, h.rank, h.world)
    m = hi - lo
    modes = ((capi.CLUSTER_SPHERE_HIZ, "sphere_hiz"), (capi.CLUSTER_SPHERE, "sphere"), (capi.CLUSTER_PASSTHROUGH, "passthrough"))
    with capi.CullContext(h.local) as A, capi.CullContext(h.local) as Bc:
        # context A: the whole dispatch list (every rank expands redundantly, no exchange); context B: this rank's range of it
        for ctx, cap, dcap in ((A, records, 16), (Bc, max(m, 1), max(m, 1))):
            ctx.set_stream(h.stream.cuda_stream)
            ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], clusters=w["clusters"], transform_id_base=w["transform_id_base"],
                             cluster_dispatch_capacity=cap, draw_capacity=dcap)
            ctx.set_view(view); ctx.set_depth(w["depth"]); ctx.build_pyramid(capi.HIZ_VK)
        A.cluster_expand()
        mw, mt = C.c_uint32(), C.c_uint32()
        A._check(A._lib.blz_cull_read_cluster_dispatch(A._h, None, 0, C.byref(mw), C.byref(mt)))
        if mw.value != records:
            raise SystemExit(f"cluster_expand produced {mw.value} records, wanted {records}")
        Bc._check(Bc._lib.blz_cull_set_cluster_dispatch(Bc._h, C.c_void_p(A.outputs().cluster_dispatch + lo * 12), m, 1))
        gather = bdist.DrawListGather(Bc, h.rank, h.world, capacity_records=records, fmt=capi.REC_VK24) if h.world > 1 else None
        epoch = [0]
        per_mode = {}
        head = None
        for mode, mname in modes:
            def step():
                with torch.cuda.stream(h.stream):
                    Bc.cluster_cull(mode, capi.REC_VK24, capi.HIZ_VK)
                    if gather:
                        epoch[0] += 1; gather.push(epoch[0])
            ms, clocks, launches = h.timed(step, Bc.launch_count)
            step(); torch.cuda.synchronize()
            mine = Bc.consume_draws(kind=1)
            s = h.sum_over_ranks([int(mine.records), int(mine.index_sum), int(mine.id_sum)])
            x = h.xor_over_ranks(int(mine.id_xor))
            res = {"ms_per_step": ms, "records_per_s": records / (ms * 1e-3), "draws_total": s[0], "index_sum": s[1]}
            if gather and h.rank == 0:
                gsum = Bc.consume_gathered(epoch[0])
                res["gathered_equals_sum_of_ranks"] = bool((int(gsum.records), int(gsum.index_sum), int(gsum.id_sum)) == (s[0], s[1], s[2]) and int(gsum.id_xor) == x and int(gsum.unsorted) == 0)
            # passthrough M*12 + M*R; sphere modes M*12 + D*40 + s*R (D = distinct owning objects, here <= n)
            nbytes = records * 12 + s[0] * 24 + (0 if mode == capi.CLUSTER_PASSTHROUGH else n * 40)
            res["algorithmic_bytes"] = nbytes
            res["frac_of_peak_per_gpu"] = nbytes / (ms * 1e-3) / 1e9 / h.world / h.peak
            per_mode[mname] = res
            if head is None:
                head = (ms, clocks, launches, nbytes)
            if h.dist is not None:
                h.dist.barrier()
        ms, clocks, launches, nbytes = head
        roof = {"bound": "hbm", "kernel": "cluster_cull_kernel<SPHERE_HIZ> (per-record bounding-sphere frustum + Hi-Z + compaction)" + (" + draw-list gather" if gather else ""),
                "achieved": nbytes / (ms * 1e-3) / 1e9 / h.world, "peak": h.peak, "peak_source": h.peak_src, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / h.world / h.peak,
                "algorithmic_bytes_per_launch": nbytes // h.world, "bytes_per_record": 12, "records_per_launch": m, "launch_ms": ms, "traffic": None}
        cfg = {"workload": f"configs[3]: cluster mode, {records} meshoptimizer cluster dispatch records (cluster_expand over an all-visible 16.7 M-object scene), "
                           f"bounding-sphere frustum + Hi-Z cull with per-cluster command compaction, sharded by record range x{h.world}" + (", draw lists gathered on rank 0" if gather else ""),
               "records_total": records, "records_per_gpu": m, "depth": [1920, 1080], "headline_mode": "sphere_hiz (the reference's own cluster shader is the passthrough mode)",
               "l2_policy": "inputs larger than L2 (3.2 GB of records)", "parallelism": f"record-range-sharded x{h.world}"}
        if h.rank == 0:
            print(json.dumps(h.line("cluster_records_culled_per_s", "records/s", records / (ms * 1e-3), ms, cfg, roof, clocks, launches, {"modes": per_mode})), flush=True)
    h.finish()
    return 0


# ------------------------------------------------------------------------------------------------------------------------------------------------
def run_cfg5(h):
    """8 views of the same 16.7 M-object scene, 4K depth each, views sharded over the ranks (rank r owns views r, r + world, ...)."""
    from blitzen_b200 import capi, scene
    a, torch = h.args, h.torch
    w = B.build_workload(a.objects, 0, 1)                   # every rank: the WHOLE scene
    n = len(w["objs"])
    cube = scene.cube_side(a.objects)
    mine = list(range(h.rank, VIEWS, h.world))
    W, H = 3840, 2160
    ctxs = []
    for v in mine:
        ctx = capi.CullContext(h.local)
        ctx.set_stream(h.stream.cuda_stream)
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], transform_id_base=w["transform_id_base"])
        ctx.set_view(scene.make_view((cube / 2, cube / 2, cube / 2), yaw=float(np.deg2rad(45.0 * v)), z_far=650.0 * (2.0 ** min(v, 4)), width=W, height=H))
        ctx.set_depth(scene.synthetic_depth(W, H, seed=0x00B1172E + v))
        ctx.clear_pyramid(capi.HIZ_VK, W, H); ctx.late(capi.REC_VK24, capi.HIZ_VK)
        ctxs.append(ctx)

    def step():
        with torch.cuda.stream(h.stream):
            for ctx in ctxs:
                ctx.early(capi.REC_VK24); ctx.build_pyramid(capi.HIZ_VK); ctx.late(capi.REC_VK24, capi.HIZ_VK)
    ms, clocks, launches = h.timed(step, lambda: sum(c.launch_count() for c in ctxs))
    per_view, nbytes = [], 0
    for v, ctx in zip(mine, ctxs):
        vis = int(ctx.read_visibility().sum())
        ctx.early(capi.REC_VK24); _, early = ctx.read_count()
        cs = ctx.consume_draws()
        ctx.build_pyramid(capi.HIZ_VK); ctx.late(capi.REC_VK24, capi.HIZ_VK); _, late = ctx.read_count()
        o = ctx.outputs()
        tex = sum(max(1, o.pyramid_width >> i) * max(1, o.pyramid_height >> i) for i in range(o.pyramid_mips))
        be, bl = B.algorithmic_bytes(n, vis, early, late)
        nbytes += be + bl + B.pyramid_bytes(W, H, tex)
        per_view.append([v, vis, early, late, int(cs.bad_object) + int(cs.bad_lod) + int(cs.unsorted)])
    allv = per_view
    if h.dist is not None:
        t = torch.full((VIEWS, 5), -1, device="cuda", dtype=torch.int64)
        for r in per_view:
            t[r[0]] = torch.tensor(r, device="cuda", dtype=torch.int64)
        h.dist.all_reduce(t, op=h.dist.ReduceOp.MAX)
        allv = [[int(x) for x in row] for row in t.tolist()]
    tot = h.sum_over_ranks([nbytes])[0]
    roof = {"bound": "hbm", "kernel": "two-phase frame per view (early_stream_kernel + pyramid_kernel at 4K + stream_cull_kernel<PASS_LATE>)",
            "achieved": tot / (ms * 1e-3) / 1e9 / h.world, "peak": h.peak, "peak_source": h.peak_src, "unit": "GB/s", "frac": tot / (ms * 1e-3) / 1e9 / h.world / h.peak,
            "algorithmic_bytes_per_launch": tot // VIEWS, "bytes_per_object": 48, "objects_per_launch": n, "launch_ms": ms / max(len(mine), 1), "traffic": None,
            "note": "per GPU: all frames of the step against the algorithmic bytes of the three kernels of every view"}
    cfg = {"workload": f"configs[4]: multi-view, {VIEWS} cameras / cascades (same position, yaw steps of 45 degrees, zFar doubled per cascade) x {n} objects, two-phase frustum + Hi-Z + LOD "
                       f"against a {W}x{H} synthetic depth image per view (2048x2048x11 pyramid), views sharded x{h.world} (scene replicated, no exchange)",
           "views": VIEWS, "objects_per_view": n, "depth": [W, H], "l2_policy": "inputs larger than L2", "parallelism": f"view-sharded x{h.world}"}
    detail = {"per_view [view, visible, early_draws, late_draws, checksum_errors]": allv}
    if h.rank == 0:
        print(json.dumps(h.line("view_objects_culled_per_s", "objects x views / s", VIEWS * n / (ms * 1e-3), ms, cfg, roof, clocks, launches, detail)), flush=True)
    for ctx in ctxs:
        ctx.close()
    h.finish()
    return 0


def run(args):
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm is defined for the contract workload only (--workload cfg2), not {args.workload}"}), flush=True)
        return 0
    h = Harness(args)
    return {"cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5}[args.workload](h)
