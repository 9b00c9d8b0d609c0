#!/usr/bin/env python
"""Device timings of the indirect-instancing and cluster kernels (BASELINE configs 3 / 4 kernels) on the bench scene, CUDA events on
the launching stream.  One JSON line per case with the SURVEY 8(d) algorithmic bytes:
    instanced        N*40 + s*4 + L*(8 + 32)
    cluster expand   N*40 + M*12
    cluster cull     passthrough M*12 + M*R;  sphere / sphere_hiz  M*12 + D*40 + s*R  (D bounded by the owning objects = frustum survivors)

    python scripts/inst_cluster_microbench.py [--objects N] [--iters K] [--dispatch-capacity M]
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
    ap.add_argument("--objects", type=int, default=B.N_OBJECTS)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dispatch-capacity", type=int, default=160_000_000)
    a = ap.parse_args()
    import torch
    from blitzen_b200 import capi, scene
    peak, _ = B.load_peaks()
    w = B.build_workload(a.objects, 0, 1)
    n = len(w["objs"])
    nl = len(w["lods"])
    stream = torch.cuda.Stream()

    def timed(ctx, fn):
        ts = []
        for it in range(a.iters + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts)), float(np.min(ts))

    def report(name, units, ms, mn, nbytes, extra=None):
        d = {"case": name, "units": int(units), "ms_mean": round(ms, 5), "ms_min": round(mn, 5), "units_per_s": units / (ms * 1e-3),
             "algorithmic_bytes": int(nbytes), "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak}
        d.update(extra or {})
        print(json.dumps(d), flush=True)

    cube = scene.cube_side(a.objects)
    half = cube / 2
    views = {"centre": scene.reference_views()[B.VIEW_NAME],
             "all_visible": scene.make_view((half, half, -4.0 * cube), z_far=1e9, width=1920, height=1080)}

    # ---- indirect instancing (drawInstCountReset / drawInstCull / drawInstCmd): bucket capacity = N per LOD would be 28 x N words;
    #      buckets are sized from a first counting run instead (capacity 1 -> counters only), as a harness would
    li = w["lodInstances"].copy()
    for vn, v in views.items():
        cap = np.full(nl, 1, dtype=np.uint32)
        li["instanceOffset"] = np.arange(nl, dtype=np.uint32)
        with capi.CullContext(0) as ctx:
            ctx.set_stream(stream.cuda_stream)
            ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], lod_instances=li, bucket_capacity=cap, transform_id_base=w["transform_id_base"])
            ctx.set_view(v)
            ctx.instanced()
            _, counters = ctx.read_instances(nl)
            cnt = counters["instanceCount"].astype(np.uint64)
        cap = np.maximum(cnt, 1).astype(np.uint32)
        li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
        with capi.CullContext(0) as ctx:
            ctx.set_stream(stream.cuda_stream)
            ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], lod_instances=li, bucket_capacity=cap, transform_id_base=w["transform_id_base"])
            ctx.set_view(v)
            ms, mn = timed(ctx, lambda: ctx.instanced())
            s = int(cnt.sum())
            report("instanced/" + vn, n, ms, mn, n * 40 + s * 4 + nl * 40, {"instances": s, "non_empty_lods": int((cnt > 0).sum())})

    # ---- cluster path (PreClusterDrawCull -> InitialClusterCull) ------------------------------------------------------------------------
    capm = a.dispatch_capacity
    with capi.CullContext(0) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], clusters=w["clusters"], transform_id_base=w["transform_id_base"],
                         cluster_dispatch_capacity=capm, draw_capacity=capm)
        ctx.set_depth(w["depth"])
        ctx.set_view(views["centre"])
        ctx.build_pyramid(capi.HIZ_VK)
        ctx.frustum_lod()
        _, surv = ctx.read_count()
        ms, mn = timed(ctx, lambda: ctx.cluster_expand())
        import ctypes as C
        mw, mt = C.c_uint32(), C.c_uint32()
        ctx._check(ctx._lib.blz_cull_read_cluster_dispatch(ctx._h, None, 0, C.byref(mw), C.byref(mt)))
        m_total = int(mt.value)
        m = min(m_total, capm)
        report("cluster_expand/centre", n, ms, mn, n * 40 + m * 12, {"records": m, "records_total": m_total, "frustum_survivors": surv})
        for mode, mname in ((capi.CLUSTER_PASSTHROUGH, "passthrough"), (capi.CLUSTER_SPHERE, "sphere"), (capi.CLUSTER_SPHERE_HIZ, "sphere_hiz")):
            ms, mn = timed(ctx, lambda: ctx.cluster_cull(mode, capi.REC_VK24, capi.HIZ_VK))
            _, tot = ctx.read_count()
            nbytes = m * 12 + m * 24 if mode == capi.CLUSTER_PASSTHROUGH else m * 12 + surv * 40 + tot * 24
            report("cluster_cull/" + mname, m, ms, mn, nbytes, {"draws": tot})


def configs345():
    """BASELINE configs 3-5 at their named sizes on ONE GPU (the multi-GPU forms shard these by object / record range / view):
    config 3: indirect instancing over 67 108 864 objects; config 4: 268 435 456 cluster dispatch records through the cluster cull;
    config 5: one view of the 16.7 M-object scene against a 3840x2160 depth image (2048x2048x11 pyramid)."""
    import torch
    from blitzen_b200 import capi, scene
    peak, _ = B.load_peaks()
    stream = torch.cuda.Stream()

    def timed(fn, iters=10):
        ts = []
        for it in range(iters + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts)), float(np.min(ts))

    def report(name, units, ms, mn, nbytes, extra=None):
        d = {"case": name, "units": int(units), "ms_mean": round(ms, 5), "ms_min": round(mn, 5), "units_per_s": units / (ms * 1e-3),
             "algorithmic_bytes": int(nbytes), "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak}
        d.update(extra or {})
        print(json.dumps(d), flush=True)

    # ---- config 5 first (same 16.7 M scene as the bench): 4K depth -> 2048 x 2048 x 11 pyramid, late pass against it -------------------
    w = B.build_workload(B.N_OBJECTS, 0, 1)
    n = len(w["objs"])
    depth4k = scene.synthetic_depth(3840, 2160)
    with capi.CullContext(0) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], transform_id_base=w["transform_id_base"])
        ctx.set_view(w["view"]); ctx.set_depth(depth4k)
        ms, mn = timed(lambda: ctx.build_pyramid(capi.HIZ_VK))
        o = ctx.outputs()
        tex = sum(max(1, o.pyramid_width >> i) * max(1, o.pyramid_height >> i) for i in range(o.pyramid_mips))
        report("config5/pyramid_4k", 3840 * 2160, ms, mn, 3840 * 2160 * 4 + tex * 4, {"pyramid": [o.pyramid_width, o.pyramid_height, o.pyramid_mips]})
        ctx.clear_pyramid(capi.HIZ_VK, 3840, 2160); ctx.reset_visibility(); ctx.late(capi.REC_VK24, capi.HIZ_VK)
        ctx.build_pyramid(capi.HIZ_VK); ctx.late(capi.REC_VK24, capi.HIZ_VK)
        ms, mn = timed(lambda: ctx.late(capi.REC_VK24, capi.HIZ_VK))
        _, tot = ctx.read_count()
        report("config5/late_steady_4k_pyramid", n, ms, mn, n * 48 + tot * 24 + 4, {"survivors": tot, "visible": int(ctx.read_visibility().sum())})
        ms, mn = timed(lambda: ctx.early(capi.REC_VK24))
        report("config5/early_steady", n, ms, mn, n * 4 + int(ctx.read_visibility().sum()) * 40 + 4)

    # ---- config 4: 268 435 456 dispatch records (the all-visible view expands to more; the dispatch capacity clamps) ----------------------
    capm = 268_435_456
    cube = scene.cube_side(n); half = cube / 2
    with capi.CullContext(0) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], clusters=w["clusters"], transform_id_base=w["transform_id_base"],
                         cluster_dispatch_capacity=capm, draw_capacity=capm)
        ctx.set_depth(w["depth"])
        ctx.set_view(scene.make_view((half, half, -1.2 * cube), z_far=1e9, width=1920, height=1080))
        ctx.build_pyramid(capi.HIZ_VK)
        ms, mn = timed(lambda: ctx.cluster_expand(), 5)
        import ctypes as C
        mw, mt = C.c_uint32(), C.c_uint32()
        ctx._check(ctx._lib.blz_cull_read_cluster_dispatch(ctx._h, None, 0, C.byref(mw), C.byref(mt)))
        m = int(mw.value)
        ctx.frustum_lod(); _, surv = ctx.read_count()
        report("config4/cluster_expand", n, ms, mn, n * 40 + m * 12, {"records_written": m, "records_total": int(mt.value), "frustum_survivors": surv})
        for mode, mname in ((capi.CLUSTER_PASSTHROUGH, "passthrough"), (capi.CLUSTER_SPHERE, "sphere"), (capi.CLUSTER_SPHERE_HIZ, "sphere_hiz")):
            ms, mn = timed(lambda: ctx.cluster_cull(mode, capi.REC_VK24, capi.HIZ_VK), 5)
            _, tot = ctx.read_count()
            nbytes = m * 12 + m * 24 if mode == capi.CLUSTER_PASSTHROUGH else m * 12 + surv * 40 + tot * 24
            report("config4/cluster_cull_" + mname, m, ms, mn, nbytes, {"draws": tot})
    del w

    # ---- config 3: indirect instancing over 67 108 864 objects ---------------------------------------------------------------------------------
    n3 = 67_108_864
    w = B.build_workload(n3, 0, 1)
    nl = len(w["lods"])
    li = w["lodInstances"].copy()
    cube = scene.cube_side(n3); half = cube / 2
    for vn, v in (("centre", scene.make_view((half, half, half), z_far=650.0, width=1920, height=1080)),
                  ("all_visible", scene.make_view((half, half, -4.0 * cube), z_far=1e9, width=1920, height=1080))):
        cap = np.full(nl, 1, dtype=np.uint32)
        li["instanceOffset"] = np.arange(nl, dtype=np.uint32)
        with capi.CullContext(0) as ctx:
            ctx.set_stream(stream.cuda_stream)
            ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], lod_instances=li, bucket_capacity=cap, transform_id_base=w["transform_id_base"])
            ctx.set_view(v); ctx.instanced()
            _, counters = ctx.read_instances(nl)
            cnt = counters["instanceCount"].astype(np.uint64)
            cap = np.maximum(cnt, 1).astype(np.uint32)
            li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
            ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], lod_instances=li, bucket_capacity=cap, transform_id_base=w["transform_id_base"])
            ctx.set_view(v)
            ms, mn = timed(lambda: ctx.instanced(), 5)
            s_ = int(cnt.sum())
            report("config3/instanced_64M/" + vn, n3, ms, mn, n3 * 40 + s_ * 4 + nl * 40, {"instances": s_, "non_empty_lods": int((cnt > 0).sum()), "largest_bucket": int(cnt.max()),
                                                                                             "note": "the reference's fixed bucket of 100 000 per LOD would overflow; buckets sized from a counting run"})


if __name__ == "__main__":
    if "--configs345" in sys.argv:
        sys.argv.remove("--configs345")
        configs345()
    else:
        main()
