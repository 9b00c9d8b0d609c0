#!/usr/bin/env python
"""Per-pass device timings of the cull kernels on the bench workload (16.7 M objects unless --objects), CUDA events on the
launching stream, inputs larger than L2.  One JSON line per case: ms (mean / min), objects/s, algorithmic GB/s, fraction
of the measured HBM peak.  Used for the tables in DESIGN.md / BASELINE.md; bench.py remains the contract line.

    python scripts/cull_microbench.py [--objects N] [--iters K] [--cases late,early,frustum,...]
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
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--cases", default="all")
    a = ap.parse_args()
    import torch
    from blitzen_b200 import capi, scene
    peak, _ = B.load_peaks()
    w = B.build_workload(a.objects, 0, 1)
    n = len(w["objs"])
    views = scene.reference_views()
    stream = torch.cuda.Stream()
    ctx = capi.CullContext(0)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], object_id_base=0, transform_id_base=w["transform_id_base"])
    ctx.set_depth(w["depth"])

    def timed(fn, prep=None):
        ts = []
        for it in range(a.iters + 3):
            if prep:
                prep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts)), float(np.min(ts))

    def report(name, ms, mn, nbytes, extra=None):
        d = {"case": name, "objects": n, "ms_mean": round(ms, 5), "ms_min": round(mn, 5), "objects_per_s": n / (ms * 1e-3),
             "algorithmic_bytes": int(nbytes), "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak}
        d.update(extra or {})
        print(json.dumps(d), flush=True)

    want = None if a.cases == "all" else set(a.cases.split(","))
    on = lambda c: want is None or c in want

    # views over the 16 M scene: the bench view (centre, ~4 % visible), a corner view, and everything-visible
    cube = scene.cube_side(a.objects)
    half = cube / 2
    vs = {"centre": views[B.VIEW_NAME],
          "all_visible": scene.make_view((half, half, -4.0 * cube), z_far=1e9, width=1920, height=1080),
          "none": scene.make_view((half, half, 4.0 * cube), z_far=10.0, width=1920, height=1080)}

    for vn, v in vs.items():
        if not on("frustum_" + vn):
            continue
        ctx.set_view(v)
        ms, mn = timed(lambda: ctx.frustum_lod())
        _, tot = ctx.read_count()
        report("frustum_lod/" + vn, ms, mn, n * 40 + tot * 24 + 4, {"survivors": tot})

    for variant, vname in ((capi.HIZ_VK, "vk"), (capi.HIZ_DX, "dx")):
        if not (on("late_" + vname) or on("early_" + vname) or on("frame0_" + vname)):
            continue
        ctx.set_view(vs["centre"])
        ctx.build_pyramid(variant)
        # frame 0: visibility 0, cleared pyramid -> every frustum survivor goes through Hi-Z and is emitted
        if on("frame0_" + vname):
            def prep0():
                ctx.reset_visibility(); ctx.clear_pyramid(variant, B.DEPTH_W, B.DEPTH_H)
            ms, mn = timed(lambda: ctx.late(capi.REC_VK24, variant), prep0)
            _, tot = ctx.read_count()
            report(f"late_frame0/{vname}", ms, mn, n * 48 + tot * 24 + 4, {"survivors": tot})
        # steady state
        ctx.reset_visibility(); ctx.clear_pyramid(variant, B.DEPTH_W, B.DEPTH_H); ctx.late(capi.REC_VK24, variant)
        ctx.build_pyramid(variant); ctx.late(capi.REC_VK24, variant)
        vis = int(ctx.read_visibility().sum())
        if on("late_" + vname):
            ms, mn = timed(lambda: ctx.late(capi.REC_VK24, variant))
            _, tot = ctx.read_count()
            report(f"late_steady/{vname}", ms, mn, n * 48 + tot * 24 + 4, {"survivors": tot, "visible": vis})
        if on("early_" + vname):
            for mode, mname in ((1, "pipelined"), (0, "streaming")):
                ctx.set_option("early_mode", mode)
                ms, mn = timed(lambda: ctx.early(capi.REC_VK24))
                _, tot = ctx.read_count()
                report(f"early_steady/{vname}/{mname}", ms, mn, n * 4 + vis * 40 + tot * 24 + 4, {"survivors": tot, "visible_prev": vis})
            if vname == "vk":
                keep = ctx.read_visibility().copy()
                rng = np.random.default_rng(1)
                for frac in (0.25, 1.0):
                    vv = (rng.random(n) < frac).astype(np.uint32)
                    ctx.write_visibility(vv)
                    nv = int(vv.sum())
                    for mode, mname in ((1, "pipelined"), (0, "streaming")):
                        ctx.set_option("early_mode", mode)
                        ms, mn = timed(lambda: ctx.early(capi.REC_VK24))
                        _, tot = ctx.read_count()
                        report(f"early_vis{int(frac * 100)}pct/{mname}", ms, mn, n * 4 + nv * 40 + tot * 24 + 4, {"survivors": tot, "visible_prev": nv})
                ctx.write_visibility(keep)
            ctx.set_option("early_mode", 1)
        if on("pyramid_" + vname):
            ms, mn = timed(lambda: ctx.build_pyramid(variant))
            o = ctx.outputs()
            tex = sum(max(1, o.pyramid_width >> i) * max(1, o.pyramid_height >> i) for i in range(o.pyramid_mips))
            report(f"pyramid/{vname}", ms, mn, B.DEPTH_W * B.DEPTH_H * 4 + tex * 4, {"pyramid": [o.pyramid_width, o.pyramid_height, o.pyramid_mips]})
    ctx.close()


if __name__ == "__main__":
    main()
