#!/usr/bin/env python
"""A/B sweep of the draw-cull kernel variants on the bench workload: for every option set (';'-separated list of
'k=v,k=v') times frustum+LOD, the steady-state late pass, frame 0 of the late pass and the early pass with CUDA events, and checks
that draws / counts / visibility are byte-identical to the FIRST option set's (the baseline variant).

    python scripts/kernel_sweep.py --opts "stream_cfg=2;stream_cfg=3;stream_cfg=6"
"""
import argparse
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
    ap.add_argument("--objects", type=int, default=B.N_OBJECTS)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--opts", default="stream_cfg=2;stream_cfg=6")
    ap.add_argument("--cases", default="frustum,late,frame0,early")
    a = ap.parse_args()
    import torch
    from blitzen_b200 import capi, scene
    peak, _ = B.load_peaks()
    w = B.build_workload(a.objects, 0, 1)
    n = len(w["objs"])
    view = scene.reference_views()[B.VIEW_NAME]
    stream = torch.cuda.Stream()
    ctx = capi.CullContext(0)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], object_id_base=0, transform_id_base=w["transform_id_base"])
    ctx.set_depth(w["depth"])
    ctx.set_view(view)
    variant = capi.HIZ_VK
    cases = set(a.cases.split(","))

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

    def digest():
        d, tot = ctx.read_draws()
        return hashlib.sha256(d.tobytes()).hexdigest()[:16], int(tot)

    # steady-state visibility, produced once by the baseline variant
    base = None
    vis_steady = None
    for oi, optset in enumerate(a.opts.split(";")):
        for kv in filter(None, optset.split(",")):
            k, v = kv.split("=")
            ctx.set_option(k.strip(), int(v))
        out = {"opts": optset}
        sig = {}
        if vis_steady is None:
            ctx.reset_visibility(); ctx.clear_pyramid(variant, B.DEPTH_W, B.DEPTH_H); ctx.late(capi.REC_VK24, variant)
            ctx.build_pyramid(variant); ctx.late(capi.REC_VK24, variant)
            vis_steady = ctx.read_visibility().copy()
        if "frustum" in cases:
            ms, mn = timed(lambda: ctx.frustum_lod())
            sig["frustum"] = digest()
            nb = n * 40 + sig["frustum"][1] * 24 + 4
            out["frustum_ms"] = round(ms, 4); out["frustum_min"] = round(mn, 4); out["frustum_frac"] = round(nb / (ms * 1e-3) / 1e9 / peak, 3)
        if "late" in cases:
            ctx.build_pyramid(variant)
            ctx.write_visibility(vis_steady)
            ms, mn = timed(lambda: ctx.late(capi.REC_VK24, variant))
            sig["late"] = digest()
            sig["late_vis"] = hashlib.sha256(ctx.read_visibility().tobytes()).hexdigest()[:16]
            nb = n * 48 + sig["late"][1] * 24 + 4
            out["late_ms"] = round(ms, 4); out["late_min"] = round(mn, 4); out["late_frac"] = round(nb / (ms * 1e-3) / 1e9 / peak, 3)
        if "frame0" in cases:
            def prep0():
                ctx.reset_visibility(); ctx.clear_pyramid(variant, B.DEPTH_W, B.DEPTH_H)
            ms, mn = timed(lambda: ctx.late(capi.REC_VK24, variant), prep0)
            sig["frame0"] = digest()
            nb = n * 48 + sig["frame0"][1] * 24 + 4
            out["frame0_ms"] = round(ms, 4); out["frame0_frac"] = round(nb / (ms * 1e-3) / 1e9 / peak, 3)
        if "early" in cases:
            ctx.write_visibility(vis_steady)
            ms, mn = timed(lambda: ctx.early(capi.REC_VK24))
            sig["early"] = digest()
            nv = int(vis_steady.sum())
            nb = n * 4 + nv * 40 + sig["early"][1] * 24 + 4
            out["early_ms"] = round(ms, 4); out["early_min"] = round(mn, 4); out["early_frac"] = round(nb / (ms * 1e-3) / 1e9 / peak, 3)
        if base is None:
            base = sig
            out["survivors"] = {k: v[1] for k, v in sig.items() if isinstance(v, tuple)}
        out["sig"] = hashlib.sha256(json.dumps(sig, sort_keys=True).encode()).hexdigest()[:12]
        out["identical_to_first"] = (sig == base)
        if sig != base:
            out["diff"] = {k: (sig[k], base[k]) for k in sig if sig[k] != base[k]}
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
