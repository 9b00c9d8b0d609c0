#!/usr/bin/env python
"""bench.py -- Blitzen cull dispatch on B200: objects culled per second + fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--objects M]

Workload (BASELINE.json configs[1]): a stress scene with the reference's mesh mix replicated to 16 777 216 objects,
two-phase frustum + Hi-Z occlusion cull against a 1920x1080 synthetic depth image (Vulkan-variant pyramid), reference
camera view "cfg2_centre_1080p".  One step = one frame of the cull path:
    early pass (InitialDrawCull)  ->  single-pass Hi-Z pyramid build  ->  late pass (LateDrawCull)
With N > 1 every rank owns a 16.7 M-object shard of an N x 16.7 M-object scene (weak scaling, no data-path collective);
the per-rank early and late draw lists are pushed into the presenting rank's buffer over NVLink peer memory inside the step.

--impl reference times the CPU restatement of the reference's cull shaders (oracle/, all host threads) on the same
workload: the reference has no CPU or CUDA implementation of this path (it exists only as GLSL/HLSL), so the multithreaded
transliteration is the reference arm (cpu_baseline.kind = "port").
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBJECTS = 16_777_216
DEPTH_W, DEPTH_H = 1920, 1080
VIEW_NAME = "cfg2_centre_1080p"
METRIC = "objects_culled_per_s"
UNIT = "objects/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=N_OBJECTS, help="objects per GPU (default: the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0, help="objects in the cpu_baseline sample (0 = the whole workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the extra device timings of the late kernel's other output regimes")
    ap.add_argument("--hiz", default="vk", choices=["vk", "dx"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1] (default, the contract line) or configs[2..4]: instancing 64 M / clusters 256 M records / 8 views x 16.7 M at 4K (bench_configs.py)")
    ap.add_argument("--depth", default="synthetic", choices=["synthetic", "raster"],
                    help="synthetic: the seeded 64-rectangle depth image of BASELINE config 2; raster: software depth splatted from the early list every frame (closed loop, csrc/raster_depth.cu)")
    ap.add_argument("--records", type=int, default=0, help="cfg4: cluster dispatch records in total (default 268 435 456)")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay measurement of the frame")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_workload(n_per_gpu, rank, world):
    """Objects [rank*n, (rank+1)*n) of a world*n-object stress scene + the transforms they reference."""
    from blitzen_b200 import scene
    total = n_per_gpu * world
    groups = scene.scaled_groups(total - 1001)
    mult = scene.cube_side(n_per_gpu)           # the per-GPU shard keeps the 16 M-object density and extent
    objs, xf = scene.generate(groups, mult, True, "counter", seed=2, first=rank * n_per_gpu, count=n_per_gpu, threads=os.cpu_count() or 8)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    view = scene.reference_views()[VIEW_NAME]
    depth = scene.synthetic_depth(DEPTH_W, DEPTH_H)
    return dict(objs=objs, transforms=transforms, transform_id_base=tbase, object_id_base=rank * n_per_gpu, view=view, depth=depth, **tables)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def algorithmic_bytes(n, vis_prev, early_draws, late_draws, rec_bytes=24):
    """BASELINE.md section 3."""
    early = n * 4 + vis_prev * 40 + early_draws * rec_bytes + 4
    late = n * 48 + late_draws * rec_bytes + 4
    return early, late


def pyramid_bytes(depth_w, depth_h, texels):
    return depth_w * depth_h * 4 + texels * 4


def run_cpu_frame(O, w, vis, pyr_variant, threads, n_sample=None):
    """One frame of the path on the CPU oracle: early + pyramid + late.  Returns seconds (wall)."""
    n = len(w["objs"]) if not n_sample else min(n_sample, len(w["objs"]))
    objs = w["objs"][:n]
    t0 = time.perf_counter()
    O.cull(objs, w["transforms"], w["surfaces"], w["lods"], w["view"], O.PASS_EARLY, vis=vis[:n], threads=threads, transform_id_base=w["transform_id_base"], object_id_base=w["object_id_base"])
    pyr = O.build_pyramid(w["depth"], pyr_variant, threads=threads)
    _, _, vis_out = O.cull(objs, w["transforms"], w["surfaces"], w["lods"], w["view"], O.PASS_LATE, hiz=pyr_variant, pyramid=pyr, vis=vis[:n], threads=threads,
                           transform_id_base=w["transform_id_base"], object_id_base=w["object_id_base"])
    return time.perf_counter() - t0, vis_out, n


def steady_visibility(O, w, variant, threads):
    """Visibility after a warm-up frame from an all-zero buffer (frame 0: cleared pyramid; frame 1: real pyramid)."""
    n = len(w["objs"])
    vis = np.zeros(n, dtype=np.uint32)
    kw = dict(threads=threads, transform_id_base=w["transform_id_base"], object_id_base=w["object_id_base"])
    _, _, vis = O.cull(w["objs"], w["transforms"], w["surfaces"], w["lods"], w["view"], O.PASS_LATE, hiz=variant, pyramid=O.cleared_pyramid(DEPTH_W, DEPTH_H, variant), vis=vis, **kw)
    pyr = O.build_pyramid(w["depth"], variant, threads=threads)
    _, _, vis = O.cull(w["objs"], w["transforms"], w["surfaces"], w["lods"], w["view"], O.PASS_LATE, hiz=variant, pyramid=pyr, vis=vis, **kw)
    return vis


def reference_arm(args):
    """CPU arm: the multithreaded transliteration of the reference's cull shaders, all host threads, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    threads = O.hardware_threads()
    variant = 0 if args.hiz == "vk" else 1
    w = build_workload(args.objects, 0, 1)
    n = len(w["objs"])
    vis = steady_visibility(O, w, variant, threads)
    times = []
    for i in range(args.warmup + args.steps):
        dt, vis2, _ = run_cpu_frame(O, w, vis.copy(), variant, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"whole workload: {n} objects, early + pyramid + late per step, {threads} std::threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, world):
    return {"workload": "configs[1]: stress-scene mesh mix replicated to 16777216 objects, two-phase frustum + Hi-Z (" + args.hiz.upper() +
                        " variant) + LOD cull, 1920x1080 synthetic depth, view " + VIEW_NAME,
            "objects_per_gpu": args.objects, "objects_total": args.objects * world, "depth": [DEPTH_W, DEPTH_H],
            "step": "early pass + " + ("software depth from the early list + " if getattr(args, "depth", "synthetic") == "raster" else "") + "Hi-Z pyramid build + late pass" + (" + peer-memory draw-list gather (early and late lists, pushed on a side stream behind each pass)" if world > 1 else ""),
            "record_format": "VK24", "l2_policy": "inputs larger than L2 (object + transform streams = 40 B x 16.7 M = 671 MB >> 126 MB)",
            "parallelism": f"object-sharded x{world}"}


def main():
    args = parse_args()
    if args.workload != "cfg2":
        import bench_configs
        return bench_configs.run(args)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from blitzen_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the cull path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    variant = capi.HIZ_VK if args.hiz == "vk" else capi.HIZ_DX
    w = build_workload(args.objects, rank, world)
    n = len(w["objs"])
    stream = torch.cuda.Stream()
    ctx = capi.CullContext(local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], object_id_base=w["object_id_base"], transform_id_base=w["transform_id_base"])
    ctx.set_view(w["view"])
    ctx.set_depth(w["depth"])

    gather = None
    if world > 1:
        from blitzen_b200 import dist as bdist
        gather = bdist.DrawListGather(ctx, rank, world, capacity_records=args.objects * world // 4, fmt=capi.REC_VK24)

    epoch = [0]

    raster = args.depth == "raster"
    # where the early list's push starts: behind the pyramid launch (the push then overlaps only the 175 us late pass and the 15 us pyramid
    # build keeps the SMs' copy/store paths to itself).  Measured at 2 / 4 / 8 GPUs against starting it right behind the early pass
    # (profiles/r02h_n8_ab.txt): better or equal everywhere once the wait for a draw buffer's previous push sits in the pass that writes it.
    # At 8 GPUs the frame is bound by the presenter's ingest either way.  BLZ_PUSH_AFTER=early|pyramid overrides.
    push_early_first = os.environ.get("BLZ_PUSH_AFTER", "early" if world > 4 else "pyramid") == "early"

    # The early list is pushed by a throttled number of small co-resident CTAs (csrc/gather.cu) on a side stream; see push_early_first.
    def frame():
        ctx.early(capi.REC_VK24)
        if raster:
            ctx.raster_depth(DEPTH_W, DEPTH_H)        # reads the early list: before the push flips the draw buffers
        if gather and push_early_first:
            epoch[0] += 1; gather.push_async(epoch[0])
        ctx.build_pyramid(variant)
        if gather and not push_early_first:
            epoch[0] += 1; gather.push_async(epoch[0])
        ctx.late(capi.REC_VK24, variant)
        if gather:
            epoch[0] += 1; gather.push_async(epoch[0])

    # frame 0 (cleared pyramid, visibility 0) then warm-up frames: establishes the steady-state visibility buffer
    ctx.clear_pyramid(variant, DEPTH_W, DEPTH_H)
    ctx.late(capi.REC_VK24, variant)
    for _ in range(max(args.warmup, 3)):
        frame()
    ctx.synchronize()
    vis_prev = int(ctx.read_visibility().sum())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K frames, CUDA events on the launching stream, per-kernel events inside -------------------
    K = args.steps
    EV_EVERY = 4
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        if dist is not None:
            # the host-side barrier above leaves the ranks' STREAMS up to a millisecond apart (host wake-up skew), which a 20-step run would
            # charge to the gather (a rank's push waits for the lower ranks' counts): one collective ON the launching stream lines the
            # devices up to within microseconds right before the first timed event.  The timed region is still exactly K steps.
            align = torch.zeros(1, device="cuda")
            dist.all_reduce(align)
        t_start.record(stream)
        for k in range(K):
            # per-kernel events on every 4th step only: an event record costs the stream ~3 us (measured: 4 per step made the step
            # 0.2486 ms, none 0.2357 ms), which is launch overhead of the harness, not of the path
            timed_k = (k % EV_EVERY) == 0
            if timed_k: ev[k][0].record(stream)
            ctx.early(capi.REC_VK24)
            if timed_k: ev[k][1].record(stream)
            if raster:
                ctx.raster_depth(DEPTH_W, DEPTH_H)      # timed with the pyramid: "depth + pyramid"
            if gather and push_early_first:
                epoch[0] += 1; gather.push_async(epoch[0])
            ctx.build_pyramid(variant)
            if timed_k: ev[k][2].record(stream)
            if gather and not push_early_first:
                epoch[0] += 1; gather.push_async(epoch[0])
            ctx.late(capi.REC_VK24, variant)
            if timed_k: ev[k][3].record(stream)
            if gather:
                epoch[0] += 1; gather.push_async(epoch[0])
        if gather:
            ctx.gather_join()            # the timed region ends when the last side-stream push has landed on the presenter
        t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    ks = range(0, K, EV_EVERY)
    t_early = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in ks]))
    t_pyr = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in ks]))
    t_late = float(np.mean([ev[k][2].elapsed_time(ev[k][3]) for k in ks]))
    if dist is not None:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / K
    value = (n * world) / (ms_per_step * 1e-3)

    # ---- one more frame, untimed: counts, the early list's device-side checksum and -- at N > 1 -- the proof that the list GATHERED on the
    #      presenter is the concatenation of the ranks' lists: the presenter reduces the gathered buffer on the device (csrc/consume.cu) and the
    #      result must equal the all-reduced sum (xor for id_xor) of the ranks' own reductions, with no pair of records out of ascending order
    gather_ok = None
    gsum = []

    def gathered_check(tag):
        if not gather:
            return
        mine = ctx.consume_draws(kind=1)
        epoch[0] += 1
        gather.push(epoch[0])
        t = torch.tensor([int(mine.records), int(mine.index_sum), int(mine.id_sum)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        x = torch.tensor([int(mine.id_xor) & 0x7FFFFFFFFFFFFFFF, int(mine.id_xor) >> 63], device="cuda", dtype=torch.int64)
        xs = [torch.zeros_like(x) for _ in range(world)]
        dist.all_gather(xs, x)
        if rank == 0:
            g = ctx.consume_gathered(epoch[0])
            xo = 0
            for v in xs:
                xo ^= int(v[0].item()) | (int(v[1].item()) << 63)
            same = (int(g.records), int(g.index_sum), int(g.id_sum)) == tuple(int(v) for v in t.tolist()) and int(g.id_xor) == xo and int(g.unsorted) == 0
            gsum.append({"list": tag, "records": int(g.records), "index_sum": int(g.index_sum), "id_xor": "%016x" % int(g.id_xor), "unsorted": int(g.unsorted), "equals_sum_of_ranks": bool(same)})

    early_written, early_total = 0, 0
    ctx.early(capi.REC_VK24); early_written, early_total = ctx.read_count()
    cs = ctx.consume_draws()          # what the indirect draw + vertex stage would read from this list, reduced on the device (csrc/consume.cu)
    early_checksum = {"records": int(cs.records), "index_sum": int(cs.index_sum), "id_xor": "%016x" % int(cs.id_xor), "bad_object": int(cs.bad_object), "bad_lod": int(cs.bad_lod), "unsorted": int(cs.unsorted)}
    if raster:
        ctx.raster_depth(DEPTH_W, DEPTH_H)
    gathered_check("early")
    ctx.build_pyramid(variant); ctx.late(capi.REC_VK24, variant); late_written, late_total = ctx.read_count()
    gathered_check("late")
    if gather and rank == 0:
        gather_ok = all(x["equals_sum_of_ranks"] for x in gsum)
    o = ctx.outputs()
    pyr_texels = sum(max(1, o.pyramid_width >> i) * max(1, o.pyramid_height >> i) for i in range(o.pyramid_mips))

    peak, peak_src = load_peaks()

    # ---- the same late kernel in its other output regimes (device-timed, NOT part of the step): frame 0 (every frustum survivor is
    #      emitted) and an all-visible frustum pass (the output-bound end).  Reported as fractions of the same peak. ------------------
    regimes = None
    if world == 1 and not args.no_regimes:
        from blitzen_b200 import scene as _scene

        def timed(fn, prep=None, iters=10):
            ts = []
            for it in range(iters + 2):
                if prep:
                    prep()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream); fn(); e1.record(stream)
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts))

        vis_keep = ctx.read_visibility().copy()

        def prep0():
            ctx.reset_visibility(); ctx.clear_pyramid(variant, DEPTH_W, DEPTH_H)
        ms0 = timed(lambda: ctx.late(capi.REC_VK24, variant), prep0)
        _, tot0 = ctx.read_count()
        cube = _scene.cube_side(args.objects)
        ctx.set_view(_scene.make_view((cube / 2, cube / 2, -4.0 * cube), z_far=1e9, width=DEPTH_W, height=DEPTH_H))
        msa = timed(lambda: ctx.frustum_lod())
        _, tota = ctx.read_count()
        ctx.set_view(w["view"])
        msf = timed(lambda: ctx.frustum_lod())
        _, totf = ctx.read_count()
        ctx.build_pyramid(variant); ctx.write_visibility(vis_keep)
        frac = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9 / peak
        regimes = {"late_frame0": {"ms": ms0, "survivors": int(tot0), "frac": frac(n * 48 + tot0 * 24 + 4, ms0)},
                   "frustum_lod_bench_view": {"ms": msf, "survivors": int(totf), "frac": frac(n * 40 + totf * 24 + 4, msf)},
                   "frustum_lod_all_visible": {"ms": msa, "survivors": int(tota), "frac": frac(n * 40 + tota * 24 + 4, msa)}}

    b_early, b_late = algorithmic_bytes(n, vis_prev, early_total, late_total)
    b_pyr = pyramid_bytes(DEPTH_W, DEPTH_H, pyr_texels)
    late_gbs = b_late / (t_late * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "stream_cull_kernel<PASS_LATE> (late pass: frustum + Hi-Z + LOD + compaction + visibility write)",
                "achieved": late_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": late_gbs / peak,
                "algorithmic_bytes_per_launch": b_late, "bytes_per_object": 48, "objects_per_launch": n, "launch_ms": t_late, "launch_ms_sampled_steps": len(list(ks)),
                "traffic": None,
                "other_kernels": {"early": {"ms": t_early, "algorithmic_bytes": b_early, "GBps": b_early / (t_early * 1e-3) / 1e9},
                                  "pyramid": {"ms": t_pyr, "algorithmic_bytes": b_pyr, "GBps": b_pyr / (t_pyr * 1e-3) / 1e9}},
                "frame_GBps": (b_early + b_pyr + b_late) / (ms_per_step * 1e-3) / 1e9}
    prof = os.path.join(ROOT, "profiles", "late_traffic.json")
    if os.path.exists(prof):
        try:
            import hashlib
            with open(prof) as f:
                tr = json.load(f)
            hh = hashlib.sha256()
            for fn in tr.get("kernel_source_files", []):
                with open(os.path.join(ROOT, fn), "rb") as f:
                    hh.update(f.read())
            same = hh.hexdigest() == tr.get("kernel_source_sha256")
            roofline["traffic"] = tr.get("dram_bytes_per_launch") if same else None      # an ncu figure of ANOTHER version of the kernel is not reported
            roofline["traffic_source"] = tr.get("source") if same else "stale: the kernel source changed since the ncu capture in profiles/late_traffic.json"
        except Exception:
            pass

    # ---- the same 3-launch frame replayed from a CUDA graph (SURVEY 8b "CUDA-graph the frame"): device time per frame and the host time it
    #      takes to enqueue a frame, against plain stream launches.  Measured, not assumed: the frame is three launches. ------------------
    graph_info = None
    if world == 1 and not args.no_graph:
        try:
            def plain():
                ctx.early(capi.REC_VK24)
                if raster:
                    ctx.raster_depth(DEPTH_W, DEPTH_H)
                ctx.build_pyramid(variant); ctx.late(capi.REC_VK24, variant)
            for _ in range(3):
                plain()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                plain()
            torch.cuda.synchronize()
            R = 50

            def measure(fn):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    t0 = time.perf_counter()
                    for _ in range(R):
                        fn()
                    host = time.perf_counter() - t0
                    e1.record(stream)
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / R, host / R * 1e6
            for _ in range(5):
                g.replay()
            s_ms, s_us = measure(plain)
            g_ms, g_us = measure(g.replay)
            graph_info = {"stream_launches": {"device_ms_per_frame": s_ms, "host_enqueue_us_per_frame": s_us},
                          "graph_replay": {"device_ms_per_frame": g_ms, "host_enqueue_us_per_frame": g_us}, "frames": R}
        except Exception as ex:          # a capture-illegal call inside the library would surface here
            graph_info = {"error": str(ex)[:300]}
            torch.cuda.synchronize()

    # ---- end to end through the C ABI with HOST buffers: scene upload + depth upload + frame + draw-list read-back --------
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory()
        h_objs, h_xf, h_depth = pin(w["objs"]), pin(w["transforms"]), pin(w["depth"])
        h_draws = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
        vis_steady = ctx.read_visibility().copy()
        h_vis = pin(vis_steady)
        from blitzen_b200 import types as T
        objs_v = h_objs.numpy().view(T.RenderObject); xf_v = h_xf.numpy().view(T.MeshTransform)
        depth_v = h_depth.numpy().view(np.float32).reshape(DEPTH_H, DEPTH_W)
        draws_v = h_draws.numpy()
        import ctypes as C
        times = []
        h2d = objs_v.nbytes + xf_v.nbytes + w["surfaces"].nbytes + w["lods"].nbytes + depth_v.nbytes + h_vis.numel() + 256
        d2h = 0
        for it in range(args.e2e_steps + 2):
            barrier()
            t0 = time.perf_counter()
            ctx.upload_scene(objs_v, xf_v, w["surfaces"], w["lods"], object_id_base=w["object_id_base"], transform_id_base=w["transform_id_base"])
            ctx.write_visibility(h_vis.numpy().view(np.uint32))     # upload_scene resets the per-object state: last frame's visibility travels with the scene
            ctx.set_view(w["view"])
            ctx.set_depth(depth_v)
            ctx.early(capi.REC_VK24)
            wv, tv = C.c_uint32(), C.c_uint32()
            ctx._check(ctx._lib.blz_cull_read_draws(ctx._h, capi.REC_VK24, C.c_void_p(draws_v.ctypes.data), n, C.byref(wv), C.byref(tv)))
            ctx.build_pyramid(variant)
            ctx.late(capi.REC_VK24, variant)
            wl, tl = C.c_uint32(), C.c_uint32()
            ctx._check(ctx._lib.blz_cull_read_draws(ctx._h, capi.REC_VK24, C.c_void_p(draws_v.ctypes.data), n, C.byref(wl), C.byref(tl)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            d2h = (wv.value + wl.value) * 24 + 16
            if it >= 2:
                times.append(dt)
        e2e_ms = 1e3 * float(np.mean(times))
        if dist is not None:
            tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        e2e = {"value": (n * world) / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
               "what": "blz_cull_upload_scene (pinned host arrays) + write_visibility + set_view + set_depth + early + read_draws + build_pyramid + late + read_draws, host wall clock"}

    # ---- second end-to-end figure: the reference's OWN per-frame host traffic (UpdateBuffers, BlitzenVulkan/vulkanDraw.cpp:46-60: view
    #      block + the dynamic transforms [0, 1000); the scene itself was uploaded once by SetupForRendering) + draw-list read-backs ------
    e2e_frame = None
    if not args.no_e2e:
        ndyn = min(1000, len(xf_v))
        ctx.write_visibility(h_vis.numpy().view(np.uint32))
        times = []
        for it in range(args.e2e_steps * 4 + 2):
            barrier()
            t0 = time.perf_counter()
            ctx.update_transforms(w["transform_id_base"], xf_v[:ndyn])       # the first 1000 transforms this rank owns (global ids)
            ctx.set_view(w["view"])
            ctx.early(capi.REC_VK24)
            wv, tv = C.c_uint32(), C.c_uint32()
            ctx._check(ctx._lib.blz_cull_read_draws(ctx._h, capi.REC_VK24, C.c_void_p(draws_v.ctypes.data), n, C.byref(wv), C.byref(tv)))
            ctx.build_pyramid(variant)
            ctx.late(capi.REC_VK24, variant)
            wl, tl = C.c_uint32(), C.c_uint32()
            ctx._check(ctx._lib.blz_cull_read_draws(ctx._h, capi.REC_VK24, C.c_void_p(draws_v.ctypes.data), n, C.byref(wl), C.byref(tl)))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= 2:
                times.append(dt)
        f_ms = 1e3 * float(np.mean(times))
        if dist is not None:
            tt = torch.tensor([f_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            f_ms = float(tt.item())
        e2e_frame = {"value": (n * world) / (f_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(ndyn * 32 + 256), "d2h_bytes_per_step": int((wv.value + wl.value) * 24 + 16),
                     "ms_per_step": f_ms, "what": "the reference's per-frame host traffic only (UpdateBuffers: view block + dynamic transforms [0,1000)), scene resident since "
                     "blz_cull_upload_scene; early + read_draws + build_pyramid + late + read_draws, host wall clock. NOT the headline: `e2e` re-uploads the whole scene every step"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores -----------------------------------------------
    cpu = None
    census = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        threads = O.hardware_threads()
        vis_s = ctx.read_visibility().copy()
        ts = []
        ns = n
        for i in range(4):
            dt, _, ns = run_cpu_frame(O, w, vis_s.copy(), 0 if args.hiz == "vk" else 1, threads, args.cpu_sample or None)
            if i >= 1:
                ts.append(dt)
        cpu = {"value": ns / float(np.median(ts)), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{ns} of {n} objects, early + pyramid + late per step, median of 3 after 1 warm-up, oracle/cull_oracle.cpp with {threads} std::threads"}
        # north_star: objects within epsilon of a frustum plane / Hi-Z texel or level boundary "are counted and reported".  These are the
        # only objects on which a real GLSL/HLSL GPU (FMA contraction, 8-bit sampler weights) may legitimately differ from the oracle;
        # the CUDA path itself is bit-exact against the oracle on all of them (tests/).
        hv = 0 if args.hiz == "vk" else 1
        cen = O.boundary_census(w["objs"][:ns], w["transforms"], w["surfaces"], w["lods"], w["view"], O.build_pyramid(w["depth"], hv, threads=threads), hv,
                                ulp_tol=4.0, texel_tol=1.0 / 256.0, transform_id_base=w["transform_id_base"], threads=threads)
        census = dict(cen, objects=int(ns), eps_plane_ulp=4.0, eps_texel=1.0 / 256.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_frame": e2e_frame, "gpu_launches": int(launches), "gather_ok": gather_ok,
                "clocks": clocks,
                "detail": {"visible_prev_frame": vis_prev, "early_draws": early_total, "late_draws": late_total, "early_list_checksum": early_checksum,
                           "kernel_ms": {"early": t_early, "pyramid": t_pyr, "late": t_late}, "late_kernel_regimes": regimes, "boundary_census": census, "gathered_lists": gsum or None, "cuda_graph": graph_info}}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
