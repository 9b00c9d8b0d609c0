#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Runs the REFERENCE'S OWN D3D12 cull shaders (SPIR-V built by oracle/Makefile `hlsl_spv` from
/root/reference/src/Renderer/HlslShaders/CS/*.hlsl with the reference's bundled glslang, HLSL front end, entry csMain) in
oracle/spirv_interp and writes inputs + outputs to tests/golden/hlsl_golden.npz.  tests/test_oracle_golden_hlsl.py then checks
oracle/cull_oracle.cpp against these vectors (CPU, no reference tree needed) and tests/test_parity_gpu.py checks the CUDA
path against the same vectors on the GPU box.  This pins the D3D12-only rows of SURVEY 8(a): OcclusionCheck (point-texel
Hi-Z, hlslMath.hlsl:58-78), the 2x2-min pyramid (depthPyramid.cs.hlsl:13-27), drawOccFirst / drawOccLate / drawOccTemporal /
drawCull, and indirect instancing (drawInstCountReset + drawInstCull + drawInstCmd).

The host side here plays the role of BlitzenDX12/dx12Draw.cpp: DrawCountReset :114-132, CullObjects :134-148, DrawCullPass
:150-184, DrawOccFirstPass :186-222, GenerateDepthPyramid :224-288 (mip i of the pyramid written from the depth target for
i == 0 and from pyramid mip i-1 otherwise -- note the root constant is set BEFORE mipLevel is incremented, :256-270),
DrawOccLatePass :290-338, DrawInstanceCullPass :340-413; group counts are GetComputeShaderGroupSize = n / 64 + 1
(BlitzenMathLibrary/blitML.h:64-67).  The pyramid extent is max(1, W >> 1) x max(1, H >> 1) and ViewData.pyramidWidth/Height hold
that extent as floats (dx12RNDResources.cpp:103-106, dx12Draw.cpp:563-564).

Run (build container only; /root/reference must exist):   make -C oracle hlsl_spv && python oracle/spirv_interp/make_hlsl_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import spirv_interp as S  # noqa: E402
from blitzen_b200 import scene  # noqa: E402

SPV = os.path.join(ROOT, "oracle", "_ref", "spv_hlsl")
f32 = np.float32


def u8(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()


def u32c(v):
    return np.array([v], dtype="<u4").view(np.uint8).copy()


def groups(n, t=64):
    return n // t + 1


def run_count_reset(count):
    """drawCountReset.cs.hlsl: one thread, rwb_drawCmdCounter[0] = 0 (dx12Draw.cpp:114-132)."""
    mc = S.Machine(S.Module(os.path.join(SPV, "drawCountReset.cs.spv")))
    mc.bind("rwb_drawCmdCounter", S.TexelBufferU32(count))
    mc.dispatch((1, 1, 1))


def run_pyramid(depth):
    """GenerateDepthPyramid, dx12Draw.cpp:224-288."""
    mod = S.Module(os.path.join(SPV, "depthPyramid.cs.spv"))
    h, w = depth.shape
    pw, ph = max(1, w >> 1), max(1, h >> 1)
    mips, a, b = 0, pw, ph
    while a > 1 or b > 1:
        mips += 1; a //= 2; b //= 2
    out = [np.zeros((max(1, ph >> i), max(1, pw >> i)), dtype=np.float32) for i in range(mips)]
    mip_level = 0
    for i in range(mips):
        lw, lh = max(1, pw >> i), max(1, ph >> i)
        mc = S.Machine(mod)
        mc.bind("PyramidMip", u32c(mip_level))
        mc.bind("rwtex_depthOut", S.StorageImage2D(out[i]))
        if i == 0:
            mc.bind("tex_depthIn", S.Texture2D([depth]))
        else:
            mc.bind("tex_depthIn", S.Texture2D(out))
            mip_level += 1
        mc.dispatch((groups(lw, 32), groups(lh, 32), 1))
    return out, (pw, ph, mips)


def bind_shared(mc, sc, objs_u8, view):
    mc.bind("ViewData", u8(view))
    mc.bind("ssbo_Renders", objs_u8)
    mc.bind("ssbo_Surfaces", sc["surf_u8"])
    mc.bind("ssbo_Transforms", sc["xf_u8"])
    mc.bind("ssbo_LODs", sc["lod_u8"])


def run_draw_cull(shader, sc, objs, view, vis=None, pyramid=None):
    """DrawCullPass / DrawOccFirstPass / DrawOccLatePass: count reset, then CullObjects(objCount)."""
    n = len(objs)
    draws = np.zeros(max(n, 1) * 32, dtype=np.uint8)
    count = u32c(0xDEAD)
    run_count_reset(count)
    visb = u8(vis) if vis is not None else None
    mc = S.Machine(S.Module(os.path.join(SPV, shader + ".spv")))
    bind_shared(mc, sc, u8(objs), view)
    mc.bind("ObjCountConstant", u32c(n))
    mc.bind("ssbo_DrawCmd", draws)
    mc.bind("rwb_DrawCmdCounter", S.TexelBufferU32(count))
    if mc.has("rwssbo_DrawVisibilityBuffer"):
        mc.bind("rwssbo_DrawVisibilityBuffer", visb)
    if mc.has("tex_HiZMap"):
        mc.bind("tex_HiZMap", S.Texture2D(pyramid))
    t0 = time.time()
    mc.dispatch((groups(n), 1, 1))
    cnt = int(count.view("<u4")[0])
    print(f"  {shader}: {n} invocations, {mc.instr_count} SPIR-V instructions, {cnt} draws, {time.time() - t0:.1f} s", flush=True)
    return draws.view("<u4").reshape(-1, 8)[:cnt].copy(), (visb.view("<u4")[:n].copy() if visb is not None else None)


def run_instanced(sc, objs, view, lod_instances, n_inst_slots):
    """DrawInstanceCullPass, dx12Draw.cpp:340-413: count reset, instance-counter reset, drawInstCull, drawInstCmd."""
    n, nl = len(objs), len(lod_instances)
    count = u32c(0xDEAD)
    run_count_reset(count)
    counters = u8(lod_instances)                      # {instanceOffset, instanceCount}; counts start dirty on purpose
    counters.view("<u4")[1::2] = 0xBEEF
    mc = S.Machine(S.Module(os.path.join(SPV, "drawInstCountReset.cs.spv")))
    mc.bind("rwssbo_InstCounter", counters)
    mc.bind("LodCount", u32c(nl))
    mc.dispatch((groups(nl), 1, 1))
    inst = np.full(n_inst_slots * 4, 0xFF, dtype=np.uint8)
    mc = S.Machine(S.Module(os.path.join(SPV, "drawInstCull.cs.spv")))
    bind_shared(mc, sc, u8(objs), view)
    mc.bind("ObjCountConstant", u32c(n))
    mc.bind("rwssbo_InstCounter", counters)
    mc.bind("rwssbo_instIndices", inst)
    t0 = time.time()
    mc.dispatch((groups(n), 1, 1))
    draws = np.zeros(max(nl, 1) * 32, dtype=np.uint8)
    mc2 = S.Machine(S.Module(os.path.join(SPV, "drawInstCmd.cs.spv")))
    mc2.bind("rwssbo_InstCounter", counters)
    mc2.bind("LodCount", u32c(nl))
    mc2.bind("ssbo_LODs", sc["lod_u8"])
    mc2.bind("ssbo_DrawCmd", draws)
    mc2.bind("rwb_DrawCmdCounter", S.TexelBufferU32(count))
    mc2.dispatch((groups(nl), 1, 1))
    cnt = int(count.view("<u4")[0])
    print(f"  drawInst*: {n} invocations, {cnt} commands, {time.time() - t0:.1f} s", flush=True)
    return inst.view("<u4").copy(), counters.view("<u4").reshape(-1, 2)[:, 1].copy(), draws.view("<u4").reshape(-1, 8)[:cnt].copy()


def main():
    tables = scene.mesh_tables()
    groups_ = ((0, 5.0, 700), (2, 1.0, 700), (1, 0.5, 300), (3, 0.2, 300))   # same scene as make_spirv_golden.py
    objs, xf = scene.generate(groups_, 260.0, True, "counter", seed=5, n_dynamic=40)
    transforms, _ = scene.assemble_transforms(objs, xf, transform_id_base=0)
    sc = dict(xf_u8=u8(transforms), surf_u8=u8(tables["surfaces"]), lod_u8=u8(tables["lods"]))
    n = len(objs)
    W, H = 320, 180
    views = {
        "inside": scene.make_view((130, 130, 130), 0.0, 0.0, 70.0, W, H, 0.1, 400.0),
        "tilted": scene.make_view((40, 200, 20), 0.7, -0.3, 70.0, W, H, 0.1, 300.0),
        "all": scene.make_view((130, 130, -900), 0.0, 0.0, 70.0, W, H, 0.1, 1e9),
        "ref_default": scene.reference_views()["default"],
    }
    out = dict(objs=objs, transforms=transforms, view_names=np.array(list(views)), views=np.concatenate([views[k] for k in views]))
    rng = np.random.default_rng(11)

    print("depth pyramid (depthPyramid.cs.hlsl)", flush=True)
    depth = scene.synthetic_depth(W, H, n_rects=18, z_min=5.0, z_max=250.0, seed=77)
    depth += rng.random((H, W), dtype=np.float32) * np.float32(1e-4)
    t0 = time.time()
    mips, (pw, ph, nm) = run_pyramid(depth)
    print(f"  {W}x{H} -> {pw}x{ph} x {nm} mips, {time.time() - t0:.1f} s", flush=True)
    out["depth"] = depth
    out["pyramid"] = np.concatenate([m.reshape(-1) for m in mips])
    out["pyramid_whm"] = np.array([pw, ph, nm], dtype=np.uint32)
    depth2 = scene.synthetic_depth(97, 61, n_rects=9, z_min=5.0, z_max=250.0, seed=78)   # odd sizes down the chain: out-of-range loads = 0
    mips2, whm2 = run_pyramid(depth2)
    out["depth_odd"] = depth2
    out["pyramid_odd"] = np.concatenate([m.reshape(-1) for m in mips2])
    out["pyramid_odd_whm"] = np.array(whm2, dtype=np.uint32)

    for vn, view in views.items():
        print(f"view {vn}", flush=True)
        d, _ = run_draw_cull("drawCull.cs", sc, objs, view)
        out[f"drawcull_{vn}"] = d
    vis0 = (rng.random(n) < 0.5).astype(np.uint32)
    out["vis0"] = vis0
    for vn in ("inside", "tilted"):
        view = np.array(views[vn], copy=True)
        d, v = run_draw_cull("drawOccFirst.cs", sc, objs, view, vis=vis0)
        out[f"occfirst_{vn}"] = d
        assert np.array_equal(v, vis0)
        view["pyramidWidth"] = f32(pw); view["pyramidHeight"] = f32(ph)       # dx12Draw.cpp:563-564
        d, v = run_draw_cull("drawOccLate.cs", sc, objs, view, vis=vis0, pyramid=mips)
        out[f"occlate_{vn}"] = d
        out[f"occlate_vis_{vn}"] = v
        d, _ = run_draw_cull("drawOccTemporal", sc, objs, view, pyramid=mips)
        out[f"occtemporal_{vn}"] = d
    # indirect instancing: LodInstanceCounter table as blitzenMeshes.cpp:159-161 builds it (instanceOffset = lodId * bucket)
    nl = len(tables["lods"])
    bucket = 1024
    li = np.zeros(nl, dtype=tables["lodInstances"].dtype)
    li["instanceOffset"] = np.arange(nl, dtype=np.uint32) * bucket
    out["inst_lod_instances"] = li
    out["inst_bucket"] = np.array([bucket], dtype=np.uint32)
    for vn in ("inside", "all"):
        idx, counts, cmds = run_instanced(sc, objs, views[vn], li, nl * bucket)
        assert counts.max() <= bucket
        out[f"inst_indices_{vn}"] = idx
        out[f"inst_counts_{vn}"] = counts
        out[f"inst_cmds_{vn}"] = cmds
    # the reference's own RenderingStressTest scene (first 4096 objects) under the reference's own cameras
    from blitzen_b200 import sceneio
    head = sceneio.read_blob(os.path.join(ROOT, "tests", "golden", "stress_head_4k.blob"))
    sch = dict(xf_u8=u8(head["transforms"]), surf_u8=u8(head["surfaces"]), lod_u8=u8(head["lods"]))
    rviews = scene.reference_views()
    vis_h = (rng.random(len(head["objs"])) < 0.3).astype(np.uint32)
    out["head_vis0"] = vis_h
    for vn in ("default", "cfg1_centre", "cfg1_tilted", "cfg1_all"):
        print(f"reference head, view {vn}", flush=True)
        view = np.array(rviews[vn], copy=True)
        view["pyramidWidth"] = f32(pw); view["pyramidHeight"] = f32(ph)
        d, v = run_draw_cull("drawOccLate.cs", sch, head["objs"], view, vis=vis_h, pyramid=mips)
        out[f"head_occlate_{vn}"] = d
        out[f"head_occlate_vis_{vn}"] = v
    path = os.path.join(ROOT, "tests", "golden", "hlsl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
