#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Runs the REFERENCE'S OWN cull shaders (SPIR-V built by oracle/Makefile `spv` from
/root/reference/src/Renderer/VulkanShaders/*.comp.glsl with the reference's bundled glslang) in oracle/spirv_interp on a
small scene and writes inputs + outputs to tests/golden/spirv_golden.npz.  tests/test_oracle_golden.py then checks
oracle/cull_oracle.cpp against these vectors (CPU, no reference tree needed) and tests/test_parity_gpu.py checks the CUDA
path against the same vectors on the GPU box.

The host side here plays the role of BlitzenVulkan/vulkanDraw.cpp: it binds the buffers the dispatch functions bind
(DrawCullFirstPass :107-158, DrawCullOcclusionPass :162-226, PreClusterDrawCull :318-370, ClusterCull :372-423,
GenerateDepthPyramid :554-622), fills the push constants (vulkanData.h:449-467) and dispatches N/64+1 groups.
The Hi-Z sampler is modelled after the reference's sampler object (BlitzenVulkan/vulkanResources.cpp:51-55, :394-429):
LINEAR filter, REDUCTION_MODE_MIN, mipmap NEAREST, CLAMP_TO_EDGE, lod clamp [0, 16], following the Vulkan texel-filtering
rules (unnormalised u = s*W, i0 = floor(u - 0.5), weights (1-a, a), MIN over the texels with non-zero weight).

Run (build container only; /root/reference must exist):   python oracle/spirv_interp/make_spirv_golden.py
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
from blitzen_b200 import scene, types as T  # noqa: E402

SPV = os.path.join(ROOT, "oracle", "_ref", "spv")
f32 = np.float32


def u8(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()


class MinSampler:
    """VkSampler of the reference's depth pyramid: LINEAR + MIN reduction + NEAREST mip + clamp-to-edge, lod in [0, 16]."""

    def __init__(self, mips):
        self.mips = mips            # list of float32 [h, w]

    def __call__(self, s, t, lod):
        lod = float(lod)
        if lod != lod or lod < 0.0:
            lod = 0.0                # NaN / -inf: clamped to minLod
        lod = min(lod, 16.0)
        level = int(min(np.floor(lod + 0.5), len(self.mips) - 1))
        img = self.mips[level]
        h, w = img.shape
        u = f32(f32(s) * f32(w)) - f32(0.5)
        v = f32(f32(t) * f32(h)) - f32(0.5)
        fu, fv = np.floor(u), np.floor(v)
        au, av = f32(u - fu), f32(v - fv)

        def clamp(i, n):
            if not (i >= 0):
                return 0
            return int(min(i, n - 1))

        i0, i1 = clamp(fu, w), clamp(fu + 1, w)
        j0, j1 = clamp(fv, h), clamp(fv + 1, h)
        cand = [img[j0, i0]]
        if au != 0:
            cand.append(img[j0, i1])
        if av != 0:
            cand.append(img[j1, i0])
            if au != 0:
                cand.append(img[j1, i1])
        d = cand[0]
        for c in cand[1:]:
            if c < d:
                d = c
        return d


def run_pyramid(depth):
    """GenerateDepthPyramid: one dispatch per mip, mip 0 samples the depth attachment, mip i samples mip i-1."""
    mod = S.Module(os.path.join(SPV, "DepthPyramidGeneration.spv"))
    h, w = depth.shape
    prev = lambda v: (lambda r: r)(max(1, 1 << (int(v - 1).bit_length() - 1)) if v > 1 else 1)   # largest power of two strictly below v (v > 1)
    pw, ph = prev(w), prev(h)
    mips = 0
    a, b = pw, ph
    while a > 1 or b > 1:
        mips += 1; a //= 2; b //= 2
    out = []
    src = depth
    for i in range(mips):
        lw, lh = max(1, pw >> i), max(1, ph >> i)
        dst = np.zeros((lh, lw), dtype=np.float32)
        mc = S.Machine(mod)
        mc.bind("inImage", S.Sampler2D(MinSampler([src])))
        mc.bind("outImage", S.StorageImage2D(dst))
        mc.bind("constants", u8(np.array([lw, lh], dtype=np.float32)))
        mc.dispatch((lw // 32 + 1, lh // 32 + 1, 1))
        out.append(dst)
        src = dst
    return out, (pw, ph, mips)


def bind_common(mc, sc, view, draws, count, vis):
    mc.bind("viewData", u8(view))
    mc.bind("transformBuffer", sc["xf_u8"])
    mc.bind("surfaceBuffer", sc["surf_u8"])
    mc.bind("lodBuffer", sc["lod_u8"])
    if mc.has("indirectDrawBuffer"):
        mc.bind("indirectDrawBuffer", draws)
    mc.bind("indirectDrawCountBuffer", count)
    mc.bind("visibilityBuffer", vis)


OBJ_ADDR, DISPATCH_ADDR, CCOUNT_ADDR = 0x10000000, 0x20000000, 0x30000000


def run_draw_cull(shader, sc, objs, view, vis=None, pyramid=None, onpc=False):
    mod = S.Module(os.path.join(SPV, shader + ".spv"))
    n = len(objs)
    draws = np.zeros(max(n, 1) * 24, dtype=np.uint8)
    count = np.zeros(4, dtype=np.uint8)            # reset before every dispatch (vulkanDraw.cpp:119, :180)
    visb = u8(vis) if vis is not None else np.zeros(max(n, 1) * 4, dtype=np.uint8)
    mc = S.Machine(mod)
    bind_common(mc, sc, view, draws, count, visb)
    obj_u8 = u8(objs)
    if onpc:
        mc.bind("onpcReflectiveObjectBuffer", obj_u8)
    mc.register_address(OBJ_ADDR, obj_u8)
    pc = np.zeros(16, dtype=np.uint8)
    pc[:8] = np.array([OBJ_ADDR], dtype="<u8").view(np.uint8)
    pc[8:12] = np.array([n], dtype="<u4").view(np.uint8)
    mc.bind("pushConstant", pc)
    if pyramid is not None:
        mc.bind("depthPyramid", S.Sampler2D(MinSampler(pyramid)))
    t0 = time.time()
    mc.dispatch((n // 64 + 1, 1, 1))
    cnt = int(count.view("<u4")[0])
    print(f"  {shader}: {n} invocations, {mc.instr_count} SPIR-V instructions, {cnt} draws, {time.time() - t0:.1f} s", flush=True)
    return draws.view("<u4").reshape(-1, 6)[:cnt].copy(), cnt, visb.view("<u4")[:n].copy()


def run_cluster_path(sc, objs, view, capacity):
    n = len(objs)
    mod = S.Module(os.path.join(SPV, "PreClusterDrawCull.spv"))
    disp = np.zeros(capacity * 12, dtype=np.uint8)
    ccount = np.zeros(4, dtype=np.uint8)
    count = np.zeros(4, dtype=np.uint8)
    mc = S.Machine(mod)
    bind_common(mc, sc, view, np.zeros(24, dtype=np.uint8), count, np.zeros(max(n, 1) * 4, dtype=np.uint8))
    obj_u8 = u8(objs)
    for base, arr in ((OBJ_ADDR, obj_u8), (DISPATCH_ADDR, disp), (CCOUNT_ADDR, ccount)):
        mc.register_address(base, arr)
    pc = np.zeros(32, dtype=np.uint8)
    pc[0:8] = np.array([OBJ_ADDR], dtype="<u8").view(np.uint8)
    pc[8:16] = np.array([DISPATCH_ADDR], dtype="<u8").view(np.uint8)
    pc[16:24] = np.array([CCOUNT_ADDR], dtype="<u8").view(np.uint8)
    pc[24:28] = np.array([n], dtype="<u4").view(np.uint8)
    mc.bind("pushConstant", pc)
    t0 = time.time()
    mc.dispatch((n // 64 + 1, 1, 1))
    m = int(ccount.view("<u4")[0])
    print(f"  PreClusterDrawCull: {n} invocations, {m} dispatch records, {time.time() - t0:.1f} s", flush=True)
    records = disp.view("<u4").reshape(-1, 3)[:m].copy()
    # ClusterCull(dispatchCount) -- vulkanDraw.cpp:372-423
    mod2 = S.Module(os.path.join(SPV, "InitialClusterCull.spv"))
    draws = np.zeros(max(m, 1) * 24, dtype=np.uint8)
    count2 = np.zeros(4, dtype=np.uint8)
    mc2 = S.Machine(mod2)
    bind_common(mc2, sc, view, draws, count2, np.zeros(max(n, 1) * 4, dtype=np.uint8))
    mc2.bind("clusterBuffer", sc["cluster_u8"])
    for base, arr in ((OBJ_ADDR, obj_u8), (DISPATCH_ADDR, disp), (CCOUNT_ADDR, ccount)):
        mc2.register_address(base, arr)
    pc2 = pc.copy()
    pc2[24:28] = np.array([m], dtype="<u4").view(np.uint8)
    mc2.bind("pushConstant", pc2)
    t0 = time.time()
    mc2.dispatch((m // 64 + 1, 1, 1))
    c2 = int(count2.view("<u4")[0])
    print(f"  InitialClusterCull: {m} invocations, {c2} draws, {time.time() - t0:.1f} s", flush=True)
    return records, draws.view("<u4").reshape(-1, 6)[:c2].copy()


def main():
    tables = scene.mesh_tables()
    # a small scene that exercises all four bundled meshes (every LOD range, cluster counts from 1 to 1046)
    groups = ((0, 5.0, 700), (2, 1.0, 700), (1, 0.5, 300), (3, 0.2, 300))
    objs, xf = scene.generate(groups, 260.0, True, "counter", seed=5, n_dynamic=40)
    transforms, _ = scene.assemble_transforms(objs, xf, transform_id_base=0)
    sc = dict(xf_u8=u8(transforms), surf_u8=u8(tables["surfaces"]), lod_u8=u8(tables["lods"]), cluster_u8=u8(tables["clusters"]))
    n = len(objs)
    W, H = 320, 180
    views = {
        "inside": scene.make_view((130, 130, 130), 0.0, 0.0, 70.0, W, H, 0.1, 400.0),
        "tilted": scene.make_view((40, 200, 20), 0.7, -0.3, 70.0, W, H, 0.1, 300.0),
        "all": scene.make_view((130, 130, -900), 0.0, 0.0, 70.0, W, H, 0.1, 1e9),
        "ref_default": scene.reference_views()["default"],
    }
    out = dict(objs=objs, transforms=transforms, view_names=np.array(list(views)), views=np.concatenate([views[k] for k in views]))
    rng = np.random.default_rng(3)

    print("depth pyramid (DepthPyramidGeneration.comp.glsl)", flush=True)
    depth = scene.synthetic_depth(W, H, n_rects=18, z_min=5.0, z_max=250.0, seed=77)
    depth += rng.random((H, W), dtype=np.float32) * np.float32(1e-4)
    t0 = time.time()
    mips, (pw, ph, nm) = run_pyramid(depth)
    print(f"  {W}x{H} -> {pw}x{ph} x {nm} mips, {time.time() - t0:.1f} s", flush=True)
    out["depth"] = depth
    out["pyramid"] = np.concatenate([m.reshape(-1) for m in mips])
    out["pyramid_whm"] = np.array([pw, ph, nm], dtype=np.uint32)
    depth2 = scene.synthetic_depth(97, 61, n_rects=9, z_min=5.0, z_max=250.0, seed=78)
    mips2, whm2 = run_pyramid(depth2)
    out["depth_odd"] = depth2
    out["pyramid_odd"] = np.concatenate([m.reshape(-1) for m in mips2])
    out["pyramid_odd_whm"] = np.array(whm2, dtype=np.uint32)

    for vn, view in views.items():
        print(f"view {vn}", flush=True)
        d, c, _ = run_draw_cull("TransparentDrawCull", sc, objs, view)
        out[f"transparent_{vn}"] = d
    vis0 = (rng.random(n) < 0.5).astype(np.uint32)
    out["vis0"] = vis0
    for vn in ("inside", "tilted"):
        view = np.array(views[vn], copy=True)
        d, c, _ = run_draw_cull("InitialDrawCull", sc, objs, view, vis=vis0)
        out[f"initial_{vn}"] = d
        view["pyramidWidth"] = f32(pw); view["pyramidHeight"] = f32(ph)     # vulkanRendererSetup.cpp:903-904
        d, c, v = run_draw_cull("LateDrawCull", sc, objs, view, vis=vis0, pyramid=mips)
        out[f"late_{vn}"] = d
        out[f"late_vis_{vn}"] = v
    onpc = objs[::29][:100].copy()
    out["onpc_objs"] = onpc
    d, c, _ = run_draw_cull("OnpcDrawCull", sc, onpc, views["inside"], onpc=True)
    out["onpc_inside"] = d
    sub = objs[:600]
    out["cluster_objs_count"] = np.array([len(sub)], dtype=np.uint32)
    rec, cd = run_cluster_path(sc, sub, views["inside"], capacity=600 * 1100)
    out["cluster_dispatch_inside"] = rec
    out["cluster_draws_inside"] = cd
    # the reference's own RenderingStressTest scene (first 4096 objects, tests/golden/stress_head_4k.blob) under the reference's own cameras
    from blitzen_b200 import sceneio
    head = sceneio.read_blob(os.path.join(ROOT, "tests", "golden", "stress_head_4k.blob"))
    sch = dict(xf_u8=u8(head["transforms"]), surf_u8=u8(head["surfaces"]), lod_u8=u8(head["lods"]), cluster_u8=u8(head["clusters"]))
    rviews = scene.reference_views()
    vis_h = (rng.random(len(head["objs"])) < 0.3).astype(np.uint32)
    out["head_vis0"] = vis_h
    for vn in ("default", "cfg1_centre", "cfg1_tilted", "cfg1_all"):
        print(f"reference head, view {vn}", flush=True)
        d, c, _ = run_draw_cull("TransparentDrawCull", sch, head["objs"], rviews[vn])
        out[f"head_transparent_{vn}"] = d
        view = np.array(rviews[vn], copy=True)
        view["pyramidWidth"] = f32(pw); view["pyramidHeight"] = f32(ph)
        d, c, v = run_draw_cull("LateDrawCull", sch, head["objs"], view, vis=vis_h, pyramid=mips)
        out[f"head_late_{vn}"] = d
        out[f"head_late_vis_{vn}"] = v
    path = os.path.join(ROOT, "tests", "golden", "spirv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
