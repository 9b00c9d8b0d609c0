"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact (draw records, counts, visibility, pyramid texels): integer / byte equality, no tolerance."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import view_at

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(built):
    from blitzen_b200 import capi
    return capi


def make_ctx(capi, sc, **kw):
    ctx = capi.CullContext(0)
    ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], clusters=sc["clusters"], lod_instances=kw.pop("lod_instances", None), **kw)
    return ctx


def recs_u32(rec):
    return rec.view(np.uint32).reshape(len(rec), rec.dtype.itemsize // 4)


SMALL_VIEWS = {
    "corner": dict(position=(20, 70, 0), z_far=650.0),
    "centre": dict(position=(380, 380, 380), z_far=2000.0),
    "outside_all": dict(position=(380, 380, -2500), z_far=1e9),
    "tilted": dict(position=(200, 500, 100), yaw=0.7, pitch=-0.3, z_far=900.0),
    "nothing": dict(position=(380, 380, 5000), z_far=100.0),
}


@pytest.mark.parametrize("vname", list(SMALL_VIEWS))
@pytest.mark.parametrize("fmt", [0, 1])
def test_frustum_lod_small(capi, small_scene, vname, fmt):
    sc = small_scene
    view = view_at(**SMALL_VIEWS[vname])
    exp, total, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, rec_words=6 if fmt == 0 else 8)
    with make_ctx(capi, sc) as ctx:
        ctx.set_view(view)
        ctx.frustum_lod(capi.LIST_OPAQUE, fmt)
        got, gtotal = ctx.read_draws(fmt)
    assert gtotal == total
    assert np.array_equal(recs_u32(got), exp)
    if vname == "outside_all":
        assert total == len(sc["objs"])
    if vname == "nothing":
        assert total == 0


def test_frustum_lod_medium_lookback(capi, medium_scene):
    sc = medium_scene
    view = view_at(position=(950, 950, 950), z_far=5000.0)
    exp, total, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, threads=8)
    with make_ctx(capi, sc) as ctx:
        ctx.set_view(view)
        for _ in range(3):   # repeated launches reuse the self-resetting control block / epoch-tagged status words
            ctx.frustum_lod()
            got, gtotal = ctx.read_draws()
            assert gtotal == total and total > 100_000
            assert np.array_equal(recs_u32(got), exp)


def test_reference_views_on_reference_head(capi, built, views):
    """Inputs produced by the reference's own frontend: first 4096 objects of its RenderingStressTest scene, reference camera."""
    import os
    from blitzen_b200 import sceneio
    sc = sceneio.read_blob(os.path.join(os.path.dirname(__file__), "golden", "stress_head_4k.blob"))
    for name in ("default", "cfg1_centre", "cfg1_all", "cfg1_tilted"):
        exp, total, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], views[name], O.PASS_FRUSTUM)
        with capi.CullContext(0) as ctx:
            ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
            ctx.set_view(views[name])
            ctx.frustum_lod()
            got, gtotal = ctx.read_draws()
        assert gtotal == total
        assert np.array_equal(recs_u32(got), exp)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("size", [(1920, 1080), (1280, 720), (640, 360), (257, 131), (64, 64), (37, 5), (4, 2)])
def test_pyramid(capi, built, variant, size):
    from blitzen_b200 import scene
    w, h = size
    depth = scene.synthetic_depth(w, h, n_rects=24, seed=w * 31 + h)
    rng = np.random.default_rng(w + h)
    depth += (rng.random((h, w), dtype=np.float32) * np.float32(1e-4))   # break ties so every min has a unique argmin
    exp = O.build_pyramid(depth, variant)
    for tma in (1, 0):
        with capi.CullContext(0) as ctx:
            ctx.set_option("pyramid_tma", tma)
            ctx.set_depth(depth)
            ctx.build_pyramid(variant)
            ctx.build_pyramid(variant)     # second build re-arms the ticket
            data, (pw, ph, mips), offs = ctx.read_pyramid()
        assert (pw, ph, mips) == (exp.width, exp.height, exp.mips)
        assert list(offs[:mips]) == exp.offsets[:mips]
        assert np.array_equal(data.view(np.uint32), exp.data[:len(data)].view(np.uint32)), f"pyramid mismatch tma={tma}"


@pytest.mark.parametrize("hiz", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
def test_two_phase_frames(capi, small_scene, hiz, fmt):
    """Frame 0 (visibility all 0, cleared pyramid) then two more frames against a synthetic depth: early + late lists,
    visibility buffer, all bit-exact (SURVEY.md 8a semantics 3)."""
    from blitzen_b200 import scene
    sc = small_scene
    W, H = 640, 360
    frame_views = [view_at(position=(380, 380, 380), z_far=2000.0, width=W, height=H)] * 2 + \
                  [view_at(position=(380, 380, 380), yaw=0.5, z_far=2000.0, width=W, height=H)]   # the camera turns in frame 2
    depth = scene.synthetic_depth(W, H, n_rects=40, z_min=20.0, z_max=400.0, seed=99)
    rw = 6 if fmt == 0 else 8
    n = len(sc["objs"])
    vis = np.zeros(n, dtype=np.uint32)
    with make_ctx(capi, sc) as ctx:
        ctx.set_view(frame_views[0])
        ctx.clear_pyramid(hiz, W, H)
        pyr = O.cleared_pyramid(W, H, hiz)
        for frame in range(3):
            view = frame_views[frame]
            ctx.set_view(view)
            e_exp, e_tot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_EARLY, rec_words=rw, vis=vis)
            ctx.early(fmt)
            e_got, e_gtot = ctx.read_draws(fmt)
            assert e_gtot == e_tot and np.array_equal(recs_u32(e_got), e_exp), f"early frame {frame}"
            if frame > 0:
                ctx.set_depth(depth)
                ctx.build_pyramid(hiz)
                pyr = O.build_pyramid(depth, hiz)
            l_exp, l_tot, vis = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_LATE, rec_words=rw, hiz=hiz, pyramid=pyr, vis=vis)
            ctx.late(fmt, hiz)
            l_got, l_gtot = ctx.read_draws(fmt)
            assert l_gtot == l_tot and np.array_equal(recs_u32(l_got), l_exp), f"late frame {frame}"
            assert np.array_equal(ctx.read_visibility(), vis), f"visibility frame {frame}"
            if frame == 0:
                assert e_tot == 0 and l_tot > 0
            if frame == 2:
                assert e_tot > 0 and l_tot > 0      # newly visible objects are emitted by the late pass
        assert 0 < int(vis.sum()) < n


def test_capacity_clamp(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, -2500), z_far=1e9)
    cap = 12_345
    exp, total, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, capacity=cap)
    with make_ctx(capi, sc, draw_capacity=cap) as ctx:
        ctx.set_view(view)
        ctx.frustum_lod()
        written, gtotal = ctx.read_count()
        got, _ = ctx.read_draws()
    assert written == cap and gtotal == total == len(sc["objs"])
    assert np.array_equal(recs_u32(got), exp)


def test_instanced(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    li = sc["lodInstances"].copy()
    nl = len(sc["lods"])
    cap = np.full(nl, 4000, dtype=np.uint32)
    cap[3] = 7
    li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
    idx_e, cnt_e, cmds_e = O.cull_instanced(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], li, cap, view)
    with make_ctx(capi, sc, lod_instances=li, bucket_capacity=cap) as ctx:
        ctx.set_view(view)
        for _ in range(2):
            ctx.instanced()
            cmds, total = ctx.read_draws(capi.REC_DX32)
            idx, counters = ctx.read_instances(int(cap.sum()))
            assert np.array_equal(counters["instanceCount"], cnt_e)
            assert np.array_equal(recs_u32(cmds), cmds_e) and total == len(cmds_e)
            for l in range(nl):
                o, c = int(li["instanceOffset"][l]), int(min(cnt_e[l], cap[l]))
                assert np.array_equal(idx[o:o + c], idx_e[o:o + c]), f"bucket {l}"
    assert cnt_e[7] > cap[7]      # the overflowing bucket really overflowed


def test_cluster_path(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=400.0)
    capacity = 3_000_000
    d_exp, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, capacity)
    assert 0 < d_tot <= capacity
    W, H = 640, 360
    from blitzen_b200 import scene
    depth = scene.synthetic_depth(W, H, n_rects=40, z_min=20.0, z_max=300.0, seed=5)
    with make_ctx(capi, sc, cluster_dispatch_capacity=capacity, draw_capacity=capacity) as ctx:
        ctx.set_view(view)
        ctx.cluster_expand()
        d_got, d_gtot = ctx.read_cluster_dispatch()
        assert d_gtot == d_tot
        assert np.array_equal(d_got.view(np.uint32).reshape(-1, 3), d_exp)
        for fmt in (0, 1):
            rw = 6 if fmt == 0 else 8
            exp, tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_exp, 0, rec_words=rw)
            ctx.cluster_cull(capi.CLUSTER_PASSTHROUGH, fmt)
            got, gtot = ctx.read_draws(fmt)
            assert gtot == tot == d_tot and np.array_equal(recs_u32(got), exp)
        exp, tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_exp, 1)
        ctx.cluster_cull(capi.CLUSTER_SPHERE, 0)
        got, gtot = ctx.read_draws(0)
        assert gtot == tot and 0 < tot < d_tot and np.array_equal(recs_u32(got), exp)
        for hiz in (0, 1):
            ctx.set_depth(depth); ctx.build_pyramid(hiz)
            pyr = O.build_pyramid(depth, hiz)
            exp, tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_exp, 1, hiz=hiz, pyramid=pyr)
            ctx.cluster_cull(capi.CLUSTER_SPHERE_HIZ, 0, hiz)
            got, gtot = ctx.read_draws(0)
            assert gtot == tot and np.array_equal(recs_u32(got), exp)


def test_onpc_quirk_and_lists(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    onpc = sc["objs"][5000:5100].copy()
    transp = sc["objs"][20000:29000].copy()
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], transparent=transp, onpc=onpc)
        ctx.set_view(view)
        exp, tot, _ = O.cull(onpc, sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, flags=O.FLAG_ONPC_LOD_QUIRK)
        ctx.frustum_lod(capi.LIST_ONPC, 0, capi.FLAG_ONPC_LOD_QUIRK)
        got, gtot = ctx.read_draws()
        assert gtot == tot and np.array_equal(recs_u32(got), exp)
        exp, tot, _ = O.cull(transp, sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM)
        ctx.frustum_lod(capi.LIST_TRANSPARENT, 0)
        got, gtot = ctx.read_draws()
        assert gtot == tot and np.array_equal(recs_u32(got), exp)


def test_update_transforms(capi, small_scene):
    sc = small_scene
    view = view_at(position=(50, 50, -200), z_far=2000.0)
    xf = sc["transforms"].copy()
    rng = np.random.default_rng(3)
    xf["pos"][:1000] += rng.standard_normal((1000, 3)).astype(np.float32) * 20
    xf["orientation"][:1000] = rng.standard_normal((1000, 4)).astype(np.float32)   # non-unit quaternions on purpose
    with make_ctx(capi, sc) as ctx:
        ctx.set_view(view)
        ctx.update_transforms(0, xf[:1000])
        ctx.frustum_lod()
        got, gtot = ctx.read_draws()
    exp, tot, _ = O.cull(sc["objs"], xf, sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM)
    assert gtot == tot and np.array_equal(recs_u32(got), exp)


def test_empty_and_tiny(capi, tables):
    from blitzen_b200 import scene
    view = view_at(position=(0, 0, -50), z_far=1e6)
    for n in (1, 31, 32, 33, 1023, 1024, 1025):
        objs, xf = scene.generate(groups=((0, 5.0, n),), multiplier=30.0, prologue=False, prng="counter", seed=n)
        transforms, base = scene.assemble_transforms(objs, xf, 0)
        exp, tot, _ = O.cull(objs, transforms, tables["surfaces"], tables["lods"], view, O.PASS_FRUSTUM)
        with capi.CullContext(0) as ctx:
            ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"])
            ctx.set_view(view)
            ctx.frustum_lod()
            got, gtot = ctx.read_draws()
        assert gtot == tot == n and np.array_equal(recs_u32(got), exp)


def test_sharded_ids(capi, small_scene):
    """object_id_base / transform_id_base: a shard produces the same records as the matching slice of the full list."""
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    full, _, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM)
    n = len(sc["objs"])
    parts = []
    for r in range(3):
        a, b = (r * n) // 3, ((r + 1) * n) // 3
        objs = sc["objs"][a:b]
        lo, hi = int(objs["transformId"].min()), int(objs["transformId"].max()) + 1
        with capi.CullContext(0) as ctx:
            ctx.upload_scene(objs, sc["transforms"][lo:hi], sc["surfaces"], sc["lods"], object_id_base=a, transform_id_base=lo)
            ctx.set_view(view)
            ctx.frustum_lod()
            got, _ = ctx.read_draws()
        parts.append(recs_u32(got))
    assert np.array_equal(np.concatenate(parts), full)


def test_against_reference_shader_outputs(capi, built, tables):
    """CUDA path vs tests/golden/spirv_golden.npz: outputs of the reference's OWN shaders (LateDrawCull, InitialDrawCull,
    TransparentDrawCull, OnpcDrawCull, PreClusterDrawCull, InitialClusterCull, DepthPyramidGeneration) executed by
    oracle/spirv_interp.  Same order on both sides (ascending id), so equality is exact."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spirv_golden.npz"))
    views = {str(n): g["views"][i:i + 1] for i, n in enumerate(g["view_names"])}
    objs, transforms = g["objs"], g["transforms"]
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], clusters=tables["clusters"], onpc=g["onpc_objs"],
                         cluster_dispatch_capacity=len(g["cluster_dispatch_inside"]) + 64)
        # depth pyramid: DepthPyramidGeneration.comp.glsl
        ctx.set_depth(g["depth"])
        ctx.build_pyramid(capi.HIZ_VK)
        data, whm, offs = ctx.read_pyramid()
        assert tuple(whm) == tuple(int(x) for x in g["pyramid_whm"])
        assert np.array_equal(data.view(np.uint32), g["pyramid"].view(np.uint32))
        for vn in ("inside", "tilted", "all", "ref_default"):
            ctx.set_view(views[vn])
            ctx.frustum_lod()
            got, tot = ctx.read_draws()
            assert np.array_equal(recs_u32(got), g[f"transparent_{vn}"]), vn
        for vn in ("inside", "tilted"):
            ctx.set_view(views[vn])
            ctx.write_visibility(g["vis0"])
            ctx.early()
            got, _ = ctx.read_draws()
            assert np.array_equal(recs_u32(got), g[f"initial_{vn}"]), vn
            ctx.late(capi.REC_VK24, capi.HIZ_VK)
            got, _ = ctx.read_draws()
            assert np.array_equal(recs_u32(got), g[f"late_{vn}"]), vn
            assert np.array_equal(ctx.read_visibility(), g[f"late_vis_{vn}"]), vn
        ctx.set_view(views["inside"])
        ctx.frustum_lod(capi.LIST_ONPC, capi.REC_VK24, capi.FLAG_ONPC_LOD_QUIRK)
        got, _ = ctx.read_draws()
        assert np.array_equal(recs_u32(got), g["onpc_inside"])
    n = int(g["cluster_objs_count"][0])
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(objs[:n], transforms, tables["surfaces"], tables["lods"], clusters=tables["clusters"],
                         cluster_dispatch_capacity=len(g["cluster_dispatch_inside"]) + 64, draw_capacity=len(g["cluster_dispatch_inside"]) + 64)
        ctx.set_view(views["inside"])
        ctx.cluster_expand()
        rec, tot = ctx.read_cluster_dispatch()
        assert np.array_equal(rec.view(np.uint32).reshape(-1, 3), g["cluster_dispatch_inside"])
        ctx.cluster_cull(capi.CLUSTER_PASSTHROUGH, capi.REC_VK24)
        got, _ = ctx.read_draws()
        assert np.array_equal(recs_u32(got), g["cluster_draws_inside"])
    # the reference's own scene head under the reference's own cameras
    from blitzen_b200 import sceneio, scene
    head = sceneio.read_blob(os.path.join(os.path.dirname(__file__), "golden", "stress_head_4k.blob"))
    rviews = scene.reference_views()
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(head["objs"], head["transforms"], head["surfaces"], head["lods"])
        ctx.set_depth(g["depth"])
        ctx.build_pyramid(capi.HIZ_VK)
        for vn in ("default", "cfg1_centre", "cfg1_tilted", "cfg1_all"):
            ctx.set_view(rviews[vn])
            ctx.frustum_lod()
            got, _ = ctx.read_draws()
            assert np.array_equal(recs_u32(got), g[f"head_transparent_{vn}"]), vn
            ctx.write_visibility(g["head_vis0"])
            ctx.late(capi.REC_VK24, capi.HIZ_VK)
            got, _ = ctx.read_draws()
            assert np.array_equal(recs_u32(got), g[f"head_late_{vn}"]), vn
            assert np.array_equal(ctx.read_visibility(), g[f"head_late_vis_{vn}"]), vn


def test_early_pass_density_switch(capi, medium_scene):
    """The early pass switches between its sparse (pipelined) and dense (streaming) kernels on the density it observed a frame earlier
    (asynchronous device -> pinned-host feedback): the list is the oracle's whichever kernel runs, through both transitions."""
    sc = medium_scene
    n = len(sc["objs"])
    view = view_at(position=(950, 950, 950), z_far=3000.0)
    S = (sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view)
    rng = np.random.default_rng(3)
    dense = np.ones(n, dtype=np.uint32)
    sparse = (rng.random(n) < 0.03).astype(np.uint32)
    exp_d, _, _ = O.cull(*S, O.PASS_EARLY, vis=dense)
    exp_s, _, _ = O.cull(*S, O.PASS_EARLY, vis=sparse)
    with make_ctx(capi, sc) as ctx:
        ctx.set_view(view)
        for vis, exp in ((dense, exp_d), (sparse, exp_s), (dense, exp_d)):
            ctx.write_visibility(vis)
            for _ in range(4):
                ctx.early()
                got, tot = ctx.read_draws()
                assert tot == len(exp) and np.array_equal(recs_u32(got), exp)


def test_against_reference_hlsl_shader_outputs(capi, built, tables):
    """CUDA path vs tests/golden/hlsl_golden.npz: outputs of the reference's OWN D3D12 shaders (depthPyramid, drawCull, drawOccFirst,
    drawOccLate with the point-texel OcclusionCheck, drawOccTemporal, drawInstCountReset + drawInstCull + drawInstCmd) executed by
    oracle/spirv_interp (oracle/spirv_interp/make_hlsl_golden.py).  Ascending id on both sides: exact equality."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hlsl_golden.npz"))
    views = {str(n): g["views"][i:i + 1] for i, n in enumerate(g["view_names"])}
    objs, transforms = g["objs"], g["transforms"]
    nl = len(tables["lods"])
    bucket = int(g["inst_bucket"][0])
    li = g["inst_lod_instances"]
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], lod_instances=li, bucket_capacity=np.full(nl, bucket, dtype=np.uint32))
        for key, dkey in (("pyramid_odd", "depth_odd"), ("pyramid", "depth")):      # depthPyramid.cs.hlsl (the second one stays bound)
            ctx.set_depth(g[dkey])
            ctx.build_pyramid(capi.HIZ_DX)
            data, whm, offs = ctx.read_pyramid()
            assert tuple(whm) == tuple(int(x) for x in g[key + "_whm"])
            assert np.array_equal(data.view(np.uint32), g[key].view(np.uint32)), key
        for vn in ("inside", "tilted", "all", "ref_default"):                         # drawCull.cs.hlsl
            ctx.set_view(views[vn])
            ctx.frustum_lod(capi.LIST_OPAQUE, capi.REC_DX32)
            got, tot = ctx.read_draws(capi.REC_DX32)
            assert tot == len(g[f"drawcull_{vn}"]) and np.array_equal(recs_u32(got), g[f"drawcull_{vn}"]), vn
        for vn in ("inside", "tilted"):
            ctx.set_view(views[vn])
            ctx.write_visibility(g["vis0"])
            ctx.early(capi.REC_DX32)                                                    # drawOccFirst.cs.hlsl
            got, _ = ctx.read_draws(capi.REC_DX32)
            assert np.array_equal(recs_u32(got), g[f"occfirst_{vn}"]), vn
            ctx.temporal(capi.LIST_OPAQUE, capi.REC_DX32, capi.HIZ_DX)                  # drawOccTemporal.hlsl
            got, _ = ctx.read_draws(capi.REC_DX32)
            assert np.array_equal(recs_u32(got), g[f"occtemporal_{vn}"]), vn
            assert np.array_equal(ctx.read_visibility(), g["vis0"]), "the temporal pass must not touch the visibility buffer"
            ctx.late(capi.REC_DX32, capi.HIZ_DX)                                        # drawOccLate.cs.hlsl
            got, _ = ctx.read_draws(capi.REC_DX32)
            assert np.array_equal(recs_u32(got), g[f"occlate_{vn}"]), vn
            assert np.array_equal(ctx.read_visibility(), g[f"occlate_vis_{vn}"]), vn
        for vn in ("inside", "all"):                                                   # drawInst*.cs.hlsl
            ctx.set_view(views[vn])
            ctx.instanced()
            cmds, tot = ctx.read_draws(capi.REC_DX32)
            idx, counters = ctx.read_instances(nl * bucket)
            assert np.array_equal(counters["instanceCount"], g[f"inst_counts_{vn}"]), vn
            assert np.array_equal(recs_u32(cmds), g[f"inst_cmds_{vn}"]), vn
            exp = g[f"inst_indices_{vn}"]
            for l in range(nl):
                a, c = int(li["instanceOffset"][l]), int(counters["instanceCount"][l])
                assert np.array_equal(idx[a:a + c], exp[a:a + c]), f"{vn} bucket {l}"
    from blitzen_b200 import sceneio, scene
    head = sceneio.read_blob(os.path.join(os.path.dirname(__file__), "golden", "stress_head_4k.blob"))
    rviews = scene.reference_views()
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(head["objs"], head["transforms"], head["surfaces"], head["lods"])
        ctx.set_depth(g["depth"])
        ctx.build_pyramid(capi.HIZ_DX)
        for vn in ("default", "cfg1_centre", "cfg1_tilted", "cfg1_all"):
            ctx.set_view(rviews[vn])
            ctx.write_visibility(g["head_vis0"])
            ctx.late(capi.REC_DX32, capi.HIZ_DX)
            got, _ = ctx.read_draws(capi.REC_DX32)
            assert np.array_equal(recs_u32(got), g[f"head_occlate_{vn}"]), vn
            assert np.array_equal(ctx.read_visibility(), g[f"head_occlate_vis_{vn}"]), vn


@pytest.mark.parametrize("hiz", [0, 1])
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("list_id", [0, 1])
def test_temporal_pass(capi, small_scene, hiz, fmt, list_id):
    """blz_cull_temporal (HlslShaders/CS/drawOccTemporal.hlsl:13-65): frustum + Hi-Z in one pass, no visibility buffer read or written;
    both Hi-Z variants, both record formats, the opaque and the transparent list."""
    from blitzen_b200 import scene
    sc = small_scene
    W, H = 640, 360
    view = view_at(position=(380, 380, 380), z_far=2000.0, width=W, height=H)
    depth = scene.synthetic_depth(W, H, n_rects=40, z_min=20.0, z_max=400.0, seed=99)
    rw = 6 if fmt == 0 else 8
    transp = sc["objs"][10000:31000].copy()
    lst = sc["objs"] if list_id == 0 else transp
    pyr = O.build_pyramid(depth, hiz)
    exp, tot, _ = O.cull(lst, sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_TEMPORAL, rec_words=rw, hiz=hiz, pyramid=pyr)
    fr, ftot, _ = O.cull(lst, sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, rec_words=rw)
    vis0 = (np.arange(len(sc["objs"])) % 3 == 0).astype(np.uint32)
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], transparent=transp)
        ctx.set_view(view)
        ctx.write_visibility(vis0)
        ctx.set_depth(depth)
        ctx.build_pyramid(hiz)
        for _ in range(2):
            ctx.temporal(list_id, fmt, hiz)
            got, gtot = ctx.read_draws(fmt)
            assert gtot == tot and np.array_equal(recs_u32(got), exp)
        assert np.array_equal(ctx.read_visibility(), vis0)
    assert 0 < tot < ftot          # the Hi-Z part rejected something


@pytest.mark.parametrize("variant", [0, 1])
def test_pyramid_4k_and_late(capi, medium_scene, variant):
    """BASELINE config 5's depth size: 3840x2160 -> 2048x2048 x 11 (Vulkan; the 64x64 tail path of pyramid.cu) / 1920x1080 x 11 (D3D12),
    then a late pass against it, both bit-exact."""
    from blitzen_b200 import scene
    sc = medium_scene
    W, H = 3840, 2160
    depth = scene.synthetic_depth(W, H, n_rects=64, z_min=50.0, z_max=1500.0, seed=4)
    rng = np.random.default_rng(8)
    depth += (rng.random((H, W), dtype=np.float32) * np.float32(1e-4))
    exp = O.build_pyramid(depth, variant, threads=8)
    view = view_at(position=(950, 950, 950), z_far=5000.0, width=W, height=H)
    n = len(sc["objs"])
    vis0 = (rng.random(n) < 0.4).astype(np.uint32)
    l_exp, l_tot, vis_exp = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_LATE, hiz=variant, pyramid=exp, vis=vis0, threads=8)
    for tma in (1, 0):
        with make_ctx(capi, sc) as ctx:
            ctx.set_option("pyramid_tma", tma)
            ctx.set_depth(depth)
            ctx.build_pyramid(variant)
            ctx.build_pyramid(variant)
            data, (pw, ph, mips), offs = ctx.read_pyramid()
            assert (pw, ph, mips) == (exp.width, exp.height, exp.mips)
            assert mips == (11 if variant == 0 else 10)
            assert np.array_equal(data.view(np.uint32), exp.data[:len(data)].view(np.uint32)), f"4K pyramid mismatch tma={tma}"
            ctx.set_view(view)
            ctx.write_visibility(vis0)
            ctx.late(capi.REC_VK24, variant)
            got, gtot = ctx.read_draws()
            assert gtot == l_tot and np.array_equal(recs_u32(got), l_exp)
            assert np.array_equal(ctx.read_visibility(), vis_exp)
    assert 0 < int(vis_exp.sum()) < n


def test_status_tag_restart(capi, medium_scene):
    """the 30-bit tag of the tile status words must not wrap into stale slots.  Option epoch_wrap_at makes the host-side
    guard restart the tag (clear the status array, epoch = 1) every few launches; passes of different tile counts stay byte-exact."""
    sc = medium_scene
    view = view_at(position=(950, 950, 950), z_far=5000.0)
    exp, total, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, threads=8)
    small = 70_001
    exp_s, total_s, _ = O.cull(sc["objs"][:small], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, threads=8)
    with make_ctx(capi, sc) as ctx, make_ctx(capi, dict(sc, objs=sc["objs"][:small])) as ctx_s:
        for c in (ctx, ctx_s):
            c.set_view(view)
            c.set_option("epoch_wrap_at", 5)
        for it in range(14):                       # several restarts, at different points of the early / late / frustum sequence
            ctx.frustum_lod()
            got, gtot = ctx.read_draws()
            assert gtot == total and np.array_equal(recs_u32(got), exp), it
            ctx_s.frustum_lod()
            got, gtot = ctx_s.read_draws()
            assert gtot == total_s and np.array_equal(recs_u32(got), exp_s), it
