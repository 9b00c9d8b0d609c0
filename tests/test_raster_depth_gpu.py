"""device-side software depth, so that  early -> depth -> pyramid -> late  is a closed loop on the device.
blz_cull_raster_depth (csrc/raster_depth.cu) against its CPU definition oracle_raster_depth (built from the reference's own
IsObjectInsideViewFrustum prologue + projectSphere, CullingShaderData.glsl:8-58): the depth image bit for bit, and a three-frame
closed loop with a moving camera in which every list, every count and the visibility buffer match the oracle frame by frame."""
import numpy as np
import pytest

import oracle_lib as O
from blitzen_b200 import capi
from conftest import view_at

pytestmark = pytest.mark.gpu

VIEWS = {
    "centre": dict(position=(380, 380, 380), z_far=2000.0),
    "tilted": dict(position=(200, 500, 100), yaw=0.7, pitch=-0.3, z_far=900.0),
    "corner": dict(position=(20, 70, 0), z_far=650.0),
}


def u32(rec):
    return rec.view(np.uint32).reshape(len(rec), rec.dtype.itemsize // 4)


@pytest.mark.parametrize("vname", list(VIEWS))
@pytest.mark.parametrize("fmt,size", [(capi.REC_VK24, (640, 360)), (capi.REC_DX32, (1920, 1080)), (capi.REC_VK24, (333, 217))])
def test_depth_image_matches_the_oracle(built, small_scene, vname, fmt, size):
    sc = small_scene
    view = view_at(**VIEWS[vname], width=size[0], height=size[1])
    W, H = size
    ctx = capi.CullContext(0)
    try:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view)
        ctx.frustum_lod(fmt=fmt)
        rec, total = ctx.read_draws()
        ctx.raster_depth(W, H)
        got = ctx.read_depth()
    finally:
        ctx.close()
    exp = O.raster_depth(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, u32(rec), W, H)
    assert got.shape == exp.shape == (H, W)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    if total:
        covered = int((exp > 0).sum())
        assert 0 < covered                                          # the list actually drew something
        assert float(exp.max()) <= 1.0                              # reverse-Z depth of something beyond the near plane


def test_large_boxes_take_the_cta_path(built, small_scene):
    """Objects next to the camera project to boxes of > 4096 pixels (the shared-memory queue); a 4K target makes many medium ones."""
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=60.0, width=3840, height=2160)
    ctx = capi.CullContext(0)
    try:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view)
        ctx.frustum_lod()
        rec, total = ctx.read_draws()
        ctx.raster_depth(3840, 2160)
        got = ctx.read_depth()
    finally:
        ctx.close()
    exp = O.raster_depth(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, u32(rec), 3840, 2160)
    assert total > 0 and np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert int((exp > 0).sum()) > 4096 * 4


@pytest.mark.parametrize("hiz", [capi.HIZ_VK, capi.HIZ_DX])
def test_closed_loop_three_frames(built, medium_scene, hiz):
    """frame k: early (last frame's visible set) -> software depth of that list -> pyramid -> late.  Camera moves every frame."""
    sc = medium_scene
    W, H = 1280, 720
    ohiz = O.HIZ_VK if hiz == capi.HIZ_VK else O.HIZ_DX
    cams = [dict(position=(950, 950, 950), z_far=5000.0), dict(position=(960, 945, 955), yaw=0.05, z_far=5000.0), dict(position=(975, 940, 960), yaw=0.11, pitch=-0.04, z_far=5000.0)]
    kw = dict(threads=8)
    ctx = capi.CullContext(0)
    try:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        # frame 0 of the reference: visibility 0, pyramid cleared -> the late pass emits every frustum survivor
        v0 = view_at(**cams[0], width=W, height=H)
        ctx.set_view(v0)
        ctx.clear_pyramid(hiz, W, H)
        ctx.late(capi.REC_VK24, hiz)
        _, _, vis = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], v0, O.PASS_LATE, hiz=ohiz, pyramid=O.cleared_pyramid(W, H, ohiz),
                           vis=np.zeros(len(sc["objs"]), dtype=np.uint32), **kw)
        assert np.array_equal(ctx.read_visibility(), vis)
        culled_by_hiz = 0
        for cam in cams:
            view = view_at(**cam, width=W, height=H)
            ctx.set_view(view)
            ctx.early(capi.REC_VK24)
            e_got, e_tot = ctx.read_draws()
            e_exp, e_etot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_EARLY, vis=vis, **kw)
            assert e_tot == e_etot and np.array_equal(u32(e_got), e_exp)
            ctx.raster_depth(W, H)
            depth = O.raster_depth(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, e_exp, W, H)
            assert np.array_equal(ctx.read_depth().view(np.uint32), depth.view(np.uint32))
            ctx.build_pyramid(hiz)
            pyr = O.build_pyramid(depth, ohiz, threads=8)
            ctx.late(capi.REC_VK24, hiz)
            l_got, l_tot = ctx.read_draws()
            l_exp, l_etot, vis_new = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_LATE, hiz=ohiz, pyramid=pyr, vis=vis, **kw)
            assert l_tot == l_etot and np.array_equal(u32(l_got), l_exp)
            assert np.array_equal(ctx.read_visibility(), vis_new)
            fr, fr_tot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, **kw)
            culled_by_hiz += fr_tot - int(vis_new.sum())
            vis = vis_new
        assert culled_by_hiz > 0                                    # the loop's own depth really occludes something
    finally:
        ctx.close()
