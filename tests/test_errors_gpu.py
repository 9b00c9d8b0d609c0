"""Error convention of the C ABI (every export returns a status and sets a last-error string; it never aborts): misuse that the
reference would hit as a BLIT_ERROR / failed VK_CHECK comes back as a negative status with a message, and the context stays usable."""
import numpy as np
import pytest

from conftest import view_at

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(built):
    from blitzen_b200 import capi
    return capi


def test_misuse_is_reported_not_fatal(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    with capi.CullContext(0) as ctx:
        with pytest.raises(capi.BlzError, match="no scene"):
            ctx.frustum_lod()
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        with pytest.raises(capi.BlzError, match="no view"):
            ctx.frustum_lod()
        ctx.set_view(view)
        with pytest.raises(capi.BlzError, match="pyramid"):
            ctx.late(capi.REC_VK24, capi.HIZ_VK)                       # no depth pyramid yet
        with pytest.raises(capi.BlzError, match="lod_instances"):
            ctx.instanced()                                            # scene uploaded without the instancing tables
        with pytest.raises(capi.BlzError, match="cluster_dispatch_capacity"):
            ctx.cluster_expand()
        with pytest.raises(capi.BlzError, match="list"):
            ctx.frustum_lod(7)
        with pytest.raises(capi.BlzError, match="record format"):
            ctx.frustum_lod(capi.LIST_OPAQUE, 9)
        with pytest.raises(capi.BlzError, match="transform range"):
            ctx.update_transforms(len(sc["transforms"]) - 1, sc["transforms"][:2])
        with pytest.raises(capi.BlzError, match="gather not set up"):
            ctx.gather_push(1)
        ctx.clear_pyramid(capi.HIZ_VK, 640, 360)
        with pytest.raises(capi.BlzError, match="variant"):
            ctx.late(capi.REC_VK24, capi.HIZ_DX)                       # pyramid of the other variant
        # the context is still good
        ctx.frustum_lod()
        _, tot = ctx.read_count()
        ctx.late(capi.REC_VK24, capi.HIZ_VK)
        _, tot_late = ctx.read_count()
        assert tot > 0 and tot_late == tot                             # frame 0: every frustum survivor is emitted


def test_malformed_scene_is_refused(capi, small_scene):
    """ids the kernels would index out of bounds with come back as an error from upload_scene; the context stays usable."""
    sc = small_scene
    ctx = capi.CullContext(0)
    try:
        bad = sc["objs"].copy()
        bad["surfaceId"][123] = len(sc["surfaces"]) + 7
        with pytest.raises(capi.BlzError, match="surfaceId"):
            ctx.upload_scene(bad, sc["transforms"], sc["surfaces"], sc["lods"])
        with pytest.raises(capi.BlzError):
            ctx.frustum_lod()                                   # nothing runs on the refused scene
        bad = sc["objs"].copy()
        bad["transformId"][5] = len(sc["transforms"]) + 100
        with pytest.raises(capi.BlzError, match="transformId"):
            ctx.upload_scene(bad, sc["transforms"], sc["surfaces"], sc["lods"])
        surf = sc["surfaces"].copy()
        surf["lodOffset"][1] = len(sc["lods"])
        with pytest.raises(capi.BlzError, match="LOD range"):
            ctx.upload_scene(sc["objs"], sc["transforms"], surf, sc["lods"])
        lods = sc["lods"].copy()
        lods["clusterOffset"][3] = len(sc["clusters"]) + 1
        with pytest.raises(capi.BlzError, match="cluster range"):
            ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], lods, clusters=sc["clusters"])
        # a good scene afterwards: the context works
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view_at(position=(380, 380, 380), z_far=2000.0))
        ctx.frustum_lod()
        _, total = ctx.read_draws()
        assert total > 0
    finally:
        ctx.close()
