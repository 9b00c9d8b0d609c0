"""Draw-list gather on ONE GPU (world = 1: the presenter pushes into its own buffer): exercises the C-ABI gather entry points, the
asynchronous push (side stream, double-buffered draw lists, two-half gather buffer) and the stream-ordered join inside the normal GPU
suite; the multi-rank form is tests/mgpu_verify_gather.py (torchrun, 2/4/8 GPUs)."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import view_at

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(built):
    from blitzen_b200 import capi
    return capi


def recs_u32(rec):
    return rec.view(np.uint32).reshape(len(rec), rec.dtype.itemsize // 4)


@pytest.mark.parametrize("fmt_name", ["VK24", "DX32"])
@pytest.mark.parametrize("desc", [1, 0], ids=["descriptors", "records"])
def test_gather_sync_and_async_single_rank(capi, small_scene, desc, fmt_name):
    """desc = 1: the ranks ship 8-byte {objectId, lodId} descriptors and the presenter expands them with its own LOD table (gather_expand_kernel);
    desc = 0: the 24-/32-byte records themselves travel.  The gathered list is the same bytes either way."""
    fmt = getattr(capi, "REC_" + fmt_name)
    rw = 6 if fmt_name == "VK24" else 8
    sc = small_scene
    n = len(sc["objs"])
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    from blitzen_b200 import scene
    depth = scene.synthetic_depth(640, 360, n_rects=40, z_min=20.0, z_max=300.0, seed=5)
    S = (sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view)
    vis0 = np.zeros(n, dtype=np.uint32)
    l0, _, vis1 = O.cull(*S, O.PASS_LATE, rec_words=rw, pyramid=O.cleared_pyramid(640, 360, O.HIZ_VK), vis=vis0)
    pyr = O.build_pyramid(depth, O.HIZ_VK)
    e1, _, _ = O.cull(*S, O.PASS_EARLY, rec_words=rw, vis=vis1)
    l1, _, vis2 = O.cull(*S, O.PASS_LATE, rec_words=rw, pyramid=pyr, vis=vis1)
    e2, _, _ = O.cull(*S, O.PASS_EARLY, rec_words=rw, vis=vis2)
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view); ctx.set_depth(depth)
        ctx.set_option("gather_desc", desc)
        blob = ctx.gather_export(n, fmt)
        ctx.gather_import(None, 0, 1, n, fmt)
        # frame 0, synchronous push
        ctx.clear_pyramid(capi.HIZ_VK, 640, 360)
        ctx.late(fmt); ctx.gather_push(1)
        got, counts = ctx.gather_read(1, 1, fmt)
        assert counts[0] == len(l0) and np.array_equal(recs_u32(got), l0)
        # frame 1, asynchronous pushes: both lists of the frame are intact on the presenter afterwards (two halves of the gather buffer)
        ctx.early(fmt); ctx.gather_push_async(2)
        ctx.build_pyramid(capi.HIZ_VK)
        ctx.late(fmt); ctx.gather_push_async(3)
        ctx.gather_join(); ctx.synchronize()
        gotE, cE = ctx.gather_read(2, 1, fmt)
        gotL, cL = ctx.gather_read(3, 1, fmt)
        assert cE[0] == len(e1) and np.array_equal(recs_u32(gotE), e1)
        assert cL[0] == len(l1) and np.array_equal(recs_u32(gotL), l1)
        assert np.array_equal(ctx.read_visibility(), vis2)
        # frame 2: the draw buffers have flipped twice; passes and pushes keep working, read_draws sees the pass that ran last
        ctx.early(fmt)
        last, _ = ctx.read_draws(fmt)
        assert np.array_equal(recs_u32(last), e2)
        ctx.gather_push_async(4)
        ctx.late(fmt); ctx.gather_push_async(5)
        ctx.gather_join()
        gotE, _ = ctx.gather_read(4, 1, fmt)
        assert np.array_equal(recs_u32(gotE), e2)


@pytest.mark.parametrize("desc", [1, 0], ids=["descriptors", "records"])
def test_absent_peer_times_out_instead_of_hanging(capi, small_scene, desc):
    """Failure detection: this context plays rank 1 of 2 and nobody plays rank 0, so its push waits on the device for rank 0's record count.
    The wait is bounded (option gather_timeout_ms): the kernels end, the next synchronising call reports BLZ_ERR_TIMEOUT naming the peer
    and the epoch, and the context stays usable (a fresh gather set-up works again)."""
    import time
    sc = small_scene
    n = len(sc["objs"])
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view)
        ctx.set_option("gather_desc", desc)
        ctx.set_option("gather_timeout_ms", 200)
        ctx.gather_export(n, capi.REC_VK24)
        ctx.gather_import(None, 1, 2, n, capi.REC_VK24)          # the exporter writes through its own pointers whatever its rank
        ctx.frustum_lod()
        expect, _ = ctx.read_draws()
        t0 = time.time()
        ctx.gather_push(1)
        with pytest.raises(capi.BlzError, match=r"error -6.*rank 0"):
            ctx.synchronize()
        assert time.time() - t0 < 5.0
        ctx.synchronize()                                          # reported once; the context is usable again
        # a later push of the same (broken) set-up does not pay the budget again before it has been reported ... and a fresh set-up works
        ctx.gather_export(n, capi.REC_VK24)
        ctx.gather_import(None, 0, 1, n, capi.REC_VK24)
        ctx.frustum_lod(); ctx.gather_push(1)
        got, counts = ctx.gather_read(1, 1)
        assert counts[0] == len(expect) and np.array_equal(recs_u32(got), recs_u32(expect))
