"""BASELINE.json configs[1] at FULL size (16 777 216 objects, 1920x1080 depth, the bench workload): the CUDA path against the oracle,
byte for byte (the multithreaded oracle finishes a frame of this size in well under a second per pass), plus the size-independent
properties of the domain: ascending ids, count == population of the visibility mask, idempotence of the late pass under a fixed
view, shard additivity (two half-scene contexts concatenate to the full list), and the same lists from every kernel variant."""
import hashlib
import os
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

N = 16_777_216


@pytest.fixture(scope="module")
def workload(built):
    import bench as B
    return B.build_workload(N, 0, 1), B


def recs_u32(rec):
    return rec.view(np.uint32).reshape(len(rec), rec.dtype.itemsize // 4)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_two_phase_frame_full_size_against_oracle(workload):
    from blitzen_b200 import capi
    w, B = workload
    threads = O.hardware_threads()
    kw = dict(threads=threads, transform_id_base=w["transform_id_base"])
    S = (w["objs"], w["transforms"], w["surfaces"], w["lods"], w["view"])
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], transform_id_base=w["transform_id_base"])
        ctx.set_view(w["view"]); ctx.set_depth(w["depth"])
        # frustum + LOD
        exp, tot, _ = O.cull(*S, O.PASS_FRUSTUM, **kw)
        ctx.frustum_lod()
        got, gtot = ctx.read_draws()
        g = recs_u32(got)
        assert gtot == tot == len(exp) and sha(g) == sha(exp)
        assert np.all(np.diff(g[:, 0].astype(np.int64)) > 0)                     # ascending objectId
        # frame 0: cleared pyramid, visibility 0 -> every frustum survivor is emitted by the late pass
        vis = np.zeros(N, dtype=np.uint32)
        exp0, tot0, vis = O.cull(*S, O.PASS_LATE, hiz=O.HIZ_VK, pyramid=O.cleared_pyramid(B.DEPTH_W, B.DEPTH_H, O.HIZ_VK), vis=vis, **kw)
        ctx.reset_visibility(); ctx.clear_pyramid(capi.HIZ_VK, B.DEPTH_W, B.DEPTH_H)
        ctx.late(capi.REC_VK24, capi.HIZ_VK)
        got, gtot = ctx.read_draws()
        assert gtot == tot0 == tot and sha(recs_u32(got)) == sha(exp0)
        gvis = ctx.read_visibility()
        assert sha(gvis) == sha(vis) and int(gvis.sum()) == gtot                 # count == population of the mask
        # frame 1: early list = last frame's visible set, real pyramid, late list = newly visible (none: fixed view), mask shrinks to the unoccluded
        pyr = O.build_pyramid(w["depth"], O.HIZ_VK, threads=threads)
        expE, totE, _ = O.cull(*S, O.PASS_EARLY, vis=vis, **kw)
        expL, totL, vis1 = O.cull(*S, O.PASS_LATE, hiz=O.HIZ_VK, pyramid=pyr, vis=vis, **kw)
        ctx.early()
        got, gtot = ctx.read_draws()
        assert gtot == totE and sha(recs_u32(got)) == sha(expE)
        ctx.build_pyramid(capi.HIZ_VK)
        ctx.late(capi.REC_VK24, capi.HIZ_VK)
        got, gtot = ctx.read_draws()
        assert gtot == totL == 0 and len(got) == 0
        gvis1 = ctx.read_visibility()
        assert sha(gvis1) == sha(vis1) and 0 < int(gvis1.sum()) < int(gvis.sum())
        # idempotence under a fixed view: another frame changes nothing
        ctx.early(); _, e2 = ctx.read_count()
        ctx.late(capi.REC_VK24, capi.HIZ_VK); _, l2 = ctx.read_count()
        assert e2 == int(gvis1.sum()) and l2 == 0 and sha(ctx.read_visibility()) == sha(vis1)
        # every kernel variant produces the same bytes (v4 pipelined kernel, generic early pass, u32 visibility words)
        ref_e = None
        for opts in ({"draw_kernel": 1, "early_mode": 3}, {"draw_kernel": 0, "early_mode": 0}, {"draw_kernel": 1, "early_mode": 1, "vis_words": 1}):
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.write_visibility(vis)
            ctx.early(); ge, _ = ctx.read_draws()
            ctx.late(capi.REC_VK24, capi.HIZ_VK); gl, glt = ctx.read_draws()
            assert sha(recs_u32(ge)) == sha(expE) and glt == 0 and sha(ctx.read_visibility()) == sha(vis1), opts


def test_shard_additivity_full_size(workload):
    """Two contexts, each owning half of the objects (global ids via object_id_base): their lists concatenate to the full list."""
    from blitzen_b200 import capi
    w, B = workload
    h = N // 2
    full = None
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(w["objs"], w["transforms"], w["surfaces"], w["lods"], transform_id_base=w["transform_id_base"])
        ctx.set_view(w["view"]); ctx.frustum_lod()
        full, ftot = ctx.read_draws()
    parts = []
    for a, b in ((0, h), (h, N)):
        with capi.CullContext(0) as ctx:
            ctx.upload_scene(w["objs"][a:b], w["transforms"], w["surfaces"], w["lods"], object_id_base=a, transform_id_base=w["transform_id_base"])
            ctx.set_view(w["view"]); ctx.frustum_lod()
            got, _ = ctx.read_draws()
            parts.append(recs_u32(got))
    assert sha(np.concatenate(parts)) == sha(recs_u32(full)) and len(full) == ftot
