"""BASELINE.json configs[1] at FULL size (16 777 216 objects, 1920x1080 depth, the bench workload): the CUDA path against the oracle,
byte for byte (the multithreaded oracle finishes a frame of this size in well under a second per pass), plus the size-independent
properties of the domain: ascending ids, count == population of the visibility mask, idempotence of the late pass under a fixed
view, shard additivity (two half-scene contexts concatenate to the full list), and the same lists from every early-pass variant."""
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
        # every kernel variant produces the same bytes (pipelined / streaming early pass, 1-bit mask / u32 visibility words as the source)
        for opts in ({"early_mode": 1}, {"early_mode": 0}, {"early_mode": 1, "early_bits": 0, "vis_words": 1}):
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


def test_instancing_and_cluster_path_full_size_against_oracle(workload):
    """The other rows of the path at the bench size: indirect instancing (16.7 M objects, buckets sized from the oracle's counts, one bucket
    deliberately too small) and the cluster path (expand -> 50 M dispatch records -> passthrough and bounding-sphere cull), byte for byte."""
    from blitzen_b200 import capi
    w, B = workload
    threads = O.hardware_threads()
    kw = dict(threads=threads, transform_id_base=w["transform_id_base"])
    S = (w["objs"], w["transforms"], w["surfaces"], w["lods"])
    view = w["view"]
    nl = len(w["lods"])
    # ---- instancing
    li = w["lodInstances"].copy()
    li["instanceOffset"] = np.arange(nl, dtype=np.uint32)
    _, cnt, _ = O.cull_instanced(*S, li, np.ones(nl, dtype=np.uint32), view, **kw)
    cap = np.maximum(cnt, 1).astype(np.uint32)
    big = int(np.argmax(cnt))
    cap[big] = cnt[big] // 2                                                    # overflowing bucket: ids beyond the capacity are dropped, the count is not
    li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
    idx_e, cnt_e, cmds_e = O.cull_instanced(*S, li, cap, view, **kw)
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(*S, lod_instances=li, bucket_capacity=cap, transform_id_base=w["transform_id_base"])
        ctx.set_view(view)
        ctx.instanced()
        cmds, total = ctx.read_draws(capi.REC_DX32)
        idx, counters = ctx.read_instances(int(cap.sum()))
    assert np.array_equal(counters["instanceCount"], cnt_e) and int(cnt_e.sum()) > 2_000_000
    assert sha(recs_u32(cmds)) == sha(cmds_e) and total == len(cmds_e)
    for l in range(nl):
        o, c = int(li["instanceOffset"][l]), int(min(cnt_e[l], cap[l]))
        assert sha(idx[o:o + c]) == sha(idx_e[o:o + c]), l
        assert c < 2 or np.all(np.diff(idx[o:o + c].astype(np.int64)) > 0)       # ascending ids inside every bucket
    # ---- cluster path
    capacity = 64_000_000
    d_exp, d_tot = O.cluster_expand(*S, view, capacity, **kw)
    assert 40_000_000 < d_tot <= capacity
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(*S, clusters=w["clusters"], transform_id_base=w["transform_id_base"], cluster_dispatch_capacity=capacity, draw_capacity=capacity)
        ctx.set_view(view)
        ctx.cluster_expand()
        got, gtot = ctx.read_cluster_dispatch()
        assert gtot == d_tot and sha(got.view(np.uint32)) == sha(d_exp)
        del got
        for mode, omode in ((capi.CLUSTER_PASSTHROUGH, 0), (capi.CLUSTER_SPHERE, 1)):
            exp, tot = O.cluster_cull(*S, w["clusters"], view, d_exp, omode, **kw)
            ctx.cluster_cull(mode, capi.REC_VK24)
            draws, dtot = ctx.read_draws(capi.REC_VK24)
            assert dtot == tot and sha(recs_u32(draws)) == sha(exp), mode
            del draws, exp
        # bounding sphere + Hi-Z (BASELINE config 4's full test), both Hi-Z variants, against the bench's 1920x1080 depth image
        ctx.set_depth(w["depth"])
        for hiz in (capi.HIZ_VK, capi.HIZ_DX):
            ctx.build_pyramid(hiz)
            pyr = O.build_pyramid(w["depth"], hiz, threads=threads)
            exp, tot = O.cluster_cull(*S, w["clusters"], view, d_exp, 1, hiz=hiz, pyramid=pyr, **kw)
            ctx.cluster_cull(capi.CLUSTER_SPHERE_HIZ, capi.REC_VK24, hiz)
            draws, dtot = ctx.read_draws(capi.REC_VK24)
            assert dtot == tot and 0 < tot < d_tot and sha(recs_u32(draws)) == sha(exp), ("sphere_hiz", hiz)
            del draws, exp
