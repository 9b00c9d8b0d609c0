"""Survivor-list pipeline (cull_stream.cu PASS_FRUSTUM with 8-byte records -> cull_list.cu) for indirect instancing and cluster
expansion: byte-exact against the oracle over the edge cases (nothing / everything visible, overflowing buckets, dispatch capacity
clamp, many tiles, odd tile counts)."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import view_at

pytestmark = pytest.mark.gpu

VIEWS = {
    "centre": dict(position=(380, 380, 380), z_far=2000.0),
    "outside_all": dict(position=(380, 380, -2500), z_far=1e9),
    "nothing": dict(position=(380, 380, 5000), z_far=100.0),
    "tilted": dict(position=(200, 500, 100), yaw=0.7, pitch=-0.3, z_far=900.0),
}


@pytest.fixture(scope="module")
def capi(built):
    from blitzen_b200 import capi
    return capi


def recs_u32(rec):
    return rec.view(np.uint32).reshape(len(rec), rec.dtype.itemsize // 4)


def make_ctx(capi, sc, **kw):
    ctx = capi.CullContext(0)
    ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], clusters=sc["clusters"], lod_instances=kw.pop("lod_instances", None), **kw)
    return ctx


def bucket_layout(sc, cap_value, tight=None):
    li = sc["lodInstances"].copy()
    nl = len(sc["lods"])
    cap = np.full(nl, cap_value, dtype=np.uint32)
    for l, c in (tight or {}).items():
        cap[l] = c
    li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
    return li, cap


@pytest.mark.parametrize("vname", list(VIEWS))
def test_instancing_views(capi, small_scene, vname):
    sc = small_scene
    view = view_at(**VIEWS[vname])
    li, cap = bucket_layout(sc, 70_000, tight={0: 5, 7: 11})
    idx_e, cnt_e, cmds_e = O.cull_instanced(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], li, cap, view)
    nl = len(cap)
    with make_ctx(capi, sc, lod_instances=li, bucket_capacity=cap) as ctx:
        ctx.set_view(view)
        for mode in (0, 1, 2):                               # repeated launches reuse the re-armed state
            ctx.instanced()
            cmds, total = ctx.read_draws(capi.REC_DX32)
            idx, counters = ctx.read_instances(int(cap.sum()))
            assert np.array_equal(counters["instanceCount"], cnt_e), (vname, mode)
            assert np.array_equal(recs_u32(cmds), cmds_e) and total == len(cmds_e), (vname, mode)
            for l in range(nl):
                o, c = int(li["instanceOffset"][l]), int(min(cnt_e[l], cap[l]))
                assert np.array_equal(idx[o:o + c], idx_e[o:o + c]), (vname, mode, l)


def test_instancing_medium_many_tiles(capi, medium_scene):
    """1 M objects, everything visible: > 512 chunks of the counting sort, every LOD bucket crosses tile boundaries."""
    sc = medium_scene
    view = view_at(position=(950, 950, -6000), z_far=1e9)
    li, cap = bucket_layout(sc, len(sc["objs"]) // 4)
    idx_e, cnt_e, cmds_e = O.cull_instanced(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], li, cap, view)
    with make_ctx(capi, sc, lod_instances=li, bucket_capacity=cap) as ctx:
        ctx.set_view(view)
        ctx.instanced()
        cmds, total = ctx.read_draws(capi.REC_DX32)
        idx, counters = ctx.read_instances(int(cap.sum()))
        assert int(cnt_e.sum()) > 900_000
        assert np.array_equal(counters["instanceCount"], cnt_e)
        assert np.array_equal(recs_u32(cmds), cmds_e) and total == len(cmds_e)
        for l in range(len(cap)):
            o, c = int(li["instanceOffset"][l]), int(min(cnt_e[l], cap[l]))
            assert np.array_equal(idx[o:o + c], idx_e[o:o + c]), l


@pytest.mark.parametrize("vname", list(VIEWS))
def test_cluster_expand_views(capi, small_scene, vname):
    sc = small_scene
    view = view_at(**VIEWS[vname])
    capacity = 6_000_000
    d_exp, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, capacity)
    with make_ctx(capi, sc, cluster_dispatch_capacity=capacity, draw_capacity=16) as ctx:
        ctx.set_view(view)
        for mode in (0, 1, 2):
            ctx.cluster_expand()
            got, gtot = ctx.read_cluster_dispatch()
            assert gtot == d_tot, (vname, mode)
            assert np.array_equal(got.view(np.uint32).reshape(-1, 3), d_exp), (vname, mode)


@pytest.mark.parametrize("capacity", [1, 16383, 16384, 16385, 100_001])
def test_cluster_expand_capacity_clamp(capi, small_scene, capacity):
    """total > capacity: exactly `capacity` records are written (slice boundaries of the write step included), total is still reported."""
    sc = small_scene
    view = view_at(**VIEWS["outside_all"])
    d_exp, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, capacity)
    assert d_tot > 100_001 and len(d_exp) == capacity
    with make_ctx(capi, sc, cluster_dispatch_capacity=capacity, draw_capacity=16) as ctx:
        ctx.set_view(view)
        ctx.cluster_expand()
        got, gtot = ctx.read_cluster_dispatch()
        assert gtot == d_tot and len(got) == capacity
        assert np.array_equal(got.view(np.uint32).reshape(-1, 3), d_exp)


def test_cluster_expand_medium_then_cull(capi, medium_scene):
    """1 M objects: heavy tiles (dragons at LOD 0) are cut into many work items; expand -> passthrough cull stays on the device."""
    sc = medium_scene
    view = view_at(position=(950, 950, 950), z_far=700.0)
    capacity = 40_000_000
    d_exp, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, capacity)
    assert 1_000_000 < d_tot <= capacity
    with make_ctx(capi, sc, cluster_dispatch_capacity=capacity, draw_capacity=capacity) as ctx:
        ctx.set_view(view)
        ctx.cluster_expand()
        got, gtot = ctx.read_cluster_dispatch()
        assert gtot == d_tot
        assert np.array_equal(got.view(np.uint32).reshape(-1, 3), d_exp)
        exp, tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_exp, 1)
        ctx.cluster_cull(capi.CLUSTER_SPHERE, capi.REC_VK24)
        draws, dtot = ctx.read_draws(capi.REC_VK24)
        assert dtot == tot and np.array_equal(recs_u32(draws), exp)


def test_cluster_cull_tight_capacity_all_modes(capi, small_scene):
    """Dispatch buffer exactly as long as the record list (regression: the per-tile status array was sized for 1024-record tiles while
    the Hi-Z variant of the cluster cull uses 768-record tiles -> out-of-bounds once the list filled more than 3/4 of the capacity)."""
    sc = small_scene
    view = view_at(**VIEWS["outside_all"])
    d_all, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, 10_000_000)
    assert d_tot == len(d_all) > 300_000
    from blitzen_b200 import scene
    depth = scene.synthetic_depth(640, 360, n_rects=40, z_min=20.0, z_max=300.0, seed=5)
    pyr = O.build_pyramid(depth, 0)
    with make_ctx(capi, sc, cluster_dispatch_capacity=d_tot, draw_capacity=d_tot) as ctx:
        ctx.set_view(view); ctx.set_depth(depth); ctx.build_pyramid(0)
        ctx.cluster_expand()
        got, gtot = ctx.read_cluster_dispatch()
        assert gtot == d_tot and np.array_equal(got.view(np.uint32).reshape(-1, 3), d_all)
        for mode, kw in ((capi.CLUSTER_SPHERE_HIZ, dict(hiz=0, pyramid=pyr)), (capi.CLUSTER_SPHERE, {}), (capi.CLUSTER_PASSTHROUGH, {})):
            exp, tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_all,
                                      0 if mode == capi.CLUSTER_PASSTHROUGH else 1, **kw)
            ctx.cluster_cull(mode, capi.REC_VK24, 0)
            draws, dtot = ctx.read_draws(capi.REC_VK24)
            assert dtot == tot and np.array_equal(recs_u32(draws), exp), mode


@pytest.mark.parametrize("n", [2048, 3000, 4095, 6144, 8191])
def test_cluster_expand_odd_tile_counts(capi, tables, n):
    """Regression: with floor(n / 2048) odd the work-item array of the expand step sat at an odd word offset and its 8-byte
    accesses faulted ('misaligned address'); every size the other tests use has an even quotient."""
    from blitzen_b200 import scene
    view = view_at(position=(0, 0, -80), z_far=1e6)
    objs, xf = scene.generate(groups=((0, 5.0, n - n // 3), (2, 1.0, n // 3)), multiplier=40.0, prologue=False, prng="counter", seed=n)
    transforms, _ = scene.assemble_transforms(objs, xf, 0)
    capacity = 4_000_000
    d_exp, d_tot = O.cluster_expand(objs, transforms, tables["surfaces"], tables["lods"], view, capacity)
    assert d_tot > 0
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(objs, transforms, tables["surfaces"], tables["lods"], clusters=tables["clusters"], cluster_dispatch_capacity=capacity, draw_capacity=16)
        ctx.set_view(view)
        for _ in range(2):
            ctx.cluster_expand()
            got, gtot = ctx.read_cluster_dispatch()
            assert gtot == d_tot and np.array_equal(got.view(np.uint32).reshape(-1, 3), d_exp)
