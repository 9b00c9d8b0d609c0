"""Draw-list consumer (csrc/consume.cu): the device-side reduction of a frame's output -- what the indirect draw and the vertex stage would
read -- equals the same reduction of the oracle's lists, for object draws, cluster draws and the instancing buckets."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import view_at

pytestmark = pytest.mark.gpu
M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


@pytest.fixture(scope="module")
def capi(built):
    from blitzen_b200 import capi
    return capi


def reduce_list(recs):
    """numpy restatement of the consumer's summary for records {objectId, indexCount, instanceCount, firstIndex, ...}."""
    ids = recs[:, 0].astype(np.uint64)
    x = np.bitwise_xor.reduce(ids * np.uint64(GOLD)) if len(ids) else np.uint64(0)       # uint64 multiplication wraps, as on the device
    return dict(records=len(recs), index_sum=int(recs[:, 1].astype(np.uint64).sum()), instance_sum=int(recs[:, 2].astype(np.uint64).sum()),
                id_sum=int(ids.sum()), id_xor=int(x), unsorted=int(np.count_nonzero(np.diff(recs[:, 0].astype(np.int64)) < 0)))


def check(summary, exp):
    for k, v in exp.items():
        assert int(getattr(summary, k)) == v, k
    assert summary.bad_object == 0 and summary.bad_lod == 0


@pytest.mark.parametrize("fmt", [0, 1])
def test_consume_object_draws(capi, small_scene, fmt):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=2000.0)
    exp, tot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, rec_words=6 if fmt == 0 else 8)
    lod_of = {(int(l["indexCount"]), int(l["firstIndex"])): i for i, l in enumerate(sc["lods"])}
    hist = np.zeros(256, dtype=np.int64)
    for key, c in zip(*np.unique(exp[:, [1, 3]], axis=0, return_counts=True)):
        hist[lod_of[(int(key[0]), int(key[1]))]] += c
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"])
        ctx.set_view(view)
        ctx.frustum_lod(capi.LIST_OPAQUE, fmt)
        s = ctx.consume_draws()
    check(s, reduce_list(exp))
    assert s.records == tot > 1000 and s.unsorted == 0
    assert np.array_equal(np.asarray(s.lod_hist[:], dtype=np.int64), hist)


def test_consume_cluster_draws_and_instances(capi, small_scene):
    sc = small_scene
    view = view_at(position=(380, 380, 380), z_far=400.0)
    d_exp, d_tot = O.cluster_expand(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, 3_000_000)
    c_exp, c_tot = O.cluster_cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], sc["clusters"], view, d_exp, 1)
    li = sc["lodInstances"].copy()
    nl = len(sc["lods"])
    cap = np.full(nl, 50_000, dtype=np.uint32)
    li["instanceOffset"] = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint32)
    idx_e, cnt_e, cmds_e = O.cull_instanced(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], li, cap, view)
    with capi.CullContext(0) as ctx:
        ctx.upload_scene(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], clusters=sc["clusters"], lod_instances=li, bucket_capacity=cap,
                         cluster_dispatch_capacity=3_000_000, draw_capacity=3_000_000)
        ctx.set_view(view)
        ctx.cluster_expand(); ctx.cluster_cull(capi.CLUSTER_SPHERE, capi.REC_VK24)
        s = ctx.consume_draws(capi.LIST_OPAQUE, 1)
        check(s, reduce_list(c_exp))
        assert s.records == c_tot > 10_000 and s.unsorted == 0
        ctx.instanced()
        t = ctx.consume_instances()
    ids = np.concatenate([idx_e[int(li["instanceOffset"][l]):int(li["instanceOffset"][l]) + int(cnt_e[l])] for l in range(nl)]).astype(np.uint64)
    assert t.records == len(ids) == int(cnt_e.sum()) and t.id_sum == int(ids.sum()) and t.id_xor == int(np.bitwise_xor.reduce(ids * np.uint64(GOLD)))
    assert t.unsorted == 0 and t.bad_object == 0 and t.bad_lod == 0
    assert np.array_equal(np.asarray(t.lod_hist[:nl], dtype=np.int64), cnt_e.astype(np.int64))
    assert t.index_sum == int(sum(int(cnt_e[l]) * int(sc["lods"][l]["indexCount"]) for l in range(nl)))
