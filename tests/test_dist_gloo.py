"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo.  Each rank culls its contiguous shard with the oracle
(standing in for the per-GPU kernels), the counts are all-gathered, exclusive-scanned and the per-rank lists concatenated in
shard order on the presenting rank -- which must be byte-identical to the single-rank list."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_stress, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from blitzen_b200 import dist as bdist, scene
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    groups = scene.scaled_groups(n_stress)
    total = 1001 + n_stress
    a, b = bdist.shard_range(total, rank, world)
    objs, xf = scene.generate(groups, 400.0, True, "counter", seed=21, first=a, count=b - a)      # only this rank's objects
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    view = scene.make_view((200, 200, 200), z_far=900.0)
    rec, tot, _ = O.cull(objs, transforms, tables["surfaces"], tables["lods"], view, O.PASS_FRUSTUM, object_id_base=a, transform_id_base=tbase)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(rec)], dtype=torch.int64))
    counts = [int(c.item()) for c in counts]
    offs = bdist.exclusive_scan(counts)
    # variable-length gather to the presenting rank (rank 0) at the scanned offsets
    if rank == 0:
        out = np.zeros((sum(counts), 6), dtype=np.uint32)
        out[offs[0]:offs[0] + counts[0]] = rec
        for r in range(1, world):
            buf = torch.zeros(counts[r] * 6, dtype=torch.int32)
            if counts[r]:
                dist.recv(buf, src=r)
            out[offs[r]:offs[r] + counts[r]] = buf.numpy().view(np.uint32).reshape(-1, 6)
        np.save(os.path.join(out_dir, "gathered.npy"), out)
        np.save(os.path.join(out_dir, "counts.npy"), np.array(counts))
    elif len(rec):
        dist.send(torch.from_numpy(rec.reshape(-1).view(np.int32).copy()), dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cull_concatenates_to_the_single_rank_list(built, tmp_path):
    import torch.multiprocessing as mp
    import oracle_lib as O
    from blitzen_b200 import scene
    n_stress, world = 30000, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_stress, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    counts = np.load(tmp_path / "counts.npy")
    objs, xf = scene.generate(scene.scaled_groups(n_stress), 400.0, True, "counter", seed=21)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    view = scene.make_view((200, 200, 200), z_far=900.0)
    exp, tot, _ = O.cull(objs, transforms, tables["surfaces"], tables["lods"], view, O.PASS_FRUSTUM, transform_id_base=tbase)
    assert counts.sum() == tot and tot > 1000 and counts.min() > 0
    assert np.array_equal(got, exp)                     # shard order == ascending global objectId
    assert np.all(np.diff(got[:, 0].astype(np.int64)) > 0)


def test_shard_ranges_cover_everything():
    from blitzen_b200 import dist as bdist
    for n in (0, 1, 7, 1000, 16_777_216, 268_435_456):
        for world in (1, 2, 3, 4, 8):
            r = [bdist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    assert list(bdist.exclusive_scan([3, 0, 5, 2])) == [0, 3, 3, 8]


def _worker_inst(rank, world, port, n_stress, out_dir):
    """Instance-list gather, host logic: per-rank per-LOD counts are all-gathered, every rank's bucket l lands behind the lower ranks' bucket l."""
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from blitzen_b200 import dist as bdist, scene
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    groups = scene.scaled_groups(n_stress)
    total = 1001 + n_stress
    a, b = bdist.shard_range(total, rank, world)
    objs, xf = scene.generate(groups, 400.0, True, "counter", seed=22, first=a, count=b - a)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    nl = len(tables["lods"])
    view = scene.make_view((200, 200, 200), z_far=900.0)
    li = tables["lodInstances"].copy()
    cap = np.full(nl, b - a, dtype=np.uint32)
    li["instanceOffset"] = (np.arange(nl, dtype=np.uint64) * (b - a)).astype(np.uint32)
    idx, cnt, _ = O.cull_instanced(objs, transforms, tables["surfaces"], tables["lods"], li, cap, view, object_id_base=a, transform_id_base=tbase)
    mine = torch.from_numpy(cnt.astype(np.int64))
    allc = [torch.zeros(nl, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, mine)
    allc = torch.stack(allc).numpy()                                   # [world][lod]
    base = allc[:rank].sum(axis=0)                                     # exclusive scan over the ranks, per LOD
    goff = np.concatenate([[0], np.cumsum(allc.sum(axis=0))[:-1]])     # global bucket layout: exact fit
    if rank == 0:
        out = np.zeros(int(allc.sum()), dtype=np.uint32)
        for l in range(nl):
            out[goff[l]:goff[l] + cnt[l]] = idx[int(li["instanceOffset"][l]):int(li["instanceOffset"][l]) + int(cnt[l])]
        for r in range(1, world):
            for l in range(nl):
                c = int(allc[r][l])
                if c:
                    buf = torch.zeros(c, dtype=torch.int32)
                    dist.recv(buf, src=r)
                    s = int(goff[l] + allc[:r, l].sum())
                    out[s:s + c] = buf.numpy().view(np.uint32)
        np.save(os.path.join(out_dir, "inst.npy"), out)
        np.save(os.path.join(out_dir, "inst_counts.npy"), allc)
    else:
        for l in range(nl):
            c = int(cnt[l])
            if c:
                o = int(li["instanceOffset"][l])
                dist.send(torch.from_numpy(idx[o:o + c].view(np.int32).copy()), dst=0)
    assert np.array_equal(base, allc[:rank].sum(axis=0))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_instancing_buckets_concatenate_per_lod(built, tmp_path):
    import torch.multiprocessing as mp
    import oracle_lib as O
    from blitzen_b200 import scene
    n_stress, world = 30000, 2
    port = _free_port()
    mp.spawn(_worker_inst, args=(world, port, n_stress, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "inst.npy")
    allc = np.load(tmp_path / "inst_counts.npy")
    objs, xf = scene.generate(scene.scaled_groups(n_stress), 400.0, True, "counter", seed=22)
    transforms, tbase = scene.assemble_transforms(objs, xf)
    tables = scene.mesh_tables()
    nl = len(tables["lods"])
    view = scene.make_view((200, 200, 200), z_far=900.0)
    tot = allc.sum(axis=0).astype(np.uint32)
    li = tables["lodInstances"].copy()
    li["instanceOffset"] = np.concatenate([[0], np.cumsum(tot.astype(np.uint64))[:-1]]).astype(np.uint32)
    idx_e, cnt_e, _ = O.cull_instanced(objs, transforms, tables["surfaces"], tables["lods"], li, np.maximum(tot, 1), view, transform_id_base=tbase)
    assert np.array_equal(cnt_e, tot) and int(tot.sum()) > 1000
    assert np.array_equal(got, idx_e[:len(got)])        # rank-order concatenation per LOD == the whole-scene buckets (ascending ids inside each)
