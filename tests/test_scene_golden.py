"""The workload generator against fixtures produced by the reference's own compiled frontend (oracle/make_fixtures.py):
the glibc-rand() path reproduces the reference's RenderingStressTest scene bit for bit.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from blitzen_b200 import scene, sceneio

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_head_matches_reference(built):
    head = sceneio.read_blob(os.path.join(GOLDEN, "stress_head_4k.blob"))
    objs, xf = scene.generate(prng="glibc", count=4096)
    assert np.array_equal(objs.view(np.uint32), head["objs"].view(np.uint32))
    tid = objs["transformId"]
    assert np.array_equal(xf.view(np.uint32), head["transforms"][tid].view(np.uint32))
    # golden from SURVEY.md 8c: the first stress object (index 1001) of the reference
    o = 1001
    assert int(objs[o]["transformId"]) == 2001 - 0 or int(objs[o]["surfaceId"]) == 0


def test_mesh_tables_match_survey(tables):
    s = tables["surfaces"]
    assert len(s) == 4 and len(tables["lods"]) == 28 and len(tables["clusters"]) == 6833
    assert list(s["lodCount"]) == [8, 8, 8, 4]
    assert abs(float(s["radius"][1]) - 7.70437) < 1e-4 and abs(float(s["radius"][3]) - 11.7257) < 1e-4
    assert int(tables["lods"]["indexCount"][0]) == 14904 and int(tables["lods"]["clusterCount"][0]) == 52


@pytest.mark.slow
def test_full_scene_checksums(built):
    with open(os.path.join(GOLDEN, "stress_checksums.json")) as f:
        sums = json.load(f)
    sc = scene.stress_scene()          # 4 101 001 objects, glibc stream, multiplier 3000: the reference's RenderingStressTest
    assert len(sc["objs"]) == sums["nObjects"] and len(sc["transforms"]) == sums["nTransforms"]
    assert hashlib.sha256(sc["objs"].tobytes()).hexdigest() == sums["objs_sha256"]
    assert hashlib.sha256(sc["transforms"].tobytes()).hexdigest() == sums["transforms_sha256"]


def test_counter_prng_ranges_are_consistent(built):
    """Any sub-range of the counter-based scene equals the same range of the whole (what sharding relies on)."""
    groups = scene.scaled_groups(20000)
    whole, wx = scene.generate(groups, 500.0, True, "counter", seed=9)
    a, ax = scene.generate(groups, 500.0, True, "counter", seed=9, first=3000, count=5000)
    assert np.array_equal(a.view(np.uint32), whole[3000:8000].view(np.uint32))
    assert np.array_equal(ax.view(np.uint32), wx[3000:8000].view(np.uint32))
