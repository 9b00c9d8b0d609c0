"""The reference's own frontend driving the CUDA backend (oracle/_ref/integration_demo, built from /root/reference by
`make -C oracle ref`) against the oracle on the bit-identical scene."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "integration_demo")


def fnv1a(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/integration_demo was not built (needs the reference tree at build time)")
def test_reference_frontend_drives_the_cuda_backend(built, views):
    from blitzen_b200 import scene
    out = subprocess.run([DEMO, os.path.join(ROOT, "blitzen_b200", "data", "stress_mesh_tables.blob"), "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    sc = scene.stress_scene()                      # bit-identical to the reference's scene (tests/test_scene_golden.py)
    assert got["objects"] == len(sc["objs"]) == 4101001
    view = views["default"]                        # SetupCamera(camera) with the engine defaults
    n = len(sc["objs"])
    vis0 = np.zeros(n, dtype=np.uint32)
    e0, e0tot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_EARLY, vis=vis0, threads=8)
    l0, l0tot, vis1 = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_LATE, hiz=O.HIZ_VK,
                             pyramid=O.cleared_pyramid(1280, 720, O.HIZ_VK), vis=vis0, threads=8)
    e1, e1tot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_EARLY, vis=vis1, threads=8)
    assert got["frame0_early"] == e0tot == 0
    assert got["frame0_late"] == l0tot and l0tot > 1000
    assert got["frame0_late_hash"] == fnv1a(l0.tobytes())
    assert got["frame0_vis_hash"] == fnv1a(np.nonzero(vis1)[0].astype(np.uint32).tobytes())
    assert got["frame1_early"] == e1tot == l0tot
    assert got["frame1_early_hash"] == fnv1a(e1.tobytes())
