"""Pins the CPU oracle (oracle/cull_oracle.cpp) against outputs of the REFERENCE'S OWN shaders.

tests/golden/spirv_golden.npz was produced by oracle/spirv_interp/make_spirv_golden.py, which executes the SPIR-V that the
reference's bundled glslang builds from /root/reference/src/Renderer/VulkanShaders/*.comp.glsl in a small interpreter.
Both sides emit in ascending invocation order, so the comparison is exact array equality (the reference on a real GPU
would give the same multiset in atomic-arrival order)."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "spirv_golden.npz")


@pytest.fixture(scope="module")
def g(built):
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def views(g):
    return {str(n): g["views"][i:i + 1] for i, n in enumerate(g["view_names"])}


def pyramid_from(g, key="pyramid"):
    pw, ph, mips = (int(x) for x in g[key + "_whm"])
    offs, off = [], 0
    for i in range(16):
        offs.append(off)
        if i < mips:
            off += max(1, pw >> i) * max(1, ph >> i)
    return O.Pyramid(g[key], pw, ph, mips, offs)


@pytest.mark.parametrize("vn", ["inside", "tilted", "all", "ref_default"])
def test_transparent_draw_cull(g, tables, views, vn):
    """VulkanShaders/TransparentDrawCull.comp.glsl == oracle PASS_FRUSTUM"""
    rec, total, _ = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_FRUSTUM)
    exp = g[f"transparent_{vn}"]
    assert total == len(exp)
    assert np.array_equal(rec, exp)
    if vn == "all":
        assert total == len(g["objs"])


@pytest.mark.parametrize("vn", ["inside", "tilted"])
def test_initial_draw_cull(g, tables, views, vn):
    """VulkanShaders/InitialDrawCull.comp.glsl == oracle PASS_EARLY"""
    rec, total, vis = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_EARLY, vis=g["vis0"])
    assert np.array_equal(rec, g[f"initial_{vn}"])
    assert np.array_equal(vis, g["vis0"])          # the early pass never writes visibility


@pytest.mark.parametrize("vn", ["inside", "tilted"])
def test_late_draw_cull(g, tables, views, vn):
    """VulkanShaders/LateDrawCull.comp.glsl (+ the reference's MIN sampler) == oracle PASS_LATE / HIZ_VK"""
    pyr = pyramid_from(g)
    rec, total, vis = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_LATE, hiz=O.HIZ_VK,
                             pyramid=pyr, vis=g["vis0"])
    assert np.array_equal(vis, g[f"late_vis_{vn}"])
    assert np.array_equal(rec, g[f"late_{vn}"])
    # the Hi-Z test must actually reject something in this fixture, otherwise it pins nothing
    fr, _, _ = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_FRUSTUM)
    assert int(vis.sum()) < len(fr)


def test_onpc_draw_cull_quirk(g, tables, views):
    """VulkanShaders/OnpcDrawCull.comp.glsl:38-40 uses the relative LOD index as absolute: oracle flag FLAG_ONPC_LOD_QUIRK"""
    rec, total, _ = O.cull(g["onpc_objs"], g["transforms"], tables["surfaces"], tables["lods"], views["inside"], O.PASS_FRUSTUM,
                           flags=O.FLAG_ONPC_LOD_QUIRK)
    assert np.array_equal(rec, g["onpc_inside"])
    plain, _, _ = O.cull(g["onpc_objs"], g["transforms"], tables["surfaces"], tables["lods"], views["inside"], O.PASS_FRUSTUM)
    assert not np.array_equal(plain, rec)          # the quirk is visible in this fixture


def test_cluster_path(g, tables, views):
    """PreClusterDrawCull.comp.glsl + InitialClusterCull.comp.glsl == oracle cluster_expand + cluster_cull(passthrough)"""
    n = int(g["cluster_objs_count"][0])
    objs = g["objs"][:n]
    exp_rec = g["cluster_dispatch_inside"]
    rec, total = O.cluster_expand(objs, g["transforms"], tables["surfaces"], tables["lods"], views["inside"], capacity=len(exp_rec) + 8)
    assert total == len(exp_rec) and total > 0
    assert np.array_equal(rec, exp_rec)
    draws, dtotal = O.cluster_cull(objs, g["transforms"], tables["surfaces"], tables["lods"], tables["clusters"], views["inside"], rec, mode=0)
    assert np.array_equal(draws, g["cluster_draws_inside"])


@pytest.mark.parametrize("key,dkey", [("pyramid", "depth"), ("pyramid_odd", "depth_odd")])
def test_depth_pyramid(g, key, dkey):
    """DepthPyramidGeneration.comp.glsl run once per mip as vulkanDraw.cpp:579-614 does == oracle_build_pyramid(variant VK)"""
    pyr = O.build_pyramid(g[dkey], O.HIZ_VK)
    pw, ph, mips = (int(x) for x in g[key + "_whm"])
    assert (pyr.width, pyr.height, pyr.mips) == (pw, ph, mips)
    n = len(g[key])
    assert np.array_equal(pyr.data[:n].view(np.uint32), g[key].view(np.uint32))


@pytest.mark.parametrize("vn", ["default", "cfg1_centre", "cfg1_tilted", "cfg1_all"])
def test_reference_scene_head(g, views, vn):
    """The reference's own scene (first 4096 objects of RenderingStressTest) under the reference's own cameras."""
    from blitzen_b200 import sceneio, scene
    head = sceneio.read_blob(os.path.join(os.path.dirname(__file__), "golden", "stress_head_4k.blob"))
    view = scene.reference_views()[vn]
    rec, total, _ = O.cull(head["objs"], head["transforms"], head["surfaces"], head["lods"], view, O.PASS_FRUSTUM)
    assert np.array_equal(rec, g[f"head_transparent_{vn}"])
    pyr = pyramid_from(g)
    rec, total, vis = O.cull(head["objs"], head["transforms"], head["surfaces"], head["lods"], view, O.PASS_LATE, hiz=O.HIZ_VK, pyramid=pyr,
                             vis=g["head_vis0"])
    assert np.array_equal(vis, g[f"head_late_vis_{vn}"])
    assert np.array_equal(rec, g[f"head_late_{vn}"])
