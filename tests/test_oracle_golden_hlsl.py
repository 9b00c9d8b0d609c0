"""Pins the D3D12 rows of the CPU oracle (oracle/cull_oracle.cpp) against outputs of the REFERENCE'S OWN HLSL shaders.

tests/golden/hlsl_golden.npz was produced by oracle/spirv_interp/make_hlsl_golden.py, which executes the SPIR-V that the
reference's bundled glslang (HLSL front end, entry csMain) builds from /root/reference/src/Renderer/HlslShaders/CS/*.hlsl in the
interpreter under oracle/spirv_interp, with the host side of BlitzenDX12/dx12Draw.cpp played around it.  Both sides emit in
ascending invocation order, so the comparison is exact array equality (a real GPU gives the same multiset)."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "hlsl_golden.npz")


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


@pytest.mark.parametrize("key,dkey", [("pyramid", "depth"), ("pyramid_odd", "depth_odd")])
def test_depth_pyramid_dx(g, key, dkey):
    """HlslShaders/CS/depthPyramid.cs.hlsl:13-27 run per mip as dx12Draw.cpp:246-277 does == oracle_build_pyramid(variant DX);
    the odd-sized chain (48x30 -> 24x15 -> 12x7 -> ...) exercises the out-of-range Load = 0 rule."""
    pyr = O.build_pyramid(g[dkey], O.HIZ_DX)
    pw, ph, mips = (int(x) for x in g[key + "_whm"])
    assert (pyr.width, pyr.height, pyr.mips) == (pw, ph, mips)
    n = len(g[key])
    assert np.array_equal(pyr.data[:n].view(np.uint32), g[key].view(np.uint32))
    if key == "pyramid_odd":
        assert (g[key] == 0).any()      # an out-of-range texel reached a stored mip


@pytest.mark.parametrize("vn", ["inside", "tilted", "all", "ref_default"])
def test_draw_cull_dx(g, tables, views, vn):
    """HlslShaders/CS/drawCull.cs.hlsl:10-52 == oracle PASS_FRUSTUM with 32-byte DrawCmd records"""
    rec, total, _ = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_FRUSTUM, rec_words=8)
    exp = g[f"drawcull_{vn}"]
    assert total == len(exp) and np.array_equal(rec, exp)


@pytest.mark.parametrize("vn", ["inside", "tilted"])
def test_draw_occ_first(g, tables, views, vn):
    """HlslShaders/CS/drawOccFirst.cs.hlsl:12-59 == oracle PASS_EARLY"""
    rec, total, vis = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_EARLY, rec_words=8, vis=g["vis0"])
    assert np.array_equal(rec, g[f"occfirst_{vn}"])
    assert np.array_equal(vis, g["vis0"])


@pytest.mark.parametrize("vn", ["inside", "tilted"])
def test_draw_occ_late(g, tables, views, vn):
    """HlslShaders/CS/drawOccLate.cs.hlsl:13-68 with hlslMath.hlsl:58-78 OcclusionCheck (one point texel) == oracle PASS_LATE / HIZ_DX"""
    pyr = pyramid_from(g)
    rec, total, vis = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_LATE, rec_words=8, hiz=O.HIZ_DX,
                             pyramid=pyr, vis=g["vis0"])
    assert np.array_equal(vis, g[f"occlate_vis_{vn}"])
    assert np.array_equal(rec, g[f"occlate_{vn}"])
    fr, _, _ = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_FRUSTUM)
    assert int(vis.sum()) < len(fr)          # the Hi-Z test rejected something, otherwise the fixture pins nothing
    # and the point-texel variant really differs from the Vulkan MIN-footprint variant on this fixture
    _, _, vis_vk = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_LATE, hiz=O.HIZ_VK, pyramid=pyr, vis=g["vis0"])
    assert not np.array_equal(vis_vk, vis)


@pytest.mark.parametrize("vn", ["inside", "tilted"])
def test_draw_occ_temporal(g, tables, views, vn):
    """HlslShaders/CS/drawOccTemporal.hlsl:13-65 (frustum + Hi-Z, no visibility buffer) == oracle PASS_TEMPORAL / HIZ_DX"""
    pyr = pyramid_from(g)
    rec, total, _ = O.cull(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], views[vn], O.PASS_TEMPORAL, rec_words=8, hiz=O.HIZ_DX, pyramid=pyr)
    assert total == len(g[f"occtemporal_{vn}"]) and np.array_equal(rec, g[f"occtemporal_{vn}"])


@pytest.mark.parametrize("vn", ["inside", "all"])
def test_indirect_instancing(g, tables, views, vn):
    """drawInstCountReset.cs.hlsl:9-19 + drawInstCull.cs.hlsl:12-43 + drawInstCmd.cs.hlsl:9-39 == oracle_cull_instanced"""
    li = g["inst_lod_instances"]
    nl = len(tables["lods"])
    bucket = int(g["inst_bucket"][0])
    cap = np.full(nl, bucket, dtype=np.uint32)
    idx, counts, cmds = O.cull_instanced(g["objs"], g["transforms"], tables["surfaces"], tables["lods"], li, cap, views[vn])
    assert np.array_equal(counts, g[f"inst_counts_{vn}"])
    assert np.array_equal(cmds, g[f"inst_cmds_{vn}"])
    exp = g[f"inst_indices_{vn}"]
    for l in range(nl):
        a, c = int(li["instanceOffset"][l]), int(counts[l])
        assert np.array_equal(idx[a:a + c], exp[a:a + c]), f"bucket {l}"
        assert (exp[a + c:a + bucket] == 0xFFFFFFFF).all()       # the shader never wrote past the count
    assert counts.sum() == len(g[f"drawcull_{vn}"])


@pytest.mark.parametrize("vn", ["default", "cfg1_centre", "cfg1_tilted", "cfg1_all"])
def test_reference_scene_head_dx(g, vn):
    """The reference's own scene (first 4096 objects of RenderingStressTest) under the reference's own cameras, drawOccLate."""
    from blitzen_b200 import sceneio, scene
    head = sceneio.read_blob(os.path.join(os.path.dirname(__file__), "golden", "stress_head_4k.blob"))
    view = scene.reference_views()[vn]
    pyr = pyramid_from(g)
    rec, total, vis = O.cull(head["objs"], head["transforms"], head["surfaces"], head["lods"], view, O.PASS_LATE, rec_words=8, hiz=O.HIZ_DX,
                             pyramid=pyr, vis=g["head_vis0"])
    assert np.array_equal(vis, g[f"head_occlate_vis_{vn}"])
    assert np.array_equal(rec, g[f"head_occlate_{vn}"])
