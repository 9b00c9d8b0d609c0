"""Hand-derived known-answer tests of the CPU oracle: boundaries of every comparison, LOD loop shape,
mip-level clamps, pyramid reductions, frame-0 behaviour, capacity clamp.  CPU only."""
import numpy as np
import pytest

import oracle_lib as O
from blitzen_b200 import scene, types as T

f32 = np.float32
IDENT = np.array([0, 0, 0, 1, 0, 0, 0, 1], dtype=np.float32)       # pos 0, scale 1, quat (0,0,0,1)


def simple_view(z_near=1.0, z_far=100.0):
    """Camera at the origin looking down +z with 45-degree half-angles on both axes: frustumRight = frustumLeft = 1/sqrt(2)."""
    v = np.zeros(1, dtype=T.CameraViewData)
    v["viewMatrix"][0] = np.eye(4, dtype=np.float32).reshape(-1)
    s = f32(0.5) ** f32(0.5)
    v["frustumRight"] = s; v["frustumLeft"] = s; v["frustumTop"] = s; v["frustumBottom"] = s
    v["proj0"] = 1.0; v["proj5"] = 1.0; v["zNear"] = z_near; v["zFar"] = z_far
    v["pyramidWidth"] = 8; v["pyramidHeight"] = 8; v["lodTarget"] = 0.01
    return v


def probe(center, radius, view, **kw):
    return O.probe(np.array(center, dtype=np.float32), radius, IDENT, view, **kw)


def test_near_and_far_planes_are_strict(built):
    v = simple_view(1.0, 100.0)
    assert probe((0, 0, 0.5), 0.5, v)[0] == 0.0         # c.z + r == zNear  -> culled (strict >)
    assert probe((0, 0, f32(0.5) + f32(2.0 ** -23)), 0.5, v)[0] == 1.0      # 0.5 + 2 ulp: the sum is the next float above 1
    assert probe((0, 0, 101.0), 1.0, v)[0] == 0.0       # c.z - r == zFar   -> culled (strict <)
    assert probe((0, 0, np.nextafter(f32(101), f32(0))), 1.0, v)[0] == 1.0


def test_side_planes_are_strict_and_symmetric(built):
    # exact arithmetic: plane coefficients 0.5 (powers of two) => c.z*0.5 - |c.x|*0.5 > -r
    v = simple_view()
    for k in ("frustumRight", "frustumLeft", "frustumTop", "frustumBottom"):
        v[k] = 0.5
    assert probe((12, 0, 10), 1.0, v)[0] == 0.0         # 5 - 6 = -1 == -r -> culled
    assert probe((-12, 0, 10), 1.0, v)[0] == 0.0        # abs(): same on the other side
    assert probe((11.5, 0, 10), 1.0, v)[0] == 1.0
    assert probe((0, 12, 10), 1.0, v)[0] == 0.0
    assert probe((0, -11.5, 10), 1.0, v)[0] == 1.0


def test_non_unit_quaternion_is_not_normalised(built):
    """RotateQuat with q = (0,0,0,2): v + 2*cross(q.xyz, ...) = v (xyz = 0) -- and with q = (0,0,1,1) (non-unit, 90 deg * 2 scale mix)."""
    v = simple_view()
    t = IDENT.copy(); t[4:8] = (0, 0, 1, 1)
    out = O.probe(np.array((1, 0, 10), dtype=np.float32), 0.1, t, v)
    # c1 = cross((0,0,1),(1,0,10)) = (0*10-0*1, 1*1-10*0, 0) = (0,1,0); t = c1 + 1*v = (1,1,10); c2 = cross((0,0,1),(1,1,10)) = (-1,1,0); r = v + 2*c2 = (-1,2,10)
    assert tuple(out[1:4]) == (-1.0, 2.0, 10.0)


def test_project_sphere_near_branch_keeps_object(built):
    v = simple_view(1.0, 100.0)
    out = probe((0, 0, 1.4), 0.5, v)                    # c.z < r + zNear -> projectSphere false -> no Hi-Z test, object kept
    assert out[0] == 1.0 and out[5] == 0.0 and out[10] == 1.0


def test_project_sphere_centred(built):
    v = simple_view(0.1, 100.0)
    out = probe((0, 0, 5), 3, v)                        # vx = sqrt(25-9) = 4; minx = (0-15)/(20+0) = -0.75, maxx = 0.75
    assert out[5] == 1.0
    assert tuple(out[6:10]) == (0.125, 0.125, 0.875, 0.875)


def lod_table(errors):
    l = np.zeros(len(errors), dtype=T.LodData)
    l["error"] = np.array(errors, dtype=np.float32)
    l["indexCount"] = np.arange(len(errors)) + 100
    return l


def test_lod_selection_takes_last_passing_index(built):
    v = simple_view()
    v["lodTarget"] = 1.0
    # distance = |c| - r = 10 - 1 = 9, threshold = 9*1/1 = 9
    assert probe((0, 0, 10), 1, v, lods=lod_table([0, 1, 2, 3]), lod_offset=0, lod_count=4)[11] == 3
    assert probe((0, 0, 10), 1, v, lods=lod_table([0, 1, 20, 3]), lod_offset=0, lod_count=4)[11] == 3   # non-monotone: LAST i that passes
    assert probe((0, 0, 10), 1, v, lods=lod_table([0, 9, 9, 9]), lod_offset=0, lod_count=4)[11] == 0    # strict <
    assert probe((0, 0, 10), 1, v, lods=lod_table([5]), lod_offset=0, lod_count=1)[11] == 0             # lodCount == 1
    assert probe((0, 0, 10), 1, v, lods=lod_table([0, 0, 0, 1, 2]), lod_offset=2, lod_count=3)[11] == 2  # relative to lodOffset
    assert probe((0, 0, 0.5), 1, v, lods=lod_table([0, 0, 0]), lod_offset=0, lod_count=3)[11] == 0      # inside the sphere: distance clamps to 0, 0 < 0 false


def test_ilog2_floor_is_exact(built):
    assert O.ilog2_floor(1.0) == 0 and O.ilog2_floor(2.0) == 1 and O.ilog2_floor(0.5) == -1
    assert O.ilog2_floor(float(np.nextafter(f32(4), f32(0)))) == 1      # glibc log2f rounds this to 2.0; the exact floor is 1
    assert O.ilog2_floor(float(np.nextafter(f32(1), f32(0)))) == -1
    assert O.ilog2_floor(1e-40) == -133                                  # denormal


def test_pyramid_constant_and_spike(built):
    d = np.full((64, 128), 0.25, dtype=np.float32)
    for variant in (O.HIZ_VK, O.HIZ_DX):
        p = O.build_pyramid(d, variant)
        n = sum(max(1, p.width >> i) * max(1, p.height >> i) for i in range(p.mips))
        assert np.all(p.data[:n] == 0.25)
    # a single low texel must reach the top (MIN), a single high texel must not
    lo = d.copy(); lo[10, 20] = 0.0
    hi = d.copy(); hi[10, 20] = 1.0
    for variant in (O.HIZ_VK, O.HIZ_DX):
        assert O.build_pyramid(lo, variant).level(O.build_pyramid(lo, variant).mips - 1).min() == 0.0
        top = O.build_pyramid(hi, variant)
        assert top.level(top.mips - 1).max() == 0.25


def test_pyramid_layouts(built):
    # BlitML::PreviousPow2 (blitML.h:54-62) returns the largest power of two r with r*2 < v... i.e. 1080 -> 1024, 1920 -> 1024 and,
    # being strict, 1024 -> 512.  (SURVEY.md / BASELINE.md quote 1024x512 for 1080p; the reference's code gives 1024x1024.)
    assert O.pyramid_layout(1920, 1080, O.HIZ_VK)[1:4] == (1024, 1024, 10)
    assert O.pyramid_layout(3840, 2160, O.HIZ_VK)[1:4] == (2048, 2048, 11)
    assert O.pyramid_layout(1024, 512, O.HIZ_VK)[1:4] == (512, 256, 9)
    assert O.pyramid_layout(1920, 1080, O.HIZ_DX)[1:4] == (960, 540, 9)             # dx12RNDResources.cpp:103-106 + GetDepthPyramidMipLevels
    assert O.pyramid_layout(1920, 1080, O.HIZ_VK)[0] == 1398100                 # 5 592 400 bytes


def test_hiz_level_clamps(built):
    """Tiny projected size -> level < 0 clamps to mip 0; huge -> clamps to the last mip."""
    v = simple_view(0.1, 1e6)
    d = np.zeros((16, 16), dtype=np.float32); d[:, :] = 0.5
    pyr = O.build_pyramid(d, O.HIZ_VK)                   # 8x8, 3 mips
    small = probe((0, 0, 1000), 0.001, v, pyramid=pyr, hiz=O.HIZ_VK)     # depthSphere = 0.1/999.999 < 0.5 -> occluded
    assert small[5] == 1.0 and small[10] == 0.0
    big = probe((0, 0, 0.3), 0.15, v, pyramid=pyr, hiz=O.HIZ_VK)         # depthSphere = 0.1/0.15 > 0.5 -> visible
    assert big[5] == 1.0 and big[10] == 1.0


def test_frame0_emits_every_frustum_survivor_and_capacity_clamps(built, tables):
    sc = scene.stress_scene(n_stress=5000, multiplier=300.0, prng="counter", seed=3)
    view = scene.make_view((150, 150, 150), z_far=500.0)
    fr, ftot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM)
    for hiz in (O.HIZ_VK, O.HIZ_DX):
        vis0 = np.zeros(len(sc["objs"]), dtype=np.uint32)
        e, etot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_EARLY, vis=vis0)
        assert etot == 0                                                           # nothing was visible last frame
        l, ltot, vis = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_LATE, hiz=hiz,
                              pyramid=O.cleared_pyramid(1280, 720, hiz), vis=vis0)
        assert ltot == ftot and np.array_equal(l, fr) and int(vis.sum()) == ftot   # cleared depth (0 = far) occludes nothing
    cap = ftot // 3
    c, ctot, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, capacity=cap)
    assert ctot == ftot and len(c) == cap and np.array_equal(c, fr[:cap])
    # threads do not change the result
    t8, _, _ = O.cull(sc["objs"], sc["transforms"], sc["surfaces"], sc["lods"], view, O.PASS_FRUSTUM, threads=8)
    assert np.array_equal(t8, fr)


def test_reference_view_constants(views):
    """Golden constants of the reference's default camera captured in SURVEY.md 8c (bit patterns)."""
    v = views["default"]
    bits = lambda k: int(np.asarray(v[k]).view(np.uint32)[0])
    assert bits("frustumRight") == 0x3F2053C6 and bits("frustumLeft") == 0x3F4793D6
    assert bits("frustumTop") == 0x3F51B3F3 and bits("frustumBottom") == 0x3F12D5E8
    assert bits("lodTarget") == 0x3AFEF013
    assert float(v["zNear"][0]) == float(f32(0.1)) and float(v["zFar"][0]) == 650.0
