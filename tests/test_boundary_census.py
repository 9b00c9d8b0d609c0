"""The epsilon-boundary census (north_star: objects whose bounds lie within a stated epsilon of a frustum plane or a Hi-Z texel /
level boundary "are counted and reported").  oracle_boundary_census (oracle/cull_oracle.cpp) is what bench.py's cpu_baseline leg
reports for the bench view; here it is checked on the golden scene of tests/golden/spirv_golden.npz: thread-count independent,
monotone in the tolerances, zero at zero tolerance except for exact ties, and a constructed tangent sphere is counted."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_golden import pyramid_from

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "spirv_golden.npz")
KEYS = ("near_frustum_plane", "near_texel_boundary", "near_mip_boundary", "near_depth_equal")


@pytest.fixture(scope="module")
def g(built):
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def views(g):
    return {str(n): g["views"][i:i + 1] for i, n in enumerate(g["view_names"])}


@pytest.mark.parametrize("hiz", [O.HIZ_VK, O.HIZ_DX])
def test_census_on_the_golden_scene(g, tables, views, hiz):
    pyr = pyramid_from(g)
    kw = dict(objs=g["objs"], transforms=g["transforms"], surfaces=tables["surfaces"], lods=tables["lods"], view=views["inside"], pyramid=pyr, hiz=hiz)
    one = O.boundary_census(**kw, threads=1)
    many = O.boundary_census(**kw, threads=5)
    assert one == many                                         # per-thread partial counts add up
    assert set(one) == set(KEYS)
    n = len(g["objs"])
    assert all(0 <= one[k] <= n for k in KEYS)
    # the stated tolerances (4 ulp, 1/256 texel) single out a small minority of this 2 041-object scene
    assert one["near_frustum_plane"] < n // 50
    wide = O.boundary_census(**kw, ulp_tol=4096.0, texel_tol=0.25, threads=3)
    assert all(wide[k] >= one[k] for k in KEYS)                # monotone in the tolerances
    assert wide["near_texel_boundary"] > one["near_texel_boundary"]
    tight = O.boundary_census(**kw, ulp_tol=0.0, texel_tol=0.0, threads=2)
    assert all(tight[k] <= one[k] for k in KEYS)


def test_tangent_sphere_is_counted(tables, views):
    """A sphere exactly tangent to the far plane (c.z - r == zFar: culled by the strict '<') is within 0 ulp of the plane."""
    v = views["inside"].copy()
    f = v.view(np.float32).reshape(-1)
    view_m = f[:16].reshape(4, 4)                              # column-major: column c at view_m[c]
    zfar = float(f[42])                                       # CameraViewData: zNear @164, zFar @168 (Game/blitCamera.h:38-64)
    # identity rotation, unit scale, place the surface-0 sphere so that its view-space z - r == zFar for a translate-only view
    surf = tables["surfaces"][:1].copy()
    sf = surf.view(np.float32).reshape(-1)
    sf[0:3] = 0.0
    sf[3] = 1.0
    if not (np.allclose(view_m[:3, :3], np.eye(3)) and view_m[3, 3] == 1.0):
        pytest.skip("golden 'inside' view is not translate-only")
    tz = float(view_m[3, 2])
    objs = np.zeros(1, dtype=g_dtype_objs())
    objs["transformId"] = 0
    objs["surfaceId"] = 0
    xf = np.zeros((1, 8), dtype=np.float32)
    xf[0, :3] = (-float(view_m[3, 0]), -float(view_m[3, 1]), np.float32(zfar) + np.float32(1.0) - np.float32(tz))
    xf[0, 3] = 1.0
    xf[0, 7] = 1.0                                             # quaternion (0,0,0,1)
    c = O.boundary_census(objs, xf, surf, tables["lods"], v, None, O.HIZ_VK)
    rec, total, _ = O.cull(objs, xf, surf, tables["lods"], v, O.PASS_FRUSTUM)
    if np.float32(np.float32(xf[0, 2]) + np.float32(tz)) - np.float32(1.0) == np.float32(zfar):
        assert total == 0                                      # strict comparison: tangent == culled
    assert c["near_frustum_plane"] == 1


def g_dtype_objs():
    return np.dtype([("transformId", np.uint32), ("surfaceId", np.uint32)])
