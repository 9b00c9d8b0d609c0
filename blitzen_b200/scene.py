"""Workload construction for the cull path: stress scenes shaped like the reference's RenderingStressTest /
InstancingStressTest, the mesh tables of the reference's four bundled meshes, the reference-generated camera views and
synthetic reverse-Z depth images.  Pure data plumbing (numpy + csrc/scene_gen.cpp); no culling happens here.
"""
import ctypes as C
import json
import os

import numpy as np

from . import sceneio
from . import types as T

_PKG = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_PKG, "data")
LIB_SCENE = os.path.join(_PKG, "libblz_scene.so")

# LoadGeometryStressTest (Renderer/Resources/RenderObject/blitzenRender.cpp:126-166): (surface, scale, count) in creation order
STRESS_GROUPS = ((0, 5.0, 2_500_000), (2, 1.0, 1_500_000), (1, 0.5, 10_000), (3, 0.2, 90_000))
STRESS_TOTAL = sum(g[2] for g in STRESS_GROUPS)          # 4 100 000
STATIC_TRANSFORM_OFFSET = 1000                            # Ce_MaxDynamicObjectCount, Core/blitzenEngine.h:74


class _Group(C.Structure):
    _fields_ = [("surfaceId", C.c_uint32), ("scale", C.c_float), ("count", C.c_uint64)]


_lib = None


def _scene_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_SCENE):
            raise RuntimeError(f"{LIB_SCENE} missing: run `python -m blitzen_b200.build`")
        _lib = C.CDLL(LIB_SCENE)
        _lib.blz_scene_generate.restype = C.c_int
        _lib.blz_scene_generate.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_Group), C.c_uint32,
                                            C.c_float, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.blz_depth_generate.restype = C.c_int
        _lib.blz_depth_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_float, C.c_float, C.c_uint32, C.c_void_p]
    return _lib


def mesh_tables():
    """Surfaces / LODs / clusters / lodInstance tables of bunny, dragon, kitten, human, produced by the reference's own
    LoadMeshFromObj pipeline (oracle/make_fixtures.py)."""
    sc = sceneio.read_blob(os.path.join(_DATA, "stress_mesh_tables.blob"))
    return {k: sc[k] for k in ("surfaces", "lods", "clusters", "lodInstances")}


def reference_views():
    views, meta = sceneio.read_views(os.path.join(_DATA, "ref_views.blob"), os.path.join(_DATA, "ref_views.json"))
    return views


def scaled_groups(n_stress):
    """The reference's mesh mix (61.0 % bunny / 36.6 % kitten / 0.24 % dragon / 2.2 % human) scaled to n_stress objects."""
    out, acc = [], 0
    for k, (sid, scale, cnt) in enumerate(STRESS_GROUPS):
        c = n_stress - acc if k == len(STRESS_GROUPS) - 1 else (cnt * n_stress) // STRESS_TOTAL
        out.append((sid, scale, int(c)))
        acc += int(c)
    return tuple(out)


def cube_side(n_objects, base=3000.0):
    """cube side 3000 * cbrt(N / 4.1e6) keeps the reference's object density."""
    return float(np.float32(base * (n_objects / 4.1e6) ** (1.0 / 3.0)))


def generate(groups=STRESS_GROUPS, multiplier=3000.0, prologue=True, prng="glibc", seed=0, first=0, count=None, threads=8,
             n_dynamic=1000, dyn_surface=2):
    """Returns (objs, transforms_of_those_objects) for objects [first, first+count) of the scene.

    prng="glibc" reproduces the reference's rand() stream bit for bit (sequential, first must be 0);
    prng="counter" is a counter-based stream with the same distributions (any range, multi-threaded)."""
    lib = _scene_lib()
    total = (1 + n_dynamic if prologue else 0) + sum(g[2] for g in groups)
    if count is None:
        count = total - first
    g = (_Group * len(groups))(*[_Group(s, sc, c) for s, sc, c in groups])
    objs = np.zeros(count, dtype=T.RenderObject)
    xf = np.zeros(count, dtype=T.MeshTransform)
    rc = lib.blz_scene_generate(0 if prng == "glibc" else 1, seed, 1 if prologue else 0, n_dynamic, dyn_surface, STATIC_TRANSFORM_OFFSET,
                                g, len(groups), np.float32(multiplier), first, count, objs.ctypes.data, xf.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"blz_scene_generate failed ({rc})")
    return objs, xf


def assemble_transforms(objs, xf_per_object, transform_id_base=None):
    """Places per-object transforms at their transformId: returns (transform_array, transform_id_base) covering
    [min transformId, max transformId]."""
    tid = objs["transformId"].astype(np.int64)
    lo = int(tid.min()) if transform_id_base is None else int(transform_id_base)
    hi = int(tid.max()) + 1
    out = np.zeros(hi - lo, dtype=T.MeshTransform)
    out["orientation"][:, 3] = 1.0
    out[tid - lo] = xf_per_object
    return out, lo


def stress_scene(n_stress=STRESS_TOTAL, multiplier=None, prng="glibc", seed=0, prologue=True, threads=8):
    """Full single-GPU scene dict (objs, transforms, tables).  n_stress=4.1M, prng glibc, multiplier 3000 is the reference's
    RenderingStressTest (config 1); other sizes use the scaled mesh mix and the density-preserving cube."""
    groups = STRESS_GROUPS if n_stress == STRESS_TOTAL else scaled_groups(n_stress)
    if multiplier is None:
        multiplier = 3000.0 if n_stress == STRESS_TOTAL else cube_side(n_stress)
    objs, xf = generate(groups, multiplier, prologue, prng, seed, threads=threads)
    transforms, base = assemble_transforms(objs, xf, transform_id_base=0)
    sc = dict(mesh_tables())
    sc.update(objs=objs, transforms=transforms, multiplier=multiplier)
    return sc


def shard(scene_objs_range, rank, world):
    """Contiguous object range of `rank`: [rank*N/world, (rank+1)*N/world)."""
    n = scene_objs_range
    return (rank * n) // world, ((rank + 1) * n) // world


def synthetic_depth(width, height, z_near=0.1, n_rects=64, z_min=50.0, z_max=600.0, seed=0x00B1172E):
    lib = _scene_lib()
    out = np.zeros((height, width), dtype=np.float32)
    rc = lib.blz_depth_generate(width, height, np.float32(z_near), n_rects, np.float32(z_min), np.float32(z_max), seed, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("blz_depth_generate failed")
    return out


def make_view(position, yaw=0.0, pitch=0.0, fov_deg=70.0, width=1280, height=720, z_near=0.1, z_far=650.0):
    """A CameraViewData block for an arbitrary camera (numpy float32; same conventions as Game/blitzenCamera.cpp:14-48,
    93-126 -- view looks down +z, reverse-Z infinite projection -- but NOT bit-identical to the reference's libm path;
    use reference_views() where bit-identity with the reference's camera code matters)."""
    f32 = np.float32
    cy, sy = np.cos(f32(yaw)), np.sin(f32(yaw))
    cp, sp = np.cos(f32(pitch)), np.sin(f32(pitch))
    # rotation = yaw about (0,-1,0) times pitch about (1,0,0); columns are the camera axes in world space
    r_yaw = np.array([[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]], dtype=np.float64)
    r_pitch = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=np.float64)
    rot = r_yaw @ r_pitch
    view3 = rot.T
    t = -view3 @ np.asarray(position, dtype=np.float64)
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = view3
    m[:3, 3] = t
    v = np.zeros(1, dtype=T.CameraViewData)
    v["viewMatrix"][0] = m.T.reshape(-1).astype(f32)       # column-major
    half = f32(1.0) / np.tan(f32(np.radians(f32(fov_deg))) / f32(2.0))
    p0, p5 = f32(half / (f32(width) / f32(height))), f32(half)
    proj = np.zeros((4, 4), dtype=np.float64)               # row-major math form of InfiniteZPerspective
    proj[0, 0] = p0; proj[1, 1] = p5; proj[3, 2] = 1.0; proj[2, 3] = z_near
    v["projectionViewMatrix"][0] = (proj @ m).T.reshape(-1).astype(f32)
    v["position"][0] = np.asarray(position, dtype=f32)
    nx = np.sqrt(np.float64(p0) ** 2 + 1.0); ny = np.sqrt(np.float64(p5) ** 2 + 1.0)
    v["frustumRight"] = f32(p0 / nx); v["frustumLeft"] = f32(1.0 / nx)
    v["frustumTop"] = f32(p5 / ny); v["frustumBottom"] = f32(1.0 / ny)
    v["proj0"] = p0; v["proj5"] = p5; v["zNear"] = f32(z_near); v["zFar"] = f32(z_far)
    v["lodTarget"] = f32((f32(2.0) / p5) * (f32(1.0) / f32(height)))
    return v
