"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE (the checker).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

PASS_FRUSTUM, PASS_EARLY, PASS_LATE, PASS_TEMPORAL = 0, 1, 2, 3
HIZ_VK, HIZ_DX = 0, 1
FLAG_ONPC_LOD_QUIRK = 1


class _Scene(C.Structure):
    _fields_ = [("objs", C.c_void_p), ("nObj", C.c_uint32), ("transforms", C.c_void_p), ("surfaces", C.c_void_p),
                ("lods", C.c_void_p), ("clusters", C.c_void_p), ("objectIdBase", C.c_uint32), ("pad", C.c_uint32)]


class _Pyr(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("mips", C.c_uint32), ("offset", C.c_uint32 * 16)]


_lib = None


def build():
    src = os.path.join(ORACLE_DIR, "cull_oracle.cpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_pyramid_layout.restype = C.c_uint64
    return _lib


def hardware_threads():
    return int(lib().oracle_hardware_threads())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Pyramid:
    def __init__(self, data, width, height, mips, offsets):
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.width, self.height, self.mips = int(width), int(height), int(mips)
        self.offsets = [int(x) for x in offsets]

    def c(self):
        p = _Pyr()
        p.data = self.data.ctypes.data
        p.width, p.height, p.mips = self.width, self.height, self.mips
        for i in range(16):
            p.offset[i] = self.offsets[i]
        return p

    def level(self, k):
        w, h = max(1, self.width >> k), max(1, self.height >> k)
        return self.data[self.offsets[k]:self.offsets[k] + w * h].reshape(h, w)


def pyramid_layout(depth_w, depth_h, variant):
    out = (C.c_uint32 * 19)()
    total = lib().oracle_pyramid_layout(C.c_uint32(depth_w), C.c_uint32(depth_h), C.c_int(variant), out)
    return int(total), int(out[0]), int(out[1]), int(out[2]), [int(out[3 + i]) for i in range(16)]


def build_pyramid(depth, variant, threads=1):
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    h, w = depth.shape
    total, pw, ph, mips, offs = pyramid_layout(w, h, variant)
    data = np.zeros(max(total, 1), dtype=np.float32)
    lib().oracle_build_pyramid(_p(depth), C.c_uint32(w), C.c_uint32(h), C.c_int(variant), _p(data), C.c_int(threads))
    return Pyramid(data, pw, ph, mips, offs)


def cleared_pyramid(depth_w, depth_h, variant):
    total, pw, ph, mips, offs = pyramid_layout(depth_w, depth_h, variant)
    return Pyramid(np.zeros(max(total, 1), dtype=np.float32), pw, ph, mips, offs)


def _scene(objs, transforms, surfaces, lods, clusters=None, object_id_base=0, transform_id_base=0):
    s = _Scene()
    s.objs = objs.ctypes.data if len(objs) else None
    s.nObj = len(objs)
    # the oracle indexes transforms by the global transformId: rebase the pointer
    s.transforms = transforms.ctypes.data - 32 * int(transform_id_base)
    s.surfaces = surfaces.ctypes.data
    s.lods = lods.ctypes.data
    s.clusters = None if clusters is None else clusters.ctypes.data
    s.objectIdBase = int(object_id_base)
    return s


def view_with_pyramid(view, pyr):
    """The backends publish the pyramid extent in the view block (vulkanRendererSetup.cpp:903-904)."""
    v = np.array(view, copy=True)
    if pyr is not None:
        v["pyramidWidth"] = np.float32(pyr.width)
        v["pyramidHeight"] = np.float32(pyr.height)
    return v


def cull(objs, transforms, surfaces, lods, view, pass_id, rec_words=6, hiz=HIZ_VK, pyramid=None, vis=None, flags=0,
         capacity=None, threads=1, object_id_base=0, transform_id_base=0):
    """Returns (records[u32 n x rec_words], total, vis_out)."""
    objs = np.ascontiguousarray(objs); transforms = np.ascontiguousarray(transforms)
    s = _scene(objs, transforms, surfaces, lods, None, object_id_base, transform_id_base)
    v = np.ascontiguousarray(view_with_pyramid(view, pyramid))
    cap = len(objs) if capacity is None else int(capacity)
    out = np.zeros((max(cap, 1), rec_words), dtype=np.uint32)
    vis_arr = None if vis is None else np.array(vis, dtype=np.uint32, copy=True)
    pyr_c = pyramid.c() if pyramid is not None else None
    written, total = C.c_uint32(), C.c_uint32()
    lib().oracle_cull(C.byref(s), _p(v), C.byref(pyr_c) if pyr_c is not None else None, C.c_int(pass_id), C.c_int(hiz), C.c_uint32(flags),
                      _p(vis_arr), _p(out), C.c_uint32(rec_words), C.c_uint64(cap), C.byref(written), C.byref(total), C.c_int(threads))
    return out[:written.value], total.value, vis_arr


def cull_instanced(objs, transforms, surfaces, lods, lod_instances, bucket_capacity, view, threads=1, object_id_base=0, transform_id_base=0):
    s = _scene(objs, transforms, surfaces, lods, None, object_id_base, transform_id_base)
    v = np.ascontiguousarray(view)
    n_l = len(lods)
    cap = np.ascontiguousarray(bucket_capacity, dtype=np.uint32)
    li = np.ascontiguousarray(lod_instances)
    size = int((li["instanceOffset"].astype(np.int64) + cap).max())
    idx = np.zeros(size, dtype=np.uint32)
    counts = np.zeros(n_l, dtype=np.uint32)
    cmds = np.zeros((n_l, 8), dtype=np.uint32)
    ncmd = C.c_uint32()
    lib().oracle_cull_instanced(C.byref(s), _p(v), _p(li), C.c_uint32(n_l), _p(cap), _p(idx), _p(counts), _p(cmds), C.byref(ncmd), C.c_int(threads))
    return idx, counts, cmds[:ncmd.value]


def cluster_expand(objs, transforms, surfaces, lods, view, capacity, threads=1, object_id_base=0, transform_id_base=0):
    s = _scene(objs, transforms, surfaces, lods, None, object_id_base, transform_id_base)
    v = np.ascontiguousarray(view)
    out = np.zeros((max(int(capacity), 1), 3), dtype=np.uint32)
    written, total = C.c_uint32(), C.c_uint32()
    lib().oracle_cluster_expand(C.byref(s), _p(v), _p(out), C.c_uint64(int(capacity)), C.byref(written), C.byref(total), C.c_int(threads))
    return out[:written.value], total.value


def cluster_cull(objs, transforms, surfaces, lods, clusters, view, dispatch, mode, rec_words=6, hiz=-1, pyramid=None, capacity=None,
                 threads=1, object_id_base=0, transform_id_base=0):
    s = _scene(objs, transforms, surfaces, lods, clusters, object_id_base, transform_id_base)
    v = np.ascontiguousarray(view_with_pyramid(view, pyramid))
    dispatch = np.ascontiguousarray(dispatch, dtype=np.uint32).reshape(-1, 3)
    cap = len(dispatch) if capacity is None else int(capacity)
    out = np.zeros((max(cap, 1), rec_words), dtype=np.uint32)
    pyr_c = pyramid.c() if pyramid is not None else None
    written, total = C.c_uint32(), C.c_uint32()
    lib().oracle_cluster_cull(C.byref(s), _p(v), C.byref(pyr_c) if pyr_c is not None else None, _p(dispatch), C.c_uint64(len(dispatch)),
                              C.c_int(mode), C.c_int(hiz), _p(out), C.c_uint32(rec_words), C.c_uint64(cap), C.byref(written), C.byref(total), C.c_int(threads))
    return out[:written.value], total.value


def boundary_census(objs, transforms, surfaces, lods, view, pyramid, hiz, ulp_tol=4.0, texel_tol=1.0 / 256.0, transform_id_base=0, threads=1):
    s = _scene(objs, transforms, surfaces, lods, None, 0, transform_id_base)
    v = np.ascontiguousarray(view_with_pyramid(view, pyramid))
    out = (C.c_uint64 * 4)()
    pyr_c = pyramid.c() if pyramid is not None else None
    lib().oracle_boundary_census(C.byref(s), _p(v), C.byref(pyr_c) if pyr_c is not None else None, C.c_int(hiz), C.c_float(ulp_tol), C.c_float(texel_tol), out, C.c_int(threads))
    return {"near_frustum_plane": int(out[0]), "near_texel_boundary": int(out[1]), "near_mip_boundary": int(out[2]), "near_depth_equal": int(out[3])}


def raster_depth(objs, transforms, surfaces, lods, view, records, width, height, object_id_base=0, transform_id_base=0):
    """Software depth from a draw list (oracle_raster_depth): H x W float32, 0 = far."""
    s = _scene(objs, transforms, surfaces, lods, None, object_id_base, transform_id_base)
    v = np.ascontiguousarray(view)
    rec = np.ascontiguousarray(records).view(np.uint32).reshape(len(records), -1)
    out = np.zeros((height, width), dtype=np.float32)
    lib().oracle_raster_depth(C.byref(s), _p(v), _p(rec), C.c_uint64(rec.shape[0]), C.c_uint32(rec.shape[1] if rec.size else 6), C.c_uint32(width), C.c_uint32(height), _p(out))
    return out


def probe(bound_center, bound_radius, transform8, view, pyramid=None, hiz=HIZ_VK, lods=None, lod_offset=0, lod_count=0):
    bc = np.ascontiguousarray(bound_center, dtype=np.float32)
    t8 = np.ascontiguousarray(transform8, dtype=np.float32)
    v = np.ascontiguousarray(view_with_pyramid(view, pyramid))
    out = np.zeros(12, dtype=np.float32)
    pyr_c = pyramid.c() if pyramid is not None else None
    lib().oracle_probe(_p(bc), C.c_float(bound_radius), _p(t8), _p(v), C.byref(pyr_c) if pyr_c is not None else None, C.c_int(hiz),
                       _p(lods) if lods is not None else None, C.c_uint32(lod_offset), C.c_uint32(lod_count), _p(out))
    return out


def ilog2_floor(x):
    return int(lib().oracle_ilog2_floor(C.c_float(x)))
