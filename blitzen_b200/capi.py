"""ctypes binding of the C ABI in include/blz_cull.h (libblitzen_cull.so).

This is the Python face of the product path.  It never falls back to a CPU implementation: if the CUDA library is
missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import types as T

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLZ_CULL_LIB") or os.path.join(_PKG, "libblitzen_cull.so")   # BLZ_CULL_LIB: developer override (A/B builds of the same library)

LIST_OPAQUE, LIST_TRANSPARENT, LIST_ONPC = 0, 1, 2
REC_VK24, REC_DX32 = 0, 1
HIZ_VK, HIZ_DX = 0, 1
CLUSTER_PASSTHROUGH, CLUSTER_SPHERE, CLUSTER_SPHERE_HIZ = 0, 1, 2
FLAG_ONPC_LOD_QUIRK = 1

REC_DTYPE = {REC_VK24: T.IndirectDrawVK, REC_DX32: T.DrawCmdDX}


class BlzError(RuntimeError):
    pass


class SceneDesc(C.Structure):
    _fields_ = [
        ("renders", C.c_void_p), ("render_count", C.c_uint32),
        ("transparent_renders", C.c_void_p), ("transparent_count", C.c_uint32),
        ("onpc_renders", C.c_void_p), ("onpc_count", C.c_uint32),
        ("transforms", C.c_void_p), ("transform_count", C.c_uint32),
        ("surfaces", C.c_void_p), ("surface_count", C.c_uint32),
        ("lods", C.c_void_p), ("lod_count", C.c_uint32),
        ("clusters", C.c_void_p), ("cluster_count", C.c_uint32),
        ("lod_instances", C.c_void_p), ("lod_instance_count", C.c_uint32),
        ("object_id_base", C.c_uint32), ("transform_id_base", C.c_uint32),
        ("draw_capacity", C.c_uint64), ("cluster_dispatch_capacity", C.c_uint64),
        ("instance_bucket_capacity", C.c_void_p),
        ("inputs_on_device", C.c_uint32), ("reserved", C.c_uint32),
    ]


class ConsumeSummary(C.Structure):
    _fields_ = [("records", C.c_uint64), ("index_sum", C.c_uint64), ("instance_sum", C.c_uint64), ("id_sum", C.c_uint64), ("id_xor", C.c_uint64),
                ("bad_object", C.c_uint32), ("bad_lod", C.c_uint32), ("unsorted", C.c_uint32), ("pad", C.c_uint32), ("lod_hist", C.c_uint32 * 256)]


class ExportedOutputs(C.Structure):
    _fields_ = [("draws_fd", C.c_int), ("counts_fd", C.c_int), ("draws_alloc_bytes", C.c_uint64), ("counts_alloc_bytes", C.c_uint64),
                ("draw_capacity_records", C.c_uint64), ("count_offset_bytes", C.c_uint64), ("generation", C.c_uint32), ("pad", C.c_uint32)]


class Outputs(C.Structure):
    _fields_ = [
        ("draws", C.c_void_p), ("draw_count", C.c_void_p), ("visibility", C.c_void_p),
        ("cluster_dispatch", C.c_void_p), ("cluster_count", C.c_void_p),
        ("instance_indices", C.c_void_p), ("instance_counts", C.c_void_p),
        ("pyramid", C.c_void_p),
        ("pyramid_width", C.c_uint32), ("pyramid_height", C.c_uint32), ("pyramid_mips", C.c_uint32),
        ("pyramid_offset", C.c_uint32 * 16),
        ("draw_capacity", C.c_uint64), ("cluster_dispatch_capacity", C.c_uint64),
    ]


# every symbol include/blz_cull.h declares (tests/test_abi.py checks the library exports exactly these)
ABI_SYMBOLS = [
    "blz_cull_abi_version", "blz_cull_last_error", "blz_cull_create", "blz_cull_destroy", "blz_cull_set_stream",
    "blz_cull_get_stream", "blz_cull_synchronize", "blz_cull_upload_scene", "blz_cull_update_transforms", "blz_cull_set_view",
    "blz_cull_reset_visibility", "blz_cull_write_visibility", "blz_cull_set_depth", "blz_cull_set_depth_device",
    "blz_cull_build_pyramid", "blz_cull_clear_pyramid", "blz_cull_frustum_lod", "blz_cull_early", "blz_cull_late",
    "blz_cull_temporal", "blz_cull_instanced", "blz_cull_cluster_expand", "blz_cull_cluster_cull",
    "blz_cull_set_cluster_dispatch", "blz_cull_get_outputs", "blz_cull_read_draws", "blz_cull_read_count",
    "blz_cull_read_visibility", "blz_cull_read_cluster_dispatch", "blz_cull_read_instances", "blz_cull_read_pyramid",
    "blz_cull_gather_export", "blz_cull_gather_import", "blz_cull_gather_configure", "blz_cull_gather_push", "blz_cull_gather_push_async", "blz_cull_gather_join",
    "blz_cull_gather_read", "blz_cull_gather_outputs", "blz_cull_instances_export", "blz_cull_instances_import", "blz_cull_instances_push", "blz_cull_instances_counts", "blz_cull_consume_draws", "blz_cull_consume_instances", "blz_cull_consume_gathered", "blz_cull_launch_count", "blz_cull_set_option",
    "blz_cull_export_outputs", "blz_cull_export_fence", "blz_cull_signal_fence", "blz_cull_import_semaphore", "blz_cull_signal_semaphore",
    "blz_cull_raster_depth", "blz_cull_read_depth",
    "blz_interop_import", "blz_interop_release", "blz_interop_wait_fence", "blz_interop_read",
]

_lib = None


def load_library():
    """Loads libblitzen_cull.so.  Raises if it has not been built (python -m blitzen_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BlzError(f"{LIB_PATH} is missing: build it with `python -m blitzen_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.blz_cull_last_error.restype = C.c_char_p
    vp, u32, u64, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "blz_cull_create": [i, C.POINTER(vp)], "blz_cull_destroy": [vp], "blz_cull_set_stream": [vp, vp],
        "blz_cull_get_stream": [vp, C.POINTER(vp)], "blz_cull_synchronize": [vp],
        "blz_cull_upload_scene": [vp, C.POINTER(SceneDesc)], "blz_cull_update_transforms": [vp, u32, u32, vp],
        "blz_cull_set_view": [vp, vp], "blz_cull_reset_visibility": [vp], "blz_cull_write_visibility": [vp, vp],
        "blz_cull_set_depth": [vp, vp, u32, u32], "blz_cull_set_depth_device": [vp, vp, u32, u32],
        "blz_cull_build_pyramid": [vp, i], "blz_cull_clear_pyramid": [vp, i, u32, u32],
        "blz_cull_frustum_lod": [vp, i, i, u32], "blz_cull_early": [vp, i], "blz_cull_late": [vp, i, i],
        "blz_cull_temporal": [vp, i, i, i], "blz_cull_instanced": [vp, i], "blz_cull_cluster_expand": [vp, i],
        "blz_cull_cluster_cull": [vp, i, i, i], "blz_cull_set_cluster_dispatch": [vp, vp, u64, i],
        "blz_cull_get_outputs": [vp, C.POINTER(Outputs)],
        "blz_cull_read_draws": [vp, i, vp, u64, C.POINTER(u32), C.POINTER(u32)],
        "blz_cull_read_count": [vp, C.POINTER(u32), C.POINTER(u32)], "blz_cull_read_visibility": [vp, vp],
        "blz_cull_read_cluster_dispatch": [vp, vp, u64, C.POINTER(u32), C.POINTER(u32)],
        "blz_cull_read_instances": [vp, vp, u64, vp], "blz_cull_read_pyramid": [vp, vp, u64, vp, vp],
        "blz_cull_gather_export": [vp, u64, i, vp], "blz_cull_gather_import": [vp, vp, i, i],
        "blz_cull_gather_configure": [vp, u64, i], "blz_cull_gather_push": [vp, u32], "blz_cull_gather_push_async": [vp, u32], "blz_cull_gather_join": [vp],
        "blz_cull_gather_read": [vp, u32, vp, u64, vp], "blz_cull_gather_outputs": [vp, C.POINTER(vp), C.POINTER(vp)],
        "blz_cull_instances_export": [vp, vp], "blz_cull_instances_import": [vp, vp, i, i], "blz_cull_instances_push": [vp, vp, vp, vp], "blz_cull_instances_counts": [vp, vp],
        "blz_cull_consume_draws": [vp, i, i, vp], "blz_cull_consume_instances": [vp, i, vp], "blz_cull_consume_gathered": [vp, u32, vp],
        "blz_cull_launch_count": [vp, C.POINTER(u64)], "blz_cull_set_option": [vp, C.c_char_p, C.c_int64],
        "blz_cull_export_outputs": [vp, C.POINTER(ExportedOutputs)], "blz_cull_export_fence": [vp, vp], "blz_cull_signal_fence": [vp],
        "blz_cull_import_semaphore": [vp, i, i], "blz_cull_signal_semaphore": [vp, u64],
        "blz_cull_raster_depth": [vp, i, u32, u32], "blz_cull_read_depth": [vp, vp, u64, vp],
        "blz_interop_import": [i, i, u64, C.POINTER(vp)], "blz_interop_release": [vp, u64], "blz_interop_wait_fence": [vp, vp],
        "blz_interop_read": [vp, vp, u64, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.blz_cull_abi_version.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype != dtype:
        if a.dtype.itemsize * a.size != dtype.itemsize * (a.size * a.dtype.itemsize // dtype.itemsize):
            raise ValueError("array size is not a multiple of the record size")
        a = a.view(dtype)
    return a


class CullContext:
    """One GPU's cull backend (mirror of the reference backend's SetupForRendering + cull dispatches)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.blz_cull_create(int(device), C.byref(h)))
        self._h = h
        self.device = device
        self.n_objects = 0
        self.n_lods = 0
        self._keep = []
        self._last_fmt = REC_VK24
        # A/B knobs for measurements (bench.py / profiling scripts), e.g. BLZ_OPTIONS="cull_items=4,pyramid_tma=0"
        for kv in filter(None, os.environ.get("BLZ_OPTIONS", "").split(",")):
            k, v = kv.split("=")
            self.set_option(k.strip(), int(v))

    def _check(self, rc):
        if rc != 0:
            raise BlzError(f"blz_cull error {rc}: {self._lib.blz_cull_last_error().decode()}")

    def close(self):
        if self._h is not None:
            self._lib.blz_cull_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- stream ----------------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.blz_cull_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def get_stream(self):
        s = C.c_void_p()
        self._check(self._lib.blz_cull_get_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self):
        self._check(self._lib.blz_cull_synchronize(self._h))

    # ---- scene -----------------------------------------------------------------------------------------------------
    def upload_scene(self, objs, transforms, surfaces, lods, clusters=None, lod_instances=None, transparent=None, onpc=None,
                     object_id_base=0, transform_id_base=0, draw_capacity=0, cluster_dispatch_capacity=0, bucket_capacity=None):
        objs = _as(objs, T.RenderObject); transforms = _as(transforms, T.MeshTransform)
        surfaces = _as(surfaces, T.PrimitiveSurface); lods = _as(lods, T.LodData)
        clusters = _as(clusters, T.Cluster); lod_instances = _as(lod_instances, T.LodInstanceCounter)
        transparent = _as(transparent, T.RenderObject); onpc = _as(onpc, T.RenderObject)
        cap = None if bucket_capacity is None else np.ascontiguousarray(bucket_capacity, dtype=np.uint32)
        d = SceneDesc()
        d.renders, d.render_count = _ptr(objs), 0 if objs is None else len(objs)
        d.transparent_renders, d.transparent_count = _ptr(transparent), 0 if transparent is None else len(transparent)
        d.onpc_renders, d.onpc_count = _ptr(onpc), 0 if onpc is None else len(onpc)
        d.transforms, d.transform_count = _ptr(transforms), len(transforms)
        d.surfaces, d.surface_count = _ptr(surfaces), len(surfaces)
        d.lods, d.lod_count = _ptr(lods), len(lods)
        d.clusters, d.cluster_count = _ptr(clusters), 0 if clusters is None else len(clusters)
        d.lod_instances, d.lod_instance_count = _ptr(lod_instances), 0 if lod_instances is None else len(lod_instances)
        d.object_id_base, d.transform_id_base = int(object_id_base), int(transform_id_base)
        d.draw_capacity, d.cluster_dispatch_capacity = int(draw_capacity), int(cluster_dispatch_capacity)
        d.instance_bucket_capacity = _ptr(cap)
        d.inputs_on_device = 0
        self._check(self._lib.blz_cull_upload_scene(self._h, C.byref(d)))
        self.n_objects = d.render_count
        self.n_lods = d.lod_count

    def update_transforms(self, first, transforms):
        transforms = _as(transforms, T.MeshTransform)
        self._check(self._lib.blz_cull_update_transforms(self._h, int(first), len(transforms), _ptr(transforms)))

    def set_view(self, view):
        v = np.ascontiguousarray(view).view(np.uint8).reshape(-1)
        if v.size != 256:
            raise ValueError("CameraViewData must be 256 bytes")
        self._check(self._lib.blz_cull_set_view(self._h, _ptr(v)))

    def reset_visibility(self):
        self._check(self._lib.blz_cull_reset_visibility(self._h))

    def write_visibility(self, vis):
        vis = np.ascontiguousarray(vis, dtype=np.uint32)
        if len(vis) != self.n_objects:
            raise ValueError("visibility length mismatch")
        self._check(self._lib.blz_cull_write_visibility(self._h, _ptr(vis)))

    # ---- depth / pyramid -------------------------------------------------------------------------------------------
    def set_depth(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        h, w = depth.shape
        self._check(self._lib.blz_cull_set_depth(self._h, _ptr(depth), w, h))

    def set_depth_device(self, dev_ptr, width, height):
        self._check(self._lib.blz_cull_set_depth_device(self._h, C.c_void_p(dev_ptr), width, height))

    def raster_depth(self, width, height, list_id=LIST_OPAQUE):
        """Software depth from the current draw list (csrc/raster_depth.cu); becomes the depth image build_pyramid reads."""
        self._check(self._lib.blz_cull_raster_depth(self._h, list_id, width, height))

    def read_depth(self):
        wh = (C.c_uint32 * 2)()
        self._check(self._lib.blz_cull_read_depth(self._h, None, 0, wh))
        out = np.zeros((wh[1], wh[0]), dtype=np.float32)
        self._check(self._lib.blz_cull_read_depth(self._h, _ptr(out), out.size, wh))
        return out

    def build_pyramid(self, variant=HIZ_VK):
        self._check(self._lib.blz_cull_build_pyramid(self._h, variant))

    def clear_pyramid(self, variant, depth_width, depth_height):
        self._check(self._lib.blz_cull_clear_pyramid(self._h, variant, depth_width, depth_height))

    # ---- passes ----------------------------------------------------------------------------------------------------
    def frustum_lod(self, list_id=LIST_OPAQUE, fmt=REC_VK24, flags=0):
        self._check(self._lib.blz_cull_frustum_lod(self._h, list_id, fmt, flags))
        self._last_fmt = fmt

    def early(self, fmt=REC_VK24):
        self._check(self._lib.blz_cull_early(self._h, fmt))
        self._last_fmt = fmt

    def late(self, fmt=REC_VK24, hiz=HIZ_VK):
        self._check(self._lib.blz_cull_late(self._h, fmt, hiz))
        self._last_fmt = fmt

    def temporal(self, list_id=LIST_OPAQUE, fmt=REC_VK24, hiz=HIZ_VK):
        self._check(self._lib.blz_cull_temporal(self._h, list_id, fmt, hiz))
        self._last_fmt = fmt

    def instanced(self, list_id=LIST_OPAQUE):
        self._check(self._lib.blz_cull_instanced(self._h, list_id))
        self._last_fmt = REC_DX32

    def cluster_expand(self, list_id=LIST_OPAQUE):
        self._check(self._lib.blz_cull_cluster_expand(self._h, list_id))

    def cluster_cull(self, mode=CLUSTER_PASSTHROUGH, fmt=REC_VK24, hiz=HIZ_VK):
        self._check(self._lib.blz_cull_cluster_cull(self._h, mode, fmt, hiz))
        self._last_fmt = fmt

    def set_cluster_dispatch(self, records):
        records = _as(records, T.ClusterDispatchData)
        self._check(self._lib.blz_cull_set_cluster_dispatch(self._h, _ptr(records), len(records), 0))

    # ---- outputs ---------------------------------------------------------------------------------------------------
    def outputs(self):
        o = Outputs()
        self._check(self._lib.blz_cull_get_outputs(self._h, C.byref(o)))
        return o

    def read_count(self):
        w, t = C.c_uint32(), C.c_uint32()
        self._check(self._lib.blz_cull_read_count(self._h, C.byref(w), C.byref(t)))
        return w.value, t.value

    def read_draws(self, fmt=None, capacity=None):
        """Returns (records, total): records is a structured array of the `written` records.  fmt defaults to the format of the pass
        this object ran last; the library refuses a format other than the one the last pass wrote (the buffer sizes differ)."""
        if fmt is None:
            fmt = self._last_fmt
        w, t = self.read_count()
        n = w if capacity is None else min(w, capacity)
        out = np.zeros(n, dtype=REC_DTYPE[fmt])
        w2, t2 = C.c_uint32(), C.c_uint32()
        self._check(self._lib.blz_cull_read_draws(self._h, int(fmt), _ptr(out) if n else None, n, C.byref(w2), C.byref(t2)))
        return out, t2.value

    def read_visibility(self):
        out = np.zeros(max(self.n_objects, 1), dtype=np.uint32)
        self._check(self._lib.blz_cull_read_visibility(self._h, _ptr(out)))
        return out[:self.n_objects]

    def read_cluster_dispatch(self):
        w, t = C.c_uint32(), C.c_uint32()
        self._check(self._lib.blz_cull_read_cluster_dispatch(self._h, None, 0, C.byref(w), C.byref(t)))
        out = np.zeros(w.value, dtype=T.ClusterDispatchData)
        if w.value:
            self._check(self._lib.blz_cull_read_cluster_dispatch(self._h, _ptr(out), w.value, C.byref(w), C.byref(t)))
        return out, t.value

    def read_instances(self, capacity):
        idx = np.zeros(capacity, dtype=np.uint32)
        counters = np.zeros(self.n_lods, dtype=T.LodInstanceCounter)
        self._check(self._lib.blz_cull_read_instances(self._h, _ptr(idx), capacity, _ptr(counters)))
        return idx, counters

    def read_pyramid(self):
        whm = np.zeros(3, dtype=np.uint32); offs = np.zeros(16, dtype=np.uint32)
        self._check(self._lib.blz_cull_read_pyramid(self._h, None, 0, _ptr(whm), _ptr(offs)))
        w, h, m = (int(x) for x in whm)
        texels = sum(max(1, w >> i) * max(1, h >> i) for i in range(m))
        data = np.zeros(texels, dtype=np.float32)
        self._check(self._lib.blz_cull_read_pyramid(self._h, _ptr(data), texels, _ptr(whm), _ptr(offs)))
        return data, (w, h, m), offs

    # ---- multi-GPU gather ------------------------------------------------------------------------------------------
    def gather_export(self, capacity_records, fmt=REC_VK24):
        blob = np.zeros(128, dtype=np.uint8)
        self._check(self._lib.blz_cull_gather_export(self._h, int(capacity_records), fmt, _ptr(blob)))
        return blob

    def gather_import(self, blob, rank, world, capacity_records, fmt=REC_VK24):
        b = None if blob is None else np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self._lib.blz_cull_gather_import(self._h, _ptr(b), rank, world))
        self._check(self._lib.blz_cull_gather_configure(self._h, int(capacity_records), fmt))

    def gather_push(self, epoch):
        self._check(self._lib.blz_cull_gather_push(self._h, int(epoch)))

    def gather_push_async(self, epoch):
        self._check(self._lib.blz_cull_gather_push_async(self._h, int(epoch)))

    def gather_join(self):
        self._check(self._lib.blz_cull_gather_join(self._h))

    def gather_read(self, epoch, world, fmt=REC_VK24, capacity=None):
        counts = np.zeros(world, dtype=np.uint32)
        self._check(self._lib.blz_cull_gather_read(self._h, int(epoch), None, 0, _ptr(counts)))
        n = int(counts.sum()) if capacity is None else min(int(counts.sum()), capacity)
        out = np.zeros(n, dtype=REC_DTYPE[fmt])
        if n:
            self._check(self._lib.blz_cull_gather_read(self._h, int(epoch), _ptr(out), n, _ptr(counts)))
        return out, counts

    def instances_export(self):
        blob = np.zeros(64, dtype=np.uint8)
        self._check(self._lib.blz_cull_instances_export(self._h, _ptr(blob)))
        return blob

    def instances_import(self, blob, rank, world):
        b = None if blob is None else np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self._lib.blz_cull_instances_import(self._h, _ptr(b), rank, world))

    def instances_counts(self, dst_device_ptr):
        self._check(self._lib.blz_cull_instances_counts(self._h, C.c_void_p(dst_device_ptr)))

    def instances_push(self, all_counts_ptr, global_offset_ptr, global_cap_ptr):
        self._check(self._lib.blz_cull_instances_push(self._h, C.c_void_p(all_counts_ptr), C.c_void_p(global_offset_ptr), C.c_void_p(global_cap_ptr)))

    # ---- draw-list consumer ------------------------------------------------------------------------------------------
    def consume_draws(self, list_id=LIST_OPAQUE, kind=0):
        out = ConsumeSummary()
        self._check(self._lib.blz_cull_consume_draws(self._h, list_id, kind, C.byref(out)))
        return out

    def consume_gathered(self, epoch):
        out = ConsumeSummary()
        self._check(self._lib.blz_cull_consume_gathered(self._h, int(epoch), C.byref(out)))
        return out

    def consume_instances(self, list_id=LIST_OPAQUE):
        out = ConsumeSummary()
        self._check(self._lib.blz_cull_consume_instances(self._h, list_id, C.byref(out)))
        return out

    # ---- instrumentation -------------------------------------------------------------------------------------------
    def launch_count(self):
        v = C.c_uint64()
        self._check(self._lib.blz_cull_launch_count(self._h, C.byref(v)))
        return v.value

    # ---- zero-copy export of the outputs (csrc/interop.cu) --------------------------------------------------------------
    def export_outputs(self):
        """Moves draws / counts into exportable allocations and returns an ExportedOutputs (file descriptors owned by the caller)."""
        e = ExportedOutputs()
        self._check(self._lib.blz_cull_export_outputs(self._h, C.byref(e)))
        return e

    def export_fence(self):
        h = (C.c_ubyte * 64)()
        self._check(self._lib.blz_cull_export_fence(self._h, h))
        return bytes(h)

    def signal_fence(self):
        self._check(self._lib.blz_cull_signal_fence(self._h))

    def set_option(self, name, value):
        self._check(self._lib.blz_cull_set_option(self._h, name.encode(), int(value)))
