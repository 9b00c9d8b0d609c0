"""Reader/writer for the flat scene blob (layout documented in oracle/ref_scene/refscene.cpp)."""
import numpy as np
from . import types as T

MAGIC = 0x535A4C42  # 'BLZS'


def read_blob(path):
    b = np.fromfile(path, dtype=np.uint8)
    hdr = b[:40].view("<u4")
    if hdr[0] != MAGIC or hdr[1] != 1:
        raise ValueError(f"{path}: not a BLZS v1 scene blob")
    n_obj, n_xf, n_sf, n_lod, n_cl, n_li, static_off, n_dyn = (int(x) for x in hdr[2:10])
    off = 40
    out = {"staticTransformOffset": static_off, "dynamicTransformCount": n_dyn}

    def take(dtype, n):
        nonlocal off
        a = b[off:off + dtype.itemsize * n].view(dtype)
        off += dtype.itemsize * n
        return a

    out["view"] = take(T.CameraViewData, 1).copy()
    out["objs"] = take(T.RenderObject, n_obj)
    out["transforms"] = take(T.MeshTransform, n_xf)
    out["surfaces"] = take(T.PrimitiveSurface, n_sf)
    out["lods"] = take(T.LodData, n_lod)
    out["clusters"] = take(T.Cluster, n_cl)
    out["lodInstances"] = take(T.LodInstanceCounter, n_li)
    if off != b.size:
        raise ValueError(f"{path}: trailing bytes ({off} != {b.size})")
    return out


def read_views(blob_path, json_path):
    import json
    b = np.fromfile(blob_path, dtype=np.uint8)
    n = int(b[:4].view("<u4")[0])
    v = b[4:4 + 256 * n].view(T.CameraViewData)
    with open(json_path) as f:
        meta = json.load(f)
    return {name: v[i:i + 1].copy() for i, name in enumerate(meta["names"])}, meta
