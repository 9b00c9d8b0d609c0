#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Regenerates every reference-derived fixture committed in this repo.

Runs ONLY where /root/reference exists (the build container).  It compiles the reference's own CPU frontend
(oracle/Makefile target `ref` -> oracle/_ref/refscene) and executes it to produce:

  blitzen_b200/data/stress_mesh_tables.blob   surfaces / LODs / clusters / lodInstance tables of the four bundled OBJ meshes
                                              (output of the reference's LoadMeshFromObj + meshoptimizer pipeline)
  blitzen_b200/data/ref_views.blob + .json    CameraViewData blocks produced by the reference's SetupCamera for the named views
  tests/golden/stress_head_4k.blob            first 4096 objects of the reference's RenderingStressTest scene
  tests/golden/stress_checksums.json          sha256 of the full 4 101 001-object scene arrays (objects, transforms)

None of these are reference SOURCES; they are outputs of running the reference's code on its bundled assets.
"""
import hashlib, json, math, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BLZ_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

CUBE_16M = 3000.0 * (16777216 / 4.1e6) ** (1.0 / 3.0)
CUBE_64M = 2000.0 * (67108864 / 4.1e6) ** (1.0 / 3.0)

# name -> (fovDeg, winW, winH, zNear, px, py, pz, zFar, yawRad, pitchRad)
VIEWS = {
    "default":            (70, 1280, 720, 0.1, 20, 70, 0, 650, 0, 0),              # Core/blitzenEngine.h:40-45
    "cfg1_centre":        (70, 1280, 720, 0.1, 1500, 1500, 1500, 1e4, 0, 0),
    "cfg1_all":           (70, 1280, 720, 0.1, 1500, 1500, -9000, 1e9, 0, 0),
    "cfg1_tilted":        (70, 1280, 720, 0.1, 900, 2100, 400, 3000, 0.7, -0.3),
    "cfg2_centre_1080p":  (70, 1920, 1080, 0.1, CUBE_16M / 2, CUBE_16M / 2, CUBE_16M / 2, 1e4, 0, 0),
    "cfg2_corner_1080p":  (70, 1920, 1080, 0.1, 20, 70, 0, 2500, 0.6, 0.35),
    "cfg3_centre":        (70, 1280, 720, 0.1, CUBE_64M / 2, CUBE_64M / 2, CUBE_64M / 2, 1e4, 0, 0),
}
for k in range(8):  # config 5: main + 7 cascades, yaw steps of 45 degrees, zFar doubles per cascade, 3840x2160
    VIEWS[f"cfg5_view{k}"] = (70, 3840, 2160, 0.1, CUBE_16M / 2, CUBE_16M / 2, CUBE_16M / 2, 650.0 * 2 ** k, k * math.pi / 4, 0)


def run(*a):
    subprocess.run(list(a), check=True)


def main():
    run("make", "-C", HERE, "ref")
    exe = os.path.join(HERE, "_ref", "refscene")
    data = os.path.join(ROOT, "blitzen_b200", "data")
    golden = os.path.join(ROOT, "tests", "golden")
    os.makedirs(data, exist_ok=True); os.makedirs(golden, exist_ok=True)

    run(exe, REF, "meshes", os.path.join(data, "stress_mesh_tables.blob"))

    spec = os.path.join(HERE, "_ref", "views.txt")
    with open(spec, "w") as f:
        for name, v in VIEWS.items():
            f.write(" ".join(repr(float(np.float32(x))) for x in v) + "\n")
    run(exe, REF, "views", os.path.join(data, "ref_views.blob"), spec)
    with open(os.path.join(data, "ref_views.json"), "w") as f:
        json.dump({"names": list(VIEWS.keys()), "params": {k: [float(np.float32(x)) for x in v] for k, v in VIEWS.items()},
                   "columns": ["fovDeg", "winW", "winH", "zNear", "px", "py", "pz", "zFar", "yawRad", "pitchRad"]}, f, indent=1)

    run(exe, REF, "stress", os.path.join(golden, "stress_head_4k.blob"), "4096")

    full = os.path.join(HERE, "_ref", "stress_full.blob")
    run(exe, REF, "stress", full)
    from blitzen_b200 import sceneio
    sc = sceneio.read_blob(full)
    sums = {"nObjects": int(len(sc["objs"])), "nTransforms": int(len(sc["transforms"])),
            "objs_sha256": hashlib.sha256(sc["objs"].tobytes()).hexdigest(),
            "transforms_sha256": hashlib.sha256(sc["transforms"].tobytes()).hexdigest()}
    with open(os.path.join(golden, "stress_checksums.json"), "w") as f:
        json.dump(sums, f, indent=1)
    os.remove(full)
    print(sums)


if __name__ == "__main__":
    main()
