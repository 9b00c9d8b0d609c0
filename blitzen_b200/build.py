"""In-tree native build of the B200 cull library (nvcc, sm_100a only) and the host-side harness libraries.

    python -m blitzen_b200.build            # builds everything that is out of date
    python -m blitzen_b200.build --force

Artifacts (git-ignored, shipped to the GPU box by gpurun):
    blitzen_b200/libblitzen_cull.so   CUDA kernels + C ABI (include/blz_cull.h)
    blitzen_b200/libblz_scene.so      host-only workload generator (csrc/scene_gen.cpp)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")

CUDA_SOURCES = ["capi.cu", "cull_stream.cu", "cull_early.cu", "cull_cluster.cu", "cull_list.cu", "pyramid.cu", "gather.cu", "consume.cu", "interop.cu", "raster_depth.cu"]
CUDA_HEADERS = ["ctx.h", "cull_types.cuh", "cull_math.cuh", "cull_kernels.cuh", "scan_lookback.cuh", os.path.join("..", "..", "include", "blz_cull.h")]
LIB_CULL = os.path.join(PKG, "libblitzen_cull.so")
LIB_SCENE = os.path.join(PKG, "libblz_scene.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # never contract a*b+c: the cull arithmetic is bit-exact against the oracle
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cull(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in CUDA_HEADERS] + [os.path.abspath(__file__)]
    deps += [os.path.join(PKG, "host", f) for f in ("blitzenCudaCull.cpp", "blitzenCudaCull.h")]
    if not force and not _stale(LIB_CULL, deps):
        return LIB_CULL
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs, logs = [], [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        log = open(o + ".log", "w")
        logs.append(log)
        procs.append(subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], stdout=log, stderr=subprocess.STDOUT))
    failed = False
    for s, pr, log in zip(srcs, procs, logs):
        rc = pr.wait()
        log.close()
        text = open(log.name).read()
        if rc != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s}\n{text}\n")
        elif verbose:
            sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    # the C++ host class above the C ABI (host/blitzenCudaCull.cpp) ships in the same library
    host_o = os.path.join(objdir, "blitzenCudaCull.o")
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-c", os.path.join(PKG, "host", "blitzenCudaCull.cpp"), "-o", host_o], check=True)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_CULL, *objs, host_o, "-cudart", "static",
                    "-Xlinker", "--no-undefined"], check=True)
    return LIB_CULL


def build_scene(force=False):
    src = os.path.join(CSRC, "scene_gen.cpp")
    if not force and not _stale(LIB_SCENE, [src]):
        return LIB_SCENE
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread", "-o", LIB_SCENE, src], check=True)
    return LIB_SCENE


def build_all(force=False, verbose=False):
    return build_cull(force, verbose), build_scene(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
