"""Build libgpann_cuda.so (sm_100a only) in-tree with nvcc.  Cross-compiles without a GPU.

Each translation unit is compiled to an object in parallel (-gencode arch=compute_100a,code=sm_100a -lineinfo),
then linked into gp-ann_b200/libgpann_cuda.so (static cudart, no other dependency: the driver entry point for
tensor maps is resolved at run time).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment hooks: GPANN_NVCC_EXTRA="-DX=1 ..." adds flags, GPANN_BUILD_TAG=name builds libgpann_cuda_name.so in build_name/
TAG = os.environ.get("GPANN_BUILD_TAG", "")
EXTRA = os.environ.get("GPANN_NVCC_EXTRA", "").split()
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
OUT = os.path.join(HERE, "libgpann_cuda" + ("_" + TAG if TAG else "") + ".so")
SOURCES = ["gpann_api.cu", "abi.cu", "multi.cu", "seg.cu", "seg_api.cu", "bkm.cu", "topk_tc.cu", "kmeans_tc.cu", "stream.cu", "kmeans.cu", "prep.cu"]
HEADERS = ["common.cuh", "distance.cuh", "params.h", "tc_ptx.cuh", "api_internal.h", os.path.join("..", "..", "include", "gpann.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(obj, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, *EXTRA, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, os.path.join(CSRC, src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(f"---- {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj, True
    return obj, False


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or not os.path.exists(OUT):
        subprocess.check_call([_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs, "-ldl", "-lpthread"])
    if not TAG:
        build_dropin(force)
    return OUT


def build_dropin(force=False):
    """libgpann_dropin.so: the C++ drop-in layer (cxx/: recall.h, kmeans_gpu.cpp over the C ABI; cxx/standalone/: the
    reference-shaped defs.h/topn.h/dist.h/kmeans.h for builds without a gp-ann checkout) behind a small C shim, so the
    tests can drive it.  Plain g++, links libgpann_cuda.so via $ORIGIN."""
    cxx = os.path.join(HERE, "cxx")
    sa = os.path.join(cxx, "standalone")
    out = os.path.join(HERE, "libgpann_dropin.so")
    srcs = [os.path.join(cxx, "dropin_shim.cpp"), os.path.join(cxx, "kmeans_gpu.cpp"), os.path.join(sa, "dist.cpp"),
            os.path.join(sa, "defs.cpp"), os.path.join(sa, "points_io.cpp")]
    deps = srcs + [os.path.join(cxx, f) for f in os.listdir(cxx)] + [os.path.join(sa, f) for f in os.listdir(sa)] + [OUT]
    if force or _stale(out, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-I" + cxx, "-I" + sa, "-o", out, *srcs,
                               "-L" + HERE, "-lgpann_cuda", "-Wl,-rpath,$ORIGIN"])
    return out


def build_integration(ref="/root/reference", force=False):
    """tests/_build/libgpann_integration.so: INTEGRATION.md's recipe executed -- the reference's UNMODIFIED
    kmeans_tree_router.cpp / defs.cpp / dist.cpp and its own headers (defs.h, topn.h, dist.h, kmeans.h), with
    `-I cxx` ahead of `-I <gp-ann>/src` so that recall.h is the drop-in's, plus cxx/kmeans_gpu.cpp and libgpann_cuda.so.
    Needs the reference checkout (absent on the GPU box: the prebuilt .so travels).  Returns the path or None."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tests", "_build", "libgpann_integration.so")
    if not os.path.exists(os.path.join(ref, "src", "kmeans_tree_router.cpp")):
        return out if os.path.exists(out) else None
    cxx = os.path.join(HERE, "cxx")
    srcs = [os.path.join(root, "tests", "integration", "router_shim.cpp"), os.path.join(cxx, "kmeans_gpu.cpp"),
            os.path.join(ref, "src", "kmeans_tree_router.cpp"), os.path.join(ref, "src", "defs.cpp"), os.path.join(ref, "src", "dist.cpp")]
    deps = srcs + [os.path.join(cxx, f) for f in os.listdir(cxx)] + [OUT]
    if force or _stale(out, deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-w",
                               "-I" + cxx, "-I" + os.path.join(ref, "src"), "-I" + os.path.join(ref, "external", "parlaylib", "include"),
                               "-o", out, *srcs, "-L" + HERE, "-lgpann_cuda", "-Wl,-rpath," + HERE + ":$ORIGIN/../../gp-ann_b200"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
