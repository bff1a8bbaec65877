"""Build the CPU oracle and, when /root/reference is mounted, the compiled reference itself.

TEST INFRASTRUCTURE (see oracle/gpann_oracle.cpp header).  Outputs:

  oracle/_build/libgpann_oracle.so   -- the restatement (oracle/gpann_oracle.cpp)
  oracle/_ref/libgpann_ref.so        -- UNMODIFIED reference sources, L2 build       (needs /root/reference)
  oracle/_ref/libgpann_ref_mips.so   -- same with -DMIPS_DISTANCE (CMakeLists.txt:54-59)

The reference is compiled from where its sources lie: /root/reference/src/{dist,defs,kmeans}.cpp plus the
headers recall.h / inverted_index.h, vendored parlaylib and hnswlib headers, with oracle/ref_shim.cpp as the
only file of ours in the link.  Flags follow the reference's release build (CMakeLists.txt:11,15:
-std=c++20 -O3 -ffast-math) with ONE deliberate change: -march=x86-64-v3 -mtune=generic instead of
-march=native, because the .so is built in the authoring container and must also run on the GPU box's host
CPU, whose ISA extensions are unknown.  x86-64-v3 defines __AVX__ and FMA, so src/dist.cpp:11-16 takes the
same 256-bit hnswlib AVX kernel with FMA contraction that -march=native selects on any AVX2 host.

Both _ref/ and _build/ are git-ignored but NOT gpurun-ignored, so the binaries travel to the GPU box,
where /root/reference does not exist.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GPANN_REFERENCE", "/root/reference")
BUILD = os.path.join(HERE, "_build")
REFOUT = os.path.join(HERE, "_ref")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "gpann_oracle.cpp")
    out = os.path.join(BUILD, "libgpann_oracle.so")
    if not force and _newer(out, [src]):
        return out
    # No -ffast-math: the oracle's arithmetic order is the documented one, deterministically.
    cmd = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-mtune=generic", "-ffp-contract=off", "-fopenmp",
           "-fPIC", "-shared", "-o", out, src]
    subprocess.check_call(cmd)
    return out


def have_reference():
    return os.path.exists(os.path.join(REF, "src", "recall.h"))


def build_ref(force=False):
    """Compile the reference hot path.  Returns the list of built libraries ([] when /root/reference is absent
    and nothing prebuilt exists)."""
    outs = []
    os.makedirs(REFOUT, exist_ok=True)
    shim = os.path.join(HERE, "ref_shim.cpp")
    for name, defs in (("libgpann_ref.so", []), ("libgpann_ref_mips.so", ["-DMIPS_DISTANCE"])):
        out = os.path.join(REFOUT, name)
        if not have_reference():
            if os.path.exists(out):
                outs.append(out)  # prebuilt, travelled with the snapshot
            continue
        srcs = [shim] + [os.path.join(REF, "src", f) for f in ("dist.cpp", "defs.cpp", "kmeans.cpp")]
        if not force and _newer(out, srcs):
            outs.append(out)
            continue
        cmd = ["g++", "-std=c++20", "-O3", "-march=x86-64-v3", "-mtune=generic", "-ffast-math", "-fPIC", "-shared",
               "-pthread", "-w", *defs, "-I" + os.path.join(REF, "src"),
               "-I" + os.path.join(REF, "external", "parlaylib", "include"), "-o", out, *srcs]
        subprocess.check_call(cmd)
        outs.append(out)
    return outs


def build_router_ref(force=False):
    """oracle/_ref/libgpann_router_ref.so: tests/integration/router_shim.cpp over the UNMODIFIED reference only
    (its recall.h, kmeans.cpp, kmeans_tree_router.cpp, defs.cpp, dist.cpp) -- the checker for tests/test_integration.py."""
    out = os.path.join(REFOUT, "libgpann_router_ref.so")
    if not have_reference():
        return out if os.path.exists(out) else None
    shim = os.path.join(os.path.dirname(HERE), "tests", "integration", "router_shim.cpp")
    srcs = [shim] + [os.path.join(REF, "src", f) for f in ("kmeans_tree_router.cpp", "kmeans.cpp", "defs.cpp", "dist.cpp")]
    if not force and _newer(out, srcs):
        return out
    os.makedirs(REFOUT, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O3", "-march=x86-64-v3", "-mtune=generic", "-ffast-math", "-fPIC", "-shared", "-pthread",
                           "-w", "-I" + os.path.join(REF, "src"), "-I" + os.path.join(REF, "external", "parlaylib", "include"), "-o", out,
                           *srcs])
    return out


def main():
    force = "--force" in sys.argv
    print(build_oracle(force))
    for o in build_ref(force):
        print(o)
    print(build_router_ref(force))


if __name__ == "__main__":
    main()
