"""Build libminimp3_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m minimp3_b200.build [--force] [--verbose]

The shared object links the CUDA runtime statically and has no Python / torch dependency: it is the
C-ABI boundary declared in include/minimp3_b200.h.  A second, test-only object (tests/_build/
libmp3b_cpu_harness.so) exposes the host parser and the __host__ __device__ math to the CPU tests.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libminimp3_b200.so")
HARNESS = os.path.join(ROOT, "tests", "_build", "libmp3b_cpu_harness.so")

CU_SOURCES = ["k_entropy.cu", "k_stereo.cu", "k_transform.cu", "k_layer12.cu", "k_sideinfo.cu", "k_order.cu", "k_scan.cu", "capi.cu"]
CPP_SOURCES = ["host_parse.cpp", "ex_api.cpp"]
HEADERS = ["l3_types.h", "l3_math.cuh", "l3_entropy.cuh", "l3_sideinfo.cuh", "l12_decode.cuh", "l3_tables_build.h", "kernels.h", "host_parse.h", "mp3_hdr.cuh", "mp3_tables.h",
           os.path.join("..", "..", "include", "minimp3_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libminimp3_b200.so cannot be built")
    return p


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES + CPP_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _newer(LIB, deps):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-ffp-contract=off", "-I", CSRC] + ARCH
    if verbose:
        common += ["-Xptxas", "-v"]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + common + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static", "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return LIB


def build_cpu_harness(force=False):
    """Test-only: host parser + HD math compiled for the CPU (g++, no CUDA)."""
    src = os.path.join(ROOT, "tests", "cpu_harness.cpp")
    deps = [src, os.path.join(CSRC, "host_parse.cpp"), os.path.join(CSRC, "host_parse.h"),
            os.path.join(CSRC, "l3_math.cuh"), os.path.join(CSRC, "mp3_tables.h"), os.path.join(CSRC, "l3_types.h"),
            os.path.join(CSRC, "l3_entropy.cuh"), os.path.join(CSRC, "l12_decode.cuh"), os.path.join(CSRC, "l3_tables_build.h"),
            os.path.join(CSRC, "l3_sideinfo.cuh"), os.path.join(CSRC, "mp3_hdr.cuh")]
    if not force and not _newer(HARNESS, deps):
        return HARNESS
    os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", CSRC, "-x", "c++", src,
           os.path.join(CSRC, "host_parse.cpp"), "-o", HARNESS, "-lpthread"]
    subprocess.check_call(cmd)
    return HARNESS


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
