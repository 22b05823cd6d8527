"""Builds libzling_b200/libzling.so IN-TREE with nvcc for sm_100a: CUDA kernels + host engine + C ABI (include/zlb.h)
+ the C++ drop-in API (include/libzling/libzling.h).  Cross-compiles without a GPU."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzling.so")
SOURCES = ["zl_engine.cu", "zl_api.cpp"]
DEPS = SOURCES + ["zl_kernels.cu", "zl_shard.cuh", "zl_mtf_walk.h", "zl_kernels.cuh", "zl_parse_v4.cuh", "zl_tables.h", "../../include/zlb.h",
                  "../../include/libzling/libzling.h", "../../include/libzling/libzling_utils.h"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


# what runs on the device and how it is launched (the C++ driver above the C ABI and the public headers cannot change a kernel's traffic)
DEVICE_DEPS = ["zl_engine.cu", "zl_kernels.cu", "zl_kernels.cuh", "zl_parse_v4.cuh", "zl_mtf_walk.h", "zl_shard.cuh", "zl_tables.h"]


def source_sha16():
    """sha256 prefix over the device-side sources of the library (the build itself is not bit-reproducible: nvcc embeds temporary
    names), in a fixed order: identifies the CODE a kernel measurement was taken from"""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEVICE_DEPS):
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(d.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-fvisibility=default", "-shared", "-cudart", "static", "-o", LIB]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if os.environ.get("ZL_V4_PROFILE"):       # profiling build: per-thread timers inside the parse kernel's decide step
        cmd += ["-DZL_V4_PROFILE=1"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
