"""Builds liblumb200.so (the C-ABI device library) in-tree with nvcc for sm_100a.

Two flag sets are used on purpose:
  * geometry TUs (bvh_build.cu, trace.cu): -fmad=false, IEEE div/sqrt. World-space vertices, camera rays and the
    watertight triangle test must be bit-identical to the CPU oracle, which is built with -ffp-contract=off.
  * shading TU (shade.cu): --use_fast_math like the reference (src/luminary/CMakeLists.txt:48).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblumb200.so")
NVCC = os.environ.get("LUMB200_NVCC", "/usr/local/cuda/bin/nvcc")

COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-I", os.path.join(HERE, "..", "include"),
]
EXACT = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"]
FAST = ["--use_fast_math"]

HOST_CC = "/usr/bin/gcc"
HOST_UNITS = ["host/light_tree.c"]

UNITS = [
    ("bvh_build.cu", EXACT),
    ("trace.cu", EXACT),
    ("shade.cu", FAST),
    ("sky.cu", FAST),
    ("device_api.cu", []),
    ("comm.cu", []),
]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(verbose=False, force=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "lumb200.h"))
    objs = []
    for name, flags in UNITS:
        src = os.path.join(CSRC, name)
        obj = os.path.join(CSRC, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    for name in HOST_UNITS:
        src = os.path.join(CSRC, name)
        obj = src.replace(".c", ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [HOST_CC, "-O2", "-std=gnu11", "-fPIC", "-Wall", "-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _newer(objs, OUT):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    build_host(verbose=verbose, force=force)
    return OUT


# The C host layer (Luminary's public API, include/luminary/luminary.h) and the headless benchmark front end. Both only
# use the C ABI of liblumb200.so (include/lumb200.h); $ORIGIN rpaths keep the three artefacts relocatable together.
HOST_LIB = os.path.join(HERE, "libluminary_b200.so")
HOST_CLI = os.path.join(HERE, "LuminaryB200")
HOST_API_UNITS = ["host/lum_host.c", "host/lum_scene_file.c", "host/lum_wavefront.c", "host/lum_png.c", "host/lum_utils.c"]


def build_host(verbose=False, force=False):
    inc = os.path.join(HERE, "..", "include")
    srcs = [os.path.join(CSRC, u) for u in HOST_API_UNITS]
    deps = srcs + [os.path.join(CSRC, "host", "lum_host_internal.h"), os.path.join(inc, "lumb200.h"), OUT]
    deps += [os.path.join(inc, "luminary", f) for f in os.listdir(os.path.join(inc, "luminary"))]
    if force or _newer(deps, HOST_LIB):
        cmd = [HOST_CC, "-O2", "-std=gnu11", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", inc, "-o", HOST_LIB] + srcs + [
            "-L", HERE, "-llumb200", "-Wl,-rpath,$ORIGIN", "-lpthread", "-ldl", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    cli_src = os.path.join(CSRC, "host", "lum_cli_main.c")
    if force or _newer([cli_src, HOST_LIB], HOST_CLI):
        cmd = [HOST_CC, "-O2", "-std=gnu11", "-Wall", "-Wextra", "-I", inc, "-o", HOST_CLI, cli_src, "-L", HERE, "-lluminary_b200", "-llumb200",
               "-Wl,-rpath,$ORIGIN", "-lpthread", "-lm"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return HOST_LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(OUT)
