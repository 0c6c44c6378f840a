"""Builds libeidola.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python build.py            # incremental
    python build.py --force

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  --fmad=false / -ffp-contract=off          the numerical contract: no implicit FMA on either side
  -lineinfo                                 ncu source pages map back to these files
"""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
_VARIANT = os.environ.get("EID_VARIANT", "")         # EID_VARIANT=x -> libeidola_x.so from build_x/ (kernel-variant sweeps)
OUT = os.path.join(HERE, "libeidola%s.so" % ("_" + _VARIANT if _VARIANT else ""))
OBJ = os.path.join(HERE, "build" + ("_" + _VARIANT if _VARIANT else ""))
SOURCES = ["accel.cu", "render.cu", "k_direct.cu", "k_indirect.cu", "k_wave.cu", "k_post.cu", "group.cu", "pipeline.cu", "gltf_import.cpp", "png_decode.cpp", "scene_host.cpp", "env_host.cpp", "sah_host.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "-I", INCLUDE, "-I", CSRC]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_library(force=False, verbose=False, extra=()):
    extra = list(extra) + os.environ.get("EID_NVCC_EXTRA", "").split()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    headers.append(os.path.abspath(__file__))
    os.makedirs(OBJ, exist_ok=True)
    procs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + COMMON + list(extra) + ["-x", "cu" if s.endswith(".cu") else "c++", "-c", src, "-o", obj]
            if s.endswith(".cu"):
                cmd += ["-Xptxas", "-v"] if verbose else []
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or force or _newer(objs, OUT):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-ldl", "-lz", "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    t = time.time()
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built %s in %.1fs" % (OUT, time.time() - t))
