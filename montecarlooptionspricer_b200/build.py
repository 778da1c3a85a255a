"""Build the in-tree native libraries with explicit nvcc / g++ commands (sm_100a only).

  montecarlooptionspricer_b200/libmcp_b200.so          CUDA kernels + the C ABI (include/mcp_b200.h)
  montecarlooptionspricer_b200/libmcp_b200_plugins.so  C++ host plugin classes with the reference's signatures

Run `python -m montecarlooptionspricer_b200.build` (or `__graft_entry__.build()`).  nvcc cross-compiles without a
GPU; the built .so files are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libmcp_b200.so")
LIB_DBG = os.path.join(PKG, "libmcp_b200_dbg.so")   # same library with MCP_DEBUG_BOUNDS index checks in the sweep kernels
PLUGINS = os.path.join(PKG, "libmcp_b200_plugins.so")
DEMO = os.path.join(HOST, "plugin_rows_demo")
LATENCY = os.path.join(HOST, "plugin_latency")

CU_SOURCES = ["ctx.cu", "pathset.cu", "gen_rbergomi.cu", "gen_gbm.cu", "lsm.cu", "pricers.cu", "estimators.cu", "surface.cu", "rows.cu", "dual.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """One object per translation unit (compiled in parallel, rebuilt only when it or a header changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "mcp_b200.h")]
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    extra = os.environ.get("MCP_NVCC_EXTRA", "").split()   # e.g. -DMCP_DEBUG_BOUNDS=1

    def one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or extra or not _newer(obj, [src] + hdrs):
            cmd = [_nvcc()] + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.run(cmd, check=True, env=env, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(one, srcs))
    if force or extra or not _newer(LIB, objs):
        subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB, "-ldl"], check=True, env=env)
    return LIB


def build_debug(force: bool = False) -> str:
    """libmcp_b200_dbg.so: lsm.cu recompiled with -DMCP_DEBUG_BOUNDS=1 (index checks + ring canaries in the asynchronous sweep
    kernels, see csrc/lsm.cu), linked with the release objects of the other translation units.  Loaded by
    tests/test_gpu_debug_bounds.py through MCP_B200_LIB; compute-sanitizer is closed on the development pool."""
    build_cuda(force=False)
    objdir, dbgdir = os.path.join(PKG, "build"), os.path.join(PKG, "build", "dbg")
    os.makedirs(dbgdir, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    src = os.path.join(CSRC, "lsm.cu")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "mcp_b200.h")]
    obj = os.path.join(dbgdir, "lsm.o")
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    if force or not _newer(obj, [src] + hdrs):
        subprocess.run([_nvcc()] + flags + ["-DMCP_DEBUG_BOUNDS=1", "-c", src, "-o", obj], check=True, env=env, cwd=CSRC)
    objs = [obj] + [os.path.join(objdir, s[:-3] + ".o") for s in CU_SOURCES if s != "lsm.cu"]
    if force or not _newer(LIB_DBG, objs):
        subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_DBG, "-ldl"], check=True, env=env)
    return LIB_DBG


def build_plugins(force: bool = False) -> str:
    """C++ host plugin classes (reference signatures) over the C ABI + the PredictionGen-shaped demo caller."""
    src = os.path.join(HOST, "mcp_plugins.cpp")
    if not os.path.exists(src):
        return ""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST]
    if force or not _newer(PLUGINS, [src, os.path.join(HOST, "mcp_plugins.hpp"), os.path.join(ROOT, "include", "mcp_b200.h"), LIB]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", *inc, src, "-o", PLUGINS, "-L", PKG, "-lmcp_b200",
                        "-Wl,-rpath,$ORIGIN"], check=True, env=env)
    demo_src = os.path.join(HOST, "plugin_rows_demo.cpp")
    if os.path.exists(demo_src) and (force or not _newer(DEMO, [demo_src, PLUGINS])):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wall", *inc, demo_src, "-o", DEMO, "-L", PKG, "-lmcp_b200_plugins",
                        "-lmcp_b200", "-Wl,-rpath,$ORIGIN/.."], check=True, env=env)
    lat_src = os.path.join(HOST, "plugin_latency.cpp")
    if os.path.exists(lat_src) and (force or not _newer(LATENCY, [lat_src, PLUGINS])):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", *inc, lat_src, "-o", LATENCY, "-L", PKG, "-lmcp_b200_plugins", "-lmcp_b200",
                        "-Wl,-rpath,$ORIGIN/.."], check=True, env=env)
    return PLUGINS


def build_all(force: bool = False, verbose: bool = False, debug: bool = True) -> None:
    build_cuda(force=force, verbose=verbose)
    build_plugins(force=force)
    if debug:
        build_debug(force=force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--no-debug" not in sys.argv)
    print(LIB)
