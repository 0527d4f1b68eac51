"""Builds graspit_b200/libb2c.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2c.so")
SOURCES = ["api.cu", "kernels.cu", "collide.cu", "contacts.cu", "contact_order.cu", "anneal.cu", "autograsp.cu", "bvh.cpp", "contacts_host.cpp", "soft_contact_host.cpp"]
# contact_order.cu shares its arithmetic with the host replay (csrc/obb_order.h): no fused multiply-adds, so that both sides round alike
EXTRA_FLAGS = {"contact_order.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-Xcompiler", "-O3"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b2c.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines/out: tuning variants (e.g. defines=("DLEAF_TRIGGER=16",), out=".../libb2c_t16.so") for A/B runs;
    the product build takes neither."""
    if out is None and not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ whose static libstdc++ misbehaves in dlopen()ed libraries
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        cmd = [_nvcc(), "-ccbin", ccbin, *NVCC_FLAGS, *EXTRA_FLAGS.get(s, []), *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd, env=env)
        objs.append(o)
    target = out or LIB
    cmd = [_nvcc(), "-ccbin", ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", target, *objs, "-lcudart"]
    subprocess.check_call(cmd, env=env)
    return target


def build_host_planner_check() -> str:
    """tests/host/b200planner_check: the C++ planner host (graspit_b200/host/b200Planner.cpp) + its test harness, linked
    with libb2c.so.  Needs nothing from the reference."""
    root = os.path.dirname(HERE)
    out = os.path.join(HERE, "build", "b200planner_check")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cudart = None
    for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(d, "libcudart.so")):
            cudart = d
    cmd = [ccbin, "-std=c++14", "-O2", "-g", "-rdynamic", "-Wall", "-I", os.path.join(root, "include"), "-I", os.path.join(HERE, "host"),
           os.path.join(HERE, "host", "b200Planner.cpp"), os.path.join(root, "tests", "host", "b200planner_check.cpp"),
           "-o", out, "-L", HERE, "-lb2c", "-Wl,-rpath,$ORIGIN/.."]
    if cudart:
        cmd += ["-L", cudart, "-lcudart", f"-Wl,-rpath,{cudart}"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
