"""Build libnanorq_b200.so in-tree: C host code (gcc) + sm_100a kernels (nvcc).

    python -m nanorq_b200.build

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the
repository snapshot (it is git-ignored, not gpurun-ignored)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnanorq_b200.so")
RT_OUT = os.path.join(HERE, "librq_roundtrip.so")
C_SOURCES = ["rqb_planner.c", "rqb_solver.c", "nanorq_api.c", "rqb_io.c"]
CU_SOURCES = ["rqb_device.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    hdrs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    inc = ["-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    for src in C_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(objdir, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [CC, "-O3", "-march=x86-64-v3", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-pthread", "-c", s, "-o", o] + inc
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    for src in CU_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(objdir, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + ARCH + os.environ.get("RQB_NVCC_EXTRA", "").split() + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c", s, "-o", o] + inc
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(OUT, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    # the nanorq.h-only round-trip harness (bench/rq_roundtrip.c), linked against the library
    rt_src = os.path.join(ROOT, "bench", "rq_roundtrip.c")
    if os.path.exists(rt_src) and (force or _stale(RT_OUT, [rt_src, OUT] + hdrs)):
        cmd = [CC, "-O2", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-shared", "-pthread", "-o", RT_OUT, rt_src,
               "-I" + os.path.join(ROOT, "include"), "-L" + HERE, "-lnanorq_b200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    # the same round trips through the batch calls of nanorq_batch.h (this library only)
    rtb_src = os.path.join(ROOT, "bench", "rq_roundtrip_batch.c")
    rtb_out = os.path.join(HERE, "librq_roundtrip_batch.so")
    if os.path.exists(rtb_src) and (force or _stale(rtb_out, [rtb_src, OUT] + hdrs)):
        cmd = [CC, "-O2", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-shared", "-pthread", "-o", rtb_out, rtb_src,
               "-I" + os.path.join(ROOT, "include"), "-L" + HERE, "-lnanorq_b200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
