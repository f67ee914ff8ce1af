"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  lidar_slam_b200/_lib/libb2ndt.so    CUDA kernels + C ABI (include/b2ndt.h), sm_100a only
  lidar_slam_b200/_lib/libb2host.so   C++ drop-in classes (NDTRegistration / VoxelFilter) over the C ABI
  lidar_slam_b200/_lib/libb2synth.so  host-only synthetic HDL-64 workload generator
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for c in ("nvcc", "/usr/local/cuda/bin/nvcc"):
        p = shutil.which(c)
        if p:
            return p
    raise RuntimeError("nvcc not found")


def build_synth(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libb2synth.so")
    src = os.path.join(CSRC, "synth_hdl64.c")
    if force or _newer(out, [src]):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src,
                               "-lm", "-lpthread"])
    return out


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False, variant=None, variant_flags=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libb2ndt.so (cross-compiles without a GPU).
    `variant` builds a tuning variant libb2ndt_<variant>.so with extra -D flags (selected at run time with
    the environment variable B2NDT_LIB; used by tools/ for kernel-shape sweeps only)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libb2ndt%s.so" % ("_" + variant if variant else ""))
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "b2ndt.h"))
    extra = os.environ.get("B2_NVCC_EXTRA", "").split() + list(variant_flags)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-shared", "-o", out] + srcs
    cmd += ["-lcudart"]
    log = os.path.join(LIBDIR, "nvcc_ptxas%s.log" % ("_" + variant if variant else ""))
    # the library is rebuilt when a source is newer OR when it was built with a different command line
    # (e.g. an experiment's -D flags): the first line of the log is the stamp
    try:
        with open(log) as f:
            same_cmd = f.readline().rstrip("\n") == " ".join(cmd)
    except OSError:
        same_cmd = False
    if force or not same_cmd or _newer(out, deps):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
        if verbose or r.returncode != 0:
            print(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed (see %s)" % log)
    return out


def build_host(force=False):
    """C++ drop-in classes (reference interface mirror) linked against libb2ndt.so."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libb2host.so")
    hdir = os.path.join(CSRC, "host")
    srcs = sorted(os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".cpp")) if os.path.isdir(hdir) else []
    if not srcs:
        return None
    hdrs = []
    for d, _, fs in os.walk(os.path.join(ROOT, "include")):
        hdrs += [os.path.join(d, f) for f in fs]
    if force or _newer(out, srcs + hdrs):
        cmd = ["g++", "-O3", "-std=c++14", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", out] + srcs
        cmd += ["-L", LIBDIR, "-lb2ndt", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return out


def build_hostcheck(force=False):
    """Test shim: csrc/b2_ndt_math.cuh compiled for the host (no GPU needed)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libb2hostcheck.so")
    src = os.path.join(CSRC, "hostcheck", "b2_hostcheck.cpp")
    if force or _newer(out, [src, os.path.join(CSRC, "b2_ndt_math.cuh")]):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
                               "-Wno-unused-function", "-Wno-unknown-pragmas", "-o", out, src])
    return out


def build_all(force=False, verbose=False):
    build_synth(force)
    build_cuda(force, verbose)
    build_host(force)
    build_hostcheck(force)
