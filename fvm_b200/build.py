"""Build libfvmgpu.so (nvcc, sm_100a) in-tree.  `python -m fvm_b200.build`

The product library is built ONLY by nvcc for sm_100a. `build_hostsim()` builds the TEST-ONLY
single-threaded simulator of the same kernel functors (tests/hostsim/, see csrc/common.cuh); it is
never loaded by the package.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfvmgpu.so")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_LIB = os.path.join(HOSTSIM_DIR, "libfvmgpu_hostsim.so")

# (source, extra flags).  Assembly kernels are compiled without FMA contraction so that the same
# sequence of IEEE operations as the reference's x86-64 build gives the same bits.
SOURCES = [
    ("runtime.cu", []),
    ("comm.cu", []),
    ("peer.cu", []),
    ("mesh.cu", ["-fmad=false"]),
    ("assemble.cu", ["-fmad=false"]),
    ("solver.cu", []),
    ("flow.cu", ["-fmad=false"]),
    ("electric.cu", ["-fmad=false"]),
    ("ilu.cu", ["-fmad=false"]),
    ("capi.cu", []),
]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
SYS_STDCXX = "/usr/lib/x86_64-linux-gnu/libstdc++.so.6"


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _cuda_lib_dir(nvcc):
    return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "lib64")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "fvmgpu.h"))
    return hs


def build_lib(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libfvmgpu.so next to this file."""
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    log = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + _headers()):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    if force or _stale(LIB, objs):
        link = ["g++", "-shared", "-o", LIB] + objs
        if os.path.exists(SYS_STDCXX):
            # the image's g++ wrapper only finds a static libstdc++; a private copy inside a python
            # process that also loads the system one is fragile, so link the shared one explicitly
            link += ["-nostdlib++", SYS_STDCXX]
        link += ["-L" + _cuda_lib_dir(nvcc), "-lcudart_static", "-ldl", "-lrt", "-lpthread", "-lm"]
        r = subprocess.run(link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    if verbose:
        sys.stderr.write("".join(log))
    with open(os.path.join(objdir, "ptxas.log"), "a") as f:
        f.write("".join(log))
    return LIB


def build_hostsim(force=False):
    """TEST-ONLY: single-threaded simulator of the kernel functors (no CUDA)."""
    os.makedirs(HOSTSIM_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s, _ in SOURCES]
    if not (force or _stale(HOSTSIM_LIB, srcs + _headers())):
        return HOSTSIM_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DFVMGPU_HOSTSIM", "-ffp-contract=off",
           "-o", HOSTSIM_LIB]
    for s in srcs:
        cmd += ["-x", "c++", s]
    if os.path.exists(SYS_STDCXX):
        cmd += ["-x", "none", "-nostdlib++", SYS_STDCXX, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostsim build failed:\n" + r.stderr[-6000:])
    return HOSTSIM_LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--hostsim" in sys.argv:
        print(build_hostsim(force=True))
