"""Build the native parts in-tree (the built .so/.exe files travel to the GPU box with the snapshot).

    python -m bayesrrcmd_b200.build
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def build_cuda(force=False, verbose=False):
    out = os.path.join(PKG, "libbrr_gpu.so")
    srcs = _sources(os.path.join(PKG, "csrc"), (".cu", ".cuh", ".inl")) + [os.path.join(ROOT, "include", "brr_gpu.h")]
    if force or _newer(out, srcs):
        cmd = [nvcc()] + NVCC_FLAGS + ["-shared", "-o", out, os.path.join(PKG, "csrc", "brr_gpu.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return out


def build_host(force=False):
    """C++ host mirror (libbrr_host.so + the brr_gpu command line), links against libbrr_gpu.so."""
    hdir = os.path.join(PKG, "host")
    srcs = _sources(hdir, (".cpp", ".h", ".hpp"))
    cpps = [s for s in srcs if s.endswith(".cpp") and not s.endswith("main.cpp")]
    if not cpps:
        return None
    lib = os.path.join(PKG, "libbrr_host.so")
    exe = os.path.join(PKG, "brr_gpu")
    if force or _newer(lib, srcs) or _newer(exe, srcs):
        common = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-deprecated-declarations",
                  "-I" + os.path.join(ROOT, "include"), "-I" + hdir]
        subprocess.check_call(common + ["-shared", "-o", lib] + cpps + ["-L" + PKG, "-lbrr_gpu", "-lz", "-Wl,-rpath,$ORIGIN"])
        main = os.path.join(hdir, "main.cpp")
        if os.path.exists(main):
            subprocess.check_call(common + ["-o", exe, main, "-L" + PKG, "-lbrr_host", "-lbrr_gpu", "-lz", "-Wl,-rpath,$ORIGIN"])
    return lib


def build_all(force=False):
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", os.path.join(PKG, "libbrr_gpu.so"))
