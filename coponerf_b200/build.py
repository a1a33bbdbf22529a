"""Build libcoponerf_b200.so (sm_100a) in-tree with nvcc.

    python -m coponerf_b200.build [--force]

The shared library lands next to this file so that it travels with the repo
snapshot; nothing is JIT-compiled at import time.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcoponerf_b200.so")
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")

# (source, extra flags). geometry.cu evaluates the reference's fp32/fp64 expressions op by op,
# so FMA contraction is off for that file only.
SOURCES = [
    ("geometry.cu", ["-fmad=false"]),
    ("gather.cu", []),
    ("gemm_simt.cu", []),
    ("gemm_tc.cu", []),
    ("attention.cu", []),
    ("weights.cu", []),
    ("render.cu", []),
    ("ufc_tail.cu", []),
    ("conv4d.cu", []),
    ("linear_attention.cu", []),
    ("ufc_ops.cu", []),
    ("pose_feat.cu", []),
]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libcoponerf_b200.so")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the C-ABI shared library. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "coponerf_b200.h"))
    objs = []
    procs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    if force or procs or _stale(LIB, objs):
        tmp = f"{LIB}.{os.getpid()}.tmp"      # link aside and rename: a concurrent dlopen never sees a partial file
        cmd = [nvcc, "-shared", "-o", tmp] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
        os.replace(tmp, LIB)
    return LIB


def build_library_locked(**kw):
    """build_library under an inter-process file lock (several ranks of a torchrun job may find the library missing)."""
    import fcntl
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build_library(**kw)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
