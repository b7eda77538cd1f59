"""Builds libsmcb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmcb200.so")
SOURCES = ["api.cu", "loglik_mm.cu", "temper.cu", "resample.cu", "mh.cu", "kinetic.cu", "dae.cu", "comm.cu", "sweep.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def header_version():
    """SMCB_ABI_VERSION of include/smcb200.h (what smcb_version() of a matching library returns)."""
    import re
    with open(os.path.join(HERE, "..", "include", "smcb200.h")) as f:
        return int(re.search(r"#define\s+SMCB_ABI_VERSION\s+(\d+)", f.read()).group(1))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, "..", "include", "smcb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source and link the shared library.  Returns its path.

    Safe under concurrent callers (one process per GPU imports the package at the same time): the build runs under
    an exclusive file lock, staleness is re-checked inside it, objects go to a per-process directory and the library
    is moved into place atomically."""
    if not force and not _stale():
        return LIB
    import fcntl
    import shutil
    import tempfile
    objroot = os.path.join(HERE, "build")
    os.makedirs(objroot, exist_ok=True)
    with open(os.path.join(objroot, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another process built it while this one waited
                return LIB
            nvcc = _nvcc()
            objdir = tempfile.mkdtemp(prefix="obj_", dir=objroot)

            def one(src):
                obj = os.path.join(objdir, src.replace(".cu", ".o"))
                cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
                if verbose:
                    sys.stderr.write(r.stderr)
                return obj

            try:
                with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
                    objs = list(ex.map(one, SOURCES))
                tmp_lib = os.path.join(objdir, "libsmcb200.so")
                cmd = [nvcc, "-shared", "-o", tmp_lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart",
                                                                 "static", "-ldl"]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
                os.replace(tmp_lib, LIB)
            finally:
                shutil.rmtree(objdir, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_user_library(src, out=None, force=False):
    """Compile a user-written likelihood (one .cu including include/smcb_user.cuh) into a shared library for sm_100a.
    Returns the path of the library."""
    src = os.path.abspath(src)
    out = out or os.path.join(os.path.dirname(src), "lib" + os.path.splitext(os.path.basename(src))[0] + ".so")
    inc = os.path.abspath(os.path.join(HERE, "..", "include"))
    deps = [src, os.path.join(inc, "smcb_user.cuh"), os.path.join(inc, "smcb200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-I", inc, src, "-o", out, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
