"""Build the C-ABI shared library (in-tree) with nvcc for sm_100a.

    python -m dfot_b200.build          # or: from dfot_b200.build import build; build()

nvcc cross-compiles without a GPU.  The resulting ``libdfot_b200.so`` is git-ignored
but travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdfot_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["abi.cu", "sampler.cu", "norm.cu", "embed.cu", "uvit.cu", "vae.cu", "dcae.cu", "gemm_tcgen05.cu", "attention_tcgen05.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the dfot_b200 kernels cannot be built")
    return nvcc


HASH_FILE = os.path.join(OBJ, "sources.sha256")     # travels with the .so (build/ is git-ignored, not gpurun-ignored)


def _dep_files():
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "dfot_b200.h"))
    return [os.path.join(CSRC, s) for s in SOURCES], headers


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, flags): mtimes do not survive a snapshot
    copy to another machine, contents do."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    srcs, headers = _dep_files()
    for f in srcs + headers:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_current() -> bool:
    """True when libdfot_b200.so exists and was built from exactly the sources now in the tree."""
    if not (os.path.exists(LIB) and os.path.exists(HASH_FILE)):
        return False
    with open(HASH_FILE) as fh:
        return fh.read().strip() == source_hash()


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs, headers = _dep_files()
    if not force and is_current():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # one builder at a time: the ranks of a torchrun launch all load the library at once, and every one of them would
    # otherwise rebuild a stale library into the same files
    import fcntl
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():       # another process finished the build while this one waited
                return LIB
            return _build_locked(srcs, headers, force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(srcs, headers, force: bool, verbose: bool) -> str:
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)                         # atomic: a concurrent loader sees the old or the new file, never half of one
    with open(HASH_FILE, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
