"""Build the C-ABI shared library in-tree with nvcc for sm_100a.

    python -m smallpebble_b200.build [--force] [--verbose]

Output: smallpebble_b200/_lib/libsmallpebble_b200.so (git-ignored; travels to the GPU box
with the gpurun snapshot).  No torch headers, no pybind: plain `extern "C"`.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libsmallpebble_b200.so")
FASTCALL_PATH = os.path.join(OUT_DIR, "_sp_fastcall.so")  # CPython binding (csrc/fastcall.c)
STAMP = os.path.join(OUT_DIR, "build.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")
    )
    files.append(os.path.join(os.path.dirname(HERE), "include", "smallpebble_b200.h"))
    files.append(os.path.join(CSRC, "fastcall.c"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build smallpebble_b200's CUDA library")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    digest = _digest()
    if (not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP)
            and os.path.exists(FASTCALL_PATH)):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, _sources()))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as fh:
        for obj, log in results:
            fh.write(f"==== {os.path.basename(obj)}\n{log}\n")
    objs = [o for o, _ in results]
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_fastcall()
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB_PATH


def build_fastcall() -> str:
    """The CPython fast-call binding: plain C, no CUDA (gcc + the interpreter's headers)."""
    import sysconfig

    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("no C compiler for smallpebble_b200/csrc/fastcall.c")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-I", sysconfig.get_paths()["include"],
           os.path.join(CSRC, "fastcall.c"), "-o", FASTCALL_PATH]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"fastcall build failed:\n{r.stdout}\n{r.stderr}")
    return FASTCALL_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
