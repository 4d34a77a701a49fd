"""Build libcvcs_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m cvcs_b200.build [--force] [--verbose]

Every .cu under cvcs_b200/csrc is compiled to an object in parallel and linked into
cvcs_b200/libcvcs_b200.so.  The .so is git-ignored but travels to the GPU box.  nvcc
cross-compiles without a GPU.  There is no fallback: if the library is missing the Python
layer raises at import of cvcs_b200._lib.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libcvcs_b200.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-fvisibility=hidden",
    "-I",
    INCLUDE,
    "-I",
    CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths: list[str], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return hs


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False, defines=(), tag: str = "") -> str:
    """defines / tag: experiment builds (`-DNAME` flags) go to libcvcs_b200_<tag>.so with their own
    object directory; the default build takes neither."""
    global OBJ_DIR, LIB_PATH
    if tag:
        OBJ_DIR = os.path.join(PKG_DIR, "build_" + tag)
        LIB_PATH = os.path.join(PKG_DIR, f"libcvcs_b200_{tag}.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    flags = NVCC_FLAGS + ARCH_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + [f"-D{d}" for d in defines]
    headers = _headers()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src] + headers, " ".join(flags))
        objs.append(obj)
        fresh = (not force and os.path.exists(obj) and os.path.exists(stamp)
                 and open(stamp).read().strip() == dig)
        if not fresh:
            jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(dig)
        return src, r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose or ptxas_info:
                    print(f"[cvcs_b200.build] compiled {os.path.relpath(src, ROOT)}", file=sys.stderr)
                    if log.strip():
                        print(log, file=sys.stderr)
    if jobs or force or not os.path.exists(LIB_PATH):
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[cvcs_b200.build] linked {os.path.relpath(LIB_PATH, ROOT)}", file=sys.stderr)
    return LIB_PATH


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas-info", action="store_true", help="pass -Xptxas -v (registers, spills, smem)")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="experiment build: extra -D define")
    ap.add_argument("--tag", default="", help="experiment build: library suffix")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, ptxas_info=a.ptxas_info, defines=a.defines, tag=a.tag))


if __name__ == "__main__":
    main()
