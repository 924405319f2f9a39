"""Build libials_b200.so (sm_100a only) in-tree with nvcc.

    python -m irspack_b200.build [--force] [--verbose]

The shared library lands in ``irspack_b200/lib/`` (git-ignored, but it travels
with the repo snapshot to the GPU box).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libials_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
    "-DIALS_BUILDING_LIBRARY",
]


def nvcc() -> str:
    exe = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the B200 backend cannot be built")
    return exe


def sources() -> List[str]:
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths: List[str]) -> str:
    h = hashlib.sha1()
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = sorted(
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
        + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    )
    jobs = []
    objs = []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha1"
        digest = _digest([path] + headers)
        objs.append(obj)
        fresh = (
            not force and os.path.exists(obj) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest
        )
        if not fresh:
            jobs.append((path, obj, stamp, digest))

    def compile_one(job):
        path, obj, stamp, digest = job
        cmd = [nvcc()] + ARCH_FLAGS + NVCC_FLAGS + ["-I", INCLUDE, "-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {path}:\n{log}")
        with open(stamp, "w") as f:
            f.write(digest)
        return path, log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for path, log in pool.map(compile_one, jobs):
                if verbose:
                    print(f"== {os.path.basename(path)}\n{log}")
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc()] + ARCH_FLAGS + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(out)
