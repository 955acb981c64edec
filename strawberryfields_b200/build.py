"""Build libb200fock.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m strawberryfields_b200.build

The shared library has a plain C ABI (include/b200fock.h) and is loaded with ctypes.
It is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200fock.so")
SOURCES = ["api.cu", "gates_gen.cu", "apply.cu", "tile.cu", "generic.cu", "exchange.cu", "inner.cu", "gram.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-Xptxas", "-v",
]
OBJ_DIR = os.path.join(HERE, "build")
HEADERS = [os.path.join(CSRC, h) for h in ("common.cuh", "blocks.cuh", "tasks.cuh", "tma.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "b200fock.h")]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + HEADERS)


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel, only the stale ones unless ``force``) and link the
    shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in HEADERS)

    def stale(src):
        o = _obj(src)
        return force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(src), newest_header)

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", _obj(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    todo = [s for s in sources() if stale(s)]
    with ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        results = list(ex.map(compile_one, todo))
    log = "".join(r[0] for r in results)
    rc = max([r[1] for r in results] + [0])
    if rc == 0:
        cmd = [nvcc, "--shared", "-o", LIB] + [_obj(s) for s in sources()]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        rc = res.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
