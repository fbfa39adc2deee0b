"""Build libavsiam_b200.so in-tree with nvcc for sm_100a (no torch dependency, plain C-ABI).

    python -m avsiam_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libavsiam_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v",
          "--expt-relaxed-constexpr"] + os.environ.get("AVS_EXTRA_NVCC_FLAGS", "").split()


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths: list[str]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(ARCH_FLAGS + CFLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "avsiam_b200.h"))
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(hdrs)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        stamp = o + ".stamp"
        want = _digest([s]) + hdr_digest
        have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(o) else ""
        objs.append(o)
        if force or have != want:
            jobs.append((s, o, stamp, want))

    def compile_one(job):
        s, o, stamp, want = job
        cmd = [NVCC, *ARCH_FLAGS, *CFLAGS, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        with open(o + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(want)
        if verbose:
            print(r.stderr)
        return o

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(OUT):
        # -cudart shared: the CUDA runtime is NOT linked in statically (no runtime code or symbol strings in the .so)
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-cudart", "shared", "-o", OUT, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print("built", p)
