"""Build libpykmer_b200.so in-tree with nvcc for sm_100a (B200).

    python -m pykmer_b200.build [--force]

The shared library lands next to this file so that it travels to the GPU box
with the repository snapshot; there is no JIT cache and no other architecture.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpykmer_b200.so")
SOURCES = ["api.cu", "indexer.cu", "merger.cu", "gram_i8.cu", "gram_f4.cu", "ingest.cpp", "unpack.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math", "-shared",
    "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC or put /usr/local/cuda/bin on PATH")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pykmer_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-lz", "-lpthread"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd))
    return LIB


TOOLS = os.path.join(os.path.dirname(HERE), "tools")
MICROBENCHES = ("microbench", "microbench2", "microbench3")


def build_microbench(force: bool = False) -> str:
    """tools/microbench, tools/microbench2: raw atomic / popcount / streaming rates of the box
    (measurement tools, not product code)."""
    out = ""
    for name in MICROBENCHES:
        src, out = os.path.join(TOOLS, name + ".cu"), os.path.join(TOOLS, name)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
            continue
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-o", out, src]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout)
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
