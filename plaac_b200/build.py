"""In-tree build of libplaac_cuda.so (sm_100a only) and the host CLI.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libplaac_cuda.so")
CLI = os.path.join(HERE, "bin", "plaac")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # Java never contracts a*b+c; neither may the device code
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(HERE, "..", "include", "plaac_cuda.h"))
    out.append(os.path.join(HERE, "..", "include", "plaac_bench.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, _sources()):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-shared", "-o", LIB,
               os.path.join(CSRC, "plaac_cuda.cu"), os.path.join(CSRC, "bench_utils.cu"),
               os.path.join(CSRC, "host_params.cpp"), os.path.join(CSRC, "host_pack.cpp")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB


def build_cli(force: bool = False) -> str:
    src = os.path.join(HERE, "host", "plaac_cli.cpp")
    if not os.path.exists(src):
        return ""
    deps = [src] + [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))]
    if force or _stale(CLI, deps + [LIB]):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cxx = "/usr/bin/g++"
        cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-o", CLI, src,
               "-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-lplaac_cuda", "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return CLI


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_cli(force)


if __name__ == "__main__":
    import sys

    build_lib(force="-f" in sys.argv, verbose="-v" in sys.argv)
    build_cli(force="-f" in sys.argv)
    print(LIB)
