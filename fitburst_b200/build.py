"""
Builds fitburst_b200/libfitburst_b200.so in-tree with nvcc for sm_100a (cross-compiles without
a GPU). Usage: `python -m fitburst_b200.build [--force]`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfitburst_b200.so")
SOURCES = ["fb_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in os.listdir(root):
            if name.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libfitburst_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
