"""In-tree build of the C-ABI shared library with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "capi.cu")
OUT = os.path.join(HERE, "libgrape_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    d = os.path.join(HERE, "csrc")
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + \
           [os.path.join(inc, f) for f in sorted(os.listdir(inc))]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


PEAKS_SRC = os.path.join(HERE, "csrc", "peaks.cu")
PEAKS_OUT = os.path.join(HERE, "libgrape_peaks.so")


def build_peaks(force=False):
    if not force and os.path.exists(PEAKS_OUT) and os.path.getmtime(PEAKS_OUT) >= os.path.getmtime(PEAKS_SRC):
        return PEAKS_OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    r = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", PEAKS_OUT, PEAKS_SRC], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libgrape_peaks.so")
    return PEAKS_OUT


def build(force=False, verbose=False):
    build_peaks(force)
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libgrape_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
