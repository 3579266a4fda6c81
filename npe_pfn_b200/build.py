"""Builds the C-ABI shared library (`include/npe_pfn_b200.h`) in-tree with nvcc for sm_100a.

The built `.so` lives at `npe_pfn_b200/_lib/libnpe_pfn_b200.so` (git-ignored, shipped to the GPU box by
`gpurun`).  nvcc cross-compiles without a GPU, so this also runs in the CPU-only build check.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libnpe_pfn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "128",
]


def _sources():
    out = []
    for root, _d, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(_HERE, "..", "include", "npe_pfn_b200.h"))
    return out


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for path in sorted(os.path.realpath(p) for p in _sources()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def files_hash(names) -> str:
    """Hash of the named files under csrc/ (+ the nvcc flags): identifies the build of ONE kernel, e.g. the item-attention
    kernel whose ncu DRAM traffic bench.py reports only while the kernel's sources are the profiled ones."""
    import hashlib
    h = hashlib.sha256()
    for n in sorted(names):
        h.update(n.encode())
        with open(os.path.join(CSRC, n), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


#: sources that define the item-attention kernel of test rows (roofline.traffic in bench.py)
ATTN_KERNEL_FILES = ["attn_tc.cuh", "attn_mma.cuh", "common.cuh"]

HASH_PATH = LIB_PATH + ".srchash"


def is_stale() -> bool:
    """Content based (file times do not survive being copied to the GPU box): the library is current when the
    hash of the sources it was built from equals the hash of the sources in the tree."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    try:
        return open(HASH_PATH).read().strip() != _source_hash()
    except OSError:
        return True


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libnpe_pfn_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = list(NVCC_FLAGS)
    if os.path.exists(os.path.join(CSRC, "attn_tc.cuh")):
        flags.append("-DPFN_WITH_ATTN_TC")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "engine.cu")]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
