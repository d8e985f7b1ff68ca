"""Builds kgl_gene_b200/libkgl_b200.so (the C ABI + sm_100a kernels) in-tree with nvcc. No torch involved."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libkgl_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "kgl_b200_api.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [os.path.join(ROOT, "include", "kgl_b200.h")]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-Xcompiler", "-pthread", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++", "-o", LIB] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libkgl_b200.so")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)          # untracked: the ptxas -v report of this build
    with open(os.path.join(ROOT, "build", "ptxas_report.txt"), "w") as f:
        f.write(proc.stderr)
    return LIB


HOST_LIB = os.path.join(HERE, "libkgl_b200_host.so")
HOST_SOURCES = [os.path.join(HERE, "host", "kgl_b200_vcf_ingest.cpp")]


def build_host(force: bool = False) -> str:
    """libkgl_b200_host.so: the host-only pieces that need neither CUDA nor the reference headers (VCF ingest)."""
    deps = HOST_SOURCES + [os.path.join(HERE, "host", "kgl_b200_vcf_ingest.h")]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", HOST_LIB] + HOST_SOURCES + ["-lz"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed building libkgl_b200_host.so")
    return HOST_LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
